set -x
# reduced round-end run: tests, smoke, bench line, reference arm, ncu launch list (the full-set page is taken by gpu_round_run.sh)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 2 --warmup 3 --frames 2 --no-cpu-baseline --configs none > gpurun_out/r02_b2.json 2> gpurun_out/r02_b2.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --frames 2 --no-cpu-baseline --configs none > gpurun_out/r02_ncu_b2.log 2>&1
gzip -f gpurun_out/r02_launches_bench_steps2.csv
tools/kernel_times.sh . > gpurun_out/r02_kernel_times.txt 2>&1
ls -la gpurun_out
