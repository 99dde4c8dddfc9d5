set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02_box.txt; nproc >> gpurun_out/r02_box.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 2 --warmup 3 --frames 2 --no-cpu-baseline --configs none > gpurun_out/r02_b2.json 2> gpurun_out/r02_b2.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --frames 2 --no-cpu-baseline --configs none > gpurun_out/r02_ncu_b2.log 2>&1
timeout 300 python tests/gpu_dec_once.py 7680 4320 1 > gpurun_out/r02_dec_once.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none -o /tmp/r02_top python tests/gpu_dec_once.py 7680 4320 1 > gpurun_out/r02_ncu_top.log 2>&1
ncu -i /tmp/r02_top.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_top_raw.csv.gz
gzip -f gpurun_out/r02_launches_bench_steps2.csv
du -sh gpurun_out
ls -la gpurun_out
