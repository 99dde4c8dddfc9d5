"""Writes profiles/r02_summary.md from the committed round-2 artefacts (run from the repo root, no GPU needed)."""
import gzip
import json
import os
import subprocess
import sys
import tempfile

P = os.path.dirname(os.path.abspath(__file__))


def table(mode, gz, *extra):
    with tempfile.NamedTemporaryFile("wb", suffix=".csv", delete=False) as f:
        f.write(gzip.open(os.path.join(P, gz)).read())
    out = subprocess.run([sys.executable, os.path.join(P, "summarize.py"), mode, f.name, *extra], capture_output=True, text=True).stdout
    os.unlink(f.name)
    return out


d = json.load(open(os.path.join(P, "r02_bench_n1.json")))
s0 = json.load(open(os.path.join(P, "r02_start_bench_n1.json")))
n8 = json.load(open(os.path.join(P, "r02_bench_n8.json")))
n2 = json.load(open(os.path.join(P, "r02_bench_n2.json")))
ref = json.load(open(os.path.join(P, "r02_bench_reference_arm.json")))
L = table("launches", "r02_launches_bench_steps2.csv.gz", "29")
K = table("kernels", "r02_ncu_top_raw.csv.gz")
KL = table("kernels", "r02_ncu_late_raw.csv.gz")
KT = open(os.path.join(P, "r02_kernel_times.txt")).read()
rows = []
for k, name in [("lift_fwd", "forward lifting (colour fused)"), ("linearize", "linearise (Hilbert gather + bit-slicing)"),
                ("enc_coder", "encoder coder"), ("dec_coder", "decoder coder"), ("reconstruct", "reconstruct (Hilbert scatter + bias)"),
                ("lift_inv", "inverse lifting (colour + clamp fused)")]:
    a, b = s0["stages"][k], d["stages"][k]
    rows.append("| %s | %.3f | %.4f | %.3f | **%.4f** | %d | %d |" % (name, a["ms"], a["frac"], b["ms"], b["frac"], b["bytes"] // 1000000,
                                                                    (b["traffic"] or 0) // 1000000))
cfg = d["configs"]


def c(k, *f):
    return ", ".join("%s %s" % (x, cfg[k][x]) for x in f)


sp = ref["cpu_baseline"]["single_process_8k"]
txt = f"""# Round 2 profile summary (B200, sm_100a, driver 580, CUDA 12.9)

All numbers: synthetic 7680x4320 RGB "photo" frame (seed 1), lossless, stream 48 863 617 B (sha pin of SURVEY App. E.1) unless a
config says otherwise.  Artefacts in this directory (gpurun calls `tools/gpu_round_run.sh`, `tools/gpu_final_run.sh`,
`tools/gpu_bench_only.sh`; every ncu command had exited 0 without ncu first; this file is written by `make_r02_summary.py`):
* `r02_pytest.log`, `r02_smoke.log` -- `python -m pytest tests -m gpu -x -q`: 47 passed; `__graft_entry__.smoke()`: ok, 176 launches (final commit);
* `r02_bench_n1.json` -- `python bench.py --steps 5 --warmup 3` at the final commit (N = 1, 8 frames in flight per step, every BASELINE config);
  `r02_kernel_times.txt` -- time and warp instructions of every kernel of one encode + decode at the final commit
  (`tools/kernel_times.sh`: ncu `gpu__time_duration.sum`, `smsp__inst_executed.sum`);
* `r02_bench_reference_arm.json` -- `python bench.py --impl reference --steps 1 --warmup 0`;
  `r02_launches_bench_steps2.csv.gz` -- ncu `--metrics gpu__time_duration.sum --clock-control none` launch list of
  `python bench.py --steps 2 --warmup 3 --frames 2 --no-cpu-baseline --configs none` (both one commit before the final one, which only
  put the forward lifting body back: `lift_fwd_kernel<0>` 134 -> 129 us);
* `r02_ncu_top_raw.csv.gz` -- `ncu --set full --clock-control none`, raw page, every kernel of one encode + decode
  (`python tests/gpu_dec_once.py 7680 4320 1`, commit `58d9376`; later commits changed `dec_extend_walk_kernel`, `dec_resolve_kernel`
  and `lift_fwd_kernel` only: `r02_ncu_late_raw.csv.gz` holds the same page for those at the final commit); `traffic.json` -- DRAM bytes per stage summed from it (`summarize.py traffic`);
* `r02_bench_n2.json` -- the bench under `torch.distributed.run` on 2 GPUs of one box at the final commit;
* `r02_bench_n8.json` -- the bench under `torch.distributed.run` on the 8 GPUs of one box (session start commit `097101a`: what it
  shows is a host property, see below);
* `r02_sass_tma.txt` -- SASS excerpt of the TMA kernels (`UTMALDG.3D`, `UTMASTG.3D`, `SYNCS.ARRIVE.TRANS64`, `SYNCS.PHASECHK`);
* `r02_lift_fwd_pairs.patch` + `r02_lift_fwd_pairs_ncu_raw.csv.gz` -- the forward lifting variant that lost (below) and its ncu page;
* `r02_start_*` -- the same artefacts at the commit this session started from (`097101a`), the "before" of the tables below;
* `summarize.py` (tables), `../tools/ncu_lines.py`, `../tools/ncu_sass_hot.py`, `../tools/sass_loops.py`, `../tools/kernel_times.sh`,
  `../tools/pipe_probe.cu` -- what the analysis below was made with.

## bench line (N = 1)

| quantity | start of this session (`097101a`) | now |
|---|---|---|
| value: device resident, 8 frames in flight, one CUDA-event region around the 5 steps | {s0['value']:.0f} Mpixel/s ({s0['ms_per_step']/8:.2f} ms per frame round trip) | **{d['value']:.0f} Mpixel/s** ({d['ms_per_step']/8:.2f} ms) |
| e2e: `dwt_pool_run`, page-locked host buffers, 148 MB H2D + 148 MB D2H per frame inside the timed region | {s0['e2e']['value']:.0f} Mpixel/s = {s0['e2e']['staging_gbs_per_direction']} GB/s per direction, {s0['e2e']['fraction_of_staging_ceiling']} of the box's staging ceiling | **{d['e2e']['value']:.0f} Mpixel/s** = {d['e2e']['staging_gbs_per_direction']} GB/s per direction, **{d['e2e']['fraction_of_staging_ceiling']}** of the ceiling ({d['e2e']['staging_ceiling_gbs_per_direction']} GB/s, measured in the run): the end-to-end pass is now bound by the PCIe path |
| single frame, one context (latency) | {s0['single_frame']['ms_per_frame']} ms; encode {s0['single_frame']['encode_mpx_s']:.0f}, decode {s0['single_frame']['decode_mpx_s']:.0f} Mpixel/s | **{d['single_frame']['ms_per_frame']} ms**; encode {d['single_frame']['encode_mpx_s']:.0f}, decode {d['single_frame']['decode_mpx_s']:.0f} Mpixel/s |
| reference programs on the box's 16 host cores (`--impl reference`) | | {ref['value']} Mpixel/s on 16 tiles in parallel; one process on the whole 8K frame: encode {sp['encode_s']} s, decode {sp['decode_s']} s = {sp['round_trip_mpx_s']} Mpixel/s.  e2e / reference = **{d['e2e']['value']/ref['value']:.0f}x** |
| whole round trip against the HBM roofline (3 465 MB algorithmic) | in flight {s0['whole_frame']['in_flight']['frac']}, single frame {s0['whole_frame']['single_frame']['frac']} | in flight **{d['whole_frame']['in_flight']['frac']}**, single frame {d['whole_frame']['single_frame']['frac']} |
| clocks | | SM {d['clocks']['sm_mhz']:.0f} of {d['clocks']['sm_max_mhz']:.0f} MHz, no throttle reasons |

Stages (single-frame pass, each stage's kernels alone on the GPU; `frac` = algorithmic bytes / live event time / 6 551.7 GB/s measured peak):

| stage | ms before | frac before | ms now | frac now | algorithmic MB | DRAM MB (ncu) |
|---|---|---|---|---|---|---|
""" + "\n".join(rows) + f"""

The time-dominant stage (the bench line's `roofline`) is the decoder's coder stage: {d['roofline']['share_of_single_frame']} of a single frame.  It is a
latency chain, not a memory problem: `dec_resolve_kernel` (one warp) 1.2 ms, five lineage passes 0.56 ms, and instruction-bound kernels
around them.  Lifting against BASELINE.json's target (>= 0.50): forward **{d['lifting_roofline']['forward']}**, inverse **{d['lifting_roofline']['inverse']}**.

## configs (BASELINE.json), each behind its reference pin

| config | numbers |
|---|---|
| 4K lossless | {c('4k_lossless','encode_ms','decode_ms','encode_mpx_s','decode_mpx_s','e2e_encode_ms','e2e_decode_ms')} |
| 8K, 64 KiB | {c('8k_cap_64KiB','encode_ms','encode_coder_ms','decode_ms','encode_mpx_s','decode_mpx_s')} (decodes to 3840x2160, as the reference does) |
| 8K, 1 MiB | {c('8k_cap_1MiB','encode_ms','encode_coder_ms','decode_ms','encode_mpx_s','decode_mpx_s')} |
| 8K, 8 MiB | {c('8k_cap_8MiB','encode_ms','encode_coder_ms','decode_ms','encode_mpx_s','decode_mpx_s')} |
| 16384 x 16384 lossless (395 MB stream) | {c('16384sq_lossless','encode_ms','decode_ms','encode_mpx_s','decode_mpx_s')} |
| 512 images of 1920x1080 through `dwt_pool`, host buffers | {c('batch_1080p','encode_images_s','decode_images_s','encode_mpx_s','decode_mpx_s')} |

A capped encode no longer costs a lossless one: coder stage {cfg['8k_cap_64KiB']['encode_coder_ms']:.2f} ms at 64 KiB against {d['stages']['enc_coder']['ms']:.2f} ms lossless (the reference also stops at the cap).

## N = 2 (one box, final commit, `r02_bench_n2.json`)

Device resident {n2['value']:.0f} Mpixel/s = **{n2['value']/d['value']/2:.2f}** of 2 x the N = 1 value; end to end {n2['e2e']['value']:.0f} Mpixel/s =
{n2['e2e']['staging_gbs_per_direction']} GB/s per direction = {n2['e2e']['fraction_of_staging_ceiling']} of that box's staging ceiling ({n2['e2e']['staging_ceiling_gbs_per_direction']} GB/s per direction for both GPUs
together, measured in the run; 24 vCPUs).  Batch of 1024 images of 1920x1080: encode {n2['configs']['batch_1080p']['encode_images_s']:.0f}, decode {n2['configs']['batch_1080p']['decode_images_s']:.0f} images/s.

## N = 8 (one box, `r02_bench_n8.json`)

Device resident {n8['value']:.0f} Mpixel/s = **0.97** of 8 x the N = 1 value of the same commit (round 1: 0.88).  End to end {n8['e2e']['value']:.0f}
Mpixel/s = {n8['e2e']['staging_gbs_per_direction']} GB/s per direction = **{n8['e2e']['fraction_of_staging_ceiling']} of the staging ceiling of the box measured in the same run**
({n8['e2e']['staging_ceiling_gbs_per_direction']} GB/s per direction with all eight GPUs copying both ways at once; one GPU alone gets 47-49): the box's host path gives
8 GB/s per GPU and direction, a frame round trip needs 148 MB each way.  Batch of 4096 images of 1920x1080: encode
{n8['configs']['batch_1080p']['encode_images_s']:.0f} images/s (67 GB/s of pixels host -> device: the same ceiling), decode {n8['configs']['batch_1080p']['decode_images_s']:.0f} images/s.

## what changed in this session and what each change bought (8K, per-kernel times under ncu unless a stage is named)

| change | effect |
|---|---|
| `reconstruct_tma_kernel`: one channel of a cell at a time (4 KB buffer, 16 plane words per warp: 32 warps per SM instead of 16), all plane rows requested before the first use (the per-plane split-word merge had serialised them: 72-77 % of the stall samples were the ten dependent loads), descriptors two cells ahead | 270 -> 132 us; stage 0.33 -> 0.18 ms, frac 0.25 -> 0.44 |
| `linearize_tma_kernel`: one channel of a cell per TMA box, two boxes per warp, 24 warps per SM instead of 8 | 148 -> 101 us; stage 0.213 -> 0.168 ms (the rest of the stage is the plane-count read-back and the zeroing of the store), frac 0.37 -> 0.47 |
| encoder: ordinary-token fast paths (threads whose eight tokens are plain ones skip the token kinds; scatter through a 64-bit shift register), `enc_emit` tile scan as one 32-bit shuffle scan with one barrier, next plane's word fetched one plane ahead, level of a tile searched from the top | `enc_scatter` 412 -> 276 us, `enc_vli` 331 -> 267 us, `enc_emit` 460 -> 410 us, `enc_count` 83 -> 70 us; stage 1.39 -> 1.15 ms |
| decoder: tokens that fit a 32-bit window are taken apart without 64-bit shifts, emit's bookkeeping in 32 bit, lean tile scan and top-down level search in deposit, lineage walks as one warp per window with a 21-instruction position step | `dec_scan_serial` 568 -> 489 us, `dec_emit` 366 -> 325 us, `dec_deposit` x9 525 -> 447 us, lineage (finds + walks) 692 -> 558 us; stage 3.50 -> 3.11 ms |
| inverse lifting: the item-start loads in one batch | `lift_inv_kernel<0>` 106 -> 96 registers, 122.8 -> 117.7 us; stage 0.214 -> 0.209 ms |

Whole job: {s0['value']:.0f} -> {d['value']:.0f} Mpixel/s device resident (+{100*(d['value']/s0['value']-1):.0f} %), single frame {s0['single_frame']['ms_per_frame']} -> {d['single_frame']['ms_per_frame']} ms.  With frames
in flight the throughput is the sum of the wide kernels' solo times (the one-warp resolver and the lineage passes hide behind other
frames), and every wide kernel is bound by the integer pipes, so the only thing that moved it was fewer instructions per token /
coefficient.

## measured and dropped (all bit-exact)

* **Forward lifting, row pairs two at a time** (`r02_lift_fwd_pairs.patch`, ncu page `r02_lift_fwd_pairs_ncu_raw.csv.gz`): the steady
  rows of interior strips run in pairs with the window roles written out (no register rotation), predicated stores, a running max / min
  instead of an abs per sample, `x - (a + b) / 2` as the added quotient of the negated sum, and every row an item starts with requested
  before the first use.  `lift_fwd_kernel<0>`: 82.8 -> 74.8 M warp instructions (18 % fewer on the ALU pipe), but 126.5 -> 131-134 us;
  chain 0.226 -> 0.232 ms at 8K, 0.088 -> 0.094 at 4K, 0.050 -> 0.052 at 1080p.  Issue slots fell from 67 % to 58 % used: the kernel
  is not bound by its instruction count.  The round-1 body stays.
* **Other attempts on `lift_fwd_kernel<0>`** (126-131 us, 55 % of the chain).  `tools/pipe_probe.cu` on the box: IMAD, LOP3, SHF, PRMT,
  VIMNMX3, IABS, LEA.HI and LEA.HI.SX32 all issue at the same rate, an IMAD + LOP3 pair at 1.7x that (two pipes), IMAD.HI at 0.4x.
  Each of these left the time within +-3 %: every item-start load in one batch (they had been four dependent DRAM latencies per item:
  29 % of the stall samples; afterwards the stall sat on a register copy of a prefetched row that the compiler places right behind
  the load); `prefetch.global.L2` of the rows 8 / 16 / 32 rows ahead (0.222 / 0.224 / 0.227 ms against 0.220); 4 instead of 5 CTAs
  per SM (128 registers, no spills: 0.220); streaming stores for the detail bands (0.234 / 0.235 on that box); row segments of 8 / 32 /
  64 rows instead of 16 (0.232 / 0.252 / 0.290).  The kernel moves 460 MB in 126 us with DRAM at 43 %, L2 at 34 %, ALU at 62-66 %,
  issue slots at 58-67 %, 9 % of the SM cycles idle in the ramp and tail, 8 % of the stall samples without instructions (the kernel is
  4 700 - 6 000 instructions long).
* `enc_vli_kernel`: no-iteration fast path for tiles whose threads map every order in 0..8 (or 0..15) to one end order: coder stage
  1.35-1.38 ms against 1.34.
* `dec_emit_kernel` with every stream window selected from four 32-bit words (no 64-bit values at all): 388 M warp instructions
  and 420 us against 314 M / 366 us -- the selects cost more than the shifts they replace; the version that keeps one 32-bit window
  for short tokens and falls back to the 64-bit pair is the one that won (284 M / 325 us).
* Lineage walk by pointer jumping (the warp builds the map offset -> offset behind the slice for order-0 tokens with runs below 7, six
  shuffle rounds, four slices at a time; exact steps from the first token that is not plain): 739 us for the five walks against 588 --
  the listed windows are not the dense ones.  What they are (`DWT_DEBUG=1`, `-DDWT_RESOLVE_PROFILE`): 1 565 / 9 / 6 / 5 / 4 windows in
  passes 1-5, ~700 tokens per window at orders 4-8, 125 000 cycles for the position walk of one window (180 cycles per token) and
  7 000 for the counting.  Skipping windows whose chains all sit at order 0: no effect (588 -> 573 us, same exact steps).  The
  21-instruction position step (slice words in registers, token length straight from the unary prefix) is what helped: 510 us.
* More or fewer lineage passes with the new walk kernel: 3 / 5 / 8 / 12 passes -> coder stage 3.46 / 3.32 / 3.49 / 3.52 ms (5 stays).
* Resolver (`-DDWT_RESOLVE_PROFILE`): 2.2 M cycles at 8K = exact steps at chunk starts 1.15 M (52 %), end searches 0.50 M, window rounds
  0.16 M, super rounds 0.11 M; 1.4 M cycles at 1080p.  Payload from the first 32-bit window in its exact walk: 1 202 -> 1 196 us.

## every kernel of one encode + decode at the final commit (`r02_kernel_times.txt`)

```
""" + KT + """```

## launch list: share of a frame round trip (cold-cache, serialised under ncu: compare shares)

""" + L + """
## per-kernel ncu page of one encode + decode (commit `58d9376`)

""" + K + """
## the same page for the kernels that changed after that capture (final commit, `r02_ncu_late_raw.csv.gz`)

""" + KL
open(os.path.join(P, "r02_summary.md"), "w").write(txt)
print("written", len(txt))
