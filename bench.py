#!/usr/bin/env python
"""bench.py -- encode/decode Mpixel/s on 8K RGB (BASELINE.json's metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over a batch of --frames synthetic 7680x4320 RGB frames: lossless encode to a
.dwt stream and decode of that stream back to pixels, the frames of a step in flight together (one codec context =
one CUDA stream + one host thread per frame; images are independent).  Mpixel/s = W*H*frames*steps*ranks / seconds.
  value : device-resident round trips (images and streams already in HBM), one region of CUDA events per step
  e2e   : the same through the public C ABI with page-locked HOST buffers (dwt_encode_into/dwt_decode_into),
          host->device and device->host copies inside the timed region
  single_frame / stages / roofline : a pass with ONE frame at a time, so that every kernel runs alone on the GPU
Images are independent, so ranks shard frames with no collective on the data path ("weak" scaling: one frame
stream per GPU); torch.distributed is only the barrier and the max-over-ranks of the timing.
  configs : every BASELINE.json config behind its reference pin -- 4K lossless, 8K at the 64 KiB / 1 MiB / 8 MiB budgets
          (encode and decode), 16384 x 16384, and the batch of 1920x1080 images through dwt_pool from host buffers
          (per-rank seeds; the one config that is also reported at N > 1)
--impl reference times the UNMODIFIED reference programs (oracle/_ref, built by oracle/Makefile) on the host
cores, on bands of the same frame (one process per band), and once on the whole 8K frame (one process).
"""
import argparse
import json
import os

# 8 + 16 codec contexts share the device: give every CUDA stream its own hardware queue (the default is 8 connections,
# more streams than that get false dependencies).  Has to be set before CUDA is initialised.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, CH = 7680, 4320, 3
WORKLOAD = "synthetic 7680x4320 RGB 'photo' frame (SURVEY App. E.2, seed = rank+1), lossless encode + decode round trip"
LIFT_BYTES_PER_PIXEL = 23.0   # u8 in (colour fused), int32 between levels, RGB
BATCH_W, BATCH_H = 1920, 1080  # BASELINE config 4: images of the batch
# reference pins of the single-image configs (tests/golden/pins_big.json): stream length, sha256[:32] of
# the stream, shape and sha256[:32] of the decoded pixels
PIN_4K = (12216090, "63f44307a0de6ba2a26ae9074d724298")
PIN_8K_CAPS = {65536: ("66fb128c746d9933afd5c104a7198fb1", (2160, 3840, 3)),
               1048576: ("eee08d5ce1e3e02bf111483c9c6cbd03", (4320, 7680, 3)),
               8388608: ("26c40605427826496473915b202408d7", (4320, 7680, 3))}
PIN_16K = (395131715, "c77d5aff9242788e9bb23170a7ee730d")
L2_NOTE = ("inputs larger than L2: a step touches ~1.5 GB per frame (126 MB L2); L2 flushed (256 MB write) before the timed "
           "region and between the frames of the single-frame pass")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s in sm if s >= 0.5 * max(sm)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------- reference arm

TILE_W, TILE_H = 1920, 1080   # 16 tiles of the 8K frame; 16:9 like the frame, so the reference's Hilbert walk over the
                              # enclosing power-of-two square (encode.c:46-49) has the frame's own overhead ratio (2.02x)


def make_tile_files(tmp, seed):
    from oracle import pyoracle as O
    img = O.synth(W, H, "photo", seed)
    paths = []
    for ty in range(H // TILE_H):
        for tx in range(W // TILE_W):
            tile = np.ascontiguousarray(img[ty * TILE_H:(ty + 1) * TILE_H, tx * TILE_W:(tx + 1) * TILE_W])
            p = os.path.join(tmp, "tile%d_%d.pnm" % (ty, tx))
            with open(p, "wb") as f:
                f.write(O.pnm_bytes(tile))
            paths.append(p)
    return paths


def reference_step(paths, workers):
    """every tile through the reference encode then decode program, `workers` processes at a time; wall seconds"""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle as O
    enc, dec = os.path.join(O.REF_DIR, "encode"), os.path.join(O.REF_DIR, "decode")

    def one(p):
        return subprocess.call("%s %s %s.dwt && %s %s.dwt %s.out" % (enc, p, p, dec, p, p), shell=True,
                               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=workers) as ex:
        rc = list(ex.map(one, paths))
    dt = time.perf_counter() - t0
    if any(rc):
        raise RuntimeError("reference program failed: %s" % rc)
    return dt


def reference_single_process_8k():
    """the stock reference programs on the whole 8K frame, one process: seconds for encode / decode"""
    from oracle import pyoracle as O
    tmp = tempfile.mkdtemp(prefix="dwtref8k", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        src, out, back = os.path.join(tmp, "f.pnm"), os.path.join(tmp, "f.dwt"), os.path.join(tmp, "b.pnm")
        with open(src, "wb") as f:
            f.write(O.pnm_bytes(O.synth(W, H, "photo", 1)))
        t0 = time.perf_counter()
        rc1 = subprocess.call([os.path.join(O.REF_DIR, "encode"), src, out], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t1 = time.perf_counter()
        rc2 = subprocess.call([os.path.join(O.REF_DIR, "decode"), out, back], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t2 = time.perf_counter()
        if rc1 or rc2 or os.path.getsize(out) != 48863617:
            raise RuntimeError("reference 8K run failed")
        npx = W * H
        return dict(encode_s=round(t1 - t0, 2), decode_s=round(t2 - t1, 2), encode_mpx_s=round(npx / (t1 - t0) / 1e6, 3),
                    decode_mpx_s=round(npx / (t2 - t1) / 1e6, 3), round_trip_mpx_s=round(npx / (t2 - t0) / 1e6, 3), cores=1,
                    sample="one reference encode + one decode process on the whole 7680x4320 frame, files in /dev/shm, CLI wall time")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def reference_bench(steps, warmup, quick=False):
    """Mpixel/s of the unmodified reference programs on the host cores: the 8K frame cut into 16 tiles of 1920x1080,
    one reference process per tile, as many at a time as there are cores (the reference is single-threaded)."""
    from oracle import pyoracle as O
    if not O.have_ref():
        raise RuntimeError("oracle/_ref is not built (oracle/Makefile needs /root/reference at build time)")
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="dwtref", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        paths = make_tile_files(tmp, 1)
        workers = max(1, min(cores, len(paths)))
        if quick:                         # bounded sample for the cpu_baseline leg of the default run: one wave of tiles
            paths = paths[:workers]
        pixels = TILE_W * TILE_H * len(paths)
        for _ in range(warmup):
            reference_step(paths, workers)
        times = [reference_step(paths, workers) for _ in range(steps)]
        # parity of the sample: the reference round trip is lossless
        assert open(paths[0], "rb").read() == open(paths[0] + ".out", "rb").read(), "reference round trip is not lossless?"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    sec = sum(times) / len(times)
    return dict(value=pixels / sec / 1e6, unit="Mpixel/s", cores=workers, kind="reference", ms_per_step=sec * 1e3,
                sample="%d tiles of %dx%d cut from the 8K frame (%.1f Mpixel), reference encode + decode program per tile, %d "
                       "processes at a time, files in /dev/shm, CLI wall time incl. PNM I/O" %
                       (len(paths), TILE_W, TILE_H, pixels / 1e6, workers))


# ----------------------------------------------------------------------------------------------- our arm

def frame_seed(rank):
    """images are independent: rank r codes its own frame stream (seed r+1), no exchange step"""
    return rank + 1


def job_mpixels_per_s(pixels_per_step, steps, world, max_ms_over_ranks):
    """whole-job throughput: every rank processed `steps` frames in the slowest rank's time"""
    return pixels_per_step * steps * world / (max_ms_over_ranks / 1e3) / 1e6


def reduce_over_ranks(total_ms, e2e_ms, launches, device):
    """max over ranks of the two timings, sum of the launch counts (the only collectives in the benchmark)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    l = torch.tensor([launches], dtype=torch.int64, device=device)
    dist.all_reduce(l, op=dist.ReduceOp.SUM)
    return float(t[0]), float(t[1]), int(l[0])


def reduce_max(values, device):
    """max over ranks of a list of floats (timings of the sharded batch config)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def staging_ceiling(device, barrier, seconds=0.4):
    """host <-> device staging rate of this rank with every rank copying at the same time: pure cudaMemcpyAsync from / to
    page-locked memory, both directions at once (what the end-to-end pass does), GB/s per direction"""
    import torch
    n = 1 << 28
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(n, dtype=torch.uint8, device=device)
    d2 = torch.empty(n, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)

    def run(reps):
        torch.cuda.synchronize(device)
        barrier()
        t = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize(device)
        return reps * n / (time.perf_counter() - t) / 1e9
    run(1)
    return run(max(2, int(seconds * 25e9 / n)))


def batch_seeds(rank, distinct):
    """BASELINE config 4: image i of the batch has seed i (0..4095); rank r owns the seeds r*512 .. r*512+511"""
    return [rank * 512 + i for i in range(distinct)]


def single_image_config(D, cod, img, caps, reps, check):
    """one image, device resident: per capacity the encode time (CUDA events around all kernels of the call), the decode
    time of that stream, and the same through the host-buffer calls (wall time incl. the copies).  check(cap, stream,
    decoded) raises when the result is not the reference's."""
    import hashlib
    h, w = img.shape[:2]
    npx = w * h
    src, o1 = D.pinned_array(img.size)
    src[:] = img.reshape(-1)
    src = src.reshape(img.shape)
    out, o2 = D.pinned_array(img.size + 4096)
    dec, o3 = D.pinned_array(img.size)
    res = {}
    for cap in caps:
        n = cod.encode_into(src, out, cap)           # also the warm-up: buffers of this geometry get allocated here
        stream = out[:n].copy()
        shp = cod.decode_into(out, n, dec)
        check(cap, stream, dec[:shp[0] * shp[1] * shp[2]].reshape(shp))
        cod.upload_image(src)
        enc_ms, coder_ms = [], []
        for _ in range(reps):
            cod.flush_l2()
            st = cod.encode_resident(cap)
            enc_ms.append(st.ms_total)
            coder_ms.append(st.ms_coder)
        cod.upload_stream(stream)
        dec_ms, dcoder_ms = [], []
        for _ in range(reps):
            cod.flush_l2()
            cod.decode_resident(-1)
            dec_ms.append(cod.stats.ms_total)
            dcoder_ms.append(cod.stats.ms_coder)
        t0 = time.perf_counter()
        cod.encode_into(src, out, cap)
        t1 = time.perf_counter()
        cod.decode_into(out, n, dec)
        t2 = time.perf_counter()
        e, d = statistics.median(enc_ms), statistics.median(dec_ms)
        res[cap] = dict(stream_bytes=int(n), decoded_shape=list(shp), parity="reference pin ok",
                        encode_ms=round(e, 4), decode_ms=round(d, 4), encode_coder_ms=round(statistics.median(coder_ms), 4),
                        decode_coder_ms=round(statistics.median(dcoder_ms), 4),
                        encode_mpx_s=round(npx / e / 1e3, 1), decode_mpx_s=round(npx / d / 1e3, 1),
                        e2e_encode_ms=round((t1 - t0) * 1e3, 3), e2e_decode_ms=round((t2 - t1) * 1e3, 3),
                        e2e_encode_mpx_s=round(npx / (t1 - t0) / 1e6, 1), e2e_decode_mpx_s=round(npx / (t2 - t1) / 1e6, 1))
    del o1, o2, o3
    return res


def single_image_suite(D, O, cod, img8k, reps=5):
    """4K lossless, 8K with the three byte budgets, 16384 x 16384 -- each behind its reference pin"""
    import hashlib

    def sha(b):
        return hashlib.sha256(bytes(b)).hexdigest()[:32]
    out = {}

    img4k = O.synth(3840, 2160, "photo", 1)

    def check4k(cap, stream, dec):
        assert (len(stream), sha(stream)) == PIN_4K, "4K stream differs from the reference pin"
        assert np.array_equal(dec, img4k), "4K round trip is not lossless"
    out["4k_lossless"] = dict(workload="synthetic 3840x2160 RGB photo (seed 1), lossless",
                              **single_image_config(D, cod, img4k, [0], reps, check4k)[0])

    with open(os.path.join(ROOT, "tests", "golden", "pins_big.json")) as f:
        big = [r for r in json.load(f) if r["spec"] == dict(kind="photo", w=W, h=H, seed=1)][0]
    dec_pin = {c["cap"]: c["decoded"] for c in big["cases"]}

    def check8k(cap, stream, dec):
        pin, shape = PIN_8K_CAPS[cap]
        assert len(stream) == cap and sha(stream) == pin, "8K stream at %d bytes differs from the reference pin" % cap
        assert tuple(dec.shape) == shape == tuple(dec_pin[cap]["shape"])
        assert hashlib.sha256(np.ascontiguousarray(dec).tobytes()).hexdigest() == dec_pin[cap]["sha"], \
            "8K decode at %d bytes differs from the reference" % cap
    r8 = single_image_config(D, cod, img8k, sorted(PIN_8K_CAPS), reps, check8k)
    for cap, name in [(65536, "8k_cap_64KiB"), (1048576, "8k_cap_1MiB"), (8388608, "8k_cap_8MiB")]:
        out[name] = dict(workload="synthetic 7680x4320 RGB photo (seed 1), capacity %d bytes" % cap, **r8[cap])

    img16 = O.synth(16384, 16384, "photo", 1)

    def check16k(cap, stream, dec):
        assert (len(stream), sha(stream)) == PIN_16K, "16384^2 stream differs from the reference pin"
        assert np.array_equal(dec, img16), "16384^2 round trip is not lossless"
    out["16384sq_lossless"] = dict(workload="synthetic 16384x16384 RGB photo (seed 1), lossless, 12 levels",
                                   **single_image_config(D, cod, img16, [0], max(2, reps // 2), check16k)[0])
    return out


def batch_config(D, O, device, rank, n_images, workers, keep):
    """BASELINE config 4 on this rank's GPU: n_images 1920x1080 images through dwt_pool from page-locked host buffers.
    Returns (encode seconds, decode seconds, info).  Gate: the streams of the pinned seeds equal the reference's
    (tests/golden/pins_batch.json) and every decoded image equals its source."""
    import hashlib
    distinct = min(n_images, 128)
    seeds = batch_seeds(rank, distinct)
    imgs = [O.synth(BATCH_W, BATCH_H, "photo", sd) for sd in seeds]
    raw = BATCH_W * BATCH_H * CH
    out_room = raw + 4096
    pool = D.Pool(device, workers)
    src, outs, decs = [], [], []
    for i in range(distinct):
        a, o = D.pinned_array(raw)
        a[:] = imgs[i].reshape(-1)
        keep.append(o)
        src.append(a)
    enc = (D.EncodeItem * n_images)()
    dec = (D.DecodeItem * n_images)()
    for i in range(n_images):
        b, o2 = D.pinned_array(out_room)
        c, o3 = D.pinned_array(raw)
        keep.extend([o2, o3])
        outs.append(b)
        decs.append(c)
        enc[i] = D.EncodeItem(src[i % distinct].ctypes.data, BATCH_W, BATCH_H, CH, 0, b.ctypes.data, b.size, 0, 0)
    warm = min(n_images, 2 * workers)
    assert pool.encode_items(enc, warm) == 0
    for i in range(warm):
        dec[i] = D.DecodeItem(outs[i].ctypes.data, enc[i].out_len, -1, decs[i].ctypes.data, raw, 0, 0, 0, 0)
    assert pool.decode_items(dec, warm) == 0
    t0 = time.perf_counter()
    bad = pool.encode_items(enc, n_images)
    t_enc = time.perf_counter() - t0
    assert bad == 0, "batch encode failed"
    for i in range(n_images):
        dec[i] = D.DecodeItem(outs[i].ctypes.data, enc[i].out_len, -1, decs[i].ctypes.data, raw, 0, 0, 0, 0)
    t0 = time.perf_counter()
    bad = pool.decode_items(dec, n_images)
    t_dec = time.perf_counter() - t0
    assert bad == 0, "batch decode failed"
    pool.close()
    with open(os.path.join(ROOT, "tests", "golden", "pins_batch.json")) as f:
        pins = {p["seed"]: p for p in json.load(f)}
    checked = 0
    for i in range(distinct):
        p = pins.get(seeds[i])
        if p:
            n = enc[i].out_len
            assert (n, hashlib.sha256(outs[i][:n].tobytes()).hexdigest()) == (p["len"], p["sha"]), "batch seed %d differs from the reference" % seeds[i]
            checked += 1
    for i in range(n_images):
        assert np.array_equal(decs[i], src[i % distinct]), "batch image %d does not round-trip" % i
    stream_bytes = sum(int(enc[i].out_len) for i in range(n_images))
    return t_enc, t_dec, dict(images_per_gpu=n_images, distinct_seeds_per_gpu=distinct, pool_workers=workers, pinned_seeds_checked=checked,
                              stream_bytes_per_gpu=stream_bytes, raw_bytes_per_gpu=n_images * raw)


def our_bench(args, rank, world, local):
    import hashlib
    import torch
    from concurrent.futures import ThreadPoolExecutor
    import dwt_b200 as D
    from oracle import pyoracle as O  # synthetic generator only (inputs); the codec never touches it

    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    F = max(1, args.frames)              # frames per step, coded concurrently on F contexts (one CUDA stream each)
    cods = [D.Codec(local) for _ in range(F)]
    cod = cods[0]
    img = O.synth(W, H, "photo", frame_seed(rank))
    npx = W * H

    # ---- page-locked host buffers for the end-to-end path (one set per context)
    keep = []
    pin_img, pin_out, pin_dec = [], [], []
    for _ in range(F):
        a, o1 = D.pinned_array(img.size)
        a[:] = img.reshape(-1)
        b, o2 = D.pinned_array(img.size * 2 + 4096)
        c, o3 = D.pinned_array(img.size)
        keep += [o1, o2, o3]
        pin_img.append(a.reshape(H, W, CH))
        pin_out.append(b)
        pin_dec.append(c)

    # ---- parity gate: no number without bit-exactness (lossless round trip + pin for rank 0's frame), every context
    for i, cd in enumerate(cods):
        n = cd.encode_into(pin_img[i], pin_out[i])
        if rank == 0:
            digest = hashlib.sha256(pin_out[i][:n].tobytes()).hexdigest()[:32]
            assert (n, digest) == (48863617, "f14ef79d0680a4ace7daa2b9e8063651"), "8K stream differs from the reference pin"
        shp = cd.decode_into(pin_out[i], n, pin_dec[i])
        assert shp == (H, W, CH) and np.array_equal(pin_dec[i], pin_img[i].reshape(-1)), "round trip is not lossless"
    stream_bytes = n

    def sync_all():
        for cd in cods:
            D.lib().dwt_ctx_sync(cd._h)

    def barrier():
        sync_all()
        if use_dist:
            dist.barrier()

    for i, cd in enumerate(cods):
        cd.upload_image(pin_img[i])
        cd.upload_stream(pin_out[i][:n])
    pool = ThreadPoolExecutor(max_workers=F)

    # ---- pass 1, one frame at a time (latency; every kernel alone on the GPU: the stage timings and the roofline)
    stage = dict(lift_fwd=[], linearize=[], enc_coder=[], dec_coder=[], reconstruct=[], lift_inv=[], enc=[], dec=[])

    def serial_step(timed):
        cod.flush_l2()
        cod.event_record(0)
        se = cod.encode_resident(0)
        e = (se.ms_lift, se.ms_linearize, se.ms_coder, se.ms_total)
        cod.decode_resident(-1)
        sd = cod.stats
        cod.event_record(1)
        ms = cod.event_elapsed_ms(0, 1)
        if timed:
            stage["lift_fwd"].append(e[0]); stage["linearize"].append(e[1]); stage["enc_coder"].append(e[2]); stage["enc"].append(e[3])
            stage["dec_coder"].append(sd.ms_coder); stage["reconstruct"].append(sd.ms_linearize)
            stage["lift_inv"].append(sd.ms_lift); stage["dec"].append(sd.ms_total)
        return ms

    for _ in range(args.warmup):
        serial_step(False)
    barrier()
    serial_ms = [serial_step(True) for _ in range(args.steps)]
    barrier()

    # ---- pass 2, the measured throughput: K steps of F frames, device resident, ONE region of CUDA events around the K
    # steps (barrier + synchronize on both sides).  Context i codes its K frames back to back; the contexts run freely
    # next to each other, so there is no idle tail between steps.  No L2 flush inside the region: a step touches
    # F x ~1.5 GB, far more than the 126 MB L2.
    def resident_frames(i, k):
        for _ in range(k):
            cods[i].encode_resident(0)
            cods[i].decode_resident(-1)

    def resident_region(k):
        cod.flush_l2()
        sync_all()
        cod.event_record(0)
        list(pool.map(lambda i: resident_frames(i, k), range(F)))
        for cd in cods[1:]:
            cod.wait_for(cd)
        cod.event_record(1)
        return cod.event_elapsed_ms(0, 1)

    for cd in cods:
        cd.set_in_flight(F)   # F contexts are busy from here on (pass 1 ran with the default: one frame at a time)
    resident_region(args.warmup)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = sum(cd.launch_count() for cd in cods)
    t_wall0 = time.perf_counter()
    total_ms = resident_region(args.steps)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sum(cd.launch_count() for cd in cods) - l0

    # ---- pass 3, end to end through the library's batch call (dwt_pool_run): host buffers in, host buffers out, copies
    # inside the timed region.  The K steps are one call: K x F encode items (frames) + K x F decode items (the streams of
    # the parity gate), interleaved on the pool's contexts so that pixel uploads overlap pixel downloads.  Every job has
    # its own output buffer.
    dpool = D.Pool(local, args.pool_workers if args.pool_workers > 0 else 2 * F)
    out_room = stream_bytes + stream_bytes // 4 + 4096

    def make_jobs(n_jobs):
        enc_items = (D.EncodeItem * n_jobs)()
        dec_items = (D.DecodeItem * n_jobs)()
        outs, decs = [], []
        for jb in range(n_jobs):
            o, own_o = D.pinned_array(out_room)
            dd, own_d = D.pinned_array(img.size)
            keep.extend([own_o, own_d])
            outs.append(o)
            decs.append(dd)
            i = jb % F
            enc_items[jb] = D.EncodeItem(pin_img[i].ctypes.data, W, H, CH, 0, o.ctypes.data, o.size, 0, 0)
            dec_items[jb] = D.DecodeItem(pin_out[i].ctypes.data, stream_bytes, -1, dd.ctypes.data, dd.size, 0, 0, 0, 0)
        return n_jobs, enc_items, dec_items, outs, decs

    def e2e_region(jobs, n_jobs=None):
        n_all, enc_items, dec_items, outs, decs = jobs
        n_jobs = n_all if n_jobs is None else n_jobs
        cod.flush_l2()
        sync_all()
        t0 = time.perf_counter()
        if dpool.run_items(enc_items, n_jobs, dec_items, n_jobs):
            raise RuntimeError("dwt_pool_run failed")
        dt = (time.perf_counter() - t0) * 1e3
        assert all(enc_items[jb].out_len == stream_bytes for jb in range(n_jobs)), "stream length changed"
        return dt

    # warm-up: every context of the pool must have coded in both directions (buffers are allocated on first use):
    # one call per direction with exactly one job per context, then mixed calls
    nwk = args.pool_workers if args.pool_workers > 0 else 2 * F
    warm_jobs = make_jobs((nwk + F - 1) // F * F)
    assert dpool.encode_items(warm_jobs[1], nwk) == 0 and dpool.decode_items(warm_jobs[2], nwk) == 0
    for _ in range(max(1, args.warmup // 2)):
        e2e_region(warm_jobs)
    # one call for the K steps; a long run (large --steps) is cut into calls of <= ~8 GB of page-locked job buffers, which are
    # reused from call to call (every call still carries all of its copies; only the drain between calls is extra)
    total_jobs = args.steps * F
    jobs = make_jobs(min(total_jobs, max(2 * nwk, int(8e9) // (out_room + img.size))))
    barrier()
    e2e_total, done = 0.0, 0
    while done < total_jobs:
        n = min(jobs[0], total_jobs - done)
        e2e_total += e2e_region(jobs, n)
        done += n
    barrier()
    clocks = sampler.stop() if sampler else None
    assert np.array_equal(jobs[4][-1], pin_img[(jobs[0] - 1) % F].reshape(-1)), "end-to-end round trip is not lossless"
    assert bytes(jobs[3][0][:stream_bytes]) == bytes(pin_out[0][:stream_bytes]), "end-to-end stream differs from the gate's"
    dpool.close()
    # what the box's host <-> device path can carry with every rank copying at once (the end-to-end number's ceiling)
    ceiling = staging_ceiling(torch.device("cuda", local), barrier)
    # the page-locked job buffers of the passes above (~8 GB) go back before the batch config takes its own
    del jobs, warm_jobs, pin_img, pin_out, pin_dec
    keep.clear()
    import gc
    gc.collect()

    # ---- per-config numbers (BASELINE.json configs), each behind its reference pin
    mode = args.configs if args.configs != "auto" else ("all" if world == 1 else "batch")
    configs = {}
    batch = None
    if mode in ("all", "batch"):
        nb = max(1, args.batch_images)
        t_enc, t_dec, binfo = batch_config(D, O, local, rank, nb, 16, keep)
        barrier()
        batch = (t_enc, t_dec, binfo)
    if mode == "all" and rank == 0:
        for cd in cods[1:]:
            cd.close()
        cod.set_in_flight(1)
        configs.update(single_image_suite(D, O, cod, img))

    # ---- max over ranks
    if use_dist:
        total_ms, e2e_total, launches = reduce_over_ranks(total_ms, e2e_total, launches, torch.device("cuda", local))
        if batch:
            t_enc, t_dec = reduce_max(batch[:2], torch.device("cuda", local))
            batch = (t_enc, t_dec, batch[2])
        ceiling = -reduce_max([-ceiling], torch.device("cuda", local))[0] * world   # slowest rank x ranks
    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return None

    if batch:
        t_enc, t_dec, binfo = batch
        n_all = binfo["images_per_gpu"] * world
        bpx = BATCH_W * BATCH_H
        configs["batch_1080p"] = dict(
            workload="batch of 1920x1080 RGB photo images (seeds rank*512+i), lossless, through dwt_pool_encode / dwt_pool_decode "
                     "from page-locked host buffers (copies inside the timed calls); one call per direction per GPU, max over ranks",
            images=n_all, n_gpus=world, parity="pinned seeds equal the reference, every image round-trips", **binfo,
            encode_s=round(t_enc, 4), decode_s=round(t_dec, 4),
            encode_images_s=round(n_all / t_enc, 1), decode_images_s=round(n_all / t_dec, 1),
            encode_mpx_s=round(n_all * bpx / t_enc / 1e6, 1), decode_mpx_s=round(n_all * bpx / t_dec / 1e6, 1),
            h2d_bytes=int(world * (binfo["raw_bytes_per_gpu"] + binfo["stream_bytes_per_gpu"])),
            d2h_bytes=int(world * (binfo["raw_bytes_per_gpu"] + binfo["stream_bytes_per_gpu"])))
    value = job_mpixels_per_s(npx * F, args.steps, world, total_ms)
    e2e_value = job_mpixels_per_s(npx * F, args.steps, world, e2e_total)
    peak, peak_kind = peaks()
    med = {k: statistics.median(v) for k, v in stage.items()}

    def roof(ms, nbytes):
        gbs = nbytes / (ms / 1e3) / 1e9
        return dict(bound="hbm", achieved=round(gbs, 1), peak=peak, unit="GB/s", frac=round(gbs / peak, 4),
                    frac_of_datasheet_8000=round(gbs / 8000.0, 4),  # also against the 8 TB/s datasheet figure
                    ms=round(ms, 4), bytes=int(nbytes))

    nsamples = npx * CH
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        traffic = {}
    how = "single-frame pass: the stage's kernels alone on the GPU, CUDA events on their stream"
    stages = dict(
        lift_fwd=dict(roof(med["lift_fwd"], LIFT_BYTES_PER_PIXEL * npx),
                      kernel="lift_fwd_kernel x levels + lift_tail_fwd_kernel (colour fused into the first level)"),
        lift_inv=dict(roof(med["lift_inv"], LIFT_BYTES_PER_PIXEL * npx),
                      kernel="lift_tail_inv_kernel + lift_inv_kernel x levels (colour + clamp fused into the last)"),
        linearize=dict(roof(med["linearize"], 4 * nsamples + nsamples * 10 / 8),
                       kernel="linearize_tma_kernel (full cells, TMA boxes) + linearize_kernel (cut cells): Hilbert gather + bit-slicing; "
                              "the stage also holds the plane-count read-back and the zeroing of the bit-sliced store"),
        enc_coder=dict(roof(med["enc_coder"], 4 * nsamples + stream_bytes), kernel="enc_* (count, scan, emit, VLI orders, bit scan, scatter)"),
        dec_coder=dict(roof(med["dec_coder"], 4 * nsamples + stream_bytes),
                       kernel="dec_* (scan, lineage passes, link, resolve, emit, prep / tilescan / deposit per plane depth)"),
        reconstruct=dict(roof(med["reconstruct"], 4 * nsamples + nsamples * 10 / 8),
                         kernel="reconstruct_tma_kernel (full cells, TMA stores) + reconstruct_kernel (cut cells): Hilbert scatter + bias"),
    )
    for k, v in stages.items():
        v.update(peak_source=peak_kind, traffic=traffic.get(k), measured_in=how)
    stage_bytes = {k: v["bytes"] for k, v in stages.items()}
    frame_bytes = sum(stage_bytes.values())          # SURVEY 8(d) algorithmic bytes of all six stages of a round trip
    single_ms = statistics.median(serial_ms)
    dominant = max(stages, key=lambda k: stages[k]["ms"])
    # the headline roofline is the stage a frame spends most of its time in (not the lifting, which is 6 % of a
    # frame); BASELINE.json's own target -- the lifting stages against the HBM roofline -- is reported next to it
    roofline = dict(stages[dominant], stage=dominant, share_of_single_frame=round(stages[dominant]["ms"] / single_ms, 3))
    lifting = dict(forward=stages["lift_fwd"]["frac"], inverse=stages["lift_inv"]["frac"], target=0.5,
                   note="BASELINE.json north_star: >= 50 % of the B200 HBM roofline for the lifting stages (23 B/pixel each way)")
    whole_frame = dict(bound="hbm", unit="GB/s", peak=peak, bytes=int(frame_bytes),
                       in_flight=dict(ms_per_frame=round(total_ms / args.steps / F, 4),
                                      achieved=round(frame_bytes / (total_ms / args.steps / F / 1e3) / 1e9, 1),
                                      frac=round(frame_bytes / (total_ms / args.steps / F / 1e3) / 1e9 / peak, 4)),
                       single_frame=dict(ms_per_frame=round(single_ms, 4), achieved=round(frame_bytes / (single_ms / 1e3) / 1e9, 1),
                                         frac=round(frame_bytes / (single_ms / 1e3) / 1e9 / peak, 4)),
                       time_dominant_stage=dominant, time_dominant_share=round(stages[dominant]["ms"] / single_ms, 3),
                       note="sum of the six stages' algorithmic bytes (lifting 23 B/pixel each way, Hilbert 4 B/sample + planes/8, "
                            "coder 4 B/sample + stream) over the time of a whole encode + decode")
    out = dict(metric="encode/decode Mpixel/s, 8K RGB", value=round(value, 2), unit="Mpixel/s", n_gpus=world, steps=args.steps,
               warmup=args.warmup, ms_per_step=round(total_ms / args.steps, 3), higher_is_better=True, scaling="weak",
               vs_baseline=None, dtype="int32", data="synthetic",
               config=dict(workload=WORKLOAD, frames_per_step_per_gpu=F, width=W, height=H, channels=CH, l2=L2_NOTE,
                           stream_bytes=stream_bytes,
                           parallelism="one frame stream per GPU, no collective; the %d frames of a step are coded concurrently "
                                       "on %d contexts (one CUDA stream and one host thread each)" % (F, F)),
               e2e=dict(value=round(e2e_value, 2), unit="Mpixel/s", ms_per_step=round(e2e_total / args.steps, 3),
                        staging_gbs_per_direction=round(world * F * (img.size + stream_bytes) / (e2e_total / args.steps / 1e3) / 1e9, 1),
                        staging_ceiling_gbs_per_direction=round(ceiling, 1),
                        fraction_of_staging_ceiling=round(world * F * (img.size + stream_bytes) / (e2e_total / args.steps / 1e3) / 1e9 / ceiling, 3),
                        staging_note="ceiling = pure cudaMemcpyAsync from / to page-locked memory, both directions at once on every rank "
                                     "at the same time (slowest rank x ranks), measured in this run",
                        h2d_bytes_per_step=int(world * F * (img.size + stream_bytes)),
                        d2h_bytes_per_step=int(world * F * (stream_bytes + img.size)),
                        api="dwt_pool_run: dwt_encode_into of %d frames + dwt_decode_into of %d streams per step, interleaved on %d contexts, "
                            "page-locked host buffers, the K steps in one call (runs of more than ~50 jobs: several calls on the same buffers)" % (F, F, 2 * F)),
               gpu_launches=int(launches), clocks=clocks, roofline=roofline, lifting_roofline=lifting, stages=stages,
               whole_frame=whole_frame, configs=configs,
               single_frame=dict(ms_per_frame=round(statistics.median(serial_ms), 3),
                                 mpixel_s=round(npx / (statistics.median(serial_ms) / 1e3) / 1e6, 1),
                                 encode_mpx_s=round(npx / (med["enc"] / 1e3) / 1e6, 1),
                                 decode_mpx_s=round(npx / (med["dec"] / 1e3) / 1e6, 1)),
               wall_ms_per_step=round(t_wall * 1e3 / args.steps, 3))
    if world == 1 and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = reference_bench(1, 0, quick=True)
        except Exception as ex:  # the oracle always exists; _ref may be missing on a box that never built it
            out["cpu_baseline"] = dict(value=None, unit="Mpixel/s", cores=0, kind="reference", sample="unavailable: %s" % ex)
    if use_dist:
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pool-workers", type=int, default=0, help="contexts of the end-to-end pool (default 2 x frames)")
    ap.add_argument("--frames", type=int, default=8, help="frames per step, coded concurrently (one context each)")
    ap.add_argument("--configs", default="auto", choices=["auto", "all", "batch", "none"],
                    help="per-config numbers: auto = all at N=1, only the sharded batch config at N>1")
    ap.add_argument("--batch-images", type=int, default=512, help="1920x1080 images per GPU in the batch config")
    ap.add_argument("--no-single-process-8k", action="store_true", help="reference arm: skip the one-process run on the whole frame")
    args = ap.parse_args()
    rank, world, local = dist_env()
    # the one JSON line goes to the real stdout; anything a library prints to fd 1 meanwhile (the NCCL version banner) is
    # sent to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = reference_bench(max(1, args.steps), max(0, args.warmup))
        single = None
        if not args.no_single_process_8k:
            try:
                single = reference_single_process_8k()
            except Exception as ex:
                single = dict(unavailable=str(ex))
        line = dict(impl="reference", metric="encode/decode Mpixel/s, 8K RGB", value=round(r["value"], 3), unit="Mpixel/s",
                    n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=round(r["ms_per_step"], 1),
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32", data="synthetic",
                    config=dict(workload=WORKLOAD, sample=r["sample"]),
                    cpu_baseline=dict(value=round(r["value"], 3), unit="Mpixel/s", cores=r["cores"], kind="reference", sample=r["sample"],
                                      single_process_8k=single),
                    e2e=dict(value=round(r["value"], 3), unit="Mpixel/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
        return 0
    if args.warmup < 3:
        args.warmup = 3
    out = our_bench(args, rank, world, local)
    if out is not None:
        real_stdout.write(json.dumps(out) + "\n")
        real_stdout.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
