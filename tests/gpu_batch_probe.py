"""throughput of the batch config (1920x1080 images through dwt_pool) -- development aid"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, dwt_b200 as D
from oracle import pyoracle as O
W, H, CH = 1920, 1080, 3
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
imgs = [O.synth(W, H, "photo", s) for s in range(8)]
pool = D.Pool(0, workers)
keep, enc, dec = [], (D.EncodeItem * N)(), (D.DecodeItem * N)()
for i in range(N):
    a, o1 = D.pinned_array(W * H * CH); a[:] = imgs[i % 8].reshape(-1)
    b, o2 = D.pinned_array(W * H * CH * 2 + 4096)
    c, o3 = D.pinned_array(W * H * CH)
    keep.append((a, b, c, o1, o2, o3))
    enc[i] = D.EncodeItem(a.ctypes.data, W, H, CH, 0, b.ctypes.data, b.size, 0, 0)
for rep in range(3):
    t = time.perf_counter(); bad = pool.encode_items(enc, N); te = time.perf_counter() - t
    for i in range(N):
        dec[i] = D.DecodeItem(keep[i][1].ctypes.data, enc[i].out_len, -1, keep[i][2].ctypes.data, keep[i][2].size, 0, 0, 0, 0)
    t = time.perf_counter(); bad2 = pool.decode_items(dec, N); td = time.perf_counter() - t
    ok = all(np.array_equal(keep[i][2], keep[i][0]) for i in range(0, N, 7))
    print("N=%d workers=%d encode %.1f ms (%.0f Mpx/s, %.0f img/s) decode %.1f ms (%.0f Mpx/s, %.0f img/s) %s %d %d" %
          (N, workers, te * 1e3, N * W * H / te / 1e6, N / te, td * 1e3, N * W * H / td / 1e6, N / td, "ok" if ok else "MISMATCH", bad, bad2), flush=True)
