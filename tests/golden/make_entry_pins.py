"""Generate tests/golden/entry_pins.json from the reference's OWN functions (oracle/_ref/libref_shim.so).

Run in the build container, where oracle/Makefile could compile the reference headers:
    python tests/golden/make_entry_pins.py
The call sequences are the seeded ones of tests/test_ref_entry_points.py; only digests are stored.  Nothing here
calls libdwt_b200.so except to produce the streams the reader cases are cut from (checked against the shim's).
"""
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import test_ref_entry_points as T  # noqa: E402


def main():
    R = T.shim()
    if R is None:
        raise SystemExit("oracle/_ref/libref_shim.so is not built: run `make -C oracle` where /root/reference exists")
    tmp = tempfile.mkdtemp()
    out = dict(writers={}, readers={}, readers_noise={}, lifting={})
    for seed in T.SEEDS:
        for cap in T.CAPS:
            out["writers"]["%d/%d" % (seed, cap)] = T.digest(T.writer_case(R, tmp, seed, cap))
        stream = T.run_writer(R, os.path.join(tmp, "s%d.bin" % seed), 0, T.write_script(seed))[0]
        for cut in T.reader_cuts(stream):
            out["readers"]["%d/%d" % (seed, cut)] = T.digest(T.reader_case(R, tmp, seed, stream, cut))
        noise = np.random.default_rng(seed).integers(0, 256, 3000).astype(np.uint8).tobytes()
        out["readers_noise"][str(seed)] = T.digest(T.reader_case(R, tmp, 50 + seed, noise, len(noise)))
    xs = list(range(-3, 70)) + [255, 256, 257, 65535, 65536, (1 << 29) - 1, 1 << 29, 0x7fffffff]
    out["ilog2"] = [R.ilog2(x) for x in xs]
    out["geometry"] = T.digest([T.geometry_of(R, w, h) for (w, h) in T.geometry_cases()])
    ip = C.POINTER(C.c_int)
    for i, (N, CH, SI, SO) in enumerate(T.lifting_cases()):
        x = T.lifting_input(i, N, CH, SI)
        xin, res = x.copy(), np.full((N - 1) * SO + CH, 77, np.int32)
        R.cdf53(res.ctypes.data_as(ip), xin.ctypes.data_as(ip), N, SO, SI, CH)
        back = np.full(x.size, 55, np.int32)
        R.icdf53(back.ctypes.data_as(ip), res.ctypes.data_as(ip), N, SI, SO, CH)
        out["lifting"][str(i)] = T.digest([res.tobytes(), xin.tobytes(), back.tobytes()])
    rng = np.random.default_rng(12)
    n = 4096
    a = rng.integers(0, 256, 3 * n).astype(np.int32)
    b = rng.integers(-700, 700, 3 * n).astype(np.int32)
    for k in range(n):
        R.rgb2ycocg(C.cast(a.ctypes.data + 12 * k, ip))
        R.ycocg2rgb(C.cast(b.ctypes.data + 12 * k, ip))
    out["colour"] = T.digest([a.tobytes(), b.tobytes()])
    with open(os.path.join(HERE, "entry_pins.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote entry_pins.json")


if __name__ == "__main__":
    main()
