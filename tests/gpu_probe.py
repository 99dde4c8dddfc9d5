"""Stage-by-stage GPU-vs-oracle probe (development aid; the real checks live in the pytest files).

usage: python tests/gpu_probe.py [enc|dec|all]
Prints one line per check and never stops at the first failure.
"""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import dwt_b200 as D  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
fails = 0


def report(name, ok, extra=""):
    global fails
    if not ok:
        fails += 1
    print(("PASS " if ok else "FAIL ") + name + (" " + extra if extra else ""), flush=True)


def first_diff(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return "shape %s vs %s" % (a.shape, b.shape)
    d = np.argwhere(a != b)
    if len(d) == 0:
        return ""
    i = tuple(d[0])
    return "ndiff=%d first@%s got=%s want=%s" % (len(d), i, a[i], b[i])


def images():
    rng = np.random.default_rng(7)
    yield "photo8x8", O.synth(8, 8, "photo", 1)
    yield "photo9x8", O.synth(9, 8, "photo", 2)
    yield "photo17x31", O.synth(17, 31, "photo", 3)
    yield "noise64x64", O.synth(64, 64, "noise", 4)
    yield "photo133x100", O.synth(133, 100, "photo", 5)
    yield "gray133x100", O.synth(133, 100, "photo", 5)[:, :, 1].copy()
    yield "photo320x240", O.synth(320, 240, "photo", 6)
    yield "sparse200x300", (rng.integers(0, 256, (300, 200, 3)) * (rng.random((300, 200, 3)) < 0.05)).astype(np.uint8)
    yield "photo8x500", O.synth(8, 500, "photo", 7)
    yield "photo3000x9", O.synth(3000, 9, "photo", 8)
    yield "photo1001x777", O.synth(1001, 777, "photo", 9)
    yield "noise1001x777", O.synth(1001, 777, "noise", 10)
    yield "flat64x64", np.full((64, 64, 3), 77, np.uint8)
    yield "photo1920x1080", O.synth(1920, 1080, "photo", 1)


def main():
    cod = D.Codec()
    if what in ("enc", "all"):
        rng = np.random.default_rng(1)
        for N, CH, SI in [(8, 1, 1), (9, 3, 3), (16, 3, 5), (33, 7, 7), (100, 4, 9)]:
            x = rng.integers(-1000, 1000, (N - 1) * SI + CH).astype(np.int32)
            want, win = O.cdf53(x, N, SI, SI, CH)
            got, gin = D.cdf53(x, N, SI, SI, CH)
            report("cdf53 N=%d CH=%d" % (N, CH), (want == got).all() and (win == gin).all(), first_diff(got, want))
            wi = O.icdf53(want, N, SI, SI, CH)
            gi = D.icdf53(want, N, SI, SI, CH)
            report("icdf53 N=%d CH=%d" % (N, CH), (wi == gi).all(), first_diff(gi, wi))
    for name, img in images():
        try:
            if what in ("enc", "all"):
                t = time.time()
                wpyr, wlin, wplanes = O.front_end(img)
                gpyr, glin, gplanes = cod.front_end(img)
                report("front_end.pyramid " + name, (wpyr.reshape(gpyr.shape) == gpyr).all(),
                       first_diff(gpyr, wpyr.reshape(gpyr.shape)))
                report("front_end.planes " + name, wplanes == gplanes, "%s vs %s" % (gplanes, wplanes))
                report("front_end.planar " + name, (wlin == glin).all(), first_diff(glin, wlin))
                if img.ndim == 3:
                    inv = D.inverse(wpyr)
                    src = img.astype(np.int32)
                    ycc = O.front_end  # noqa: F841
                    # inverse of the oracle pyramid must give back the YCoCg image the oracle started from
                    y = D.ycocg_from_rgb(src.reshape(-1))
                    report("inverse2d " + name, (inv.reshape(-1) == y).all(), first_diff(inv.reshape(-1), y))
                full, st = O.encode(img)
                for cap in [0, 7, 100, len(full) // 3, len(full) - 1, len(full), len(full) + 5]:
                    want = full if cap == 0 else full[:cap]
                    got = cod.encode(img, cap)
                    ok = got == want
                    extra = ""
                    if not ok:
                        n = min(len(got), len(want))
                        diff = [i for i in range(n) if got[i] != want[i]][:3]
                        extra = "len %d vs %d firstdiff %s prefixbits=%d" % (len(got), len(want), diff,
                                                                             st.meta_bits + st.root_bits)
                    report("encode cap=%d %s" % (cap, name), ok, extra)
                print("   stats: lift %.3f ms linearize %.3f ms coder %.3f ms total %.3f ms  (%.1fs wall incl oracle)" %
                      (cod.stats.ms_lift, cod.stats.ms_linearize, cod.stats.ms_coder, cod.stats.ms_total,
                       time.time() - t), flush=True)
            if what in ("dec", "all"):
                full, _ = O.encode(img)
                npx = img.shape[0] * img.shape[1]
                for cap, pm in [(0, -1), (len(full) // 2, -1), (len(full) // 5, -1), (len(full) // 50 + 7, -1), (6, -1),
                                (0, npx // 4), (0, npx // 100), (0, 0), (len(full) - 1, -1)]:
                    s = full if cap == 0 else full[:cap]
                    want = O.decode(s, pm)
                    got = cod.decode(s, pm)
                    ok = (want is None and got is None) or (want is not None and got is not None and
                                                            want.shape == got.shape and (want == got).all())
                    extra = ""
                    if not ok:
                        extra = "got None" if got is None else ("want None" if want is None else first_diff(got, want))
                    report("decode cap=%d pix=%d %s" % (cap, pm, name), ok, extra)
                print("   stats: total %.3f ms coder %.3f ms windows %d short %d" % (cod.stats.ms_total, cod.stats.ms_coder, cod.stats.parse_windows, cod.stats.parse_exact), flush=True)
        except Exception:
            fails_local = traceback.format_exc()
            report("exception " + name, False, fails_local.splitlines()[-1])
            print(fails_local, flush=True)
    print("probe done: %d failures" % fails)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
