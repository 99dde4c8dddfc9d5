"""Generate tests/golden/pins.json from the UNMODIFIED reference programs (oracle/_ref/encode, decode).

Run in the build container (where /root/reference exists and oracle/Makefile has built oracle/_ref):
    python tests/golden/make_golden.py [--big | --batch]

Inputs are the integer-only synthetic images of SURVEY.md App. E.2 (oracle.pyoracle.synth), so they can be
regenerated anywhere; only hashes, sizes and a few short streams are stored.  Every record is produced by
running the reference binaries on files, never by our own code.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as O  # noqa: E402


def sha(b):
    return hashlib.sha256(b).hexdigest()


def make_image(spec):
    kind = spec["kind"]
    w, h = spec["w"], spec["h"]
    if kind in ("photo", "noise"):
        img = O.synth(w, h, kind, spec["seed"])
    elif kind == "flat":
        img = np.full((h, w, 3), spec["seed"] & 255, np.uint8)
    elif kind == "sparse":
        rng = np.random.default_rng(spec["seed"])
        img = (rng.integers(0, 256, (h, w, 3)) * (rng.random((h, w, 3)) < 0.05)).astype(np.uint8)
    else:
        raise ValueError(kind)
    if spec.get("gray"):
        img = np.ascontiguousarray(img[:, :, 1])
    return img


def record(spec, caps, pixel_args, keep_stream=False):
    img = make_image(spec)
    full = O.ref_encode(img)
    rec = dict(spec=spec, pixels_sha=sha(img.tobytes()), full_len=len(full), full_sha=sha(full), cases=[])
    for cap in caps:
        if isinstance(cap, float):
            cap = max(1, int(len(full) * cap))
        stream = full if cap is None else O.ref_encode(img, cap)
        if cap is not None:
            assert stream == full[:cap], "reference capped stream is not a prefix?"
        for pm in pixel_args:
            dec = O.ref_decode(stream, pm)
            case = dict(cap=cap, pixels_max=pm, stream_len=len(stream), stream_sha=sha(stream))
            if dec is None:
                case.update(decoded=None)
            else:
                case.update(decoded=dict(shape=list(dec.shape), sha=sha(np.ascontiguousarray(dec).tobytes())))
            rec["cases"].append(case)
    if keep_stream:
        rec["full_hex"] = full.hex()
    return rec


def batch_pins():
    """BASELINE config 4: 64 of the 4096 per-image seeds of the 1080p batch through the reference encoder (8 at a time)"""
    from concurrent.futures import ThreadPoolExecutor

    def one(seed):
        img = O.synth(1920, 1080, "photo", seed)
        full = O.ref_encode(img)
        rec = dict(seed=seed, len=len(full), sha=sha(full))
        if seed < 8:
            cut = O.ref_encode(img, 200000)
            assert cut == full[:200000]
            rec["sha_cap200000"] = sha(cut)
        return rec
    seeds = list(range(8)) + [8 + 73 * k for k in range(56)]   # spread over 0..4095
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        out = list(ex.map(one, seeds))
    with open(os.path.join(HERE, "pins_batch.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote pins_batch.json", len(out), "records")


def main():
    if "--batch" in sys.argv:
        if not O.have_ref():
            raise SystemExit("oracle/_ref is not built: run `make -C oracle` where /root/reference exists")
        return batch_pins()
    big = "--big" in sys.argv
    if not O.have_ref():
        raise SystemExit("oracle/_ref is not built: run `make -C oracle` where /root/reference exists")
    out = []
    small = [
        dict(kind="photo", w=8, h=8, seed=1), dict(kind="photo", w=9, h=8, seed=2), dict(kind="photo", w=15, h=15, seed=3),
        dict(kind="photo", w=16, h=16, seed=4), dict(kind="photo", w=17, h=31, seed=5), dict(kind="photo", w=8, h=500, seed=6),
        dict(kind="photo", w=3000, h=9, seed=7), dict(kind="photo", w=133, h=100, seed=8),
        dict(kind="photo", w=133, h=100, seed=8, gray=True), dict(kind="noise", w=64, h=64, seed=9),
        dict(kind="sparse", w=200, h=300, seed=10), dict(kind="flat", w=64, h=64, seed=77),
        dict(kind="photo", w=320, h=240, seed=11), dict(kind="noise", w=320, h=240, seed=12),
    ]
    for i, spec in enumerate(small):
        caps = [None, 1, 5, 6, 7, 8, 20, 0.1, 0.33, 0.5, 0.9]
        npx = spec["w"] * spec["h"]
        out.append(record(spec, caps, [None, 0, 64, npx // 16, npx // 4, npx - 1, npx], keep_stream=i in (1, 4, 7)))
        print("golden", spec, flush=True)
    medium = [dict(kind="photo", w=1001, h=777, seed=13), dict(kind="noise", w=1001, h=777, seed=14),
              dict(kind="photo", w=1920, h=1080, seed=1), dict(kind="noise", w=1920, h=1080, seed=1)]
    for spec in medium:
        npx = spec["w"] * spec["h"]
        out.append(record(spec, [None, 100, 4096, 0.02, 0.25, 0.75], [None, npx // 4]))
        print("golden", spec, flush=True)
    if big:
        out.append(record(dict(kind="photo", w=3840, h=2160, seed=1), [None], [None]))
        print("golden 4K", flush=True)
        out.append(record(dict(kind="photo", w=7680, h=4320, seed=1), [None, 65536, 1048576, 8388608], [None]))
        print("golden 8K photo", flush=True)
        out.append(record(dict(kind="noise", w=7680, h=4320, seed=1), [None], [None]))
        print("golden 8K noise", flush=True)
    name = "pins_big.json" if big else "pins.json"
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", name, len(out), "records")


if __name__ == "__main__":
    main()
