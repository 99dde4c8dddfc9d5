"""Corrupted-payload cases for the decoder (decode.c:67-100, rle.h:91-103): seeded, so that the CPU test (oracle
against the reference program) and the GPU test (CUDA decoder against the reference program) see the same streams."""
import os
import subprocess
import tempfile

import numpy as np

from oracle import pyoracle as O

N_CASES = 240


def base_image(rng):
    w, h = int(rng.integers(8, 401)), int(rng.integers(8, 301))
    kind = int(rng.integers(0, 4))
    if kind == 0:
        img = O.synth(w, h, "photo", int(rng.integers(0, 1000)))
    elif kind == 1:
        img = O.synth(w, h, "noise", int(rng.integers(0, 1000)))
    elif kind == 2:   # sparse: long zero runs, high Rice orders
        img = (rng.integers(0, 256, (h, w, 3)) * (rng.random((h, w, 3)) < 0.03)).astype(np.uint8)
    else:             # smooth: streams that are mostly refinement bits
        img = np.clip(np.cumsum(rng.integers(-2, 3, (h, w, 3)), axis=1) + 128, 0, 255).astype(np.uint8)
    if rng.random() < 0.15:
        img = np.ascontiguousarray(img[:, :, 1])
    return img


def fuzz_case(i):
    """(description, corrupted stream): bit flips strictly inside the bit-plane payload and / or a random tail"""
    rng = np.random.default_rng(77000 + i)
    img = base_image(rng)
    stream, st = O.encode(img)
    s = bytearray(stream)
    # header + root image + plane counts end a few bits after meta + root (three short VLIs): stay clear of them
    first = (st.meta_bits + st.root_bits + 64 + 7) // 8
    mode = i % 4
    what = []
    if mode in (0, 1, 3) and first < len(s):
        if mode == 1:   # cut first, then corrupt what is left
            s = s[: max(first + 1, int(len(s) * rng.uniform(0.1, 0.9)))]
        nflip = int(rng.integers(1, 51))
        for _ in range(nflip):
            pos = int(rng.integers(first, len(s)))
            s[pos] ^= 1 << int(rng.integers(0, 8))
        what.append("%d flips" % nflip)
    if mode in (2, 3):
        tail = rng.integers(0, 256, int(rng.integers(1, 400))).astype(np.uint8).tobytes()
        s += tail
        what.append("tail %d" % len(tail))
    shape = "%dx%d%s" % (img.shape[1], img.shape[0], "g" if img.ndim == 2 else "")
    return "%s %s" % (shape, " + ".join(what)), bytes(s)


def ref_decode_guarded(stream, timeout=20):
    """the unmodified reference decoder on `stream`: ('ok', image) / ('exit1', None) / ('undefined', None) when it crashes
    or does not finish (corrupted streams can drive it into C undefined behaviour, e.g. shifts by >= 32 in vli.h:89-90)"""
    with tempfile.TemporaryDirectory() as d:
        pin, pout = os.path.join(d, "i.dwt"), os.path.join(d, "o.pnm")
        with open(pin, "wb") as f:
            f.write(stream)
        try:
            r = subprocess.run([os.path.join(O.REF_DIR, "decode"), pin, pout], capture_output=True, timeout=timeout)
        except subprocess.TimeoutExpired:
            return "undefined", None
        if r.returncode < 0:
            return "undefined", None
        if r.returncode != 0 or not os.path.exists(pout):
            return "exit1", None
        with open(pout, "rb") as f:
            return "ok", O.parse_pnm(f.read())
