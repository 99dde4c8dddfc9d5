"""CPU: the oracle restatement against (a) golden pins made by the unmodified reference binaries,
(b) the reference binaries themselves when oracle/_ref is present, (c) the pins of SURVEY.md App. E.1."""
import os

import numpy as np
import pytest

from tests.golden_util import make_image, sha, spec_id


def test_geometry_matches_survey(oracle):
    g = oracle.geometry(320, 240)
    assert g["levels"] == 6 and g["widths"][0] == 5 and g["heights"][0] == 4 and g["lengths"][-1] == 512
    g = oracle.geometry(7680, 4320)
    assert g["levels"] == 10 and (g["widths"][0], g["heights"][0]) == (8, 5) and g["lengths"][-1] == 8192
    assert oracle.geometry(16384, 16384)["levels"] == 12


def test_synth_pins(oracle):
    # SURVEY.md App. E.1 "pixels" hashes: a mismatch is a generator bug, not a codec bug
    assert sha(oracle.synth(1920, 1080, "photo", 1).tobytes())[:32] == "8f437f32b68133621ae2e3011df78442"
    assert sha(oracle.synth(1920, 1080, "noise", 1).tobytes())[:32] == "d7112f34d7891c42ccd6908637a36ff5"


def _small(pins):
    return [r for r in pins if r["spec"]["w"] * r["spec"]["h"] <= 320 * 240]


def test_oracle_encode_matches_golden(oracle, pins):
    for rec in _small(pins):
        img = make_image(rec["spec"])
        assert sha(img.tobytes()) == rec["pixels_sha"]
        full, st = oracle.encode(img)
        assert len(full) == rec["full_len"] and sha(full) == rec["full_sha"], spec_id(rec["spec"])
        if "full_hex" in rec:
            assert full.hex() == rec["full_hex"]
        for case in rec["cases"]:
            if case["pixels_max"] is not None:
                continue
            cap = case["cap"]
            got = full if cap is None else oracle.encode(img, cap)[0]
            assert sha(got) == case["stream_sha"], (spec_id(rec["spec"]), cap)
            assert got == full[: len(got)]  # a capped stream is a byte prefix (bytes.h:75-85)


def test_oracle_decode_matches_golden(oracle, pins):
    for rec in _small(pins):
        img = make_image(rec["spec"])
        full, _ = oracle.encode(img)
        for case in rec["cases"]:
            stream = full if case["cap"] is None else full[: case["cap"]]
            pm = -1 if case["pixels_max"] is None else case["pixels_max"]
            dec = oracle.decode(stream, pm)
            if case["decoded"] is None:
                assert dec is None, (spec_id(rec["spec"]), case)
            else:
                assert dec is not None and list(dec.shape) == case["decoded"]["shape"], (spec_id(rec["spec"]), case)
                assert sha(np.ascontiguousarray(dec).tobytes()) == case["decoded"]["sha"], (spec_id(rec["spec"]), case)


def test_oracle_medium_pin(oracle, pins):
    rec = [r for r in pins if r["spec"] == dict(kind="photo", w=1001, h=777, seed=13)][0]
    img = make_image(rec["spec"])
    full, st = oracle.encode(img)
    assert sha(full) == rec["full_sha"]
    dec = oracle.decode(full)
    assert np.array_equal(dec, img)


def test_oracle_against_reference_binaries(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    rng = np.random.default_rng(3)
    for (w, h) in [(8, 8), (23, 9), (40, 41), (100, 64)]:
        img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
        full = oracle.ref_encode(img)
        assert oracle.encode(img)[0] == full
        for cap in (9, len(full) // 2, len(full) - 1):
            s = oracle.ref_encode(img, cap)
            assert oracle.encode(img, cap)[0] == s
            a, b = oracle.decode(s), oracle.ref_decode(s)
            assert (a is None) == (b is None)
            if a is not None:
                assert a.shape == b.shape and np.array_equal(a, b)


def test_oracle_sweep_against_reference_binaries(oracle):
    """a wider sweep than the pins: odd geometries (incl. one dimension at the minimum of 8), gray and colour, photo / noise /
    sparse / smooth content, every capacity class and PIXELS arguments -- restatement against the unmodified programs"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    rng = np.random.default_rng(20261018)
    shapes = [(8, 8), (9, 8), (8, 31), (17, 16), (33, 65), (64, 64), (129, 47), (200, 8), (96, 135), (255, 33)]
    n = 0
    for i, (w, h) in enumerate(shapes):
        kind = i % 4
        if kind == 0:
            img = oracle.synth(w, h, "photo", 100 + i)
        elif kind == 1:
            img = oracle.synth(w, h, "noise", 100 + i)
        elif kind == 2:   # sparse: long zero runs, high Rice orders
            img = (rng.integers(0, 256, (h, w, 3)) * (rng.random((h, w, 3)) < 0.04)).astype(np.uint8)
        else:             # smooth: mostly refinement bits
            img = np.clip(np.cumsum(rng.integers(-2, 3, (h, w, 3)), axis=1) + 128, 0, 255).astype(np.uint8)
        if i % 3 == 2:
            img = np.ascontiguousarray(img[:, :, 1])  # 'W5' gray
        full = oracle.ref_encode(img)
        assert oracle.encode(img)[0] == full, (w, h, kind)
        caps = sorted({1, 5, 6, 7, 8, 12, len(full) // 3, len(full) - 1, len(full), len(full) + 7})
        for cap in caps:
            if cap <= 0:
                continue
            s = oracle.ref_encode(img, cap)
            assert oracle.encode(img, cap)[0] == s, (w, h, kind, cap)
            for px in (None, 16, 64, w * h // 4, w * h):
                if cap < 6 and px is not None:
                    continue
                a = oracle.decode(s) if px is None else oracle.decode(s, px)
                b = oracle.ref_decode(s) if px is None else oracle.ref_decode(s, px)
                assert (a is None) == (b is None), (w, h, kind, cap, px)
                if a is not None:
                    assert a.shape == b.shape and np.array_equal(a, b), (w, h, kind, cap, px)
                n += 1
    assert n > 300


def test_smpte_pins(oracle):
    path = os.path.join(os.path.dirname(oracle.LIB_PATH), "_ref", "smpte.pnm")
    if not os.path.exists(path):
        pytest.skip("smpte.pnm travels with oracle/_ref only")
    data = open(path, "rb").read()
    # header with a comment line: 'P6\n# ...\n320 240\n255\n'
    body = data[len(data) - 320 * 240 * 3:]
    img = np.frombuffer(body, dtype=np.uint8).reshape(240, 320, 3)
    full, st = oracle.encode(img)
    assert len(full) == 10147 and sha(full)[:32] == "2ac1d6b75498f2982c2fbf80edc9ea74"
    assert (st.meta_bits, st.root_bits, st.total_bits) == (48, 559, 81174)
    for cap, s_pin, d_pin in [(100, "b6ab8b137f09ad1a7714dcef31051740", "73316eafa90bbf4fa8f432785af60100"),
                              (1024, "84b4f91bcc4614f12e3d4a964793666a", "1f1215ea2af3e42ee1a22a7f4df67184"),
                              (4096, "a1db9b49cae90c93c2a613f67126febf", "af191909a85b877ae6c1b7f037c1ccff")]:
        s = oracle.encode(img, cap)[0]
        assert sha(s)[:32] == s_pin
        d = oracle.decode(s)
        assert sha(oracle.pnm_bytes(d))[:32] == d_pin


def test_oracle_corrupted_payloads_match_reference_program(oracle):
    """bit flips inside the payload and random tails (tests/fuzz_util.py): the restatement must decode what the unmodified
    reference program decodes -- this is what lets the GPU fuzz test use the oracle as its checker"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    from tests import fuzz_util as F
    compared = 0
    for i in range(0, F.N_CASES, 2):
        desc, s = F.fuzz_case(i)
        kind, ref = F.ref_decode_guarded(s)
        if kind == "undefined":
            continue
        mine = oracle.decode(s)
        assert (mine is None) == (ref is None), (i, desc)
        if ref is not None:
            assert mine.shape == ref.shape and np.array_equal(mine, ref), (i, desc)
        compared += 1
    assert compared >= 100


def test_pnm_reader_accepts_what_the_reference_accepts(oracle, tmp_path):
    """dwt_b200/host/pnm.c (a tokenizer) against read_pnm of pnm.h:14-80 through the reference encoder's exit status and
    output: comment lines, odd separators, 15-digit fields, missing fields, wrong maxval, truncated pixel data"""
    import ctypes as C
    import subprocess
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = tmp_path / "pnm.so"
    subprocess.run(["gcc", "-std=c99", "-O1", "-shared", "-fPIC", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "dwt_b200", "host", "pnm.c"), "-o", str(so)], check=True)
    L = C.CDLL(str(so))
    L.dwt_read_pnm.restype = C.POINTER(C.c_uint8)
    L.dwt_read_pnm.argtypes = [C.c_char_p] + [C.POINTER(C.c_int)] * 3
    px = bytes(range(192))
    gray = bytes(range(64))
    cases = [b"P6\n8 8\n255\n" + px, b"P6 8 8 255\n" + px, b"P6\n# one\n# two\n8 8\n255\n" + px, b"P6\n8\n# c\n8\n# d\n255\n" + px,
             b"P6\t8\r\n8  255 " + px, b"P6\n 8 8 255\n" + px, b"P6x8y8z255q" + px, b"P5\n8 8\n255\n" + gray, b"P6\n08 008\n0255\n" + px,
             b"P6\n8 8\n65535\n" + px, b"P6\n8 8\n254\n" + px, b"P6\n0 8\n255\n" + px, b"P6\n8 8\n255\n" + px[:100], b"P6\n8 8\n255",
             b"P6\n8 8\n", b"P6", b"P", b"", b"P7\n8 8\n255\n" + px, b"P6\n8 8\n255\n\n" + px, b"P6\n  # not a comment 9 9\n8 8\n255\n" + px + px,
             b"P6\n000000000000008 8 255\n" + px, b"P6\n8 8 255#" + px, b"P6#\n8 8 255\n" + px]
    for i, data in enumerate(cases):
        f, o = tmp_path / ("c%d.pnm" % i), tmp_path / ("c%d.dwt" % i)
        f.write_bytes(data)
        r = subprocess.run([os.path.join(oracle.REF_DIR, "encode"), str(f), str(o)], capture_output=True)
        w, h, ch = C.c_int(), C.c_int(), C.c_int()
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2)
        os.dup2(devnull, 2)
        try:
            p = L.dwt_read_pnm(str(f).encode(), C.byref(w), C.byref(h), C.byref(ch))
        finally:
            os.dup2(saved, 2)
            os.close(saved)
            os.close(devnull)
        ref_ok = r.returncode == 0
        ours_ok = bool(p) and w.value >= 8 and h.value >= 8   # the 8x8 minimum is main()'s test (encode.c:144-146)
        assert ours_ok == ref_ok, (i, data[:24])
        if ref_ok:
            img = np.ctypeslib.as_array(p, shape=(w.value * h.value * ch.value,)).copy()
            shape = (h.value, w.value, 3) if ch.value == 3 else (h.value, w.value)
            assert o.read_bytes() == oracle.encode(img.reshape(shape))[0], i
