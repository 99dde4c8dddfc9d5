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
