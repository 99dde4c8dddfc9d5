"""GPU parity, second file (run on the B200 box: python -m pytest tests -m gpu): the reference's own sample image through the
programs and the C ABI, the stderr counters next to the reference program's, corrupted payloads, pinned batch seeds, 8K with
several frames in flight, the multi-device pool.  Bit-exact bar as in test_gpu_parity.py."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests.golden_util import sha

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENC, DEC = os.path.join(ROOT, "encode"), os.path.join(ROOT, "decode")


def smpte_path(oracle):
    p = os.path.join(oracle.REF_DIR, "smpte.pnm")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/smpte.pnm travels with the built reference only")
    return p


def test_smpte_through_programs_and_abi(codec, oracle, tmp_path):
    """BASELINE config 1 (README.md:5-15,23-29 of the reference) with the pins of SURVEY.md App. E.1"""
    src = smpte_path(oracle)
    data = open(src, "rb").read()
    img = np.frombuffer(data[len(data) - 320 * 240 * 3:], dtype=np.uint8).reshape(240, 320, 3)
    out, back = tmp_path / "s.dwt", tmp_path / "s.pnm"
    # the file as shipped (header with a '# CREATOR' comment line) through ./encode, then ./decode
    r = subprocess.run([ENC, src, str(out)], capture_output=True)
    assert r.returncode == 0, r.stderr
    full = out.read_bytes()
    assert len(full) == 10147 and sha(full)[:32] == "2ac1d6b75498f2982c2fbf80edc9ea74"
    assert r.stderr.decode().splitlines() == ["48 bits for meta data", "559 bits for root image", "81174 bits (10 KiB) encoded"]
    r = subprocess.run([DEC, str(out), str(back)], capture_output=True)
    assert r.returncode == 0, r.stderr
    assert len(back.read_bytes()) == 230415 and sha(back.read_bytes())[:32] == "d5aa54cc941a4f09ee8944c191cc678c"
    # the same through dwt_encode / dwt_decode
    assert codec.encode(img) == full
    assert np.array_equal(codec.decode(full), img)
    for cap, s_pin, d_pin, shape in [(100, "b6ab8b137f09ad1a7714dcef31051740", "73316eafa90bbf4fa8f432785af60100", (15, 20, 3)),
                                     (1024, "84b4f91bcc4614f12e3d4a964793666a", "1f1215ea2af3e42ee1a22a7f4df67184", (240, 320, 3)),
                                     (4096, "a1db9b49cae90c93c2a613f67126febf", "af191909a85b877ae6c1b7f037c1ccff", (240, 320, 3))]:
        r = subprocess.run([ENC, src, str(out), str(cap)], capture_output=True)
        assert r.returncode == 0
        s = out.read_bytes()
        assert s == full[:cap] and sha(s)[:32] == s_pin, cap
        assert codec.encode(img, cap) == s
        r = subprocess.run([DEC, str(out), str(back)], capture_output=True)
        assert r.returncode == 0 and sha(back.read_bytes())[:32] == d_pin, cap
        d = codec.decode(s)
        assert d.shape == shape and sha(oracle.pnm_bytes(d))[:32] == d_pin, cap   # cap 100 drops to 20x15 (App. A.6)


def test_stderr_counters_next_to_the_reference_program(codec, oracle, tmp_path):
    """encode.c:175-180,226-230 under capacities that land inside the header, the root image, the plane counts and the payload,
    and on a stream whose length is 512 mod 1024 with a partial last byte (the KiB figure rounds on the padded byte count)"""
    if not oracle.have_ref():
        pytest.skip("needs the built reference programs (oracle/_ref)")
    ref_enc = os.path.join(oracle.REF_DIR, "encode")
    src = smpte_path(oracle)
    odd = tmp_path / "odd.pnm"
    img = oracle.synth(56, 44, "photo", 433)
    odd.write_bytes(oracle.pnm_bytes(img))
    s, st = oracle.encode(img)
    assert len(s) % 1024 == 512 and st.total_bits % 8 != 0
    for path, caps in [(src, [None, 0, -5, 1, 3, 5, 6, 7, 8, 20, 50, 75, 76, 77, 100, 1024, 10146, 10147, 10148, 65536]),
                       (str(odd), [None, 3583, 3584, 3585])]:
        for cap in caps:
            a, b = tmp_path / "a.dwt", tmp_path / "b.dwt"
            extra = [] if cap is None else [str(cap)]
            ra = subprocess.run([ENC, path, str(a)] + extra, capture_output=True)
            rb = subprocess.run([ref_enc, path, str(b)] + extra, capture_output=True)
            assert ra.returncode == rb.returncode == 0, (path, cap)
            assert a.read_bytes() == b.read_bytes(), (path, cap)
            assert ra.stderr == rb.stderr, (path, cap, ra.stderr, rb.stderr)


def test_reject_codes_and_messages(codec, oracle, tmp_path):
    """decode.c:143-159,180-186: silent exit 1 on a wrong magic or a size below 8; 'reached end of file' (bytes.h:99-103)
    when the stream ends inside the header, the root image or the plane counts"""
    img = oracle.synth(64, 48, "photo", 9)
    s, _ = oracle.encode(img)
    ref_dec = os.path.join(oracle.REF_DIR, "decode") if oracle.have_ref() else None
    for name, bad, code in [("empty", b"", 1), ("w", b"W", 1), ("magic", b"X6" + s[2:], 2), ("digit", b"W7" + s[2:], 2),
                            ("short5", s[:5], 1), ("hdr", s[:6], 1), ("root", s[:9], 1), ("tiny", b"W6\x03\x00\x03\x00" + s[6:], 2)]:
        assert codec.decode(bad) is None and codec.reject_code == code, name
        f, o = tmp_path / (name + ".dwt"), tmp_path / (name + ".pnm")
        f.write_bytes(bad)
        r = subprocess.run([DEC, str(f), str(o)], capture_output=True)
        assert r.returncode == 1 and not o.exists(), name
        assert (b"reached end of file" in r.stderr) == (code == 1), (name, r.stderr)
        if ref_dec:
            rr = subprocess.run([ref_dec, str(f), str(o)], capture_output=True)
            assert rr.returncode == 1 and (b"reached end of file" in rr.stderr) == (code == 1), (name, rr.stderr)


def test_corrupted_payloads_decode_like_the_reference(codec, oracle):
    """bit flips inside the bit-plane payload and random tails (decode.c:67-100; rle.h:91-103 'ret != 1 -> -1'): the pixels must
    be the reference's, with both scan kernels, and nothing may fault.  tests/test_oracle.py checks the same cases between
    the oracle and the reference program on the CPU."""
    import dwt_b200 as D
    from tests import fuzz_util as F
    use_ref = oracle.have_ref()
    other = D.Codec(0)
    other.set_scan("parallel")
    codec.set_scan("serial")
    compared = 0
    try:
        for i in range(F.N_CASES):
            desc, s = F.fuzz_case(i)
            want = oracle.decode(s)
            if use_ref and i % 3 == 0:   # a third of the cases also straight against the program (the CPU suite does all)
                kind, ref = F.ref_decode_guarded(s)
                if kind == "undefined":
                    continue
                assert (ref is None) == (want is None) and (ref is None or np.array_equal(ref, want)), (i, desc)
            for cd in (codec, other):
                got = cd.decode(s)
                assert (got is None) == (want is None), (i, desc)
                if want is not None:
                    assert got.shape == want.shape and np.array_equal(got, want), (i, desc)
            compared += 1
    finally:
        codec.set_scan("auto")
        other.close()
    assert compared >= 200


def test_batch_seeds_match_reference_pins(codec, oracle):
    """BASELINE config 4: per-image seeds of the 1080p batch; 64 of them pinned from the reference programs
    (tests/golden/pins_batch.json, made by tests/golden/make_golden.py --batch)"""
    import dwt_b200 as D
    with open(os.path.join(ROOT, "tests", "golden", "pins_batch.json")) as f:
        pins = json.load(f)
    assert len(pins) >= 64
    imgs = [oracle.synth(1920, 1080, "photo", p["seed"]) for p in pins]
    pool = D.Pool(0, 8)
    try:
        streams = pool.encode_batch(imgs)
        for p, s in zip(pins, streams):
            assert len(s) == p["len"] and sha(s) == p["sha"], p["seed"]
        dec = pool.decode_batch(streams, [im.shape for im in imgs])
        for im, d in zip(imgs, dec):
            assert np.array_equal(im, d)
        cut = pool.encode_batch(imgs[:8], 200000)
        for p, s in zip(pins, cut):
            assert sha(s) == p["sha_cap200000"], p["seed"]
    finally:
        pool.close()


def test_8k_with_frames_in_flight(oracle, pins_big):
    """the bench's configuration: several contexts announced as busy (dwt_ctx_set_in_flight(8)) at 8K, lossless and at the
    1 MiB budget, against the reference pins"""
    import dwt_b200 as D
    rec = [r for r in pins_big if r["spec"] == dict(kind="photo", w=7680, h=4320, seed=1)][0]
    img = oracle.synth(7680, 4320, "photo", 1)
    cods = [D.Codec(0) for _ in range(2)]
    try:
        for cd in cods:
            cd.set_in_flight(8)
        for cd in cods:
            full = cd.encode(img)
            assert len(full) == rec["full_len"] and sha(full) == rec["full_sha"]
            assert np.array_equal(cd.decode(full), img)
            case = [c for c in rec["cases"] if c["cap"] == 1048576][0]
            cut = cd.encode(img, 1048576)
            assert sha(cut) == case["stream_sha"]
            dec = cd.decode(cut)
            assert list(dec.shape) == case["decoded"]["shape"] and sha(np.ascontiguousarray(dec).tobytes()) == case["decoded"]["sha"]
    finally:
        for cd in cods:
            cd.close()


def test_multi_device_pool_and_batch_cli(oracle, tmp_path):
    """dwt_pool_create_multi / dwtbatch -g: item i is coded on devices[i mod G] (no exchange between GPUs).  On a
    one-GPU box the device list names GPU 0 twice, which runs the same sharding code; with more GPUs "all" spans them."""
    import dwt_b200 as D
    imgs = [oracle.synth(160 + 31 * i, 96 + 17 * i, "photo" if i % 2 else "noise", 80 + i) for i in range(11)]
    want = [oracle.encode(im)[0] for im in imgs]
    for dev in ([0, 0], "all"):
        pool = D.Pool(dev, 2)
        try:
            assert len(pool.devices()) >= 1
            assert pool.encode_batch(imgs) == want
            assert pool.encode_batch(imgs, 999) == [w[:999] for w in want]
            dec = pool.decode_batch(want, [im.shape for im in imgs])
            for a, b in zip(dec, imgs):
                assert np.array_equal(a, b)
        finally:
            pool.close()
    tool = os.path.join(ROOT, "dwtbatch")
    src, enc_dir, dec_dir = tmp_path / "src", tmp_path / "enc", tmp_path / "dec"
    for d in (src, enc_dir, dec_dir):
        d.mkdir()
    for i, im in enumerate(imgs):
        (src / ("im%02d.pnm" % i)).write_bytes(oracle.pnm_bytes(im))
    names = sorted(str(p) for p in src.iterdir())
    r = subprocess.run([tool, "encode", "-j", "2", "-g", "all", str(enc_dir)] + names, capture_output=True)
    assert r.returncode == 0, r.stderr
    for i, w in enumerate(want):
        assert (enc_dir / ("im%02d.dwt" % i)).read_bytes() == w
    r = subprocess.run([tool, "decode", "-g", "0,0", str(dec_dir)] + sorted(str(p) for p in enc_dir.iterdir()), capture_output=True)
    assert r.returncode == 0, r.stderr
    for i, im in enumerate(imgs):
        assert (dec_dir / ("im%02d.pnm" % i)).read_bytes() == oracle.pnm_bytes(im)


def test_hilbert_ballot_and_tma_paths_agree(codec, oracle):
    """full 32x32 cells go through the TMA-staged kernels when rows are 16-byte aligned (width % 4 == 0) and through the ballot
    kernels otherwise (or with DWT_HILBERT=ballot): both against the oracle, on aligned and unaligned widths, colour and gray"""
    cases = [oracle.synth(640, 360, "photo", 3), oracle.synth(642, 361, "photo", 4), oracle.synth(1024, 1024, "noise", 5),
             oracle.synth(516, 300, "photo", 6)[:, :, 1].copy(), oracle.synth(1001, 777, "photo", 13)]
    want = [oracle.encode(im)[0] for im in cases]
    old = os.environ.get("DWT_HILBERT")
    try:
        for mode in ("tma", "ballot"):
            os.environ["DWT_HILBERT"] = mode
            for im, w in zip(cases, want):
                assert codec.encode(im) == w, (mode, im.shape)
                assert np.array_equal(codec.decode(w), im), (mode, im.shape)
                cut = w[: len(w) // 2]
                a, b = codec.decode(cut), oracle.decode(cut)
                assert a.shape == b.shape and np.array_equal(a, b), (mode, im.shape)
                pyr, lin, planes = codec.front_end(im)
                wpyr, wlin, wplanes = oracle.front_end(im)
                assert np.array_equal(lin, wlin) and planes == wplanes, (mode, im.shape)
    finally:
        if old is None:
            os.environ.pop("DWT_HILBERT", None)
        else:
            os.environ["DWT_HILBERT"] = old


def test_lineage_passes_on_small_streams(oracle):
    """the decoder's lineage passes (dec_extend_find / dec_extend_walk: one warp per window) only run on images of 20 Mpixel
    and more by default; DWT_LINEAGE forces them (read once per process, hence the child process), so that small streams --
    sparse images with long runs, noise, truncated and corrupted streams -- exercise them against the oracle as well"""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import dwt_b200 as D
from oracle import pyoracle as O
cod = D.Codec()
rng = np.random.default_rng(7)
bad = 0
for (w, h, kind, seed) in [(1920, 1080, "photo", 1), (1024, 1024, "noise", 2), (1500, 900, "sparse", 3), (640, 360, "smooth", 4)]:
    if kind == "sparse":    # long zero runs, high Rice orders: where seeded chains synchronise slowly
        img = (rng.integers(0, 256, (h, w, 3)) * (rng.random((h, w, 3)) < 0.03)).astype(np.uint8)
    elif kind == "smooth":  # streams that are mostly refinement bits
        img = np.clip(np.cumsum(rng.integers(-2, 3, (h, w, 3)), axis=1) + 128, 0, 255).astype(np.uint8)
    else:
        img = O.synth(w, h, kind, seed)
    s, _ = O.encode(img)
    for cut in (len(s), len(s) // 2, len(s) // 7, 4096):
        t = s[:cut]
        a, b = cod.decode(t), O.decode(t)
        ok = (a is None and b is None) or (a is not None and b is not None and a.shape == b.shape and np.array_equal(a, b))
        bad += not ok
    t = bytearray(s)
    for _ in range(20):
        t[int(rng.integers(64, len(t)))] ^= 1 << int(rng.integers(0, 8))
    a, b = cod.decode(bytes(t)), O.decode(bytes(t))
    ok = (a is None and b is None) or (a is not None and b is not None and a.shape == b.shape and np.array_equal(a, b))
    bad += not ok
print("BAD", bad)
''' % ROOT
    for passes in ("3", "6"):
        env = dict(os.environ, DWT_LINEAGE=passes)
        r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "BAD 0" in r.stdout, (passes, r.stdout[-500:], r.stderr[-500:])
