"""GPU parity tests (run on the B200 box: python -m pytest tests -m gpu).

Everything goes through the C ABI (ctypes -> libdwt_b200.so); the oracle is only the checker.
Bar: bit-exact -- identical .dwt bytes (lossless and every capacity) and identical decoded pixels.
"""
import os
import subprocess

import numpy as np
import pytest

from tests.golden_util import make_image, sha, spec_id

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ transform entry points

def test_cdf53_entry_points(codec, oracle):
    import dwt_b200 as D
    rng = np.random.default_rng(1)
    for N, CH, S in [(8, 1, 1), (9, 3, 3), (16, 3, 5), (33, 7, 7), (100, 4, 9), (2, 1, 1), (3, 2, 2), (1000, 3, 3)]:
        x = rng.integers(-2000, 2000, (N - 1) * S + CH).astype(np.int32)
        want, win = oracle.cdf53(x, N, S, S, CH)
        got, gin = D.cdf53(x, N, S, S, CH)
        assert np.array_equal(want, got) and np.array_equal(win, gin), (N, CH, S)
        assert np.array_equal(oracle.icdf53(want, N, S, S, CH), D.icdf53(want, N, S, S, CH)), (N, CH, S)


def test_colour_entry_points(codec, oracle):
    import dwt_b200 as D
    rng = np.random.default_rng(2)
    rgb = rng.integers(0, 256, 3 * 5000).astype(np.int32)
    ycc = D.ycocg_from_rgb(rgb)
    ref = rgb.copy()
    oracle.lib().orc_rgb_to_ycocg(ref.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_int)), 5000)
    assert np.array_equal(ycc, ref)
    wild = rng.integers(-600, 600, 3 * 5000).astype(np.int32)  # out-of-range values exercise the clamps
    back = D.rgb_from_ycocg(wild)
    ref = wild.copy()
    oracle.lib().orc_ycocg_to_rgb(ref.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_int)), 5000)
    assert np.array_equal(back, ref)


@pytest.mark.parametrize("shape", [(8, 8), (9, 8), (15, 15), (17, 31), (8, 500), (3000, 9), (133, 100), (320, 240), (1001, 777)])
def test_front_end_stages(codec, oracle, shape):
    import dwt_b200 as D
    w, h = shape
    img = oracle.synth(w, h, "photo", 21)
    wpyr, wlin, wplanes = oracle.front_end(img)
    gpyr, glin, gplanes = codec.front_end(img)
    assert np.array_equal(gpyr, wpyr), "Mallat pyramid"
    assert gplanes == wplanes, "plane counts"
    assert np.array_equal(glin, wlin), "linearised coefficients"
    # dwt_forward / dwt_inverse: the two `transformation` drivers on host int buffers
    ycc = D.ycocg_from_rgb(img.astype(np.int32).reshape(-1)).reshape(h, w, 3)
    assert np.array_equal(D.forward(ycc), wpyr)
    assert np.array_equal(D.inverse(wpyr), ycc)


def test_front_end_gray(codec, oracle):
    img = oracle.synth(200, 120, "photo", 5)[:, :, 0].copy()
    wpyr, wlin, wplanes = oracle.front_end(img)
    gpyr, glin, gplanes = codec.front_end(img)
    assert np.array_equal(gpyr, wpyr) and np.array_equal(glin, wlin) and gplanes == wplanes


# ------------------------------------------------------------------ golden vectors (made by the reference binaries)

def test_encode_matches_golden(codec, pins):
    for rec in pins:
        img = make_image(rec["spec"])
        full = codec.encode(img)
        assert len(full) == rec["full_len"] and sha(full) == rec["full_sha"], spec_id(rec["spec"])
        for case in rec["cases"]:
            if case["pixels_max"] is not None or case["cap"] is None:
                continue
            got = codec.encode(img, case["cap"])
            assert len(got) == case["stream_len"] and sha(got) == case["stream_sha"], (spec_id(rec["spec"]), case["cap"])


def test_decode_matches_golden(codec, pins):
    for rec in pins:
        img = make_image(rec["spec"])
        full = codec.encode(img)
        assert sha(full) == rec["full_sha"]
        for case in rec["cases"]:
            stream = full if case["cap"] is None else full[: case["cap"]]
            pm = -1 if case["pixels_max"] is None else case["pixels_max"]
            dec = codec.decode(stream, pm)
            if case["decoded"] is None:
                assert dec is None, (spec_id(rec["spec"]), case)
            else:
                assert dec is not None and list(dec.shape) == case["decoded"]["shape"], (spec_id(rec["spec"]), case)
                assert sha(np.ascontiguousarray(dec).tobytes()) == case["decoded"]["sha"], (spec_id(rec["spec"]), case)


def test_decode_of_committed_reference_streams(codec, pins):
    """streams stored verbatim from the reference encoder: the decoder is checked without our encoder"""
    n = 0
    for rec in pins:
        if "full_hex" not in rec:
            continue
        stream = bytes.fromhex(rec["full_hex"])
        img = make_image(rec["spec"])
        assert np.array_equal(codec.decode(stream), img)
        n += 1
    assert n >= 3


# ------------------------------------------------------------------ oracle on seeded inputs

@pytest.mark.parametrize("seed", range(6))
def test_random_images_against_oracle(codec, oracle, seed):
    rng = np.random.default_rng(100 + seed)
    w, h = int(rng.integers(8, 400)), int(rng.integers(8, 300))
    kind = seed % 3
    if kind == 0:
        img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    elif kind == 1:
        img = (rng.integers(0, 256, (h, w, 3)) * (rng.random((h, w, 3)) < 0.02)).astype(np.uint8)  # long zero runs
    else:
        img = np.clip(np.cumsum(rng.integers(-3, 4, (h, w, 3)), axis=1) + 128, 0, 255).astype(np.uint8)
    want, st = oracle.encode(img)
    got = codec.encode(img)
    assert got == want
    assert (codec.stats.meta_bits, codec.stats.root_bits, codec.stats.total_bits) == (st.meta_bits, st.root_bits, st.total_bits)
    for cap in [int(c) for c in rng.integers(1, len(want) + 10, 6)]:
        assert codec.encode(img, cap) == want[:cap]
        a, b = codec.decode(want[:cap]), oracle.decode(want[:cap])
        assert (a is None) == (b is None)
        if a is not None:
            assert a.shape == b.shape and np.array_equal(a, b), (w, h, cap)
    for pm in [0, 50, w * h // 7, w * h]:
        a, b = codec.decode(want, pm), oracle.decode(want, pm)
        assert a.shape == b.shape and np.array_equal(a, b), (w, h, pm)


def test_degenerate_streams(codec, oracle):
    img = oracle.synth(64, 48, "photo", 9)
    s, _ = oracle.encode(img)
    for bad in [b"", b"W", b"X6" + s[2:], b"W7" + s[2:], s[:5], s[:6], s[:7], b"W6\x03\x00\x03\x00" + s[6:]]:
        assert codec.decode(bad) is None and oracle.decode(bad) is None
    flat = np.full((64, 64, 3), 77, np.uint8)  # all-zero detail: SURVEY.md App. D-1
    fs, _ = oracle.encode(flat)
    assert codec.encode(flat) == fs
    a, b = codec.decode(fs), oracle.decode(fs)
    assert a.shape == b.shape and np.array_equal(a, b)


# ------------------------------------------------------------------ BASELINE.json configs at full size

def test_4k_and_8k_pins(codec, pins_big):
    for rec in pins_big:
        img = make_image(rec["spec"])
        assert sha(img.tobytes()) == rec["pixels_sha"]
        full = codec.encode(img)
        assert len(full) == rec["full_len"] and sha(full) == rec["full_sha"], spec_id(rec["spec"])
        for case in rec["cases"]:
            cap = case["cap"]
            stream = full if cap is None else codec.encode(img, cap)
            assert sha(stream) == case["stream_sha"], (spec_id(rec["spec"]), cap)
            dec = codec.decode(stream)
            assert list(dec.shape) == case["decoded"]["shape"]
            assert sha(np.ascontiguousarray(dec).tobytes()) == case["decoded"]["sha"], (spec_id(rec["spec"]), cap)
            if cap is None:
                assert np.array_equal(dec, img)


def test_survey_pins(codec, oracle):
    # SURVEY.md App. E.1 (first 32 hex digits), measured there with the unmodified reference
    for (w, h, kind, size, pin) in [(1920, 1080, "photo", 3048215, "c5b662f845a94d038fd830a8c95451bc"),
                                    (1920, 1080, "noise", 6807283, "9148bc6e7891eb2d2e5815b651cf1aab"),
                                    (3840, 2160, "photo", 12216090, "63f44307a0de6ba2a26ae9074d724298"),
                                    (7680, 4320, "photo", 48863617, "f14ef79d0680a4ace7daa2b9e8063651"),
                                    (7680, 4320, "noise", 108895348, "9d30508a3cda1eaefceeb7e814adda31")]:
        s = codec.encode(oracle.synth(w, h, kind, 1))
        assert len(s) == size and sha(s)[:32] == pin, (w, h, kind)


def test_16k_stress_roundtrip(codec, oracle):
    """16384x16384 RGB, 12 levels: pin of SURVEY.md App. E.1 + the size-independent properties
    decode(encode(x)) == x and encode(x, cap) == encode(x)[:cap]"""
    img = oracle.synth(16384, 16384, "photo", 1)
    assert sha(img.tobytes())[:32] == "83ea32ea85eed7b384992340ee22a777"
    s = codec.encode(img)
    assert len(s) == 395131715 and sha(s)[:32] == "c77d5aff9242788e9bb23170a7ee730d"
    cap = 50_000_000
    assert codec.encode(img, cap) == s[:cap]
    dec = codec.decode(s)
    assert dec.shape == img.shape and np.array_equal(dec, img)


def test_batch_of_1080p_images(codec, oracle):
    """per-image seeds of the batch config; every stream must round-trip and a sample must match the oracle"""
    for seed in range(8):
        img = oracle.synth(1920, 1080, "photo", seed)
        s = codec.encode(img)
        if seed in (0, 5):
            assert s == oracle.encode(img)[0]
        assert np.array_equal(codec.decode(s), img)


def test_adversarial_patterns(codec, oracle):
    """checkerboards and stripes at several scales: largest coefficient growth through the levels, every token kind
    (long runs, order excursions), streams that are mostly refinement bits"""
    yy, xx = np.mgrid[0:256, 0:384]
    pats = [((xx ^ yy) & 1) * 255, ((xx >> 1) ^ (yy >> 2)) % 2 * 255, (xx & 1) * 255, ((xx // 7 + yy // 5) % 2) * 255,
            ((xx * yy) % 251)]
    for k, p2 in enumerate(pats):
        img = np.stack([p2, np.roll(p2, k + 1, axis=0), 255 - p2], axis=2).astype(np.uint8)
        want, _ = oracle.encode(img)
        assert codec.encode(img) == want, k
        for cap in (len(want) // 3, len(want) - 1):
            assert codec.encode(img, cap) == want[:cap]
            a, b = codec.decode(want[:cap]), oracle.decode(want[:cap])
            assert a.shape == b.shape and np.array_equal(a, b), (k, cap)
        assert np.array_equal(codec.decode(want), img), k


def test_concurrent_contexts(oracle):
    """several contexts (one CUDA stream each) driven from their own host threads, as bench.py does: every thread's
    streams and pixels must equal the oracle's"""
    import threading
    import dwt_b200 as D
    imgs = [oracle.synth(640 + 64 * i, 360 + 24 * i, "photo" if i % 2 == 0 else "noise", 40 + i) for i in range(4)]
    want = [oracle.encode(im)[0] for im in imgs]
    cods = [D.Codec(0) for _ in imgs]
    errors = []

    def work(i):
        try:
            for _ in range(6):
                s = cods[i].encode(imgs[i])
                if s != want[i]:
                    errors.append("stream %d" % i)
                d = cods[i].decode(s)
                if d is None or not np.array_equal(d, imgs[i]):
                    errors.append("pixels %d" % i)
                cap = len(want[i]) // 2
                a, b = cods[i].decode(want[i][:cap]), oracle.decode(want[i][:cap])
                if a.shape != b.shape or not np.array_equal(a, b):
                    errors.append("truncated %d" % i)
        except Exception as ex:  # noqa: BLE001
            errors.append("%d: %r" % (i, ex))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(imgs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for c in cods:
        c.close()
    assert not errors, errors


def test_throughput_scan_matches_latency_scan(codec, oracle):
    """the decoder has two scan kernels (parallel per window / one thread per window and chain, chosen by the in-flight
    hint for streams of >= 2048 windows): both must give the same pixels, lossless and truncated"""
    import dwt_b200 as D
    img = oracle.synth(3840, 2160, "photo", 1)
    s = codec.encode(img)
    assert len(s) == 12216090  # 4K pin: 2 983 scan windows
    other = D.Codec(0)
    other.set_in_flight(8)
    try:
        for cap in (len(s), len(s) - 1, len(s) // 2, 9_000_001, 70_000):
            a, b = other.decode(s[:cap]), codec.decode(s[:cap])
            assert a.shape == b.shape and np.array_equal(a, b), cap
        assert np.array_equal(other.decode(s), img)
        noise = oracle.synth(2048, 2048, "noise", 3)
        sn = codec.encode(noise)
        assert np.array_equal(other.decode(sn), noise)
        assert np.array_equal(other.decode(sn[:len(sn) // 3]), codec.decode(sn[:len(sn) // 3]))
    finally:
        other.close()


def test_pool_batch_matches_oracle(oracle):
    """dwt_pool: a batch of independent images of different sizes through 3 workers; streams, truncated streams
    and decoded pixels must equal the oracle's, item by item"""
    import dwt_b200 as D
    imgs = [oracle.synth(200 + 37 * i, 120 + 29 * i, "photo" if i % 3 else "noise", 60 + i) for i in range(7)]
    imgs.append(oracle.synth(133, 100, "photo", 5)[:, :, 1].copy())  # gray
    want = [oracle.encode(im)[0] for im in imgs]
    pool = D.Pool(0, 3)
    try:
        assert pool.encode_batch(imgs) == want
        cut = pool.encode_batch(imgs, 777)
        assert cut == [w[:777] for w in want]
        shapes = [im.shape if im.ndim == 3 else im.shape + (1,) for im in imgs]
        dec = pool.decode_batch(want, shapes)
        for a, b in zip(dec, imgs):
            assert a.shape == b.shape and np.array_equal(a, b)
        part = pool.decode_batch(cut, shapes)
        for a, s in zip(part, cut):
            b = oracle.decode(s)
            assert a.shape == b.shape and np.array_equal(a, b)
    finally:
        pool.close()


# ------------------------------------------------------------------ the drop-in programs

def test_cli_roundtrip_matches_reference_behaviour(codec, oracle, tmp_path):
    enc, dec = os.path.join(ROOT, "encode"), os.path.join(ROOT, "decode")
    img = oracle.synth(320, 240, "photo", 11)
    src = tmp_path / "in.pnm"
    src.write_bytes(b"P6\n# a comment line\n320 240\n255\n" + img.tobytes())
    want, st = oracle.encode(img)
    out = tmp_path / "o.dwt"
    r = subprocess.run([enc, str(src), str(out)], capture_output=True)
    assert r.returncode == 0 and out.read_bytes() == want
    assert r.stderr.decode().splitlines() == ["%d bits for meta data" % st.meta_bits, "%d bits for root image" % st.root_bits,
                                              "%d bits (%d KiB) encoded" % (st.total_bits, (len(want) + 512) // 1024)]
    cap = 4096
    r = subprocess.run([enc, str(src), str(out), str(cap)], capture_output=True)
    assert r.returncode == 0 and out.read_bytes() == want[:cap]
    assert r.stderr.decode().splitlines()[-1] == "%d bits (%d KiB) encoded" % (8 * cap, (cap + 512) // 1024)
    back = tmp_path / "b.pnm"
    r = subprocess.run([dec, str(out), str(back)], capture_output=True)
    assert r.returncode == 0
    ref = oracle.decode(want[:cap])
    assert back.read_bytes() == oracle.pnm_bytes(ref)
    # stdin / stdout
    r = subprocess.run("%s %s - | %s - -" % (enc, src, dec), shell=True, capture_output=True)
    assert r.returncode == 0 and r.stdout == oracle.pnm_bytes(img)
    # PIXELS argument
    r = subprocess.run([dec, str(tmp_path / "full.dwt"), str(back), "5000"], capture_output=True)
    assert r.returncode == 1  # missing input
    (tmp_path / "full.dwt").write_bytes(want)
    r = subprocess.run([dec, str(tmp_path / "full.dwt"), str(back), "5000"], capture_output=True)
    assert r.returncode == 0 and back.read_bytes() == oracle.pnm_bytes(oracle.decode(want, 5000))
    # truncated inside the root image: exit 1, no output (decode.c:180-186)
    (tmp_path / "short.dwt").write_bytes(want[:9])
    gone = tmp_path / "none.pnm"
    r = subprocess.run([dec, str(tmp_path / "short.dwt"), str(gone)], capture_output=True)
    assert r.returncode == 1 and not gone.exists()


def test_batch_cli_matches_per_file_programs(oracle, tmp_path):
    """dwtbatch codes every file exactly as encode / decode would"""
    tool = os.path.join(ROOT, "dwtbatch")
    src, enc_dir, dec_dir = tmp_path / "src", tmp_path / "enc", tmp_path / "dec"
    for d in (src, enc_dir, dec_dir):
        d.mkdir()
    imgs = {}
    for i, (w, h, kind) in enumerate([(320, 240, "photo"), (133, 100, "noise"), (640, 360, "photo"), (64, 48, "photo"),
                                      (257, 255, "noise"), (1920, 1080, "photo"), (96, 96, "gray")]):
        if kind == "gray":
            img = oracle.synth(w, h, "photo", 40 + i)[:, :, 1].copy()
            hdr = b"P5\n%d %d\n255\n" % (w, h)
        else:
            img = oracle.synth(w, h, kind, 40 + i)
            hdr = b"P6\n%d %d\n255\n" % (w, h)
        imgs["im%d" % i] = img
        (src / ("im%d.pnm" % i)).write_bytes(hdr + img.tobytes())
    (src / "bad.pnm").write_bytes(b"P6\n4 4\n255\n" + bytes(48))  # rejected: smaller than 8x8 (encode.c:144-146)
    names = sorted(str(p) for p in src.iterdir())
    r = subprocess.run([tool, "encode", "-j", "3", str(enc_dir)] + names, capture_output=True)
    assert r.returncode == 1 and b"1 of %d files failed" % len(names) in r.stderr, r.stderr
    assert not (enc_dir / "bad.dwt").exists()
    for k, img in imgs.items():
        assert (enc_dir / (k + ".dwt")).read_bytes() == oracle.encode(img)[0], k
    # capacity for every file, paths on stdin
    cap_dir = tmp_path / "cap"
    cap_dir.mkdir()
    good = [n for n in names if not n.endswith("bad.pnm")]
    r = subprocess.run([tool, "encode", "-c", "3000", str(cap_dir)], input="\n".join(good).encode(), capture_output=True)
    assert r.returncode == 0, r.stderr
    for k, img in imgs.items():
        assert (cap_dir / (k + ".dwt")).read_bytes() == oracle.encode(img, 3000)[0], k
    # decode: full streams, then truncated streams with a PIXELS bound
    r = subprocess.run([tool, "decode", str(dec_dir)] + sorted(str(p) for p in enc_dir.iterdir()), capture_output=True)
    assert r.returncode == 0, r.stderr
    for k, img in imgs.items():
        assert (dec_dir / (k + ".pnm")).read_bytes() == oracle.pnm_bytes(img), k
    low_dir = tmp_path / "low"
    low_dir.mkdir()
    r = subprocess.run([tool, "decode", "-p", "20000", "-j", "2", str(low_dir)] + sorted(str(p) for p in cap_dir.iterdir()),
                       capture_output=True)
    assert r.returncode == 0, r.stderr
    for k, img in imgs.items():
        want = oracle.pnm_bytes(oracle.decode(oracle.encode(img, 3000)[0], 20000))
        assert (low_dir / (k + ".pnm")).read_bytes() == want, k


@pytest.mark.parametrize("shape", [(4096, 8), (8, 4096), (4096, 17), (2500, 9)])
def test_thin_strips(codec, oracle, shape):
    """one dimension at the minimum: a single level whose Hilbert square is mostly empty (the reference walks all of it,
    which is why this stays at 4096: its time grows with the square of the longer side, and 65536 overflows its int counters)"""
    w, h = shape
    img = oracle.synth(w, h, "photo", 7)
    want, st = oracle.encode(img)
    assert codec.encode(img) == want
    assert np.array_equal(codec.decode(want), img)
    cap = len(want) // 3
    assert codec.encode(img, cap) == want[:cap]
    for cut, pm in [(cap, -1), (len(want) - 5, -1), (len(want), 10000)]:
        a, b = codec.decode(want[:cut], pm), oracle.decode(want[:cut], pm)
        assert (a is None) == (b is None)  # the root image of a strip is large: a cut inside it is "exit 1, no output"
        if a is not None:
            assert a.shape == b.shape and np.array_equal(a, b)
