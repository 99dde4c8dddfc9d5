"""Function-level pins: the entry points of libdwt_b200.so next to the reference's OWN functions of the same name.

oracle/_ref/libref_shim.so is the reference's header-only modules compiled unmodified (oracle/ref_shim.c only
names the headers; built by oracle/Makefile where /root/reference exists, travels to the GPU box with the built
files).  Both libraries are driven with the same seeded call sequences:

  bytes / bits / vli / rle writers and readers   bytes.h:23-118, bits.h:23-106, vli.h:67-101, rle.h:37-103
  ilog2 / compute_lengths                        utils.h:9-40
  cdf53 / icdf53 (GPU)                           cdf53.h:9,36
  colour transforms (GPU)                        image.h:39-65

tests/golden/entry_pins.json (made by tests/golden/make_entry_pins.py from the shim) keeps the same checks alive on
a checkout that has no shim.
"""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "_ref", "libref_shim.so")
PINS = os.path.join(ROOT, "tests", "golden", "entry_pins.json")

VP = C.c_void_p
STREAM_API = [("bytes_reader", VP, [C.c_char_p]), ("bytes_writer", VP, [C.c_char_p, C.c_int]), ("bytes_count", C.c_int, [VP]),
              ("close_bytes_reader", None, [VP]), ("close_bytes_writer", None, [VP]), ("put_byte", C.c_int, [VP, C.c_int]),
              ("write_bytes", C.c_int, [VP, C.c_int, C.c_int]), ("get_byte", C.c_int, [VP]),
              ("read_bytes", C.c_int, [VP, C.POINTER(C.c_int), C.c_int]),
              ("bits_reader", VP, [VP]), ("bits_writer", VP, [VP]), ("bits_count", C.c_int, [VP]),
              ("close_bits_reader", None, [VP]), ("close_bits_writer", None, [VP]), ("put_bit", C.c_int, [VP, C.c_int]),
              ("write_bits", C.c_int, [VP, C.c_int, C.c_int]), ("get_bit", C.c_int, [VP]),
              ("read_bits", C.c_int, [VP, C.POINTER(C.c_int), C.c_int]),
              ("vli_reader", VP, [VP]), ("vli_writer", VP, [VP]), ("delete_vli_reader", None, [VP]),
              ("delete_vli_writer", None, [VP]), ("vli_put_bit", C.c_int, [VP, C.c_int]), ("vli_get_bit", C.c_int, [VP]),
              ("vli_write_bits", C.c_int, [VP, C.c_int, C.c_int]), ("vli_read_bits", C.c_int, [VP, C.POINTER(C.c_int), C.c_int]),
              ("put_vli", C.c_int, [VP, C.c_int]), ("get_vli", C.c_int, [VP]),
              ("rle_reader", VP, [VP]), ("rle_writer", VP, [VP]), ("rle_flush", C.c_int, [VP]),
              ("delete_rle_reader", None, [VP]), ("delete_rle_writer", None, [VP]), ("put_rle", C.c_int, [VP, C.c_int]),
              ("get_rle", C.c_int, [VP]), ("rle_put_bit", C.c_int, [VP, C.c_int]), ("rle_get_bit", C.c_int, [VP]),
              ("ilog2", C.c_int, [C.c_int]),
              ("compute_lengths", C.c_int, [C.POINTER(C.c_int)] * 4 + [C.c_int] * 3),
              ("cdf53", None, [C.POINTER(C.c_int)] * 2 + [C.c_int] * 4), ("icdf53", None, [C.POINTER(C.c_int)] * 2 + [C.c_int] * 4)]


def typed(L, extra=()):
    for name, res, args in list(STREAM_API) + list(extra):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    return L


def ours():
    import dwt_b200
    return typed(C.CDLL(dwt_b200.LIB_PATH))   # a private handle: the typed prototypes of other tests stay untouched


def shim():
    if not os.path.exists(SHIM):
        return None
    return typed(C.CDLL(SHIM), [("rgb2ycocg", None, [C.POINTER(C.c_int)]), ("ycocg2rgb", None, [C.POINTER(C.c_int)]),
                                ("ref_hilbert_xy", None, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)])])


def pins():
    with open(PINS) as f:
        return json.load(f)


# ------------------------------------------------------------------ seeded call sequences

def write_script(seed, n=4000):
    """a mix of every writer call, with run lengths and values from tiny to huge"""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(n):
        k = int(rng.integers(0, 10))
        if k < 5:
            ops.append(("put_rle", int(rng.random() < 0.15)))
        elif k < 7:
            ops.append(("rle_put_bit", int(rng.integers(0, 2))))
        elif k == 7:
            mag = int(rng.integers(0, 27))
            ops.append(("put_vli", int(rng.integers(0, 1 << mag))))
        elif k == 8:
            nb = int(rng.integers(1, 20))
            ops.append(("vli_write_bits", int(rng.integers(0, 1 << nb)), nb))
        else:
            ops.append(("vli_put_bit", int(rng.integers(0, 2))))
    if seed % 3 == 0:   # a very long run, like the 66 M zero run of the 8K photo stream
        ops += [("put_rle", 0)] * 70000 + [("put_rle", 1)]
    return ops


def run_writer(L, path, cap, ops, header=True):
    """returns (file bytes, return codes, bit counts seen on the way)"""
    bw = L.bytes_writer(path.encode(), cap)
    rets, counts = [], []
    if header:
        rets.append(L.put_byte(bw, ord("W")))
        rets.append(L.write_bytes(bw, 0x1234, 2))
    bits = L.bits_writer(bw)
    vli = L.vli_writer(bits)
    rle = L.rle_writer(vli)
    for i, op in enumerate(ops):
        if op[0] == "put_rle":
            rets.append(L.put_rle(rle, op[1]))
        elif op[0] == "rle_put_bit":
            rets.append(L.rle_put_bit(rle, op[1]))
        elif op[0] == "put_vli":
            rets.append(L.put_vli(vli, op[1]))
        elif op[0] == "vli_write_bits":
            rets.append(L.vli_write_bits(vli, op[1], op[2]))
        else:
            rets.append(L.vli_put_bit(vli, op[1]))
        if i % 97 == 0:
            counts.append(L.bits_count(bits))
    rets.append(L.rle_flush(rle))
    counts.append(L.bits_count(bits))
    # (delete_rle_writer would complain on stderr about a latched error; both libraries do)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)
    try:
        L.delete_rle_writer(rle)
    finally:
        os.dup2(saved, 2)
        os.close(saved)
        os.close(devnull)
    L.delete_vli_writer(vli)
    L.close_bits_writer(bits)
    counts.append(L.bytes_count(bw))
    L.close_bytes_writer(bw)
    with open(path, "rb") as f:
        return f.read(), rets, counts


def read_script(seed, n=6000):
    rng = np.random.default_rng(1000 + seed)
    kinds = ["get_rle"] * 6 + ["rle_get_bit"] * 2 + ["get_vli", "vli_read_bits", "vli_get_bit", "get_bit"]
    return [(kinds[int(rng.integers(0, len(kinds)))], int(rng.integers(1, 18))) for _ in range(n)]


def run_reader(L, path, ops, header=True):
    devnull = os.open(os.devnull, os.O_WRONLY)   # get_byte prints "reached end of file" at EOF (bytes.h:99-103)
    saved = os.dup(2)
    os.dup2(devnull, 2)
    try:
        br = L.bytes_reader(path.encode())
        out = []
        if header:
            v = C.c_int(-7)
            out.append(L.get_byte(br))
            out.append((L.read_bytes(br, C.byref(v), 2), v.value))
        bits = L.bits_reader(br)
        vli = L.vli_reader(bits)
        rle = L.rle_reader(vli)
        for kind, nb in ops:
            if kind == "get_rle":
                out.append(L.get_rle(rle))
            elif kind == "rle_get_bit":
                out.append(L.rle_get_bit(rle))
            elif kind == "get_vli":
                out.append(L.get_vli(vli))
            elif kind == "vli_read_bits":
                v = C.c_int(-7)
                r = L.vli_read_bits(vli, C.byref(v), nb)
                out.append((r, v.value if r == 0 else None))
            elif kind == "vli_get_bit":
                out.append(L.vli_get_bit(vli))
            else:
                out.append(L.get_bit(bits))
        L.delete_rle_reader(rle)
        L.delete_vli_reader(vli)
        L.close_bits_reader(bits)
        L.close_bytes_reader(br)
    finally:
        os.dup2(saved, 2)
        os.close(saved)
        os.close(devnull)
    return out


def digest(obj):
    return hashlib.sha256(json.dumps(obj, sort_keys=True, default=lambda b: b.hex()).encode()).hexdigest()[:32]


CAPS = [0, 1, 2, 3, 40, 1000]
SEEDS = list(range(6))


def writer_case(L, tmp, seed, cap):
    data, rets, counts = run_writer(L, os.path.join(tmp, "w_%d_%d.bin" % (seed, cap)), cap, write_script(seed))
    return dict(data=data, rets=rets, counts=counts)


def reader_case(L, tmp, seed, stream, cut):
    path = os.path.join(tmp, "r_%d_%d.bin" % (seed, cut))
    with open(path, "wb") as f:
        f.write(stream[:cut])
    return run_reader(L, path, read_script(seed))


def reader_cuts(stream):
    n = len(stream)
    return sorted(set([0, 1, 2, 3, 4, n // 7, n // 2, max(0, n - 1), n]))


# ------------------------------------------------------------------ CPU: stream + geometry entry points

def test_stream_writers_match_reference_functions(built, tmp_path):
    A, R, P = ours(), shim(), pins()
    for seed in SEEDS:
        for cap in CAPS:
            mine = writer_case(A, str(tmp_path), seed, cap)
            assert digest(mine) == P["writers"]["%d/%d" % (seed, cap)], (seed, cap)
            if R is not None:
                ref = writer_case(R, str(tmp_path), seed, cap)
                assert mine["data"] == ref["data"], (seed, cap)
                assert mine["rets"] == ref["rets"], (seed, cap)       # 0 / -2 at the same calls, latched in rle->cnt
                assert mine["counts"] == ref["counts"], (seed, cap)   # bits_count / bytes_count incl. after the cut


def test_stream_readers_match_reference_functions(built, tmp_path):
    A, R, P = ours(), shim(), pins()
    for seed in SEEDS:
        stream = run_writer(A, str(tmp_path / ("s%d.bin" % seed)), 0, write_script(seed))[0]
        for cut in reader_cuts(stream):
            mine = reader_case(A, str(tmp_path), seed, stream, cut)
            assert digest(mine) == P["readers"]["%d/%d" % (seed, cut)], (seed, cut)
            if R is not None:
                assert mine == reader_case(R, str(tmp_path), seed, stream, cut), (seed, cut)
        # a reader let loose on noise: values, -1 at EOF and rle_get_bit's "ret != 1 -> -1" (rle.h:91-103)
        noise = np.random.default_rng(seed).integers(0, 256, 3000).astype(np.uint8).tobytes()
        mine = reader_case(A, str(tmp_path), 50 + seed, noise, len(noise))
        assert digest(mine) == P["readers_noise"][str(seed)]
        if R is not None:
            assert mine == reader_case(R, str(tmp_path), 50 + seed, noise, len(noise))


def geometry_cases():
    rng = np.random.default_rng(7)
    sizes = [(8, 8), (9, 8), (15, 15), (16, 16), (17, 31), (8, 500), (3000, 9), (320, 240), (1001, 777), (1920, 1080),
             (3840, 2160), (7680, 4320), (16384, 16384), (65536, 8), (65536, 65536), (4, 4), (1, 1), (8, 7)]
    sizes += [(int(rng.integers(1, 70000)), int(rng.integers(1, 70000))) for _ in range(200)]
    return sizes


def geometry_of(L, w, h):
    arrs = [(C.c_int * 16)(*([-1] * 16)) for _ in range(4)]
    levels = L.compute_lengths(arrs[0], arrs[1], arrs[2], arrs[3], w, h, 8)
    return [levels] + [list(a)[:levels + 1] for a in arrs]


def test_geometry_matches_reference_functions(built):
    A, R, P = ours(), shim(), pins()
    xs = list(range(-3, 70)) + [255, 256, 257, 65535, 65536, (1 << 29) - 1, 1 << 29, 0x7fffffff]
    assert [A.ilog2(x) for x in xs] == P["ilog2"]
    geo = [geometry_of(A, w, h) for (w, h) in geometry_cases()]
    assert digest(geo) == P["geometry"]
    if R is not None:
        assert [R.ilog2(x) for x in xs] == P["ilog2"]
        assert geo == [geometry_of(R, w, h) for (w, h) in geometry_cases()]


def test_oracle_hilbert_matches_reference_function(built, oracle):
    """the oracle's curve (used to check the GPU's 32x32 table + cell orientations) against hilbert.h:15-34"""
    R = shim()
    if R is None:
        pytest.skip("oracle/_ref/libref_shim.so is only built where /root/reference exists")
    x, y, rx, ry = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rng = np.random.default_rng(3)
    for n in [1, 2, 4, 8, 32, 512, 8192, 16384]:
        for d in [0, 1, 2, 3, n * n - 1] + [int(v) for v in rng.integers(0, n * n, 300)]:
            oracle.lib().orc_hilbert(n, d, C.byref(x), C.byref(y))
            R.ref_hilbert_xy(n, d, C.byref(rx), C.byref(ry))
            assert (x.value, y.value) == (rx.value, ry.value), (n, d)


# ------------------------------------------------------------------ GPU: transform entry points

def lifting_cases():
    rng = np.random.default_rng(11)
    cases = [(8, 1, 1, 1), (9, 3, 3, 3), (16, 3, 5, 4), (33, 7, 7, 9), (100, 4, 9, 4), (2, 1, 1, 1), (3, 2, 2, 2), (1000, 3, 3, 3),
             (4320, 6, 6, 8), (7680, 3, 3, 3)]
    for _ in range(30):
        ch = int(rng.integers(1, 12))
        cases.append((int(rng.integers(2, 700)), ch, ch + int(rng.integers(0, 5)), ch + int(rng.integers(0, 5))))
    return cases   # (N, CH, SI, SO)


def lifting_input(i, N, CH, SI):
    rng = np.random.default_rng(500 + i)
    lim = [2000, 1 << 20, 255][i % 3]
    return rng.integers(-lim, lim + 1, (N - 1) * SI + CH).astype(np.int32)


@pytest.mark.gpu
def test_cdf53_icdf53_match_reference_functions(built, codec):
    A, R, P = ours(), shim(), pins()
    ip = C.POINTER(C.c_int)
    for i, (N, CH, SI, SO) in enumerate(lifting_cases()):
        x = lifting_input(i, N, CH, SI)
        span_out = (N - 1) * SO + CH
        res = {}
        for name, L in (("ours", A), ("ref", R)):
            if L is None:
                continue
            xin, out = x.copy(), np.full(span_out, 77, np.int32)   # gaps between the CH lanes must survive
            L.cdf53(out.ctypes.data_as(ip), xin.ctypes.data_as(ip), N, SO, SI, CH)
            # inverse: strides swapped so that the deinterleaved buffer is read with SO and written with SI
            back = np.full(x.size, 55, np.int32)
            keep = out.copy()
            L.icdf53(back.ctypes.data_as(ip), out.ctypes.data_as(ip), N, SI, SO, CH)
            assert np.array_equal(out, keep), "icdf53 must leave `in` alone (cdf53.h:36-61)"
            res[name] = (out, xin, back)
        mine = res["ours"]
        assert digest([a.tobytes() for a in mine]) == P["lifting"][str(i)], (N, CH, SI, SO)
        if "ref" in res:
            for a, b, what in zip(mine, res["ref"], ("out", "clobbered in (cdf53.h:12-23)", "icdf53 out")):
                assert np.array_equal(a, b), (N, CH, SI, SO, what)


@pytest.mark.gpu
def test_colour_transforms_match_reference_functions(built, codec):
    import dwt_b200 as D
    R, P = shim(), pins()
    rng = np.random.default_rng(12)
    n = 4096
    rgb = rng.integers(0, 256, 3 * n).astype(np.int32)
    wild = rng.integers(-700, 700, 3 * n).astype(np.int32)   # beyond the clamps of image.h:41-43
    fwd, inv = D.ycocg_from_rgb(rgb), D.rgb_from_ycocg(wild)
    assert digest([fwd.tobytes(), inv.tobytes()]) == P["colour"]
    if R is not None:
        ip = C.POINTER(C.c_int)
        a, b = rgb.copy(), wild.copy()
        for k in range(n):
            R.rgb2ycocg(C.cast(a.ctypes.data + 12 * k, ip))
            R.ycocg2rgb(C.cast(b.ctypes.data + 12 * k, ip))
        assert np.array_equal(fwd, a) and np.array_equal(inv, b)
