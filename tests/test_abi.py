"""CPU: the C-ABI library loads and exports every symbol include/dwt_b200.h declares; the host-side stream
entry points (bytes / bits / vli / rle) behave like the reference's; the CLIs keep the reference's argv rules.
No GPU compute is attempted here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "dwt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_library_exports_every_declared_symbol(built):
    import dwt_b200
    lib = dwt_b200.lib()
    names = header_functions()
    assert len(names) > 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(set(dwt_b200.ABI_SYMBOLS)) == names


def test_no_cpu_fallback_message(built):
    """without a CUDA device the product must fail loudly, never compute on the CPU"""
    import dwt_b200
    lib = dwt_b200.lib()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    assert not lib.dwt_ctx_create(-1)
    assert b"no CPU fallback" in lib.dwt_last_error()


def test_compute_lengths_matches_oracle(built, oracle):
    import dwt_b200
    for (w, h) in [(8, 8), (9, 8), (320, 240), (1001, 777), (1920, 1080), (7680, 4320), (16384, 16384), (65536, 8)]:
        assert dwt_b200.geometry(w, h) == oracle.geometry(w, h)


def _lib():
    import dwt_b200
    L = dwt_b200.lib()
    vp = C.c_void_p
    for name, res, args in [("bytes_writer_mem", vp, [C.c_int]), ("bytes_writer_data", C.POINTER(C.c_uint8), [vp, C.POINTER(C.c_size_t)]),
                            ("bytes_reader_mem", vp, [C.POINTER(C.c_uint8), C.c_size_t]), ("bits_writer", vp, [vp]),
                            ("bits_reader", vp, [vp]), ("vli_writer", vp, [vp]), ("vli_reader", vp, [vp]),
                            ("rle_writer", vp, [vp]), ("rle_reader", vp, [vp]), ("put_vli", C.c_int, [vp, C.c_int]),
                            ("get_vli", C.c_int, [vp]), ("put_rle", C.c_int, [vp, C.c_int]), ("get_rle", C.c_int, [vp]),
                            ("rle_put_bit", C.c_int, [vp, C.c_int]), ("rle_get_bit", C.c_int, [vp]), ("rle_flush", C.c_int, [vp]),
                            ("bits_count", C.c_int, [vp]), ("close_bits_writer", None, [vp]), ("close_bits_reader", None, [vp]),
                            ("close_bytes_writer", None, [vp]), ("close_bytes_reader", None, [vp]),
                            ("delete_vli_writer", None, [vp]), ("delete_vli_reader", None, [vp]),
                            ("delete_rle_writer", None, [vp]), ("delete_rle_reader", None, [vp]),
                            ("put_byte", C.c_int, [vp, C.c_int]), ("bytes_count", C.c_int, [vp])]:
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    return L


def py_vli(vals):
    """bit string of the adaptive Rice code, restated from vli.h:67-84"""
    bits, order = [], 0
    for v in vals:
        while v >= (1 << order):
            bits.append(0)
            v -= 1 << order
            order += 1
        bits.append(1)
        bits += [(v >> i) & 1 for i in range(order)]
        order = max(order - 2, 0)
    return bits


def pack(bits):
    out = bytearray((len(bits) + 7) // 8)
    for i, b in enumerate(bits):
        out[i >> 3] |= b << (i & 7)
    return bytes(out)


def test_vli_writer_reader_roundtrip(built):
    L = _lib()
    rng = np.random.default_rng(0)
    vals = [int(v) for v in np.concatenate([rng.integers(0, 4, 200), rng.integers(0, 100000, 50), [0, 1, 2, 3, 66254184]])]
    bw = L.bytes_writer_mem(0)
    bits = L.bits_writer(bw)
    vli = L.vli_writer(bits)
    for v in vals:
        assert L.put_vli(vli, v) == 0
    nbits = L.bits_count(bits)
    L.delete_vli_writer(vli)
    L.close_bits_writer(bits)
    n = C.c_size_t()
    data = C.string_at(L.bytes_writer_data(bw, C.byref(n)), n.value)
    L.close_bytes_writer(bw)
    want = py_vli(vals)
    assert nbits == len(want) and data == pack(want)
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    br = L.bytes_reader_mem(buf, len(data))
    bits = L.bits_reader(br)
    vli = L.vli_reader(bits)
    assert [L.get_vli(vli) for _ in vals] == vals
    L.delete_vli_reader(vli)
    L.close_bits_reader(bits)
    L.close_bytes_reader(br)


def test_rle_semantics_and_capacity(built):
    L = _lib()
    # symbols: runs of zeros closed by ones, raw bits in between (rle.h:56-103)
    seq = [("s", 0)] * 5 + [("s", 1), ("r", 1), ("s", 1), ("r", 0)] + [("s", 0)] * 3 + [("r", 1), ("r", 0)] + [("s", 0)] * 2
    bw = L.bytes_writer_mem(0)
    bits = L.bits_writer(bw)
    vli = L.vli_writer(bits)
    rle = L.rle_writer(vli)
    for kind, b in seq:
        assert (L.put_rle(rle, b) if kind == "s" else L.rle_put_bit(rle, b)) == 0
    assert L.rle_flush(rle) == 0
    L.delete_rle_writer(rle)
    L.delete_vli_writer(vli)
    L.close_bits_writer(bits)
    n = C.c_size_t()
    data = C.string_at(L.bytes_writer_data(bw, C.byref(n)), n.value)
    L.close_bytes_writer(bw)
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    br = L.bytes_reader_mem(buf, len(data))
    bits = L.bits_reader(br)
    vli = L.vli_reader(bits)
    rle = L.rle_reader(vli)
    for kind, b in seq:
        assert (L.get_rle(rle) if kind == "s" else L.rle_get_bit(rle)) == b
    L.delete_rle_reader(rle)
    L.delete_vli_reader(vli)
    L.close_bits_reader(bits)
    L.close_bytes_reader(br)
    # capacity: put_byte refuses with -2 once cnt >= cap and writes nothing (bytes.h:77-78)
    bw = L.bytes_writer_mem(3)
    assert [L.put_byte(bw, i) for i in range(5)] == [0, 0, 0, -2, -2]
    assert L.bytes_count(bw) == 3
    L.close_bytes_writer(bw)


def test_cli_usage_and_rejections(built, tmp_path):
    enc, dec = os.path.join(ROOT, "encode"), os.path.join(ROOT, "decode")
    assert os.path.exists(enc) and os.path.exists(dec)
    r = subprocess.run([enc], capture_output=True)
    assert r.returncode == 1 and b"usage:" in r.stderr and b"input.pnm output.dwt [CAPACITY]" in r.stderr
    r = subprocess.run([dec, "a", "b", "c", "d"], capture_output=True)
    assert r.returncode == 1 and b"input.dwt output.pnm [PIXELS]" in r.stderr
    # too small an image: exit 1 and no output file is created (encode.c:139-146,166)
    small = tmp_path / "s.pnm"
    small.write_bytes(b"P6 4 4 255\n" + bytes(48))
    out = tmp_path / "s.dwt"
    r = subprocess.run([enc, str(small), str(out)], capture_output=True)
    assert r.returncode == 1 and not out.exists()
    # not a PNM
    bad = tmp_path / "bad.pnm"
    bad.write_bytes(b"hello")
    r = subprocess.run([enc, str(bad), str(out)], capture_output=True)
    assert r.returncode == 1 and not out.exists()


def test_batch_cli_usage(built):
    tool = os.path.join(ROOT, "dwtbatch")
    assert os.path.exists(tool)
    for argv in ([tool], [tool, "transcode", "out"], [tool, "encode"], [tool, "encode", "-j", "0", "out"]):
        r = subprocess.run(argv, capture_output=True)
        assert r.returncode == 1 and b"usage:" in r.stderr and b"OUTDIR" in r.stderr
