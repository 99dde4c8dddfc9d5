"""ctypes view of oracle/liboracle.so and oracle/_ref/{encode,decode} -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (dwt_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


class Stats(C.Structure):
    _fields_ = [("meta_bits", C.c_longlong), ("root_bits", C.c_longlong), ("total_bits", C.c_longlong),
                ("bytes", C.c_longlong), ("planes", C.c_int * 3), ("levels", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        ip = C.POINTER(C.c_int)
        up = C.POINTER(C.c_uint8)
        L.orc_geometry.argtypes = [C.c_int, C.c_int, ip, ip, ip, ip]
        L.orc_geometry.restype = C.c_int
        L.orc_cdf53.argtypes = [ip, ip, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_icdf53.argtypes = [ip, ip, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_front_end.argtypes = [up, C.c_int, C.c_int, C.c_int, ip, ip, ip]
        L.orc_front_end.restype = C.c_int
        L.orc_encode.argtypes = [up, C.c_int, C.c_int, C.c_int, C.c_int, up, C.c_longlong, C.POINTER(Stats)]
        L.orc_encode.restype = C.c_longlong
        L.orc_decode.argtypes = [up, C.c_longlong, C.c_int, C.POINTER(up), ip, ip, ip]
        L.orc_decode.restype = C.c_int
        L.orc_decode_coeffs.argtypes = [up, C.c_longlong, C.c_int, C.POINTER(ip), ip, ip, ip, ip, ip]
        L.orc_decode_coeffs.restype = C.c_int
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_synth.argtypes = [up, C.c_int, C.c_int, C.c_int, C.c_uint32]
        L.orc_hilbert.argtypes = [C.c_int, C.c_int, ip, ip]
        _lib = L
    return _lib


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def geometry(w, h):
    arrs = [(C.c_int * 16)() for _ in range(4)]
    levels = lib().orc_geometry(w, h, *arrs)
    lengths, pixels, widths, heights = [list(a)[:levels + 1] for a in arrs]
    return dict(levels=levels, lengths=lengths, pixels=pixels, widths=widths, heights=heights)


def synth(w, h, kind="photo", seed=1):
    out = np.empty((h, w, 3), dtype=np.uint8)
    lib().orc_synth(_u8p(out), w, h, 1 if kind == "noise" else 0, seed)
    return out


def cdf53(x, N, SO, SI, CH, out_len=None):
    """returns (out, clobbered_in) like the reference cdf53(out, in, N, SO, SI, CH)."""
    x = np.ascontiguousarray(x, dtype=np.int32).copy()
    out = np.zeros(out_len if out_len else x.size, dtype=np.int32)
    lib().orc_cdf53(_ip(out), _ip(x), N, SO, SI, CH)
    return out, x


def icdf53(x, N, SO, SI, CH, out_len=None):
    x = np.ascontiguousarray(x, dtype=np.int32).copy()
    out = np.zeros(out_len if out_len else x.size, dtype=np.int32)
    lib().orc_icdf53(_ip(out), _ip(x), N, SO, SI, CH)
    return out


def front_end(img):
    """img: (h,w,ch) or (h,w) uint8 -> (pyramid int32 (h,w,ch), planar int32 (ch, w*h), planes list)"""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    pyr = np.zeros((h, w, ch), dtype=np.int32)
    lin = np.zeros((ch, w * h), dtype=np.int32)
    planes = (C.c_int * 3)()
    if lib().orc_front_end(_u8p(img), w, h, ch, _ip(pyr), _ip(lin), planes):
        raise ValueError("orc_front_end failed")
    return pyr, lin, list(planes)[:ch]


def encode(img, capacity=0):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    room = int(img.size * 2 + 4096)
    out = np.empty(room, dtype=np.uint8)
    st = Stats()
    n = lib().orc_encode(_u8p(img), w, h, ch, int(capacity), _u8p(out), room, C.byref(st))
    if n < 0:
        raise ValueError("orc_encode failed")
    return out[:n].tobytes(), st


def decode(stream, pixels_max=-1):
    """returns uint8 image (h,w,ch) / (h,w), or None where the reference exits 1 without output."""
    buf = np.frombuffer(bytes(stream), dtype=np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, dtype=np.uint8)
        n = 0
    else:
        n = buf.size
    pix = C.POINTER(C.c_uint8)()
    w, h, ch = C.c_int(), C.c_int(), C.c_int()
    r = lib().orc_decode(_u8p(buf), n, int(pixels_max), C.byref(pix), C.byref(w), C.byref(h), C.byref(ch))
    if r:
        return None
    shape = (h.value, w.value, 3) if ch.value == 3 else (h.value, w.value)
    arr = np.ctypeslib.as_array(pix, shape=(int(np.prod(shape)),)).copy().reshape(shape)
    lib().orc_free(pix)
    return arr


def decode_coeffs(stream, pixels_max=-1):
    buf = np.frombuffer(bytes(stream), dtype=np.uint8)
    planar = C.POINTER(C.c_int)()
    missing = (C.c_int * 48)()
    level, w, h, ch = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    r = lib().orc_decode_coeffs(_u8p(buf), buf.size, int(pixels_max), C.byref(planar), missing,
                                C.byref(level), C.byref(w), C.byref(h), C.byref(ch))
    if r:
        return None
    g = geometry(w.value, h.value)
    total = g["pixels"][level.value + 1]
    arr = np.ctypeslib.as_array(planar, shape=(ch.value * total,)).copy().reshape(ch.value, total)
    lib().orc_free(planar)
    return dict(coeffs=arr, missing=list(missing), level=level.value, w=w.value, h=h.value, ch=ch.value)


# ---------------------------------------------------------------- the unmodified reference programs

def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "encode")) and os.path.exists(os.path.join(REF_DIR, "decode"))


def pnm_bytes(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    kind = 5 if img.ndim == 2 else 6
    return b"P%d %d %d 255\n" % (kind, w, h) + img.tobytes()


def parse_pnm(data):
    """minimal P5/P6 parser for the writer's own header form 'P6 W H 255\\n'."""
    head, rest = data.split(b"\n", 1)
    parts = head.split()
    w, h = int(parts[1]), int(parts[2])
    if parts[0] == b"P6":
        return np.frombuffer(rest, dtype=np.uint8)[: w * h * 3].reshape(h, w, 3)
    return np.frombuffer(rest, dtype=np.uint8)[: w * h].reshape(h, w)


def ref_encode(img, capacity=None):
    """run oracle/_ref/encode on img; returns the .dwt bytes."""
    with tempfile.TemporaryDirectory() as d:
        pin, pout = os.path.join(d, "i.pnm"), os.path.join(d, "o.dwt")
        with open(pin, "wb") as f:
            f.write(pnm_bytes(img))
        cmd = [os.path.join(REF_DIR, "encode"), pin, pout]
        if capacity is not None:
            cmd.append(str(capacity))
        r = subprocess.run(cmd, capture_output=True)
        if r.returncode != 0:
            return None
        with open(pout, "rb") as f:
            return f.read()


def ref_decode(stream, pixels_max=None):
    """run oracle/_ref/decode; returns decoded image array or None when it exits 1."""
    with tempfile.TemporaryDirectory() as d:
        pin, pout = os.path.join(d, "i.dwt"), os.path.join(d, "o.pnm")
        with open(pin, "wb") as f:
            f.write(stream)
        cmd = [os.path.join(REF_DIR, "decode"), pin, pout]
        if pixels_max is not None:
            cmd.append(str(pixels_max))
        r = subprocess.run(cmd, capture_output=True)
        if r.returncode != 0 or not os.path.exists(pout):
            return None
        with open(pout, "rb") as f:
            return parse_pnm(f.read())
