"""dwt_b200 -- ctypes binding of libdwt_b200.so (the C ABI in include/dwt_b200.h).

This module is the Python-side mirror used by the tests and by bench.py; the product is the shared
library and the `encode` / `decode` programs built from dwt_b200/host/.  There is no CPU fallback: if
the library is missing, or no CUDA device is usable, every call raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DWT_B200_LIB") or os.path.join(HERE, "libdwt_b200.so")  # the override is an A/B aid for builds with other flags

# every symbol include/dwt_b200.h declares (tests/test_abi.py checks the library exports all of them)
ABI_SYMBOLS = [
    "dwt_ctx_create", "dwt_ctx_destroy", "dwt_last_error", "dwt_encode", "dwt_decode", "dwt_free",
    "dwt_ctx_upload_image", "dwt_ctx_encode_resident", "dwt_ctx_download_stream", "dwt_ctx_upload_stream",
    "dwt_ctx_decode_resident", "dwt_ctx_download_image", "dwt_ctx_launch_count", "dwt_ctx_sync", "dwt_ctx_set_in_flight",
    "dwt_ctx_set_decoder_scan",
    "dwt_host_alloc", "dwt_host_free", "dwt_encode_into", "dwt_decode_into", "dwt_ctx_flush_l2",
    "dwt_ctx_event_record", "dwt_ctx_event_elapsed_ms", "dwt_ctx_wait_for",
    "dwt_pool_create", "dwt_pool_create_multi", "dwt_pool_devices", "dwt_pool_destroy", "dwt_pool_workers", "dwt_pool_last_error", "dwt_pool_encode", "dwt_pool_decode", "dwt_pool_run",
    "cdf53", "icdf53", "dwt_forward", "dwt_inverse", "dwt_ycocg_from_rgb", "dwt_rgb_from_ycocg",
    "compute_lengths", "ilog2", "dwt_debug_front_end",
    "bytes_reader", "bytes_writer", "bytes_count", "close_bytes_reader", "close_bytes_writer", "put_byte",
    "write_bytes", "get_byte", "read_bytes", "bytes_writer_mem", "bytes_writer_data", "bytes_reader_mem",
    "bits_reader", "bits_writer", "bits_count", "close_bits_reader", "close_bits_writer", "put_bit", "write_bits",
    "get_bit", "read_bits",
    "vli_reader", "vli_writer", "delete_vli_reader", "delete_vli_writer", "vli_put_bit", "vli_get_bit",
    "vli_write_bits", "vli_read_bits", "put_vli", "get_vli",
    "rle_reader", "rle_writer", "rle_flush", "delete_rle_reader", "delete_rle_writer", "put_rle", "get_rle",
    "rle_put_bit", "rle_get_bit",
]


class Stats(C.Structure):
    _fields_ = [("meta_bits", C.c_longlong), ("root_bits", C.c_longlong), ("total_bits", C.c_longlong),
                ("kib", C.c_longlong), ("full_bits", C.c_longlong), ("levels", C.c_int), ("planes", C.c_int * 3),
                ("ms_h2d", C.c_float), ("ms_lift", C.c_float), ("ms_linearize", C.c_float), ("ms_coder", C.c_float),
                ("ms_d2h", C.c_float), ("ms_total", C.c_float), ("level_reached", C.c_int),
                ("parse_windows", C.c_longlong), ("parse_jumps", C.c_longlong), ("parse_exact", C.c_longlong)]


class EncodeItem(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("capacity", C.c_int),
                ("out", C.c_void_p), ("out_room", C.c_size_t), ("out_len", C.c_size_t), ("status", C.c_int)]


class DecodeItem(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("len", C.c_size_t), ("pixels_max", C.c_int), ("pixels", C.c_void_p),
                ("pixels_room", C.c_size_t), ("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("status", C.c_int)]


class DwtError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libdwt_b200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DwtError("libdwt_b200.so is not built: run `make` (or __graft_entry__.build()) first")
    L = C.CDLL(LIB_PATH)
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    u8p = C.POINTER(C.c_uint8)
    L.dwt_ctx_create.argtypes = [C.c_int]
    L.dwt_ctx_create.restype = vp
    L.dwt_ctx_destroy.argtypes = [vp]
    L.dwt_last_error.restype = C.c_char_p
    L.dwt_encode.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(u8p), C.POINTER(C.c_size_t),
                             C.POINTER(Stats)]
    L.dwt_decode.argtypes = [vp, u8p, C.c_size_t, C.c_int, C.POINTER(u8p), ip, ip, ip, C.POINTER(Stats)]
    L.dwt_free.argtypes = [vp]
    L.dwt_ctx_upload_image.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int]
    L.dwt_ctx_encode_resident.argtypes = [vp, C.c_int, C.POINTER(Stats)]
    L.dwt_ctx_download_stream.argtypes = [vp, C.POINTER(u8p), C.POINTER(C.c_size_t)]
    L.dwt_ctx_upload_stream.argtypes = [vp, u8p, C.c_size_t]
    L.dwt_ctx_decode_resident.argtypes = [vp, C.c_int, C.POINTER(Stats)]
    L.dwt_ctx_download_image.argtypes = [vp, C.POINTER(u8p), ip, ip, ip]
    L.dwt_ctx_launch_count.argtypes = [vp]
    L.dwt_ctx_launch_count.restype = C.c_longlong
    L.dwt_ctx_sync.argtypes = [vp]
    L.dwt_host_alloc.argtypes = [C.c_size_t]
    L.dwt_host_alloc.restype = vp
    L.dwt_host_free.argtypes = [vp]
    L.dwt_host_free.restype = None
    L.dwt_encode_into.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_size_t, C.POINTER(C.c_size_t),
                                  C.POINTER(Stats)]
    L.dwt_decode_into.argtypes = [vp, u8p, C.c_size_t, C.c_int, u8p, C.c_size_t, ip, ip, ip, C.POINTER(Stats)]
    L.dwt_ctx_flush_l2.argtypes = [vp]
    L.dwt_ctx_event_record.argtypes = [vp, C.c_int]
    L.dwt_ctx_event_elapsed_ms.argtypes = [vp, C.c_int, C.c_int]
    L.dwt_ctx_event_elapsed_ms.restype = C.c_float
    L.dwt_ctx_wait_for.argtypes = [vp, vp]
    L.dwt_ctx_set_in_flight.argtypes = [vp, C.c_int]
    L.dwt_ctx_set_decoder_scan.argtypes = [vp, C.c_int]
    L.dwt_pool_create.argtypes = [C.c_int, C.c_int]
    L.dwt_pool_create.restype = vp
    L.dwt_pool_create_multi.argtypes = [ip, C.c_int, C.c_int]
    L.dwt_pool_create_multi.restype = vp
    L.dwt_pool_devices.argtypes = [vp, ip, C.c_int]
    L.dwt_pool_destroy.argtypes = [vp]
    L.dwt_pool_destroy.restype = None
    L.dwt_pool_workers.argtypes = [vp]
    L.dwt_pool_last_error.argtypes = [vp]
    L.dwt_pool_last_error.restype = C.c_char_p
    L.dwt_pool_encode.argtypes = [vp, C.POINTER(EncodeItem), C.c_int]
    L.dwt_pool_decode.argtypes = [vp, C.POINTER(DecodeItem), C.c_int]
    L.dwt_pool_run.argtypes = [vp, C.POINTER(EncodeItem), C.c_int, C.POINTER(DecodeItem), C.c_int]
    L.cdf53.argtypes = [ip, ip, C.c_int, C.c_int, C.c_int, C.c_int]
    L.cdf53.restype = None
    L.icdf53.argtypes = [ip, ip, C.c_int, C.c_int, C.c_int, C.c_int]
    L.icdf53.restype = None
    L.dwt_forward.argtypes = [ip, ip, C.c_int, C.c_int, C.c_int]
    L.dwt_inverse.argtypes = [ip, ip, C.c_int, C.c_int, C.c_int]
    L.dwt_ycocg_from_rgb.argtypes = [ip, C.c_int]
    L.dwt_rgb_from_ycocg.argtypes = [ip, C.c_int]
    L.compute_lengths.argtypes = [ip, ip, ip, ip, C.c_int, C.c_int, C.c_int]
    L.dwt_debug_front_end.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, ip, ip, ip]
    _lib = L
    return L


def last_error():
    return (lib().dwt_last_error() or b"").decode()


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _shape(img):
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    return w, h, ch


class Codec:
    """One device context (dwt_ctx): a CUDA device, a stream and reusable buffers."""

    def __init__(self, device=-1):
        self._h = lib().dwt_ctx_create(device)
        if not self._h:
            raise DwtError("dwt_ctx_create failed: " + last_error())
        self.stats = Stats()

    def close(self):
        if self._h:
            lib().dwt_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- whole-call API (host buffers in, host buffers out)
    def encode(self, img, capacity=0):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        w, h, ch = _shape(img)
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        r = lib().dwt_encode(self._h, _u8p(img), w, h, ch, int(capacity), C.byref(out), C.byref(n), C.byref(self.stats))
        if r:
            raise DwtError("dwt_encode failed: " + last_error())
        data = C.string_at(out, n.value)
        lib().dwt_free(out)
        return data

    def decode(self, stream, pixels_max=-1):
        """returns the decoded uint8 image, or None where the reference program exits 1 without output"""
        buf = np.frombuffer(bytes(stream), dtype=np.uint8)
        if buf.size == 0:
            buf = np.zeros(1, dtype=np.uint8)
            n = 0
        else:
            n = buf.size
        pix, w, h, ch = C.POINTER(C.c_uint8)(), C.c_int(), C.c_int(), C.c_int()
        r = lib().dwt_decode(self._h, _u8p(buf), n, int(pixels_max), C.byref(pix), C.byref(w), C.byref(h), C.byref(ch),
                             C.byref(self.stats))
        if r > 0:  # 1: stream ends inside the prefix, 2: bad magic / size -- the reference exits 1 without output
            self.reject_code = r
            return None
        if r:
            raise DwtError("dwt_decode failed: " + last_error())
        shape = (h.value, w.value, 3) if ch.value == 3 else (h.value, w.value)
        arr = np.frombuffer(C.string_at(pix, int(np.prod(shape))), dtype=np.uint8).reshape(shape).copy()
        lib().dwt_free(pix)
        return arr

    # ---- device-resident API (bench.py: inputs already in HBM)
    def upload_image(self, img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        w, h, ch = _shape(img)
        if lib().dwt_ctx_upload_image(self._h, _u8p(img), w, h, ch) or lib().dwt_ctx_sync(self._h):
            raise DwtError("upload_image failed: " + last_error())

    def encode_resident(self, capacity=0):
        if lib().dwt_ctx_encode_resident(self._h, int(capacity), C.byref(self.stats)):
            raise DwtError("encode_resident failed: " + last_error())
        return self.stats

    def download_stream(self):
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        if lib().dwt_ctx_download_stream(self._h, C.byref(out), C.byref(n)):
            raise DwtError("download_stream failed: " + last_error())
        data = C.string_at(out, n.value)
        lib().dwt_free(out)
        return data

    def upload_stream(self, stream):
        buf = np.frombuffer(bytes(stream), dtype=np.uint8)
        if lib().dwt_ctx_upload_stream(self._h, _u8p(buf), buf.size) or lib().dwt_ctx_sync(self._h):
            raise DwtError("upload_stream failed: " + last_error())

    def decode_resident(self, pixels_max=-1):
        r = lib().dwt_ctx_decode_resident(self._h, int(pixels_max), C.byref(self.stats))
        if r < 0:
            raise DwtError("decode_resident failed: " + last_error())
        return r

    def download_image(self):
        pix, w, h, ch = C.POINTER(C.c_uint8)(), C.c_int(), C.c_int(), C.c_int()
        if lib().dwt_ctx_download_image(self._h, C.byref(pix), C.byref(w), C.byref(h), C.byref(ch)):
            raise DwtError("download_image failed: " + last_error())
        shape = (h.value, w.value, 3) if ch.value == 3 else (h.value, w.value)
        arr = np.frombuffer(C.string_at(pix, int(np.prod(shape))), dtype=np.uint8).reshape(shape).copy()
        lib().dwt_free(pix)
        return arr

    def launch_count(self):
        return int(lib().dwt_ctx_launch_count(self._h))

    # ---- caller-owned (page-locked) buffers: the end-to-end path of bench.py
    def encode_into(self, img, out, capacity=0):
        """img, out: uint8 numpy arrays (ideally views of pinned_array()); returns the stream length"""
        w, h, ch = _shape(img)
        n = C.c_size_t()
        if lib().dwt_encode_into(self._h, _u8p(img), w, h, ch, int(capacity), _u8p(out), out.size, C.byref(n),
                                 C.byref(self.stats)):
            raise DwtError("dwt_encode_into failed: " + last_error())
        return n.value

    def decode_into(self, stream, nbytes, out, pixels_max=-1):
        """stream: uint8 array holding nbytes stream bytes; out: uint8 array for the pixels; returns (h, w, ch)"""
        w, h, ch = C.c_int(), C.c_int(), C.c_int()
        r = lib().dwt_decode_into(self._h, _u8p(stream), nbytes, int(pixels_max), _u8p(out), out.size, C.byref(w),
                                  C.byref(h), C.byref(ch), C.byref(self.stats))
        if r:
            raise DwtError("dwt_decode_into failed (%d): %s" % (r, last_error()))
        return h.value, w.value, ch.value

    def flush_l2(self):
        if lib().dwt_ctx_flush_l2(self._h):
            raise DwtError("flush_l2 failed: " + last_error())

    def event_record(self, slot):
        if lib().dwt_ctx_event_record(self._h, slot):
            raise DwtError("event_record failed: " + last_error())

    def event_elapsed_ms(self, a, b):
        return float(lib().dwt_ctx_event_elapsed_ms(self._h, a, b))

    def set_in_flight(self, contexts):
        """how many contexts the caller keeps busy on this device at once (throughput- vs latency-oriented kernels)"""
        if lib().dwt_ctx_set_in_flight(self._h, int(contexts)):
            raise DwtError("set_in_flight failed")

    def set_scan(self, mode):
        """pin the decoder's scan kernel: "auto", "parallel" or "serial" (dwt_ctx_set_decoder_scan)"""
        if lib().dwt_ctx_set_decoder_scan(self._h, {"auto": 0, "parallel": 1, "serial": 2}[mode]):
            raise DwtError("set_decoder_scan failed")

    def wait_for(self, other):
        """this context's stream waits for the work queued so far on `other`'s stream"""
        if lib().dwt_ctx_wait_for(self._h, other._h):
            raise DwtError("wait_for failed: " + last_error())

    # ---- parity taps
    def front_end(self, img):
        """(pyramid int32 (h,w,ch) in the reference's interleaved Mallat layout, planar (ch, w*h), planes)"""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        w, h, ch = _shape(img)
        pyr = np.zeros((h, w, ch), dtype=np.int32)
        lin = np.zeros((ch, w * h), dtype=np.int32)
        planes = (C.c_int * 3)()
        if lib().dwt_debug_front_end(self._h, _u8p(img), w, h, ch, _ip(pyr), _ip(lin), planes):
            raise DwtError("dwt_debug_front_end failed: " + last_error())
        return pyr, lin, list(planes)[:ch]


class Pool:
    """dwt_pool: `workers` contexts on one device coding the items of a batch concurrently (one host thread each)"""

    def __init__(self, device=0, workers=4):
        """device: one device index, a list of them, or "all" (dwt_pool_create_multi: item i -> devices[i mod G])"""
        if device == "all":
            self._h = lib().dwt_pool_create_multi(None, 0, int(workers))
        elif isinstance(device, (list, tuple)):
            arr = (C.c_int * len(device))(*[int(d) for d in device])
            self._h = lib().dwt_pool_create_multi(arr, len(device), int(workers))
        else:
            self._h = lib().dwt_pool_create(int(device), int(workers))
        if not self._h:
            raise DwtError("dwt_pool_create failed: " + last_error())

    def devices(self):
        arr = (C.c_int * 64)()
        n = lib().dwt_pool_devices(self._h, arr, 64)
        return list(arr)[:n]

    def close(self):
        if self._h:
            lib().dwt_pool_destroy(self._h)
            self._h = None

    def encode_items(self, items, n):
        """items: ctypes array of EncodeItem (buffers owned by the caller); returns the number of failed items"""
        return lib().dwt_pool_encode(self._h, items, n)

    def decode_items(self, items, n):
        return lib().dwt_pool_decode(self._h, items, n)

    def run_items(self, enc_items, n_enc, dec_items, n_dec):
        """encode and decode items in one call, interleaved on the workers"""
        return lib().dwt_pool_run(self._h, enc_items, n_enc, dec_items, n_dec)

    def encode_batch(self, images, capacity=0):
        """list of uint8 images -> list of .dwt byte strings"""
        n = len(images)
        items = (EncodeItem * n)()
        keep = []
        for i, img in enumerate(images):
            img = np.ascontiguousarray(img, dtype=np.uint8)
            w, h, ch = _shape(img)
            src, o1 = pinned_array(img.size)
            src[:] = img.reshape(-1)
            dst, o2 = pinned_array(img.size * 2 + 4096)
            keep.append((src, dst, o1, o2))
            items[i] = EncodeItem(src.ctypes.data, w, h, ch, int(capacity), dst.ctypes.data, dst.size, 0, 0)
        bad = self.encode_items(items, n)
        if bad:
            raise DwtError("dwt_pool_encode: %d items failed" % bad)
        return [bytes(keep[i][1][:items[i].out_len]) for i in range(n)]

    def decode_batch(self, streams, shapes, pixels_max=-1):
        """list of streams + upper bounds (h, w, ch) of their decoded sizes -> list of uint8 arrays"""
        n = len(streams)
        items = (DecodeItem * n)()
        keep = []
        for i, s in enumerate(streams):
            src, o1 = pinned_array(max(1, len(s)))
            src[:len(s)] = np.frombuffer(s, dtype=np.uint8)
            h, w, ch = shapes[i]
            dst, o2 = pinned_array(h * w * ch)
            keep.append((src, dst, o1, o2))
            items[i] = DecodeItem(src.ctypes.data, len(s), int(pixels_max), dst.ctypes.data, dst.size, 0, 0, 0, 0)
        bad = self.decode_items(items, n)
        if bad:
            raise DwtError("dwt_pool_decode: %d items failed" % bad)
        out = []
        for i in range(n):
            it = items[i]
            shape = (it.height, it.width, it.channels) if it.channels > 1 else (it.height, it.width)
            out.append(np.array(keep[i][1][:it.height * it.width * it.channels]).reshape(shape))
        return out


def pinned_array(nbytes):
    """uint8 numpy view of page-locked host memory (dwt_host_alloc); keep the returned owner alive"""
    p = lib().dwt_host_alloc(nbytes)
    if not p:
        raise DwtError("dwt_host_alloc failed: " + last_error())
    arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib().dwt_host_free(self.ptr)
            except Exception:
                pass
    return arr, _Owner(p)


# ---- transform entry points with the reference's argument meaning (host buffers)

def cdf53(x, N, SO, SI, CH, out_len=None):
    """reference cdf53(out, in, N, SO, SI, CH): returns (out, in_after) -- `in` is clobbered like the reference"""
    x = np.ascontiguousarray(x, dtype=np.int32).copy()
    out = np.zeros(out_len if out_len else x.size, dtype=np.int32)
    lib().cdf53(_ip(out), _ip(x), N, SO, SI, CH)
    return out, x


def icdf53(x, N, SO, SI, CH, out_len=None):
    x = np.ascontiguousarray(x, dtype=np.int32).copy()
    out = np.zeros(out_len if out_len else x.size, dtype=np.int32)
    lib().icdf53(_ip(out), _ip(x), N, SO, SI, CH)
    return out


def forward(img_int):
    """`transformation` of encode.c:16-30 on an interleaved int32 (h,w,ch) buffer -> Mallat pyramid"""
    a = np.ascontiguousarray(img_int, dtype=np.int32)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, ch = a.shape
    out = np.zeros_like(a)
    if lib().dwt_forward(_ip(out), _ip(a), w, h, ch):
        raise DwtError("dwt_forward failed: " + last_error())
    return out


def inverse(pyr):
    a = np.ascontiguousarray(pyr, dtype=np.int32)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, ch = a.shape
    out = np.zeros_like(a)
    if lib().dwt_inverse(_ip(out), _ip(a), w, h, ch):
        raise DwtError("dwt_inverse failed: " + last_error())
    return out


def ycocg_from_rgb(buf):
    a = np.ascontiguousarray(buf, dtype=np.int32).copy()
    if lib().dwt_ycocg_from_rgb(_ip(a), a.size // 3):
        raise DwtError(last_error())
    return a


def rgb_from_ycocg(buf):
    a = np.ascontiguousarray(buf, dtype=np.int32).copy()
    if lib().dwt_rgb_from_ycocg(_ip(a), a.size // 3):
        raise DwtError(last_error())
    return a


def geometry(w, h):
    arrs = [(C.c_int * 16)() for _ in range(4)]
    levels = lib().compute_lengths(arrs[0], arrs[1], arrs[2], arrs[3], w, h, 8)
    lengths, pixels, widths, heights = [list(a)[:levels + 1] for a in arrs]
    return dict(levels=levels, lengths=lengths, pixels=pixels, widths=widths, heights=heights)
