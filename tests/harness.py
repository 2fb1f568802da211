"""ctypes harness shared by the tests, bench.py's CPU legs and __graft_entry__.smoke().

Three libraries are wrapped behind the same small Python interface:
  * Ref      -- the UNMODIFIED reference compiled into oracle/_ref/libicer_ref{,15}.so (oracle/Makefile)
  * Oracle   -- our plain-C restatement, oracle/liboracle.so
  * (the product library is wrapped in icer_compression_b200/, not here)

Nothing in this file reads /root/reference at run time.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

OK, OVERFLOW, BUF_SMALL, TOO_MANY_SEGMENTS, TOO_MANY_STAGES, QUOTA = 0, -1, -2, -3, -4, -5

FILTERS = dict(A=0, B=1, C=2, D=3, E=4, F=5, Q=6)


class _OutBuf(C.Structure):  # icer_output_data_buf_typedef, lib_icer/inc/icer.h:307-312
    _fields_ = [("size_used", C.c_size_t), ("size_allocated", C.c_size_t),
                ("data_start", C.POINTER(C.c_uint8)), ("rearrange_start", C.POINTER(C.c_uint8))]


def _u16p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class Ref:
    """The compiled reference.  planes15=True loads the 7-stage / 15-plane build (BASELINE config 4)."""

    def __init__(self, planes15=False):
        name = "libicer_ref15.so" if planes15 else "libicer_ref.so"
        path = os.path.join(ORACLE_DIR, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        self.lib = C.CDLL(path)
        self.planes = 15 if planes15 else 9
        self.max_stages = 7 if planes15 else 6
        self.lib.icer_init()
        L = self.lib
        L.icer_init_output_struct.argtypes = [C.POINTER(_OutBuf), C.POINTER(C.c_uint8), C.c_size_t, C.c_size_t]
        L.icer_compress_image_uint16.argtypes = [C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_uint8, C.c_int,
                                                 C.c_uint8, C.POINTER(_OutBuf)]
        L.icer_compress_image_yuv_uint16.argtypes = [C.POINTER(C.c_uint16)] * 3 + [C.c_size_t, C.c_size_t, C.c_uint8,
                                                                                    C.c_int, C.c_uint8, C.POINTER(_OutBuf)]
        L.icer_decompress_image_uint16.argtypes = [C.POINTER(C.c_uint16), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                                   C.c_size_t, C.POINTER(C.c_uint8), C.c_size_t, C.c_uint8, C.c_int,
                                                   C.c_uint8]
        L.icer_decompress_image_yuv_uint16.argtypes = [C.POINTER(C.c_uint16)] * 3 + [
            C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(C.c_uint8), C.c_size_t, C.c_uint8,
            C.c_int, C.c_uint8]
        L.icer_wavelet_transform_stages_uint16.argtypes = [C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_uint8, C.c_int]
        L.icer_inverse_wavelet_transform_stages_uint16.argtypes = L.icer_wavelet_transform_stages_uint16.argtypes
        L.icer_wavelet_transform_1d_uint16.argtypes = [C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_int]
        L.icer_inverse_wavelet_transform_1d_uint16.argtypes = L.icer_wavelet_transform_1d_uint16.argtypes
        L.icer_compute_bin.argtypes = [C.c_uint32, C.c_uint32]
        # the native uint8 codec (same shapes, uint8 samples)
        u8 = C.POINTER(C.c_uint8)
        L.icer_compress_image_uint8.argtypes = [u8, C.c_size_t, C.c_size_t, C.c_uint8, C.c_int, C.c_uint8, C.POINTER(_OutBuf)]
        L.icer_compress_image_yuv_uint8.argtypes = [u8] * 3 + [C.c_size_t, C.c_size_t, C.c_uint8, C.c_int, C.c_uint8, C.POINTER(_OutBuf)]
        L.icer_decompress_image_uint8.argtypes = [u8, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_size_t, u8, C.c_size_t,
                                                  C.c_uint8, C.c_int, C.c_uint8]
        L.icer_decompress_image_yuv_uint8.argtypes = [u8] * 3 + [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_size_t, u8,
                                                                 C.c_size_t, C.c_uint8, C.c_int, C.c_uint8]
        L.icer_wavelet_transform_stages_uint8.argtypes = [u8, C.c_size_t, C.c_size_t, C.c_uint8, C.c_int]
        L.icer_inverse_wavelet_transform_stages_uint8.argtypes = L.icer_wavelet_transform_stages_uint8.argtypes
        L.icer_interleave_uint8.argtypes = [u8, C.c_size_t, C.c_size_t]
        L.icer_interleave_uint8.restype = None

    # ---- whole image -------------------------------------------------------------------------------
    def compress(self, planes_in, stages, filt, segments, quota, dtype=np.uint16):
        """planes_in: list of 1 or 3 arrays [h, w] (dtype uint16, or uint8 for the native uint8 codec).
        Returns dict(rc, size_used, stream, prio, coefs)."""
        imgs = [np.ascontiguousarray(p, dtype=dtype).copy() for p in planes_in]
        h, w = imgs[0].shape
        buf = np.zeros(2 * quota + 64, dtype=np.uint8)
        ob = _OutBuf()
        rc0 = self.lib.icer_init_output_struct(C.byref(ob), _u8p(buf), 2 * quota, quota)
        assert rc0 == 0
        if dtype == np.uint8:
            if len(imgs) == 1:
                rc = self.lib.icer_compress_image_uint8(_u8p(imgs[0]), w, h, stages, filt, segments, C.byref(ob))
            else:
                rc = self.lib.icer_compress_image_yuv_uint8(_u8p(imgs[0]), _u8p(imgs[1]), _u8p(imgs[2]), w, h, stages, filt,
                                                            segments, C.byref(ob))
        elif len(imgs) == 1:
            rc = self.lib.icer_compress_image_uint16(_u16p(imgs[0]), w, h, stages, filt, segments, C.byref(ob))
        else:
            rc = self.lib.icer_compress_image_yuv_uint16(_u16p(imgs[0]), _u16p(imgs[1]), _u16p(imgs[2]), w, h, stages,
                                                         filt, segments, C.byref(ob))
        n = ob.size_used
        return dict(rc=rc, size_used=n, stream=bytes(buf[quota:quota + n]), prio=bytes(buf[:n]), coefs=imgs)

    def decompress(self, stream, chan, bufsize_px, stages, filt, segments, dtype=np.uint16):
        # Zero slack behind the stream.  lib_icer's scanner reads header fields past the end, and its bit reader never
        # stops at a segment's bit count (decoded_bits_total is never advanced, icer_decoding.c:66,88): when encoder and
        # decoder contexts disagree (coefficients with bits above the top coded plane, e.g. 12-bit data at 9 planes) the
        # decoder keeps pulling bits far behind the buffer and its output depends on that memory.  Oracle and product
        # define those bytes as zero, so the reference gets zeros there too.
        data = np.frombuffer(bytes(stream) + b"\0" * (4096 + 4 * max(bufsize_px, 1) * chan), dtype=np.uint8).copy()
        imgs = [np.zeros(max(bufsize_px, 1), dtype=dtype) for _ in range(chan)]
        w = C.c_size_t(0)
        h = C.c_size_t(0)
        if dtype == np.uint8:
            if chan == 1:
                rc = self.lib.icer_decompress_image_uint8(_u8p(imgs[0]), C.byref(w), C.byref(h), bufsize_px, _u8p(data),
                                                          len(stream), stages, filt, segments)
            else:
                rc = self.lib.icer_decompress_image_yuv_uint8(_u8p(imgs[0]), _u8p(imgs[1]), _u8p(imgs[2]), C.byref(w),
                                                              C.byref(h), bufsize_px, _u8p(data), len(stream), stages, filt,
                                                              segments)
        elif chan == 1:
            rc = self.lib.icer_decompress_image_uint16(_u16p(imgs[0]), C.byref(w), C.byref(h), bufsize_px, _u8p(data),
                                                       len(stream), stages, filt, segments)
        else:
            rc = self.lib.icer_decompress_image_yuv_uint16(_u16p(imgs[0]), _u16p(imgs[1]), _u16p(imgs[2]), C.byref(w),
                                                           C.byref(h), bufsize_px, _u8p(data), len(stream), stages, filt,
                                                           segments)
        W, H = w.value, h.value
        out = [im[:W * H].reshape(H, W) for im in imgs] if rc == 0 and W * H <= bufsize_px else None
        return dict(rc=rc, w=W, h=H, images=out)

    # ---- sub-APIs ----------------------------------------------------------------------------------
    def dwt_forward(self, img, stages, filt):
        a = np.ascontiguousarray(img, dtype=np.uint16).copy()
        rc = self.lib.icer_wavelet_transform_stages_uint16(_u16p(a), a.shape[1], a.shape[0], stages, filt)
        return rc, a

    def dwt_inverse(self, img, stages, filt):
        a = np.ascontiguousarray(img, dtype=np.uint16).copy()
        rc = self.lib.icer_inverse_wavelet_transform_stages_uint16(_u16p(a), a.shape[1], a.shape[0], stages, filt)
        return rc, a

    def dwt_1d(self, line, filt, inverse=False):
        a = np.ascontiguousarray(line, dtype=np.uint16).copy()
        f = self.lib.icer_inverse_wavelet_transform_1d_uint16 if inverse else self.lib.icer_wavelet_transform_1d_uint16
        rc = f(_u16p(a), a.size, 1, filt)
        return rc, a

    def dwt8(self, img, stages, filt, inverse=False):
        """The native uint8 transform (icer_wavelet_transform_stages_uint8 / inverse)."""
        a = np.ascontiguousarray(img, dtype=np.uint8).copy()
        f = self.lib.icer_inverse_wavelet_transform_stages_uint8 if inverse else self.lib.icer_wavelet_transform_stages_uint8
        rc = f(_u8p(a), a.shape[1], a.shape[0], stages, filt)
        return rc, a

    def interleave8_map(self, n):
        """dst[i] = where icer_interleave_uint8 puts element i of a length-n array (two passes: low / high index byte)."""
        idx = np.arange(n)
        lo = (idx & 0xFF).astype(np.uint8)
        hi = (idx >> 8).astype(np.uint8)
        self.lib.icer_interleave_uint8(_u8p(lo), n, 1)
        self.lib.icer_interleave_uint8(_u8p(hi), n, 1)
        at = lo.astype(np.int64) | (hi.astype(np.int64) << 8)      # at[pos] = original index now at pos
        dst = np.zeros(n, dtype=np.int64)
        dst[at] = idx
        return dst

    def bin_of(self, zero, total):
        return self.lib.icer_compute_bin(zero, total)


class _Limits(C.Structure):
    _fields_ = [("planes", C.c_int), ("max_stages", C.c_int), ("max_packets", C.c_int)]


class _Rect(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int)]


class Oracle:
    """Our plain-C restatement (oracle/liboracle.so)."""

    def __init__(self, planes=9, max_stages=6, max_packets=800):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle liboracle.so`)")
        self.lib = C.CDLL(path)
        self.lim = _Limits(planes, max_stages, max_packets)
        self.planes = planes
        L = self.lib
        L.orc_init()
        P3 = C.POINTER(C.c_uint16) * 3
        self._P3 = P3
        L.orc_compress_u16.argtypes = [P3, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(_Limits), C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_decompress_u16.argtypes = [P3, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_size_t,
                                         C.POINTER(C.c_uint8), C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(_Limits)]
        L.orc_dwt_forward_u16.argtypes = [C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_int, C.c_int]
        L.orc_dwt_inverse_u16.argtypes = L.orc_dwt_forward_u16.argtypes
        L.orc_dwt_forward_1d_u16.argtypes = [C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_int]
        L.orc_dwt_inverse_1d_u16.argtypes = L.orc_dwt_forward_1d_u16.argtypes
        L.orc_partition_rects.argtypes = [C.c_size_t, C.c_size_t, C.c_int, C.POINTER(_Rect)]
        L.orc_bin_of.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_crc32.argtypes = [C.POINTER(C.c_uint8), C.c_size_t]
        L.orc_crc32.restype = C.c_uint32
        L.orc_synth_frame.argtypes = [C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_synth_frame.restype = None
        L.orc_encode_segment_plane.argtypes = [C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                               C.c_int, C.POINTER(C.c_uint8), C.c_size_t]
        L.orc_encode_segment_plane.restype = C.c_long
        # the native uint8 codec: limits = the reference's ICER_BITPLANES_TO_COMPRESS_8 / ICER_MAX_PACKETS (icer.h:35-43)
        P3b = C.POINTER(C.c_uint8) * 3
        self._P3b = P3b
        self.lim8 = _Limits(7, max_stages, 300)
        L.orc_compress_u8.argtypes = [P3b, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(_Limits), C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_decompress_u8.argtypes = [P3b, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_size_t,
                                        C.POINTER(C.c_uint8), C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(_Limits)]
        L.orc_dwt_forward_u8.argtypes = [C.POINTER(C.c_uint8), C.c_size_t, C.c_size_t, C.c_int, C.c_int]
        L.orc_dwt_inverse_u8.argtypes = L.orc_dwt_forward_u8.argtypes
        L.orc_interleave8_map.argtypes = [C.c_size_t, C.POINTER(C.c_uint32)]
        L.orc_interleave8_map.restype = None

    def compress(self, planes_in, stages, filt, segments, quota, dtype=np.uint16):
        imgs = [np.ascontiguousarray(p, dtype=dtype).copy() for p in planes_in]
        h, w = imgs[0].shape
        buf = np.zeros(2 * quota + 64, dtype=np.uint8)
        used = C.c_size_t(0)
        if dtype == np.uint8:
            ptrs = self._P3b(*[_u8p(imgs[min(i, len(imgs) - 1)]) for i in range(3)])
            rc = self.lib.orc_compress_u8(ptrs, len(imgs), w, h, stages, filt, segments, C.byref(self.lim8), _u8p(buf),
                                          quota, C.byref(used))
            n = used.value
            return dict(rc=rc, size_used=n, stream=bytes(buf[quota:quota + n]), prio=bytes(buf[:n]), coefs=imgs)
        ptrs = self._P3(*[_u16p(imgs[min(i, len(imgs) - 1)]) for i in range(3)])
        rc = self.lib.orc_compress_u16(ptrs, len(imgs), w, h, stages, filt, segments, C.byref(self.lim), _u8p(buf),
                                       quota, C.byref(used))
        n = used.value
        return dict(rc=rc, size_used=n, stream=bytes(buf[quota:quota + n]), prio=bytes(buf[:n]), coefs=imgs)

    def decompress(self, stream, chan, bufsize_px, stages, filt, segments, dtype=np.uint16):
        data = np.frombuffer(bytes(stream) + b"\0" * 64, dtype=np.uint8).copy()
        imgs = [np.zeros(max(bufsize_px, 1), dtype=dtype) for _ in range(chan)]
        w = C.c_size_t(0)
        h = C.c_size_t(0)
        if dtype == np.uint8:
            ptrs = self._P3b(*[_u8p(imgs[min(i, chan - 1)]) for i in range(3)])
            rc = self.lib.orc_decompress_u8(ptrs, chan, C.byref(w), C.byref(h), bufsize_px, _u8p(data), len(stream),
                                            stages, filt, segments, C.byref(self.lim8))
            W, H = w.value, h.value
            out = [im[:W * H].reshape(H, W) for im in imgs] if rc == 0 and W * H <= bufsize_px else None
            return dict(rc=rc, w=W, h=H, images=out)
        ptrs = self._P3(*[_u16p(imgs[min(i, chan - 1)]) for i in range(3)])
        rc = self.lib.orc_decompress_u16(ptrs, chan, C.byref(w), C.byref(h), bufsize_px, _u8p(data), len(stream),
                                         stages, filt, segments, C.byref(self.lim))
        W, H = w.value, h.value
        out = [im[:W * H].reshape(H, W) for im in imgs] if rc == 0 and W * H <= bufsize_px else None
        return dict(rc=rc, w=W, h=H, images=out)

    def dwt_forward(self, img, stages, filt):
        a = np.ascontiguousarray(img, dtype=np.uint16).copy()
        rc = self.lib.orc_dwt_forward_u16(_u16p(a), a.shape[1], a.shape[0], stages, filt)
        return rc, a

    def dwt_inverse(self, img, stages, filt):
        a = np.ascontiguousarray(img, dtype=np.uint16).copy()
        rc = self.lib.orc_dwt_inverse_u16(_u16p(a), a.shape[1], a.shape[0], stages, filt)
        return rc, a

    def dwt_1d(self, line, filt, inverse=False):
        a = np.ascontiguousarray(line, dtype=np.uint16).copy()
        f = self.lib.orc_dwt_inverse_1d_u16 if inverse else self.lib.orc_dwt_forward_1d_u16
        rc = f(_u16p(a), a.size, 1, filt)
        return rc, a

    def dwt8(self, img, stages, filt, inverse=False):
        a = np.ascontiguousarray(img, dtype=np.uint8).copy()
        f = self.lib.orc_dwt_inverse_u8 if inverse else self.lib.orc_dwt_forward_u8
        rc = f(_u8p(a), a.shape[1], a.shape[0], stages, filt)
        return rc, a

    def interleave8_map(self, n):
        d = np.zeros(n, dtype=np.uint32)
        self.lib.orc_interleave8_map(n, d.ctypes.data_as(C.POINTER(C.c_uint32)))
        return d.astype(np.int64)

    def partition(self, w, h, segments):
        r = (_Rect * 40)()
        n = self.lib.orc_partition_rects(w, h, segments, r)
        return n, [(r[i].x, r[i].y, r[i].w, r[i].h) for i in range(max(n, 0))]

    def bin_of(self, zero, total):
        return self.lib.orc_bin_of(zero, total)

    def crc32(self, data):
        a = np.frombuffer(bytes(data) + b"\0", dtype=np.uint8).copy()
        return self.lib.orc_crc32(_u8p(a), len(data))

    def synth(self, w, h, frame, seed, kind):
        a = np.empty((h, w), dtype=np.uint16)
        self.lib.orc_synth_frame(_u16p(a), w, h, frame, seed, kind)
        return a


def have_ref():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libicer_ref.so"))


def load_boat():
    """512x512 8-bit grayscale test image (tests/golden/boat512.u8.gz, from the reference's boat.512.bmp)."""
    import gzip
    with gzip.open(os.path.join(ROOT, "tests", "golden", "boat512.u8.gz"), "rb") as f:
        return np.frombuffer(f.read(), dtype=np.uint8).reshape(512, 512).astype(np.uint16)


def load_boatcolor_rgb():
    """512x512x3 RGB pixels of the reference's boatcolor.512.bmp (tests/golden/boatcolor512.rgb.u8.gz)."""
    import gzip
    with gzip.open(os.path.join(ROOT, "tests", "golden", "boatcolor512.rgb.u8.gz"), "rb") as f:
        return np.frombuffer(f.read(), dtype=np.uint8).reshape(512, 512, 3)


def rgb_to_yuv(rgb):
    """The reference tools' integer RGB -> YCbCr (example/inc/color_util.h:27-29, example/src/icer_util.c:69-92)."""
    x = rgb.astype(np.int64)
    r, g, b = x[..., 0], x[..., 1], x[..., 2]
    clip = lambda v: np.clip(v, 0, 255)
    y = clip((19595 * r + 38470 * g + 7471 * b) >> 16)
    cb = clip(((36962 * (b - y)) >> 16) + 128)
    cr = clip(((46727 * (r - y)) >> 16) + 128)
    return [y.astype(np.uint16), cb.astype(np.uint16), cr.astype(np.uint16)]


class ColorRef:
    """The reference tools' colour macros, compiled (oracle/_ref/libcolor_ref.so = oracle/color_ref.c + the reference's own
    example/inc/color_util.h)."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "_ref", "libcolor_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        u8, u16 = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16)
        self.lib.colorref_rgb888_to_yuv.argtypes = [u8, u16, u16, u16, C.c_size_t]
        self.lib.colorref_rgb888_to_yuv.restype = None
        self.lib.colorref_yuv_to_rgb888.argtypes = [u16, u16, u16, u8, C.c_size_t]
        self.lib.colorref_yuv_to_rgb888.restype = None

    def rgb_to_yuv(self, rgb):
        """rgb: uint8 [..., 3] -> three uint16 arrays of the leading shape."""
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        n = rgb.size // 3
        out = [np.zeros(n, dtype=np.uint16) for _ in range(3)]
        self.lib.colorref_rgb888_to_yuv(_u8p(rgb), _u16p(out[0]), _u16p(out[1]), _u16p(out[2]), n)
        return [o.reshape(rgb.shape[:-1]) for o in out]

    def yuv_to_rgb(self, y, u, v):
        y, u, v = [np.ascontiguousarray(a, dtype=np.uint16) for a in (y, u, v)]
        rgb = np.zeros(y.shape + (3,), dtype=np.uint8)
        self.lib.colorref_yuv_to_rgb888(_u16p(y), _u16p(u), _u16p(v), _u8p(rgb), y.size)
        return rgb


def have_color_ref():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libcolor_ref.so"))


def load_boatcolor_yuv():
    """512x512 Y,U,V planes (uint16) of the reference's boatcolor.512.bmp after the tools' RGB->YCbCr."""
    return rgb_to_yuv(load_boatcolor_rgb())


def write_bmp24(path, rgb):
    """24-bit bottom-up BMP (what stb_image reads back verbatim)."""
    h, w, _ = rgb.shape
    stride = (w * 3 + 3) & ~3
    rows = np.zeros((h, stride), dtype=np.uint8)
    rows[:, :w * 3] = rgb[::-1, :, ::-1].reshape(h, w * 3)
    hdr = b"BM" + (54 + stride * h).to_bytes(4, "little") + b"\0\0\0\0" + (54).to_bytes(4, "little")
    hdr += (40).to_bytes(4, "little") + w.to_bytes(4, "little") + h.to_bytes(4, "little") + (1).to_bytes(2, "little")
    hdr += (24).to_bytes(2, "little") + b"\0" * 4 + (stride * h).to_bytes(4, "little") + (2835).to_bytes(4, "little") * 2 + b"\0" * 8
    with open(path, "wb") as f:
        f.write(hdr + rows.tobytes())
