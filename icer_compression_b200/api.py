"""ctypes mirror of include/icer.h and include/icer_b200.h (same names, argument meaning and status codes)."""
import ctypes as C
import os

import numpy as np

ICER_RESULT_OK = 0
ICER_INTEGER_OVERFLOW = -1
ICER_OUTPUT_BUF_TOO_SMALL = -2
ICER_TOO_MANY_SEGMENTS = -3
ICER_TOO_MANY_STAGES = -4
ICER_BYTE_QUOTA_EXCEEDED = -5
ICER_FATAL_ERROR = -10
FILTERS = dict(A=0, B=1, C=2, D=3, E=4, F=5, Q=6)
MEM_HOST, MEM_DEVICE = 0, 1
PIX_U16, PIX_U8, PIX_RGB888, PIX_NATIVE_U8 = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))


class IcerError(RuntimeError):
    pass


def lib_path():
    # ICER_B200_LIB: development hook to benchmark an alternative build of the same library (profiles/variants.sh)
    return os.environ.get("ICER_B200_LIB") or os.path.join(_HERE, "lib", "libicer_b200.so")


class _OutBuf(C.Structure):  # icer_output_data_buf_typedef (include/icer.h)
    _fields_ = [("size_used", C.c_size_t), ("size_allocated", C.c_size_t),
                ("data_start", C.POINTER(C.c_uint8)), ("rearrange_start", C.POINTER(C.c_uint8))]


class _Params(C.Structure):  # icer_b200_params (include/icer_b200.h)
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("channels", C.c_uint32), ("stages", C.c_uint32),
                ("filter", C.c_uint32), ("segments", C.c_uint32), ("byte_quota", C.c_uint64)]


_LIB = None


def load_library():
    """Load libicer_b200.so (built by __graft_entry__.build() / csrc/Makefile).  Fails loudly if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise IcerError("%s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                        "there is no CPU fallback" % p)
    L = C.CDLL(p)
    vp, u8p, u16p, szp = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_size_t)
    L.icer_init.restype = C.c_int
    L.icer_init_output_struct.argtypes = [C.POINTER(_OutBuf), u8p, C.c_size_t, C.c_size_t]
    L.icer_get_image_dimensions.argtypes = [u8p, C.c_size_t, szp, szp]
    L.icer_compress_image_uint16.argtypes = [u16p, C.c_size_t, C.c_size_t, C.c_uint8, C.c_int, C.c_uint8, C.POINTER(_OutBuf)]
    L.icer_compress_image_yuv_uint16.argtypes = [u16p] * 3 + [C.c_size_t, C.c_size_t, C.c_uint8, C.c_int, C.c_uint8, C.POINTER(_OutBuf)]
    L.icer_decompress_image_uint16.argtypes = [u16p, szp, szp, C.c_size_t, u8p, C.c_size_t, C.c_uint8, C.c_int, C.c_uint8]
    L.icer_decompress_image_yuv_uint16.argtypes = [u16p] * 3 + [szp, szp, C.c_size_t, u8p, C.c_size_t, C.c_uint8, C.c_int, C.c_uint8]
    L.icer_wavelet_transform_stages_uint16.argtypes = [u16p, C.c_size_t, C.c_size_t, C.c_uint8, C.c_int]
    L.icer_inverse_wavelet_transform_stages_uint16.argtypes = [u16p, C.c_size_t, C.c_size_t, C.c_uint8, C.c_int]
    L.icer_compress_image_uint8.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_uint8, C.c_int, C.c_uint8, C.POINTER(_OutBuf)]
    L.icer_compress_image_yuv_uint8.argtypes = [u8p] * 3 + [C.c_size_t, C.c_size_t, C.c_uint8, C.c_int, C.c_uint8, C.POINTER(_OutBuf)]
    L.icer_decompress_image_uint8.argtypes = [u8p, szp, szp, C.c_size_t, u8p, C.c_size_t, C.c_uint8, C.c_int, C.c_uint8]
    L.icer_decompress_image_yuv_uint8.argtypes = [u8p] * 3 + [szp, szp, C.c_size_t, u8p, C.c_size_t, C.c_uint8, C.c_int, C.c_uint8]
    L.icer_wavelet_transform_stages_uint8.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_uint8, C.c_int]
    L.icer_inverse_wavelet_transform_stages_uint8.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_uint8, C.c_int]
    L.icer_b200_set_limits_u8.argtypes = [vp, C.c_int, C.c_int]
    L.icer_b200_dwt_batch_u8.argtypes = [vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, vp, C.c_int, vp]
    L.icer_b200_u8_interleave_map.argtypes = [C.c_uint32, vp]
    L.icer_b200_create.argtypes = [C.POINTER(vp), C.c_int]
    L.icer_b200_destroy.argtypes = [vp]
    L.icer_b200_destroy.restype = None
    L.icer_b200_set_limits.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.icer_b200_set_scratch_bits.argtypes = [vp, C.c_int]
    L.icer_b200_compress_batch_u16.argtypes = [vp, C.POINTER(_Params), vp, C.c_size_t, vp, C.c_size_t, vp, vp, vp, vp, C.c_int, vp]
    L.icer_b200_decompress_batch_u16.argtypes = [vp, C.POINTER(_Params), vp, C.c_size_t, vp, C.c_size_t, vp, vp, vp, C.c_int, vp]
    L.icer_b200_compress_batch_pix.argtypes = [vp, C.POINTER(_Params), vp, C.c_int, C.c_size_t, vp, C.c_size_t, vp, vp, C.c_int, vp]
    L.icer_b200_decompress_batch_pix.argtypes = [vp, C.POINTER(_Params), vp, C.c_size_t, vp, C.c_size_t, vp, C.c_int, vp, vp, C.c_int, vp]
    L.icer_b200_dwt_batch_u16.argtypes = [vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, vp, C.c_int, vp]
    L.icer_b200_compress_batch_multi.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(_Params), vp, C.c_int, C.c_size_t, vp, C.c_size_t, vp, vp, C.c_int]
    L.icer_b200_decompress_batch_multi.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(_Params), vp, C.c_size_t, vp, C.c_size_t, vp, C.c_int, vp, vp, C.c_int]
    L.icer_b200_device_count.restype = C.c_int
    L.icer_b200_launch_count.argtypes = [vp]
    L.icer_b200_launch_count.restype = C.c_uint64
    L.icer_b200_last_phase_ms.argtypes = [vp, C.c_int]
    L.icer_b200_last_phase_ms.restype = C.c_float
    L.icer_b200_enable_timing.argtypes = [vp, C.c_int]
    L.icer_b200_enable_timing.restype = None
    L.icer_b200_version.restype = C.c_char_p
    L.icer_b200_synth_frames_u16.argtypes = [vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, vp]
    _LIB = L
    return L


def _ptr(x):
    """Raw address of a numpy array or a torch tensor (host or device)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()


class IcerB200:
    """One codec context on one CUDA device (include/icer_b200.h)."""

    def __init__(self, device=-1, planes=9, max_stages=6, max_packets=800):
        self.h = None
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.icer_b200_create(C.byref(h), device)
        if rc != 0:
            raise IcerError("icer_b200_create failed (%d): no usable CUDA device; there is no CPU fallback" % rc)
        self.h = h
        rc = self.lib.icer_b200_set_limits(h, planes, max_stages, max_packets)
        if rc != 0:
            raise IcerError("icer_b200_set_limits(%d, %d, %d) -> %d" % (planes, max_stages, max_packets, rc))
        self.planes = planes

    def close(self):
        if self.h:
            self.lib.icer_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_scratch_bits(self, bits):
        return self.lib.icer_b200_set_scratch_bits(self.h, bits)

    def enable_timing(self, on=True):
        self.lib.icer_b200_enable_timing(self.h, 1 if on else 0)

    def phase_ms(self, phase):
        return float(self.lib.icer_b200_last_phase_ms(self.h, phase))

    def launch_count(self):
        return int(self.lib.icer_b200_launch_count(self.h))

    # ---- batch, raw pointers (numpy arrays or torch tensors; mem = MEM_HOST / MEM_DEVICE) ---------------
    def compress_batch_raw(self, params, frames, n, out, out_stride, out_prio, coefs, size_used, rc, mem, stream=None):
        p = _Params(*params)
        e = self.lib.icer_b200_compress_batch_u16(self.h, C.byref(p), _ptr(frames), n, _ptr(out), out_stride, _ptr(out_prio),
                                                  _ptr(coefs), _ptr(size_used), _ptr(rc), mem, stream)
        if e != 0:
            raise IcerError("icer_b200_compress_batch_u16 -> %d" % e)

    def decompress_batch_raw(self, params, streams, stream_stride, stream_len, n, frames, dims, rc, mem, stream=None):
        p = _Params(*params)
        e = self.lib.icer_b200_decompress_batch_u16(self.h, C.byref(p), _ptr(streams), stream_stride, _ptr(stream_len), n,
                                                    _ptr(frames), _ptr(dims), _ptr(rc), mem, stream)
        if e != 0:
            raise IcerError("icer_b200_decompress_batch_u16 -> %d" % e)

    # ---- batch, numpy convenience ---------------------------------------------------------------------
    def compress_batch(self, frames, stages, filt, segments, quota, want_prio=False, want_coefs=False):
        """frames: uint16 [n, chan, h, w] (chan 1 or 3).  Returns dict(rc[n], size_used[n], streams[list of bytes], ...)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint16)
        n, chan, h, w = frames.shape
        out = np.zeros((n, quota), dtype=np.uint8)
        prio = np.zeros((n, quota), dtype=np.uint8) if want_prio else None
        coefs = np.zeros_like(frames) if want_coefs else None
        used = np.zeros(n, dtype=np.uint64)
        rc = np.zeros(n, dtype=np.int32)
        self.compress_batch_raw((w, h, chan, stages, filt, segments, quota), frames, n, out, quota, prio, coefs, used, rc, MEM_HOST)
        res = dict(rc=rc, size_used=used, streams=[bytes(out[i, :int(used[i])]) for i in range(n)])
        if want_prio:
            res["prio"] = [bytes(prio[i, :int(used[i])]) for i in range(n)]
        if want_coefs:
            res["coefs"] = coefs
        return res

    def decompress_batch(self, streams, chan, w, h, stages, filt, segments):
        """streams: list of bytes.  Returns dict(rc[n], dims[n,2], frames uint16 [n, chan, h*w])."""
        n = len(streams)
        stride = max(max((len(s) for s in streams), default=1), 1)
        buf = np.zeros((n, stride), dtype=np.uint8)
        lens = np.zeros(n, dtype=np.uint64)
        for i, s in enumerate(streams):
            buf[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
            lens[i] = len(s)
        frames = np.zeros((n, chan, h * w), dtype=np.uint16)
        dims = np.zeros((n, 2), dtype=np.uint32)
        rc = np.zeros(n, dtype=np.int32)
        self.decompress_batch_raw((w, h, chan, stages, filt, segments, 0), buf, stride, lens, n, frames, dims, rc, MEM_HOST)
        return dict(rc=rc, dims=dims, frames=frames)

    # ---- 8-bit / RGB frames at the boundary (widening and colour conversion on the device) -------------------------
    def compress_batch_pix(self, frames, pixfmt, w, h, chan, stages, filt, segments, quota):
        """frames: uint8 [n, chan, h, w] (PIX_U8) or uint8 [n, h, w, 3] (PIX_RGB888)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        n = frames.shape[0]
        out = np.zeros((n, quota), dtype=np.uint8)
        used = np.zeros(n, dtype=np.uint64)
        rc = np.zeros(n, dtype=np.int32)
        p = _Params(w, h, chan, stages, filt, segments, quota)
        e = self.lib.icer_b200_compress_batch_pix(self.h, C.byref(p), _ptr(frames), pixfmt, n, _ptr(out), quota, _ptr(used), _ptr(rc), MEM_HOST, None)
        if e != 0:
            raise IcerError("icer_b200_compress_batch_pix -> %d" % e)
        return dict(rc=rc, size_used=used, streams=[bytes(out[i, :int(used[i])]) for i in range(n)])

    def decompress_batch_pix(self, streams, pixfmt, chan, w, h, stages, filt, segments):
        n = len(streams)
        stride = max(max((len(s) for s in streams), default=1), 1)
        buf = np.zeros((n, stride), dtype=np.uint8)
        lens = np.zeros(n, dtype=np.uint64)
        for i, s in enumerate(streams):
            buf[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
            lens[i] = len(s)
        frames = np.zeros((n, h * w, 3) if pixfmt == PIX_RGB888 else (n, chan, h * w), dtype=np.uint8)
        dims = np.zeros((n, 2), dtype=np.uint32)
        rc = np.zeros(n, dtype=np.int32)
        p = _Params(w, h, chan, stages, filt, segments, 0)
        e = self.lib.icer_b200_decompress_batch_pix(self.h, C.byref(p), _ptr(buf), stride, _ptr(lens), n, _ptr(frames), pixfmt, _ptr(dims),
                                                    _ptr(rc), MEM_HOST, None)
        if e != 0:
            raise IcerError("icer_b200_decompress_batch_pix -> %d" % e)
        return dict(rc=rc, dims=dims, frames=frames)

    def synth_frames(self, dst, n, w, h, first_frame, seed, kind, stream=None):
        """Fill a DEVICE buffer (torch tensor) with n synthetic planes (bench utility)."""
        e = self.lib.icer_b200_synth_frames_u16(self.h, _ptr(dst), n, w, h, first_frame, seed, kind, stream)
        if e != 0:
            raise IcerError("icer_b200_synth_frames_u16 -> %d" % e)

    def dwt_batch(self, planes, stages, filt, inverse=False, dtype=np.uint16):
        """dtype uint16: icer_b200_dwt_batch_u16; uint8: the native uint8 transform (icer_b200_dwt_batch_u8)."""
        a = np.ascontiguousarray(planes, dtype=dtype).copy()
        n, h, w = a.shape
        rc = np.zeros(n, dtype=np.int32)
        f = self.lib.icer_b200_dwt_batch_u8 if dtype == np.uint8 else self.lib.icer_b200_dwt_batch_u16
        e = f(self.h, _ptr(a), n, w, h, stages, filt, 1 if inverse else 0, _ptr(rc), MEM_HOST, None)
        if e != 0:
            raise IcerError("icer_b200_dwt_batch -> %d" % e)
        return rc, a

    def set_limits_u8(self, planes=0, max_packets=0):
        rc = self.lib.icer_b200_set_limits_u8(self.h, planes, max_packets)
        if rc != 0:
            raise IcerError("icer_b200_set_limits_u8 -> %d" % rc)


class IcerMulti:
    """One batch over several contexts (icer_b200_compress_batch_multi / icer_b200_decompress_batch_multi): context g codes
    the frames [g*n/G, (g+1)*n/G), one host thread per context inside the library.  devices: list of CUDA ordinals (a device
    may appear more than once: every entry gets its own context)."""

    def __init__(self, devices, planes=9, max_stages=6, max_packets=800):
        self.ctxs = [IcerB200(device=d, planes=planes, max_stages=max_stages, max_packets=max_packets) for d in devices]
        self.lib = self.ctxs[0].lib
        self._arr = (C.c_void_p * len(self.ctxs))(*[c.h for c in self.ctxs])

    def close(self):
        for c in self.ctxs:
            c.close()

    def launch_count(self):
        return sum(c.launch_count() for c in self.ctxs)

    def compress_batch_raw(self, params, frames, n, out, out_stride, size_used, rc, mem, pixfmt=PIX_U16):
        p = _Params(*params)
        e = self.lib.icer_b200_compress_batch_multi(self._arr, len(self.ctxs), C.byref(p), _ptr(frames), pixfmt, n, _ptr(out), out_stride,
                                                    _ptr(size_used), _ptr(rc), mem)
        if e != 0:
            raise IcerError("icer_b200_compress_batch_multi -> %d" % e)

    def decompress_batch_raw(self, params, streams, stream_stride, stream_len, n, frames, dims, rc, mem, pixfmt=PIX_U16):
        p = _Params(*params)
        e = self.lib.icer_b200_decompress_batch_multi(self._arr, len(self.ctxs), C.byref(p), _ptr(streams), stream_stride, _ptr(stream_len),
                                                      n, _ptr(frames), pixfmt, _ptr(dims), _ptr(rc), mem)
        if e != 0:
            raise IcerError("icer_b200_decompress_batch_multi -> %d" % e)

    def compress_batch(self, frames, stages, filt, segments, quota):
        frames = np.ascontiguousarray(frames, dtype=np.uint16)
        n, chan, h, w = frames.shape
        out = np.zeros((n, quota), dtype=np.uint8)
        used = np.zeros(n, dtype=np.uint64)
        rc = np.zeros(n, dtype=np.int32)
        self.compress_batch_raw((w, h, chan, stages, filt, segments, quota), frames, n, out, quota, used, rc, MEM_HOST)
        return dict(rc=rc, size_used=used, streams=[bytes(out[i, :int(used[i])]) for i in range(n)])

    def decompress_batch(self, streams, chan, w, h, stages, filt, segments):
        n = len(streams)
        stride = max(max((len(s) for s in streams), default=1), 1)
        buf = np.zeros((n, stride), dtype=np.uint8)
        lens = np.zeros(n, dtype=np.uint64)
        for i, s in enumerate(streams):
            buf[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
            lens[i] = len(s)
        frames = np.zeros((n, chan, h * w), dtype=np.uint16)
        dims = np.zeros((n, 2), dtype=np.uint32)
        rc = np.zeros(n, dtype=np.int32)
        self.decompress_batch_raw((w, h, chan, stages, filt, segments, 0), buf, stride, lens, n, frames, dims, rc, MEM_HOST)
        return dict(rc=rc, dims=dims, frames=frames)


def device_count():
    return int(load_library().icer_b200_device_count())


class LegacyAPI:
    """The reference's single-image entry points of libicer_b200.so (icer_init, icer_compress_image_uint16, ...),
    called through the C ABI exactly like a C caller would.  The library's default context takes its plane count from
    ICER_B200_PLANES / ICER_B200_MAX_STAGES (environment, read at first use)."""

    def __init__(self):
        self.lib = load_library()
        rc = self.lib.icer_init()
        if rc != 0:
            raise IcerError("icer_init() -> %d: no usable CUDA device; there is no CPU fallback" % rc)

    def compress(self, planes_in, stages, filt, segments, quota, dtype=np.uint16):
        imgs = [np.ascontiguousarray(p, dtype=dtype).copy() for p in planes_in]
        h, w = imgs[0].shape
        buf = np.zeros(2 * quota + 64, dtype=np.uint8)
        ob = _OutBuf()
        u8 = buf.ctypes.data_as(C.POINTER(C.c_uint8))
        assert self.lib.icer_init_output_struct(C.byref(ob), u8, 2 * quota, quota) == 0
        pp = [im.ctypes.data_as(C.POINTER(C.c_uint8 if dtype == np.uint8 else C.c_uint16)) for im in imgs]
        if dtype == np.uint8:
            if len(imgs) == 1:
                rc = self.lib.icer_compress_image_uint8(pp[0], w, h, stages, filt, segments, C.byref(ob))
            else:
                rc = self.lib.icer_compress_image_yuv_uint8(pp[0], pp[1], pp[2], w, h, stages, filt, segments, C.byref(ob))
        elif len(imgs) == 1:
            rc = self.lib.icer_compress_image_uint16(pp[0], w, h, stages, filt, segments, C.byref(ob))
        else:
            rc = self.lib.icer_compress_image_yuv_uint16(pp[0], pp[1], pp[2], w, h, stages, filt, segments, C.byref(ob))
        n = ob.size_used
        return dict(rc=rc, size_used=n, stream=bytes(buf[quota:quota + n]), prio=bytes(buf[:n]), coefs=imgs)

    def decompress(self, stream, chan, bufsize_px, stages, filt, segments, dtype=np.uint16):
        data = np.frombuffer(bytes(stream) + b"\0" * 64, dtype=np.uint8).copy()
        imgs = [np.zeros(max(bufsize_px, 1), dtype=dtype) for _ in range(chan)]
        w, h = C.c_size_t(0), C.c_size_t(0)
        pp = [im.ctypes.data_as(C.POINTER(C.c_uint8 if dtype == np.uint8 else C.c_uint16)) for im in imgs]
        u8 = data.ctypes.data_as(C.POINTER(C.c_uint8))
        if dtype == np.uint8:
            if chan == 1:
                rc = self.lib.icer_decompress_image_uint8(pp[0], C.byref(w), C.byref(h), bufsize_px, u8, len(stream), stages, filt, segments)
            else:
                rc = self.lib.icer_decompress_image_yuv_uint8(pp[0], pp[1], pp[2], C.byref(w), C.byref(h), bufsize_px, u8,
                                                              len(stream), stages, filt, segments)
        elif chan == 1:
            rc = self.lib.icer_decompress_image_uint16(pp[0], C.byref(w), C.byref(h), bufsize_px, u8, len(stream), stages, filt, segments)
        else:
            rc = self.lib.icer_decompress_image_yuv_uint16(pp[0], pp[1], pp[2], C.byref(w), C.byref(h), bufsize_px, u8,
                                                           len(stream), stages, filt, segments)
        W, H = w.value, h.value
        out = [im[:W * H].reshape(H, W) for im in imgs] if rc == 0 and W * H <= bufsize_px else None
        return dict(rc=rc, w=W, h=H, images=out)

    def dwt8(self, img, stages, filt, inverse=False):
        a = np.ascontiguousarray(img, dtype=np.uint8).copy()
        f = self.lib.icer_inverse_wavelet_transform_stages_uint8 if inverse else self.lib.icer_wavelet_transform_stages_uint8
        rc = f(a.ctypes.data_as(C.POINTER(C.c_uint8)), a.shape[1], a.shape[0], stages, filt)
        return rc, a

    def dwt_forward(self, img, stages, filt):
        a = np.ascontiguousarray(img, dtype=np.uint16).copy()
        rc = self.lib.icer_wavelet_transform_stages_uint16(a.ctypes.data_as(C.POINTER(C.c_uint16)), a.shape[1], a.shape[0], stages, filt)
        return rc, a

    def dwt_inverse(self, img, stages, filt):
        a = np.ascontiguousarray(img, dtype=np.uint16).copy()
        rc = self.lib.icer_inverse_wavelet_transform_stages_uint16(a.ctypes.data_as(C.POINTER(C.c_uint16)), a.shape[1], a.shape[0], stages, filt)
        return rc, a


def u8_interleave_map(n):
    """Host-only planning helper of the library: where lib_icer's icer_interleave_uint8 puts element i (length n)."""
    dst = np.zeros(n, dtype=np.uint32)
    rc = load_library().icer_b200_u8_interleave_map(n, dst.ctypes.data)
    if rc != 0:
        raise IcerError("icer_b200_u8_interleave_map -> %d" % rc)
    return dst
