/*
 * icer_b200.h -- batch extension of the drop-in C ABI (libicer_b200.so).
 *
 * lib_icer's per-image entry points (include/icer.h) are synchronous and single-frame; throughput on a
 * B200 needs thousands of independent frames in flight.  These calls are what a reference-side binding
 * would use for batches; the seven legacy entry points are thin wrappers around them with n = 1.
 *
 * Plain C ABI: pointers, sizes and PODs only -- no CUDA or torch types in the signatures (`stream` is an
 * opaque cudaStream_t passed as void*, NULL = the default stream).
 *
 * Reference interfaces replaced (lib_icer/inc/icer.h):
 *   icer_b200_compress_batch_u16    <- icer_compress_image_uint16 :440, icer_compress_image_yuv_uint16 :442
 *   icer_b200_decompress_batch_u16  <- icer_decompress_image_uint16 :461, icer_decompress_image_yuv_uint16 :463
 *   icer_b200_set_limits            <- compile-time ICER_BITPLANES_TO_COMPRESS_16 / ICER_MAX_DECOMP_STAGES /
 *                                      ICER_MAX_PACKETS_16 (:29-46)
 */
#ifndef ICER_B200_H
#define ICER_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct icer_b200_ctx icer_b200_ctx;

/* where the caller's buffers live */
enum { ICER_B200_MEM_HOST = 0, ICER_B200_MEM_DEVICE = 1 };

typedef struct {
    uint32_t width, height;
    uint32_t channels;      /* 1 = gray, 3 = planar Y,U,V (same size) */
    uint32_t stages;        /* wavelet decomposition stages */
    uint32_t filter;        /* enum icer_filter_types */
    uint32_t segments;      /* error-containment segments per subband, 1..32 */
    uint64_t byte_quota;    /* per frame, same meaning as icer_init_output_struct's byte_quota */
} icer_b200_params;

/* Create / destroy a codec context bound to one CUDA device (device < 0: the current device).  A context
 * owns its device workspace and cached plans; it is not thread-safe (use one per host thread / rank). */
int icer_b200_create(icer_b200_ctx **ctx, int device);
void icer_b200_destroy(icer_b200_ctx *ctx);

/* Run-time equivalents of the reference's compile-time limits.  planes: bit planes coded (1..15, default 9);
 * max_stages default 6; max_packets default 800.  Pass 0 to keep a value. */
int icer_b200_set_limits(icer_b200_ctx *ctx, int planes, int max_stages, int max_packets);

/* The same for the native uint8 codec: planes 1..7 (default ICER_BITPLANES_TO_COMPRESS_8 = 7), max_packets (default
 * ICER_MAX_PACKETS = 300).  Pass 0 to keep a value. */
int icer_b200_set_limits_u8(icer_b200_ctx *ctx, int planes, int max_packets);

/* Per-stream payload scratch budget in bits per pixel per plane (default 3).  If any stream needs more the
 * encoder transparently retries with a doubled budget; this knob exists so tests can exercise that path. */
int icer_b200_set_scratch_bits(icer_b200_ctx *ctx, int bits_per_px_plane);

/*
 * Encode n frames.
 *   frames      [n][channels][height][width] uint16 (host or device per `mem`); NOT modified
 *   out         [n][out_stride] bytes; frame i's rearranged stream (the bytes lib_icer leaves at
 *               rearrange_start) is written at out + i*out_stride; out_stride >= byte_quota
 *   out_prio    optional (NULL to skip): same layout, the priority-ordered stream (lib_icer's data_start)
 *   coefs       optional (NULL to skip): [n][channels][height][width] uint16, receives what lib_icer leaves in
 *               the caller's image buffer (LL-mean-removed sign-magnitude wavelet coefficients)
 *   size_used   [n] uint64, rc [n] int32 (enum icer_status per frame; ICER_BYTE_QUOTA_EXCEEDED = normal lossy)
 * All pointers live in the memory space `mem`.  Returns ICER_RESULT_OK or a call-level error.
 * With device memory the call is asynchronous on `stream` unless a scratch retry is needed.
 */
int icer_b200_compress_batch_u16(icer_b200_ctx *ctx, const icer_b200_params *p, const uint16_t *frames, size_t n,
                                 uint8_t *out, size_t out_stride, uint8_t *out_prio, uint16_t *coefs,
                                 uint64_t *size_used, int32_t *rc, int mem, void *stream);

/*
 * Decode n streams produced with the same (channels, stages, filter, segments); width/height of `p` give the
 * capacity of each output slot (pixels per channel = width*height); the actual dimensions come from the
 * stream headers like in lib_icer and are returned in dims.
 *   streams     [n][stream_stride] bytes, stream i occupies stream_len[i] bytes
 *   frames      [n][channels][height*width] uint16 output
 *   dims        [n][2] uint32 (w, h) from the headers;  rc [n] int32
 */
int icer_b200_decompress_batch_u16(icer_b200_ctx *ctx, const icer_b200_params *p, const uint8_t *streams,
                                   size_t stream_stride, const uint64_t *stream_len, size_t n, uint16_t *frames,
                                   uint32_t *dims, int32_t *rc, int mem, void *stream);

/* Pixel formats at the batch boundary.  lib_icer itself only sees uint16 planes; its tools widen 8-bit samples and
 * convert packed RGB to planar Y,Cb,Cr on the host right before the call and back right after it
 * (example/src/icer_util.c:69-121,163-168,322-325, example/inc/color_util.h:27-34).  The *_pix calls do those steps on the
 * device, bit-identically, so 8-bit / RGB batches travel at 1-3 bytes per pixel:
 *   ICER_B200_PIX_U16     frames are uint16 [n][channels][h][w]                 (same as the calls above)
 *   ICER_B200_PIX_U8      frames are uint8  [n][channels][h][w]; decode clamps to 255 like the tools' display path
 *   ICER_B200_PIX_RGB888  frames are uint8  [n][h][w][3] packed R,G,B; params.channels must be 3
 *   ICER_B200_PIX_NATIVE_U8  frames are uint8 [n][channels][h][w] coded by lib_icer's NATIVE uint8 codec
 *                         (icer_compress_image_uint8 / _yuv_uint8, icer.h:386-390): int8 wavelet arithmetic (only <= 7-bit
 *                         data survives), ICER_BITPLANES_TO_COMPRESS_8 = 7 planes, ICER_MAX_PACKETS = 300, its own colour
 *                         segment order and its odd-length interleave defect -- all reproduced bit for bit */
enum { ICER_B200_PIX_U16 = 0, ICER_B200_PIX_U8 = 1, ICER_B200_PIX_RGB888 = 2, ICER_B200_PIX_NATIVE_U8 = 3 };
int icer_b200_compress_batch_pix(icer_b200_ctx *ctx, const icer_b200_params *p, const void *frames, int pixfmt, size_t n,
                                 uint8_t *out, size_t out_stride, uint64_t *size_used, int32_t *rc, int mem, void *stream);
int icer_b200_decompress_batch_pix(icer_b200_ctx *ctx, const icer_b200_params *p, const uint8_t *streams, size_t stream_stride,
                                   const uint64_t *stream_len, size_t n, void *frames, int pixfmt, uint32_t *dims, int32_t *rc,
                                   int mem, void *stream);

/*
 * One batch over several contexts -- normally one context per GPU of the box (frames are independent, the
 * batch shards by frame index, context g codes frames [g*n/nctx, (g+1)*n/nctx), no exchange between devices).  The call
 * runs one host thread per context and returns when all of them are done.  Contexts must be distinct.
 *   mem = ICER_B200_MEM_HOST    every context moves its slice over its own PCIe link
 *   mem = ICER_B200_MEM_DEVICE  all buffers live on the device of ctxs[0]; the other devices fetch their slice and return
 *                               their results with peer copies over NVLink (one scatter in, one gather out)
 * Arguments as in icer_b200_compress_batch_pix / icer_b200_decompress_batch_pix (pixfmt ICER_B200_PIX_U16 for uint16 frames).
 * No reference counterpart (lib_icer is single-threaded, src/icer_util.c:16-32); this is the C entry point of BASELINE's
 * "batch sharded across 8 GPUs" configuration. */
int icer_b200_compress_batch_multi(icer_b200_ctx *const *ctxs, int nctx, const icer_b200_params *p, const void *frames, int pixfmt,
                                   size_t n, uint8_t *out, size_t out_stride, uint64_t *size_used, int32_t *rc, int mem);
int icer_b200_decompress_batch_multi(icer_b200_ctx *const *ctxs, int nctx, const icer_b200_params *p, const uint8_t *streams,
                                     size_t stream_stride, const uint64_t *stream_len, size_t n, void *frames, int pixfmt,
                                     uint32_t *dims, int32_t *rc, int mem);
/* number of usable CUDA devices (0 if none) */
int icer_b200_device_count(void);

/* Wavelet transform only (parity hook for icer_wavelet_transform_stages_uint16 :446 / inverse :468) over a
 * batch of [n][height][width] planes, in place.  rc[n] receives ICER_INTEGER_OVERFLOW per plane (forward). */
int icer_b200_dwt_batch_u16(icer_b200_ctx *ctx, uint16_t *planes, size_t n, uint32_t width, uint32_t height,
                            uint32_t stages, uint32_t filter, int inverse, int32_t *rc, int mem, void *stream);

/* The native uint8 transform (icer_wavelet_transform_stages_uint8 :392 / inverse :414) over [n][height][width] uint8 planes. */
int icer_b200_dwt_batch_u8(icer_b200_ctx *ctx, uint8_t *planes, size_t n, uint32_t width, uint32_t height, uint32_t stages,
                           uint32_t filter, int inverse, int32_t *rc, int mem, void *stream);

/* Planning helper (host only, no device needed): dst[i] = where lib_icer's icer_interleave_uint8 (icer_wavelet.c:570-628) puts
 * element i of a length-len lows|highs array; for odd len this is NOT the perfect shuffle (the defect the native uint8 decoder
 * has to reproduce). */
int icer_b200_u8_interleave_map(uint32_t len, uint32_t *dst);

/* Number of kernels launched by this context so far (bench.py's gpu_launches), and the elapsed device time of
 * the last batch call's named phase in milliseconds (measured with CUDA events on the call's stream):
 * phase 0 = whole call, 1 = wavelet, 2 = entropy coder kernel, 3 = assemble/parse. */
uint64_t icer_b200_launch_count(const icer_b200_ctx *ctx);
float icer_b200_last_phase_ms(const icer_b200_ctx *ctx, int phase);
/* enable/disable phase timing (adds event records + one sync at the end of each batch call) */
void icer_b200_enable_timing(icer_b200_ctx *ctx, int on);

/* Benchmark utility: fill n planes of width*height uint16 in DEVICE memory with the deterministic synthetic frames of
 * SURVEY.md 8(d) (kind 0 natural-like 8-bit, 1 noise 8-bit, 2 natural-like 12-bit, 3 noise 12-bit; plane i is frame
 * first_frame + i).  Not part of lib_icer's interface. */
int icer_b200_synth_frames_u16(icer_b200_ctx *ctx, uint16_t *dst_device, size_t n, uint32_t width, uint32_t height,
                               uint32_t first_frame, uint32_t seed, int kind, void *stream);

const char *icer_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
