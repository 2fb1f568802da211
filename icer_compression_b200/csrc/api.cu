// api.cu -- the C ABI of libicer_b200.so: codec context, device workspace, batch encode/decode orchestration and
// the lib_icer-compatible single-image entry points (include/icer.h) implemented as batch-of-one wrappers.
//
// There is NO CPU codec path in this file: every compress/decompress call runs the CUDA kernels of wavelet.cu,
// wavelet_tma.cu, coder.cu, parse.cu and pixfmt.cu; if no CUDA device is usable the calls fail with ICER_FATAL_ERROR.
// Host-side work is limited to planning (plan.cu), the byte-stream scanner that locates CRC-valid segments
// (lib_icer does this on the host too: icer_compress.c:443-459,:569-588) and buffer management.
#include "common.h"
#include <chrono>
#include <algorithm>
#include <string.h>
#include <string>
#include <thread>
#include <mutex>

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return ICER_RESULT_OK;
        if (p) cudaFree(p);            // contents are never carried over; freeing first lets the allocator reuse the range
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            (void) cudaGetLastError();
            p = nullptr;
            want = bytes;              // retry without the growth margin before giving up
            CUDA_TRY(cudaMalloc(&p, want));
        }
        cap = want;
        return ICER_RESULT_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

struct icer_b200_ctx {
    int device = 0;
    int planes = ICER_BITPLANES_TO_COMPRESS_16, max_stages = ICER_MAX_DECOMP_STAGES, max_packets = ICER_MAX_PACKETS_16;
    int planes8 = ICER_BITPLANES_TO_COMPRESS_8, max_packets8 = ICER_MAX_PACKETS;      // native uint8 codec (icer.h:35-43)
    std::map<uint32_t, const uint32_t *> perms8;                                       // device tables, see u8_interleave_dst
    int scratch_bits = 3;
    std::map<std::tuple<uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, int, int>, Plan *> plans;
    LaunchStats ls;
    DevBuf coefs, tmp, scratch, bits, crc, rings, out, out_prio, size_used, rc, keep, offp, offr, mean, ovf, misc,
        streams, segoff, dims, frame_ok, cand, ncand, pinfo, list, pix, fin, tmp2, slices, pstat, vplanes;
    uint32_t ring_threads = 0;
    bool timing = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float phase_ms[4] = {0, 0, 0, 0};
    size_t workspace_budget = (size_t) 64 << 30;
    // priority-prefix encoding: statistics of the last quota walk (pinned, read lazily), the plan / quota they belong to
    uint32_t *h_stat = nullptr;
    cudaEvent_t ev_stat = nullptr;
    Plan *stat_plan = nullptr;
    uint64_t stat_quota = 0;
    uint32_t force_prefix = 0;        // ICER_B200_FORCE_PREFIX=n: always code the first n streams of the priority order first (tests)
    // host-buffer pipeline: chunks of `pipe_chunk` frames, H2D of chunk i+1 and D2H of chunk i-1 overlap the kernels of chunk i
    bool bulk_copy = true;        // strided bulk copies (one cudaMemcpy2DAsync per direction) on the host paths; ICER_B200_BULK_COPY=0 falls back to one copy per frame
    size_t pipe_chunk = 0;        // 0 = off (default): on B200 the per-chunk kernel tails cost more than the overlap saves; ICER_B200_PIPE_CHUNK=n enables it
    cudaStream_t cs_in = nullptr, cs_out = nullptr, cs_own = nullptr;    // cs_own: the stream of the *_multi calls
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    DevBuf p_in[2], p_out[2], p_used[2], p_rc[2], p_dims[2];
};

int ctx_device(const icer_b200_ctx *c) { return c->device; }

namespace {

Plan *get_plan(icer_b200_ctx *c, uint32_t w, uint32_t h, uint32_t chan, uint32_t stages, uint32_t segments, bool native8 = false) {
    const uint32_t planes = (uint32_t) (native8 ? c->planes8 : c->planes);
    auto key = std::make_tuple(w, h, chan, stages, segments, planes, c->scratch_bits, native8 ? 1 : 0);
    auto it = c->plans.find(key);
    if (it != c->plans.end()) return it->second;
    if (c->plans.size() > 64) {   // keep the cache bounded
        c->stat_plan = nullptr;
        for (auto &kv : c->plans) { free_plan_device(*kv.second); delete kv.second; }
        c->plans.clear();
    }
    Plan *pl = new Plan();
    build_plan(*pl, w, h, chan, stages, segments, planes, (uint32_t) c->max_stages,
               (uint32_t) (native8 ? c->max_packets8 : c->max_packets), c->scratch_bits, native8);
    if (upload_plan(*pl) != ICER_RESULT_OK) pl->status = ICER_FATAL_ERROR;
    c->plans[key] = pl;
    return pl;
}

// device copies of icer_interleave_uint8's placement for the odd line lengths of a w x h, `stages`-stage inverse transform
int ensure_perms8(icer_b200_ctx *c, uint32_t w, uint32_t h, uint32_t stages) {
    for (uint32_t s = 0; s < stages; s++)
        for (uint32_t n : {(uint32_t) dim_low(w, s), (uint32_t) dim_low(h, s)}) {
            if (!(n & 1) || c->perms8.count(n)) continue;
            std::vector<uint32_t> dst;
            u8_interleave_dst(n, dst);
            uint32_t *d = nullptr;
            CUDA_TRY(cudaMalloc(&d, (size_t) n * 4));
            CUDA_TRY(cudaMemcpy(d, dst.data(), (size_t) n * 4, cudaMemcpyHostToDevice));
            c->perms8[n] = d;
        }
    return ICER_RESULT_OK;
}

// ICER_B200_TRACE=1: host timestamps (ms since the first call) at the synchronisation points of the batch calls, on stderr
void trace(icer_b200_ctx *c, const char *what) {
    static const bool on = getenv("ICER_B200_TRACE") != nullptr;
    if (!on) return;
    static const auto t0 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[trace] %p %-18s %10.3f\n", (void *) c, what, ms);
}
void ev_record(icer_b200_ctx *c, int i, cudaStream_t st) {
    if (!c->timing) return;
    if (!c->ev[i]) cudaEventCreate(&c->ev[i]);
    cudaEventRecord(c->ev[i], st);
}

int copy_in(void *dst, const void *src, size_t bytes, int mem, cudaStream_t st) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, mem == ICER_B200_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    return ICER_RESULT_OK;
}
int copy_out(void *dst, const void *src, size_t bytes, int mem, cudaStream_t st) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, mem == ICER_B200_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    return ICER_RESULT_OK;
}

}  // namespace

extern "C" {

const char *icer_b200_version(void) { return "icer_b200 0.1 (sm_100a)"; }

int icer_b200_create(icer_b200_ctx **out, int device) {
    if (!out) return ICER_INVALID_INPUT;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        fprintf(stderr, "[icer_b200] no usable CUDA device (%s); this library has no CPU fallback\n", cudaGetErrorString(e));
        return ICER_FATAL_ERROR;
    }
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    CUDA_TRY(cudaSetDevice(device));
    icer_b200_ctx *c = new icer_b200_ctx();
    c->device = device;
    { const char *pc = getenv("ICER_B200_PIPE_CHUNK"); if (pc) c->pipe_chunk = (size_t) atol(pc); }
    { const char *bc = getenv("ICER_B200_BULK_COPY"); if (bc) c->bulk_copy = atoi(bc) != 0; }
    { const char *fp = getenv("ICER_B200_FORCE_PREFIX"); if (fp) c->force_prefix = (uint32_t) atoi(fp); }
    {   // encoder workspace per chunk: 55 % of the device memory by default (a B200 has 180 GB: a batch of 256 frames of 4096^2 -- ~300 MB of
        // planes, bit slices and payload scratch each -- then runs as ONE chunk; two chunks would pay the coder's serial tail twice)
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) c->workspace_budget = (size_t) ((double) total_b * 0.55);
        else (void) cudaGetLastError();
    }
    const char *gb = getenv("ICER_B200_WORKSPACE_GB");
    if (gb && atof(gb) > 0) c->workspace_budget = (size_t) (atof(gb) * (double) ((size_t) 1 << 30));
    *out = c;
    return ICER_RESULT_OK;
}

void icer_b200_destroy(icer_b200_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto &kv : c->plans) { free_plan_device(*kv.second); delete kv.second; }
    for (auto &kv : c->perms8) cudaFree((void *) kv.second);
    DevBuf *all[] = {&c->coefs, &c->tmp, &c->scratch, &c->bits, &c->crc, &c->rings, &c->out, &c->out_prio, &c->size_used, &c->rc,
                     &c->keep, &c->offp, &c->offr, &c->mean, &c->ovf, &c->misc, &c->streams, &c->segoff, &c->dims, &c->frame_ok,
                     &c->cand, &c->ncand, &c->pinfo, &c->list, &c->pix, &c->fin, &c->tmp2, &c->slices, &c->pstat, &c->vplanes};
    for (auto b : all) b->release();
    for (int k = 0; k < 2; k++) {
        c->p_in[k].release(); c->p_out[k].release(); c->p_used[k].release(); c->p_rc[k].release(); c->p_dims[k].release();
        if (c->ev_in[k]) cudaEventDestroy(c->ev_in[k]);
        if (c->ev_out[k]) cudaEventDestroy(c->ev_out[k]);
    }
    if (c->cs_in) cudaStreamDestroy(c->cs_in);
    if (c->cs_out) cudaStreamDestroy(c->cs_out);
    if (c->cs_own) cudaStreamDestroy(c->cs_own);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->ev_stat) cudaEventDestroy(c->ev_stat);
    if (c->h_stat) cudaFreeHost(c->h_stat);
    delete c;
}

static void drop_plans(icer_b200_ctx *c) {
    c->stat_plan = nullptr;
    for (auto &kv : c->plans) { free_plan_device(*kv.second); delete kv.second; }
    c->plans.clear();
}

int icer_b200_set_limits(icer_b200_ctx *c, int planes, int max_stages, int max_packets) {
    if (!c) return ICER_INVALID_INPUT;
    drop_plans(c);
    if (planes) { if (planes < 1 || planes > 15) return ICER_BITPLANE_OUT_OF_RANGE; c->planes = planes; }
    if (max_stages) { if (max_stages < 1 || max_stages > 15) return ICER_TOO_MANY_STAGES; c->max_stages = max_stages; }
    if (max_packets) c->max_packets = max_packets;
    return ICER_RESULT_OK;
}

int icer_b200_set_limits_u8(icer_b200_ctx *c, int planes, int max_packets) {
    if (!c) return ICER_INVALID_INPUT;
    drop_plans(c);
    if (planes) { if (planes < 1 || planes > 7) return ICER_BITPLANE_OUT_OF_RANGE; c->planes8 = planes; }   // icer_context_modeller.c:41-42
    if (max_packets) c->max_packets8 = max_packets;
    return ICER_RESULT_OK;
}

int icer_b200_u8_interleave_map(uint32_t len, uint32_t *dst) {
    if (!dst || len < 2) return ICER_INVALID_INPUT;
    std::vector<uint32_t> d;
    u8_interleave_dst(len, d);
    memcpy(dst, d.data(), (size_t) len * 4);
    return ICER_RESULT_OK;
}

int icer_b200_set_scratch_bits(icer_b200_ctx *c, int bits) {
    if (!c || bits < 1 || bits > 24) return ICER_INVALID_INPUT;
    c->scratch_bits = bits;
    return ICER_RESULT_OK;
}

uint64_t icer_b200_launch_count(const icer_b200_ctx *c) { return c ? c->ls.launches : 0; }
float icer_b200_last_phase_ms(const icer_b200_ctx *c, int phase) { return (c && phase >= 0 && phase < 4) ? c->phase_ms[phase] : 0.f; }
void icer_b200_enable_timing(icer_b200_ctx *c, int on) { if (c) c->timing = on != 0; }

// ---------------------------------------------------------------------------------------------------------
// batch encode
// ---------------------------------------------------------------------------------------------------------
static int compress_batch_any(icer_b200_ctx *c, const icer_b200_params *p, const void *frames_any, int pixfmt, size_t n, uint8_t *out,
                              size_t out_stride, uint8_t *out_prio, uint16_t *coefs_out, uint64_t *size_used, int32_t *rc,
                              int mem, void *stream);
static int decompress_batch_any(icer_b200_ctx *c, const icer_b200_params *p, const uint8_t *streams, size_t stream_stride,
                                const uint64_t *stream_len, size_t n, void *frames_any, int pixfmt, uint32_t *dims, int32_t *rc, int mem,
                                void *stream);

static int pipe_setup(icer_b200_ctx *c) {
    if (c->cs_in) return ICER_RESULT_OK;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->cs_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->cs_out, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_in[k], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_out[k], cudaEventDisableTiming));
    }
    return ICER_RESULT_OK;
}

// Host buffers, large batch: chunks of pipe_chunk frames run through the DEVICE-pointer path; the upload of chunk i+1
// (copy stream cs_in) and the download of chunk i-1 (cs_out) overlap the kernels of chunk i (caller's stream).
static int compress_host_pipelined(icer_b200_ctx *c, const icer_b200_params *p, const void *frames_any, int pixfmt, size_t n, uint8_t *out,
                                   size_t out_stride, uint64_t *size_used, int32_t *rc, cudaStream_t st) {
    int e = pipe_setup(c);
    if (e) return e;
    const size_t CH = c->pipe_chunk;
    const size_t frame_bytes = (size_t) p->width * p->height * p->channels * (pixfmt == ICER_B200_PIX_U16 ? 2 : 1);
    const size_t nchunks = (n + CH - 1) / CH;
    for (int k = 0; k < 2; k++) {
        if ((e = c->p_in[k].ensure(CH * frame_bytes))) return e;
        if ((e = c->p_out[k].ensure(CH * out_stride))) return e;
        if ((e = c->p_used[k].ensure(CH * 8))) return e;
        if ((e = c->p_rc[k].ensure(CH * 4))) return e;
    }
    const uint8_t *src = (const uint8_t *) frames_any;
    CUDA_TRY(cudaMemcpyAsync(c->p_in[0].p, src, std::min(CH, n) * frame_bytes, cudaMemcpyHostToDevice, c->cs_in));
    CUDA_TRY(cudaEventRecord(c->ev_in[0], c->cs_in));
    for (size_t i = 0; i < nchunks; i++) {
        const int b = (int) (i & 1);
        const size_t f0 = i * CH, nb = std::min(CH, n - f0);
        if (i + 1 < nchunks) {      // p_in[b^1] was consumed by chunk i-1, whose call has returned: upload the next chunk now
            const size_t f1 = (i + 1) * CH, nb1 = std::min(CH, n - f1);
            CUDA_TRY(cudaMemcpyAsync(c->p_in[b ^ 1].p, src + f1 * frame_bytes, nb1 * frame_bytes, cudaMemcpyHostToDevice, c->cs_in));
            CUDA_TRY(cudaEventRecord(c->ev_in[b ^ 1], c->cs_in));
        }
        CUDA_TRY(cudaStreamWaitEvent(st, c->ev_in[b], 0));
        if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_out[b], 0));       // p_out[b] still being downloaded (chunk i-2)
        e = compress_batch_any(c, p, c->p_in[b].p, pixfmt, nb, c->p_out[b].as<uint8_t>(), out_stride, nullptr, nullptr,
                               c->p_used[b].as<uint64_t>(), c->p_rc[b].as<int32_t>(), ICER_B200_MEM_DEVICE, st);
        if (e) return e;
        CUDA_TRY(cudaMemcpyAsync(size_used + f0, c->p_used[b].p, nb * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(rc + f0, c->p_rc[b].p, nb * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        size_t maxlen = 0;
        for (size_t k = 0; k < nb; k++) maxlen = std::max<size_t>(maxlen, (size_t) size_used[f0 + k]);
        if (maxlen) CUDA_TRY(cudaMemcpy2DAsync(out + f0 * out_stride, out_stride, c->p_out[b].p, out_stride, maxlen, nb, cudaMemcpyDeviceToHost, c->cs_out));
        CUDA_TRY(cudaEventRecord(c->ev_out[b], c->cs_out));
    }
    CUDA_TRY(cudaStreamSynchronize(c->cs_out));
    return ICER_RESULT_OK;
}

static int decompress_host_pipelined(icer_b200_ctx *c, const icer_b200_params *p, const uint8_t *streams, size_t stream_stride,
                                     const uint64_t *stream_len, size_t n, void *frames_any, int pixfmt, uint32_t *dims, int32_t *rc,
                                     cudaStream_t st) {
    int e = pipe_setup(c);
    if (e) return e;
    const size_t CH = c->pipe_chunk;
    const size_t cap_px = (size_t) p->width * p->height;
    const size_t frame_bytes = cap_px * p->channels * (pixfmt == ICER_B200_PIX_U16 ? 2 : 1);
    const size_t nchunks = (n + CH - 1) / CH;
    for (int k = 0; k < 2; k++) {
        if ((e = c->p_in[k].ensure(CH * stream_stride + 16))) return e;
        if ((e = c->p_out[k].ensure(CH * frame_bytes))) return e;
        if ((e = c->p_used[k].ensure(CH * 8))) return e;
        if ((e = c->p_rc[k].ensure(CH * 4))) return e;
        if ((e = c->p_dims[k].ensure(CH * 8))) return e;
    }
    auto upload = [&](size_t i, int b) -> int {
        const size_t f0 = i * CH, nb = std::min(CH, n - f0);
        size_t maxlen = 0;
        for (size_t k = 0; k < nb; k++) maxlen = std::max<size_t>(maxlen, (size_t) std::min<uint64_t>(stream_len[f0 + k], stream_stride));
        if (maxlen) CUDA_TRY(cudaMemcpy2DAsync(c->p_in[b].p, stream_stride, streams + f0 * stream_stride, stream_stride, maxlen, nb, cudaMemcpyHostToDevice, c->cs_in));
        CUDA_TRY(cudaMemcpyAsync(c->p_used[b].p, stream_len + f0, nb * 8, cudaMemcpyHostToDevice, c->cs_in));
        CUDA_TRY(cudaEventRecord(c->ev_in[b], c->cs_in));
        return ICER_RESULT_OK;
    };
    if ((e = upload(0, 0))) return e;
    for (size_t i = 0; i < nchunks; i++) {
        const int b = (int) (i & 1);
        const size_t f0 = i * CH, nb = std::min(CH, n - f0);
        if (i + 1 < nchunks && (e = upload(i + 1, b ^ 1))) return e;
        CUDA_TRY(cudaStreamWaitEvent(st, c->ev_in[b], 0));
        if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_out[b], 0));
        e = decompress_batch_any(c, p, c->p_in[b].as<uint8_t>(), stream_stride, c->p_used[b].as<uint64_t>(), nb, c->p_out[b].p, pixfmt,
                                 c->p_dims[b].as<uint32_t>(), c->p_rc[b].as<int32_t>(), ICER_B200_MEM_DEVICE, st);
        if (e) return e;
        CUDA_TRY(cudaMemcpyAsync(rc + f0, c->p_rc[b].p, nb * 4, cudaMemcpyDeviceToHost, st));
        if (dims) CUDA_TRY(cudaMemcpyAsync(dims + 2 * f0, c->p_dims[b].p, nb * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpyAsync((uint8_t *) frames_any + f0 * frame_bytes, c->p_out[b].p, nb * frame_bytes, cudaMemcpyDeviceToHost, c->cs_out));
        CUDA_TRY(cudaEventRecord(c->ev_out[b], c->cs_out));
    }
    CUDA_TRY(cudaStreamSynchronize(c->cs_out));
    return ICER_RESULT_OK;
}

static int compress_batch_any(icer_b200_ctx *c, const icer_b200_params *p, const void *frames_any, int pixfmt, size_t n, uint8_t *out,
                              size_t out_stride, uint8_t *out_prio, uint16_t *coefs_out, uint64_t *size_used, int32_t *rc,
                              int mem, void *stream) {
    const uint16_t *frames = (const uint16_t *) frames_any;
    if (pixfmt != ICER_B200_PIX_U16 && pixfmt != ICER_B200_PIX_U8 && pixfmt != ICER_B200_PIX_RGB888 && pixfmt != ICER_B200_PIX_NATIVE_U8) return ICER_INVALID_INPUT;
    if (pixfmt == ICER_B200_PIX_RGB888 && p && p->channels != 3) return ICER_INVALID_INPUT;
    if (!c || !p || !frames || !out || !size_used || !rc) return ICER_INVALID_INPUT;
    const bool native8 = pixfmt == ICER_B200_PIX_NATIVE_U8;
    if (p->channels != 1 && p->channels != 3) return ICER_INVALID_INPUT;
    if (p->filter > ICER_FILTER_Q || p->width == 0 || p->height == 0) return ICER_INVALID_INPUT;
    if (out_stride < p->byte_quota) return ICER_OUTPUT_BUF_TOO_SMALL;
    if (n == 0) return ICER_RESULT_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t) stream;
    const uint32_t w = p->width, h = p->height, chan = p->channels;
    Plan *pl = get_plan(c, w, h, chan, p->stages, p->segments, native8);
    if (pl->status == ICER_FATAL_ERROR) return ICER_FATAL_ERROR;
    const bool host = mem == ICER_B200_MEM_HOST;
    // a packet-count violation is found by lib_icer only AFTER the transform (icer_compress.c:327 vs :22): frames that also
    // overflow report ICER_INTEGER_OVERFLOW, so that case runs the transform and skips only the coder
    const int pstatus = pl->status;
    if (host && c->pipe_chunk && n > c->pipe_chunk && pstatus == ICER_RESULT_OK && !out_prio && !coefs_out)
        return compress_host_pipelined(c, p, frames_any, pixfmt, n, out, out_stride, size_used, rc, st);
    if (pstatus != ICER_RESULT_OK) {   // every frame of this shape fails the same way, nothing is written
        std::vector<uint64_t> z(n, 0);
        std::vector<int32_t> r(n, pstatus);
        CUDA_TRY(cudaMemcpyAsync(size_used, z.data(), n * 8, host ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(rc, r.data(), n * 4, host ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return ICER_RESULT_OK;
    }
    const size_t plane_px = (size_t) w * h, frame_px = plane_px * chan;
    if (c->ring_threads == 0) c->ring_threads = encode_resident_threads(c->device);
    if (c->stat_plan && c->ev_stat && cudaEventQuery(c->ev_stat) == cudaSuccess) {
        // the previous batch's quota walk: if its frames kept well under 3/4 of their streams, the next batch of that shape and quota
        // codes a prefix (what was kept + 1/8 + 32 streams) first
        Plan *sp = c->stat_plan;
        c->stat_plan = nullptr;
        bool alive = false;
        for (auto &kv : c->plans) alive = alive || kv.second == sp;
        if (alive) {
            const uint32_t kept = c->h_stat[0];
            const uint64_t want = (uint64_t) kept + kept / 8 + 32;
            sp->enc_prefix = (want * 4 < (uint64_t) sp->nstreams * 3) ? (uint32_t) want : 0;
            sp->enc_prefix_quota = c->stat_quota;
        }
    } else if (c->stat_plan) {
        (void) cudaGetLastError();
    }

    ev_record(c, 0, st);
    float ms_dwt = 0, ms_enc = 0, ms_asm = 0; trace(c, "enc:enter");
    size_t done = 0;
    while (done < n) {
        // chunk size under the workspace budget: every per-frame buffer counts (coefs, tmp rounded up to 32-frame groups, tmp2
        // and fin at 2 B/sample each, the 8-bit staging, payload scratch, per-stream arrays, host staging of the output slots)
        const size_t slice_bytes_per_frame = enc_slice_words(pl->planes, pl->nslots) * 4 / 32;
        const size_t per_frame = frame_px * 2 * 4 + slice_bytes_per_frame + (pixfmt != ICER_B200_PIX_U16 ? frame_px : 0) + pl->scratch_words * 4 +
                                 (size_t) pl->nstreams * 24 + (host ? out_stride * (out_prio ? 2 : 1) : 0) + 64;
        size_t nb = std::max<size_t>(1, c->workspace_budget / per_frame);
        if (nb > 32) nb &= ~(size_t) 31;             // whole coder groups
        if (nb < n - done) {                         // several chunks: equal ones (a short last chunk pays the coder's serial tail alone)
            const size_t chunks = (n - done + nb - 1) / nb;
            nb = std::min(nb, (((n - done + chunks - 1) / chunks) + 31) & ~(size_t) 31);
        }
        nb = std::min(nb, n - done);
        int e;
        uint32_t ngroups = 0;
        for (;;) {      // a failed allocation halves the chunk instead of failing the call
            ngroups = (uint32_t) ((nb + 31) / 32);
            e = c->coefs.ensure(nb * frame_px * 2);
            if (!e) e = c->tmp.ensure(nb * frame_px * 2);
            if (!e) e = c->slices.ensure((size_t) ngroups * enc_slice_words(pl->planes, pl->nslots) * 4);
            if (!e) e = c->fin.ensure(host || pixfmt != ICER_B200_PIX_U16 ? nb * frame_px * 2 : 0);
            if (!e) e = c->scratch.ensure(nb * pl->scratch_words * 4);
            if (!e) e = c->bits.ensure(nb * pl->nstreams * 4);
            if (!e) e = c->crc.ensure(nb * pl->nstreams * 4);
            if (!e) e = c->offp.ensure(nb * pl->nstreams * 8);
            if (!e) e = c->offr.ensure(nb * pl->nstreams * 8);
            if (!e) break;
            if (nb <= 1) return e;
            nb = (nb + 1) / 2;
        }
        if ((e = c->rings.ensure((size_t) c->ring_threads * ICER_RING * 2))) return e;
        if ((e = c->mean.ensure(nb * chan * 2))) return e;
        if ((e = c->ovf.ensure(nb * chan * 4))) return e;
        if ((e = c->keep.ensure(nb * 4))) return e;
        if ((e = c->misc.ensure(64))) return e;
        uint8_t *d_out = out + done * out_stride, *d_prio = out_prio ? out_prio + done * out_stride : nullptr;
        uint64_t *d_used = size_used + done;
        int32_t *d_rc = rc + done;
        if (host) {
            if ((e = c->out.ensure(nb * out_stride))) return e;
            if (out_prio && (e = c->out_prio.ensure(nb * out_stride))) return e;
            if ((e = c->size_used.ensure(nb * 8))) return e;
            if ((e = c->rc.ensure(nb * 4))) return e;
            d_out = c->out.as<uint8_t>(); d_prio = out_prio ? c->out_prio.as<uint8_t>() : nullptr;
            d_used = c->size_used.as<uint64_t>(); d_rc = c->rc.as<int32_t>();
        }
        uint16_t *d_coefs = c->coefs.as<uint16_t>();
        // the planes to transform: the caller's device frames as they are, or a staging buffer (host input / 8-bit / RGB)
        const uint16_t *d_in = frames + done * frame_px;
        if (pixfmt == ICER_B200_PIX_U16) {
            if (host) {
                if ((e = c->fin.ensure(nb * frame_px * 2))) return e;
                CUDA_TRY(cudaMemcpyAsync(c->fin.p, frames + done * frame_px, nb * frame_px * 2, cudaMemcpyHostToDevice, st));
                d_in = c->fin.as<uint16_t>();
            }
        } else {
            // 8-bit samples / packed RGB: one byte per sample crosses PCIe, widening / colour conversion happen here
            const uint8_t *src8 = (const uint8_t *) frames_any + done * frame_px;
            if (host) {
                if ((e = c->pix.ensure(nb * frame_px))) return e;
                CUDA_TRY(cudaMemcpyAsync(c->pix.p, src8, nb * frame_px, cudaMemcpyHostToDevice, st));
                src8 = c->pix.as<uint8_t>();
            }
            if ((e = c->fin.ensure(nb * frame_px * 2))) return e;
            if (pixfmt == ICER_B200_PIX_U8) e = launch_u8_to_u16(src8, c->fin.as<uint16_t>(), nb * frame_px, st, c->ls);
            else if (native8) e = launch_s8_to_i16(src8, c->fin.as<uint16_t>(), nb * frame_px, st, c->ls);
            else e = launch_rgb_to_yuv(src8, c->fin.as<uint16_t>(), nb, plane_px, st, c->ls);
            if (e) return e;
            d_in = c->fin.as<uint16_t>();
        }
        CUDA_TRY(cudaMemsetAsync(c->ovf.p, 0, nb * chan * 4, st));
        ev_record(c, 1, st);
        static const bool no_tma = getenv("ICER_B200_NO_TMA") != nullptr;
        if (!native8 && !no_tma && dwt_fused_applicable(d_in, w, h, plane_px)) {
            // fused row+column stage, tiles staged by TMA; reads d_in, leaves it untouched
            if ((e = c->tmp2.ensure(nb * frame_px * 2))) return e;
            if ((e = launch_dwt_forward_fused(d_in, plane_px, d_coefs, c->tmp.as<uint16_t>(), c->tmp2.as<uint16_t>(), nb * chan, w, h,
                                              p->stages, p->filter, c->ovf.as<int>(), st, c->ls))) return e;
        } else {
            CUDA_TRY(cudaMemcpyAsync(d_coefs, d_in, nb * frame_px * 2, cudaMemcpyDeviceToDevice, st));
            if ((e = launch_dwt_forward(d_coefs, c->tmp.as<uint16_t>(), nb * chan, w, h, p->stages, p->filter, c->ovf.as<int>(), st, c->ls, native8))) return e;
        }
        // sign-magnitude is applied on the way into the coder's layout unless the caller wants the planar coefficients back
        const bool fuse_sm = !native8 && !coefs_out;
        if ((e = launch_ll_mean_signmag(d_coefs, nb * chan, w, h, (uint32_t) dim_low(w, p->stages), (uint32_t) dim_low(h, p->stages),
                                        c->mean.as<uint16_t>(), c->ovf.as<int>(), st, c->ls, native8, fuse_sm))) return e;
        if (coefs_out && !native8) if ((e = copy_out(coefs_out + done * frame_px, d_coefs, nb * frame_px * 2, mem, st))) return e;
        if (coefs_out && native8) {     // the uint8 codec leaves 8-bit sign-magnitude coefficients behind
            if ((e = c->pix.ensure(nb * frame_px))) return e;
            if ((e = launch_sm16_to_sm8(d_coefs, c->pix.as<uint8_t>(), nb * frame_px, st, c->ls))) return e;
            if ((e = copy_out((uint8_t *) coefs_out + done * frame_px, c->pix.p, nb * frame_px, mem, st))) return e;
        }
        // bit slices of the coefficients for the coder (lane = frame of a group of 32)
        uint32_t *d_bs = c->slices.as<uint32_t>();
        if ((e = launch_bitslice(d_coefs, d_bs, pl->d_slots, pl->nslots, (uint32_t) nb, ngroups, chan, w, h, pl->planes, fuse_sm, st, c->ls))) return e;
        ev_record(c, 2, st);
        // entropy coder + segment assembly.
        //  * a stream that outgrows its payload slot (rare) restarts the coder with a larger scratch budget;
        //  * priority-prefix encoding: with a byte quota far below the lossless size most streams are coded only to be dropped by
        //    the quota walk (lib_icer itself stops coding at the first packet that does not fit, icer_compress.c:404-406).  When the
        //    previous batch of this shape kept only a fraction of its streams, the coder first produces a prefix of the priority
        //    order, the quota walk tells whether any frame wants more, and only then the rest is coded.  Results do not depend
        //    on the prefix length (tests/test_gpu_parity.py::test_priority_prefix_encoding).
        uint32_t *d_stat = c->misc.as<uint32_t>() + 8;
        AssembleArgs aa;
        auto fill_aa = [&]() {
            aa.nframes = (uint32_t) nb; aa.nstreams = pl->nstreams; aa.planes = pl->planes; aa.chan = chan; aa.w = w; aa.h = h;
            aa.units = pl->d_units; aa.prio = pl->d_prio; aa.rearr = pl->d_rearr; aa.soff = pl->d_soff; aa.scratch_words = pl->scratch_words;
            aa.scratch = c->scratch.as<uint32_t>(); aa.bits = c->bits.as<uint32_t>(); aa.crc = c->crc.as<uint32_t>();
            aa.mean = c->mean.as<uint16_t>(); aa.ovf = c->ovf.as<int>(); aa.plan_status = pl->enc_status; aa.quota = p->byte_quota;
            aa.out = d_out; aa.out_stride = out_stride; aa.out_prio = d_prio; aa.size_used = d_used; aa.rc = d_rc;
            aa.keep = c->keep.as<uint32_t>(); aa.off_prio = c->offp.as<uint64_t>(); aa.off_rearr = c->offr.as<uint64_t>();
            aa.stat = d_stat; aa.limit = pl->nstreams; aa.phase = 0;
        };
        // run the coder over list[lo, hi) of the plan's prefix / remainder lists (nullptr = every stream); returns the overflow flag
        auto run_coder = [&](const uint32_t *list, uint32_t nlist, int *ovf) -> int {
            if ((e = c->scratch.ensure(nb * pl->scratch_words * 4))) return e;
            CUDA_TRY(cudaMemsetAsync(c->misc.p, 0, 32, st));      // work counter + overflow flag (the quota statistics live behind them)
            EncodeArgs ea;
            ea.bs = d_bs; ea.nslots = pl->nslots; ea.nframes = (uint32_t) nb; ea.ngroups = ngroups; ea.planes = pl->planes;
            ea.units = pl->d_units; ea.nunits = (uint32_t) pl->units.size(); ea.soff = pl->d_soff; ea.scap = pl->d_scap;
            ea.scratch_words = pl->scratch_words; ea.scratch = c->scratch.as<uint32_t>(); ea.bits = c->bits.as<uint32_t>();
            ea.rings = c->rings.as<uint16_t>(); ea.nring_threads = c->ring_threads;
            ea.work_counter = c->misc.as<unsigned long long>(); ea.overflow = c->misc.as<int>() + 4;
            ea.list = list; ea.nlist = nlist;
            if ((e = launch_encode(ea, st, c->ls))) return e;
            CUDA_TRY(cudaMemcpyAsync(ovf, c->misc.as<int>() + 4, 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st)); trace(c, "enc:coded");
            return ICER_RESULT_OK;
        };
        static const bool no_prefix = getenv("ICER_B200_NO_PREFIX") != nullptr;
        for (bool coded = false; !coded && pl->enc_status == ICER_RESULT_OK;) {
            uint32_t prefix = 0;
            if (!no_prefix && pl->enc_prefix && pl->enc_prefix < pl->nstreams && pl->enc_prefix_quota == p->byte_quota && p->byte_quota < ((uint64_t) 1 << 28))
                prefix = pl->enc_prefix;
            if (c->force_prefix && c->force_prefix < pl->nstreams && p->byte_quota < ((uint64_t) 1 << 28)) prefix = c->force_prefix;     // test hook
            int ovf = 0;
            if (!prefix) {
                if ((e = run_coder(nullptr, 0, &ovf))) return e;
            } else {
                if ((e = upload_prefix_lists(*pl, prefix, st))) return e;
                CUDA_TRY(cudaMemsetAsync(c->bits.p, 0xF0, nb * pl->nstreams * 4, st));      // ICER_BITS_NOT_CODED everywhere
                if ((e = run_coder(pl->d_list, prefix, &ovf))) return e;
                if (!ovf) {
                    // does any frame keep the whole prefix?  then it may keep more: code the remainder as well
                    fill_aa();
                    aa.limit = prefix; aa.phase = 1;
                    CUDA_TRY(cudaMemsetAsync(d_stat, 0, 8, st));
                    if ((e = launch_assemble(aa, st, c->ls))) return e;
                    uint32_t stat[2] = {0, 0};
                    CUDA_TRY(cudaMemcpyAsync(stat, d_stat, 8, cudaMemcpyDeviceToHost, st));
                    CUDA_TRY(cudaStreamSynchronize(st));
                    if (stat[1]) { if ((e = run_coder(pl->d_list + prefix, pl->nstreams - prefix, &ovf))) return e; }
                }
            }
            if (!ovf) { coded = true; break; }
            if (c->scratch_bits >= 24) { fprintf(stderr, "[icer_b200] payload scratch overflow at the worst-case bound\n"); return ICER_FATAL_ERROR; }
            c->scratch_bits = std::min(24, c->scratch_bits * 2);
            const uint32_t keep_prefix = pl->enc_prefix; const uint64_t keep_q = pl->enc_prefix_quota;
            pl = get_plan(c, w, h, chan, p->stages, p->segments, native8);
            if (pl->status != ICER_RESULT_OK) return ICER_FATAL_ERROR;
            pl->enc_prefix = keep_prefix; pl->enc_prefix_quota = keep_q;
        }
        ev_record(c, 3, st);
        fill_aa();
        CUDA_TRY(cudaMemsetAsync(d_stat, 0, 8, st));
        if ((e = launch_assemble(aa, st, c->ls))) return e;
        // learn the prefix for the next batch of this shape and quota from how many streams the frames kept (read lazily: the copy is
        // consumed at the start of the next call)
        if (pl->enc_status == ICER_RESULT_OK && !no_prefix) {
            if (!c->h_stat) { CUDA_TRY(cudaMallocHost((void **) &c->h_stat, 16)); CUDA_TRY(cudaEventCreateWithFlags(&c->ev_stat, cudaEventDisableTiming)); }
            CUDA_TRY(cudaMemcpyAsync(c->h_stat, d_stat, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaEventRecord(c->ev_stat, st));
            c->stat_plan = pl; c->stat_quota = p->byte_quota;
        }
        ev_record(c, 4, st);
        if (host) {
            CUDA_TRY(cudaMemcpyAsync(size_used + done, d_used, nb * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(rc + done, d_rc, nb * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st)); trace(c, "enc:assembled");
            // only the used part of every slot travels back
            size_t maxlen = 0;
            for (size_t i = 0; i < nb; i++) maxlen = std::max<size_t>(maxlen, (size_t) size_used[done + i]);
            if (c->bulk_copy && maxlen) {      // one strided copy: every slot up to the longest stream (bytes behind size_used are unspecified)
                CUDA_TRY(cudaMemcpy2DAsync(out + done * out_stride, out_stride, d_out, out_stride, maxlen, nb, cudaMemcpyDeviceToHost, st));
                if (out_prio) CUDA_TRY(cudaMemcpy2DAsync(out_prio + done * out_stride, out_stride, d_prio, out_stride, maxlen, nb, cudaMemcpyDeviceToHost, st));
            } else
            for (size_t i = 0; i < nb; i++) {
                const size_t len = (size_t) size_used[done + i];
                if (len) CUDA_TRY(cudaMemcpyAsync(out + (done + i) * out_stride, d_out + i * out_stride, len, cudaMemcpyDeviceToHost, st));
                if (len && out_prio) CUDA_TRY(cudaMemcpyAsync(out_prio + (done + i) * out_stride, d_prio + i * out_stride, len, cudaMemcpyDeviceToHost, st));
            }
            CUDA_TRY(cudaStreamSynchronize(st)); trace(c, "enc:copied");
        }
        if (c->timing) {
            CUDA_TRY(cudaStreamSynchronize(st));
            float t;
            cudaEventElapsedTime(&t, c->ev[1], c->ev[2]); ms_dwt += t;
            cudaEventElapsedTime(&t, c->ev[2], c->ev[3]); ms_enc += t;
            cudaEventElapsedTime(&t, c->ev[3], c->ev[4]); ms_asm += t;
        }
        done += nb;
    }
    if (c->timing) {
        cudaEventElapsedTime(&c->phase_ms[0], c->ev[0], c->ev[4]);
        c->phase_ms[1] = ms_dwt; c->phase_ms[2] = ms_enc; c->phase_ms[3] = ms_asm;
    }
    return ICER_RESULT_OK;
}

int icer_b200_compress_batch_u16(icer_b200_ctx *c, const icer_b200_params *p, const uint16_t *frames, size_t n, uint8_t *out,
                                 size_t out_stride, uint8_t *out_prio, uint16_t *coefs_out, uint64_t *size_used, int32_t *rc,
                                 int mem, void *stream) {
    return compress_batch_any(c, p, frames, ICER_B200_PIX_U16, n, out, out_stride, out_prio, coefs_out, size_used, rc, mem, stream);
}

int icer_b200_compress_batch_pix(icer_b200_ctx *c, const icer_b200_params *p, const void *frames, int pixfmt, size_t n, uint8_t *out,
                                 size_t out_stride, uint64_t *size_used, int32_t *rc, int mem, void *stream) {
    return compress_batch_any(c, p, frames, pixfmt, n, out, out_stride, nullptr, nullptr, size_used, rc, mem, stream);
}

// ---------------------------------------------------------------------------------------------------------
// stream scanner (host): lib_icer's icer_find_packet_in_bytestream, icer_compress.c:569-588
// ---------------------------------------------------------------------------------------------------------
static inline uint32_t rd16(const uint8_t *p) { return p[0] | (p[1] << 8); }
static inline uint32_t rd32(const uint8_t *p) { return rd16(p) | (rd16(p + 2) << 16); }

uint32_t icer_calculate_packet_crc32(const icer_image_segment_typedef *pkt) {
    return host_crc32((const uint8_t *) pkt, ICER_HDR_BYTES - 4);
}
uint32_t icer_calculate_segment_crc32(const icer_image_segment_typedef *pkt) {
    return host_crc32((const uint8_t *) pkt + ICER_HDR_BYTES, (pkt->data_length + 7) / 8);
}

static bool segment_at(const uint8_t *p, size_t left, size_t *len) {
    if (left < ICER_HDR_BYTES) return false;   // lib_icer reads the header without this check (past the buffer)
    if (rd16(p) != ICER_PACKET_PREAMBLE) return false;
    if (rd32(p + 24) != host_crc32(p, 24)) return false;
    const size_t nbytes = ((size_t) rd32(p + 16) + 7) / 8;
    if (nbytes > left - ICER_HDR_BYTES) return false;
    if (rd32(p + 20) != host_crc32(p + ICER_HDR_BYTES, nbytes)) return false;
    *len = ICER_HDR_BYTES + nbytes;
    return true;
}

int icer_find_packet_in_bytestream(const icer_image_segment_typedef **seg, const uint8_t *datastream, size_t data_length,
                                   size_t *offset) {
    *offset = 0;
    *seg = nullptr;
    while (*offset < data_length) {
        size_t len;
        if (segment_at(datastream + *offset, data_length - *offset, &len)) {
            *seg = (const icer_image_segment_typedef *) (datastream + *offset);
            *offset += len;
            return ICER_RESULT_OK;
        }
        (*offset)++;
    }
    return ICER_DECODER_OUT_OF_DATA;
}

int icer_get_image_dimensions(const uint8_t *datastream, size_t data_length, size_t *image_w, size_t *image_h) {
    if (!datastream || !image_w || !image_h) return ICER_INVALID_INPUT;   // icer_compress.c:541-566: FIRST valid segment
    size_t off = 0;
    while (off < data_length) {
        size_t len;
        if (segment_at(datastream + off, data_length - off, &len)) {
            *image_w = rd32(datastream + off + 8);
            *image_h = rd32(datastream + off + 12);
            return ICER_RESULT_OK;
        }
        off++;
    }
    return ICER_DECODER_OUT_OF_DATA;
}

namespace {
struct SegRec { uint32_t off; uint8_t level, subband, seg, lsb_chan; };
struct Parsed { std::vector<SegRec> segs; uint32_t w = 0, h = 0; uint16_t mean[3] = {0, 0, 0}; bool any = false; };

void parse_stream(const uint8_t *s, size_t len, uint32_t chan, Parsed &out) {
    size_t off = 0;
    while (off < len) {
        size_t slen;
        if (segment_at(s + off, len - off, &slen)) {
            const uint8_t *p = s + off;
            out.segs.push_back(SegRec{(uint32_t) off, p[4], p[5], p[6], p[7]});
            out.w = rd32(p + 8); out.h = rd32(p + 12);               // the LAST valid header wins (icer_compress.c:454-456)
            const uint32_t ch = chan == 1 ? 0 : (p[7] >> 4);
            if (ch < 3) out.mean[ch] = (uint16_t) rd16(p + 2);
            out.any = true;
            off += slen;
        } else {
            off++;
        }
    }
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------
// batch decode
// ---------------------------------------------------------------------------------------------------------
static int decompress_batch_any(icer_b200_ctx *c, const icer_b200_params *p, const uint8_t *streams, size_t stream_stride,
                                const uint64_t *stream_len, size_t n, void *frames_any, int pixfmt, uint32_t *dims, int32_t *rc, int mem,
                                void *stream) {
    uint16_t *frames = (uint16_t *) frames_any;
    if (pixfmt != ICER_B200_PIX_U16 && pixfmt != ICER_B200_PIX_U8 && pixfmt != ICER_B200_PIX_RGB888 && pixfmt != ICER_B200_PIX_NATIVE_U8) return ICER_INVALID_INPUT;
    if (pixfmt == ICER_B200_PIX_RGB888 && p && p->channels != 3) return ICER_INVALID_INPUT;
    const bool native8 = pixfmt == ICER_B200_PIX_NATIVE_U8;
    if (!c || !p || !streams || !stream_len || !frames || !rc) return ICER_INVALID_INPUT;
    if (p->channels != 1 && p->channels != 3) return ICER_INVALID_INPUT;
    if (p->filter > ICER_FILTER_Q) return ICER_INVALID_INPUT;
    if (n == 0) return ICER_RESULT_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t) stream;
    const bool host = mem == ICER_B200_MEM_HOST;
    if (host && c->pipe_chunk && n > c->pipe_chunk)
        return decompress_host_pipelined(c, p, streams, stream_stride, stream_len, n, frames_any, pixfmt, dims, rc, st);
    const uint32_t chan = p->channels;
    const size_t cap_px = (size_t) p->width * p->height;
    ev_record(c, 0, st); trace(c, "dec:enter");

    // 1. the streams and their lengths on the device
    int e;
    const uint8_t *d_streams = streams;
    const uint64_t *d_lens = stream_len;
    const size_t total_bytes = n * stream_stride;
    if (host) {
        if ((e = c->streams.ensure(total_bytes + 16))) return e;
        if ((e = c->dims.ensure(n * 8))) return e;
        size_t maxlen = 0;
        for (size_t i = 0; i < n; i++) maxlen = std::max<size_t>(maxlen, (size_t) std::min<uint64_t>(stream_len[i], stream_stride));
        // only the used part of every slot travels: the kernels never read a frame's slot behind its stream_len
        if (c->bulk_copy && maxlen && maxlen < stream_stride)
            CUDA_TRY(cudaMemcpy2DAsync(c->streams.p, stream_stride, streams, stream_stride, maxlen, n, cudaMemcpyHostToDevice, st));
        else if (maxlen)
            CUDA_TRY(cudaMemcpyAsync(c->streams.p, streams, (n - 1) * stream_stride + maxlen, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c->dims.p, stream_len, n * 8, cudaMemcpyHostToDevice, st));
        d_streams = c->streams.as<uint8_t>();
        d_lens = c->dims.as<uint64_t>();
    }
    // 2. scan every stream for CRC-valid segments on the device (parse.cu)
    uint32_t cap = (uint32_t) std::min<uint64_t>(8192, stream_stride / ICER_HDR_BYTES + 1);
    { const char *pc = getenv("ICER_B200_PARSE_CAP"); if (pc && atoi(pc) > 0) cap = (uint32_t) atoi(pc); }   // test hook
    uint32_t cap_pow2 = 1;
    while (cap_pow2 < cap) cap_pow2 <<= 1;
    if ((e = c->cand.ensure((size_t) n * cap * 8))) return e;
    if ((e = c->ncand.ensure(n * 4))) return e;
    if ((e = c->pinfo.ensure(n * sizeof(ParseInfo)))) return e;
    ParseArgs pa;
    pa.streams = d_streams; pa.stride = stream_stride; pa.lens = d_lens; pa.n = (uint32_t) n; pa.chan = chan;
    pa.cand = c->cand.as<uint2>(); pa.cap = cap; pa.cap_pow2 = cap_pow2; pa.ncand = c->ncand.as<uint32_t>();
    pa.info = c->pinfo.as<ParseInfo>();
    if ((e = launch_parse(pa, st, c->ls))) return e;
    std::vector<ParseInfo> info(n);
    CUDA_TRY(cudaMemcpyAsync(info.data(), c->pinfo.p, n * sizeof(ParseInfo), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st)); trace(c, "dec:parsed");
    // frames with more candidates than the device path holds are scanned on the host (never seen with real streams)
    std::map<size_t, Parsed> host_parsed;
    for (size_t i = 0; i < n; i++) {
        if (!info[i].overflow) continue;
        std::vector<uint8_t> tmp;
        const uint8_t *hs = streams + i * stream_stride;
        uint64_t len_i = 0;
        if (host) len_i = stream_len[i];
        else {
            CUDA_TRY(cudaMemcpy(&len_i, stream_len + i, 8, cudaMemcpyDeviceToHost));
            tmp.resize(stream_stride);
            CUDA_TRY(cudaMemcpy(tmp.data(), streams + i * stream_stride, stream_stride, cudaMemcpyDeviceToHost));
            hs = tmp.data();
        }
        Parsed ps;
        parse_stream(hs, (size_t) std::min<uint64_t>(len_i, stream_stride), chan, ps);
        info[i].w = ps.w; info[i].h = ps.h; info[i].any = ps.any;
        for (int k = 0; k < 3; k++) info[i].mean[k] = ps.mean[k];
        host_parsed[i] = std::move(ps);
    }
    ev_record(c, 1, st);
    // 3. per-frame status; frames are grouped by their header dimensions (one plan per group)
    std::vector<int32_t> hrc(n, ICER_RESULT_OK);
    std::vector<uint32_t> hdims(n * 2, 0);
    std::map<std::pair<uint32_t, uint32_t>, std::vector<size_t>> groups;
    for (size_t i = 0; i < n; i++) {
        hdims[2 * i] = info[i].w; hdims[2 * i + 1] = info[i].h;
        if (!info[i].any) { hrc[i] = ICER_DECODER_OUT_OF_DATA; continue; }   // lib_icer: dimensions left uninitialised
        if (cap_px < (size_t) info[i].w * info[i].h) { hrc[i] = ICER_BYTE_QUOTA_EXCEEDED; continue; }   // icer_compress.c:461-463
        groups[{info[i].w, info[i].h}].push_back(i);
    }
    const bool own16 = host || pixfmt != ICER_B200_PIX_U16;     // decode into our own uint16 planes, convert / copy afterwards
    if (own16 && (e = c->coefs.ensure(n * chan * cap_px * 2))) return e;
    if ((e = c->mean.ensure(n * chan * 2))) return e;
    if ((e = c->misc.ensure(64))) return e;
    uint16_t *d_frames = own16 ? c->coefs.as<uint16_t>() : frames;
    float ms_dec = 0, ms_dwt = 0;
    for (auto &g : groups) {
        const uint32_t w = g.first.first, h = g.first.second;
        Plan *pl = get_plan(c, w, h, chan, p->stages, p->segments, native8);
        if (pl->status == ICER_FATAL_ERROR) return ICER_FATAL_ERROR;
        if (pl->status != ICER_RESULT_OK) {
            // lib_icer's decoder reports the partition / stage errors per image (icer_compress.c:480, icer_wavelet.c:92-94)
            for (size_t i : g.second) hrc[i] = pl->status;
            continue;
        }
        // frames of this dimension group, 32 per coder group (lane = frame); segment table; LL means
        const size_t plane_px = (size_t) w * h;
        const uint32_t ngroups = (uint32_t) ((g.second.size() + 31) / 32);
        std::vector<int32_t> frame_of((size_t) ngroups * 32, -1);
        std::vector<uint32_t> list;
        std::vector<uint16_t> hmean(n * chan, 0);
        size_t k = 0;
        for (size_t i : g.second) {
            frame_of[k++] = (int32_t) i;
            if (!info[i].overflow) list.push_back((uint32_t) i);
            for (uint32_t ch = 0; ch < chan; ch++) hmean[i * chan + ch] = info[i].mean[ch];
        }
        if ((e = c->tmp.ensure(n * chan * plane_px * 2))) return e;
        if ((e = c->slices.ensure((size_t) ngroups * dec_slice_words(pl->planes, pl->nslots) * 4))) return e;
        if ((e = c->pstat.ensure(n * pl->units.size() * pl->planes + 16))) return e;
        if ((e = c->vplanes.ensure(n * pl->units.size() + 16))) return e;
        if ((e = c->segoff.ensure(n * (size_t) pl->nstreams * 4))) return e;
        if ((e = c->frame_ok.ensure(frame_of.size() * 4))) return e;
        if ((e = c->list.ensure(list.size() * 4 + 4))) return e;
        CUDA_TRY(cudaMemsetAsync(c->segoff.p, 0, n * (size_t) pl->nstreams * 4, st));          // entries are offset + 1, 0 = absent
        CUDA_TRY(cudaMemcpyAsync(c->list.p, list.data(), list.size() * 4, cudaMemcpyHostToDevice, st));
        TableArgs ta;
        ta.list = c->list.as<uint32_t>(); ta.nlist = (uint32_t) list.size(); ta.cand = c->cand.as<uint2>(); ta.cap = cap;
        ta.streams = d_streams; ta.stride = stream_stride; ta.info = c->pinfo.as<ParseInfo>();
        ta.seg_off = c->segoff.as<uint32_t>(); ta.nstreams = pl->nstreams; ta.planes = pl->planes; ta.chan = chan;
        ta.max_level = pl->max_stages; ta.unit_lut = pl->d_unit_lut;
        if ((e = launch_table(ta, st, c->ls))) return e;
        for (size_t i : g.second) {     // host-scanned frames (overflow fallback): their table rows come from the host
            auto hp = host_parsed.find(i);
            if (hp == host_parsed.end()) continue;
            std::vector<uint32_t> rowv(pl->nstreams, 0u);
            for (auto &r : hp->second.segs) {
                const int ch = chan == 1 ? 0 : (r.lsb_chan >> 4);
                const int lsb = r.lsb_chan & 0x0F;
                if (lsb >= (int) pl->planes) continue;
                auto it = pl->unit_of.find(std::make_tuple(ch, (int) r.level, (int) r.subband, (int) r.seg));
                if (it == pl->unit_of.end()) continue;
                rowv[it->second * pl->planes + lsb] = r.off + 1;
            }
            CUDA_TRY(cudaMemcpyAsync(c->segoff.as<uint32_t>() + i * pl->nstreams, rowv.data(), rowv.size() * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        CUDA_TRY(cudaMemcpyAsync(c->mean.p, hmean.data(), hmean.size() * 2, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c->frame_ok.p, frame_of.data(), frame_of.size() * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemsetAsync(c->misc.p, 0, 64, st));
        ev_record(c, 2, st);
        DecodeArgs da;
        da.bs = c->slices.as<uint32_t>(); da.nslots = pl->nslots; da.pstat = c->pstat.as<uint8_t>(); da.ngroups = ngroups; da.planes = pl->planes;
        da.units = pl->d_units; da.nunits = (uint32_t) pl->units.size();
        da.frame_of = c->frame_ok.as<int32_t>();
        da.streams = d_streams; da.stream_stride = stream_stride; da.stream_len = d_lens;
        da.seg_off = c->segoff.as<uint32_t>();
        da.work_counter = c->misc.as<unsigned long long>();
        if ((e = launch_decode(da, st, c->ls))) return e;
        ev_record(c, 3, st);
        // 4. back to planar frames (from sign-magnitude, + LL mean), inverse wavelet, clamp
        const uint32_t llw = (uint32_t) dim_low(w, p->stages), llh = (uint32_t) dim_low(h, p->stages);
        const bool dense = plane_px == cap_px;
        static const bool no_tma = getenv("ICER_B200_NO_TMA") != nullptr;
        bool fused = false;
        if (!native8 && !no_tma && dense && g.second.size() == n && p->filter == ICER_FILTER_A) {
            // fused inverse stages (TMA-staged sub-band tiles): coefficients go to their own planar buffer, the intermediate LL regions
            // ping-pong between the (now free) T buffer and `fin`, the last stage writes the caller-visible planes
            if ((e = c->tmp2.ensure(n * chan * plane_px * 2))) return e;
            if ((e = c->fin.ensure(n * chan * plane_px * 2))) return e;
            fused = dwt_inverse_fused_applicable(c->tmp2.p, d_frames, w, h, p->stages, p->filter) && ((uintptr_t) c->fin.p % 16) == 0;
        }
        uint16_t *d_planar = fused ? c->tmp2.as<uint16_t>() : d_frames;
        // every pixel of the frames of this group is written (icer_compress.c:470: the image starts at zero)
        if ((e = launch_unslice(c->slices.as<uint32_t>(), d_planar, pl->d_slots, c->pstat.as<uint8_t>(), c->vplanes.as<uint8_t>(), n, c->frame_ok.as<int32_t>(),
                                c->mean.as<uint16_t>(), pl->nslots, (uint32_t) pl->units.size(), ngroups, chan, w, h, cap_px, pl->planes, llw, llh,
                                st, c->ls, native8))) return e;
        if (native8 && (e = ensure_perms8(c, w, h, p->stages))) return e;
        if (fused) {
            if ((e = launch_dwt_inverse_fused(d_planar, c->tmp.as<uint16_t>(), c->fin.as<uint16_t>(), d_frames, n * chan, w, h, p->stages,
                                              p->filter, true, st, c->ls))) return e;
        } else if (dense && g.second.size() == n) {
            if ((e = launch_dwt_inverse(d_frames, c->tmp.as<uint16_t>(), n * chan, w, h, p->stages, p->filter, true, st, c->ls, native8, &c->perms8))) return e;
        } else {
            for (size_t i : g.second)
                for (uint32_t ch = 0; ch < chan; ch++) {
                    uint16_t *pp = d_frames + (i * chan + ch) * cap_px;
                    if ((e = launch_dwt_inverse(pp, c->tmp.as<uint16_t>(), 1, w, h, p->stages, p->filter, true, st, c->ls, native8, &c->perms8))) return e;
                }
        }
        ev_record(c, 4, st);
        if (c->timing) {
            CUDA_TRY(cudaStreamSynchronize(st));
            float t;
            cudaEventElapsedTime(&t, c->ev[2], c->ev[3]); ms_dec += t;
            cudaEventElapsedTime(&t, c->ev[3], c->ev[4]); ms_dwt += t;
        }
    }
    // 5. results back (8-bit clamp / Y,Cb,Cr -> RGB on the device when asked for)
    const size_t out_bytes_per_frame = pixfmt == ICER_B200_PIX_U16 ? chan * cap_px * 2 : chan * cap_px;
    uint8_t *d_out8 = nullptr;
    if (pixfmt != ICER_B200_PIX_U16) {
        d_out8 = (uint8_t *) frames_any;
        if (host) { if ((e = c->pix.ensure(n * out_bytes_per_frame))) return e; d_out8 = c->pix.as<uint8_t>(); }
        for (auto &g : groups) {
            const size_t px = (size_t) g.first.first * g.first.second;
            for (size_t i : g.second) {
                if (hrc[i] != ICER_RESULT_OK) continue;
                if (native8) {
                    for (uint32_t ch = 0; ch < chan; ch++)
                        if ((e = launch_i16_to_b8(d_frames + (i * chan + ch) * cap_px, d_out8 + (i * chan + ch) * cap_px, px, st, c->ls))) return e;
                } else if (pixfmt == ICER_B200_PIX_U8) {
                    for (uint32_t ch = 0; ch < chan; ch++)
                        if ((e = launch_u16_to_u8(d_frames + (i * chan + ch) * cap_px, d_out8 + (i * chan + ch) * cap_px, px, st, c->ls))) return e;
                } else {
                    if ((e = launch_yuv_to_rgb(d_frames + i * 3 * cap_px, d_out8 + i * 3 * cap_px, 1, px, cap_px, st, c->ls))) return e;
                }
            }
        }
    }
    if (host) {
        bool all_ok = c->bulk_copy;
        for (size_t i = 0; i < n && all_ok; i++) all_ok = hrc[i] == ICER_RESULT_OK;
        if (all_ok) {        // every frame decoded: one copy for the whole batch
            if (pixfmt == ICER_B200_PIX_U16) CUDA_TRY(cudaMemcpyAsync(frames, d_frames, n * chan * cap_px * 2, cudaMemcpyDeviceToHost, st));
            else CUDA_TRY(cudaMemcpyAsync(frames_any, d_out8, n * out_bytes_per_frame, cudaMemcpyDeviceToHost, st));
        } else
        for (size_t i = 0; i < n; i++)
            if (hrc[i] == ICER_RESULT_OK) {
                if (pixfmt == ICER_B200_PIX_U16)
                    CUDA_TRY(cudaMemcpyAsync(frames + i * chan * cap_px, d_frames + i * chan * cap_px, chan * cap_px * 2, cudaMemcpyDeviceToHost, st));
                else
                    CUDA_TRY(cudaMemcpyAsync((uint8_t *) frames_any + i * out_bytes_per_frame, d_out8 + i * out_bytes_per_frame, out_bytes_per_frame, cudaMemcpyDeviceToHost, st));
            }
        memcpy(rc, hrc.data(), n * 4);
        if (dims) memcpy(dims, hdims.data(), n * 8);
    } else {
        CUDA_TRY(cudaMemcpyAsync(rc, hrc.data(), n * 4, cudaMemcpyHostToDevice, st));
        if (dims) CUDA_TRY(cudaMemcpyAsync(dims, hdims.data(), n * 8, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));   // host vectors above must outlive the copies
    trace(c, "dec:done");
    if (c->timing) {
        ev_record(c, 4, st);
        CUDA_TRY(cudaStreamSynchronize(st));
        cudaEventElapsedTime(&c->phase_ms[0], c->ev[0], c->ev[4]);
        float t = 0;
        cudaEventElapsedTime(&t, c->ev[0], c->ev[1]);
        c->phase_ms[1] = ms_dwt; c->phase_ms[2] = ms_dec; c->phase_ms[3] = t;
    }
    return ICER_RESULT_OK;
}

int icer_b200_decompress_batch_u16(icer_b200_ctx *c, const icer_b200_params *p, const uint8_t *streams, size_t stream_stride,
                                   const uint64_t *stream_len, size_t n, uint16_t *frames, uint32_t *dims, int32_t *rc, int mem,
                                   void *stream) {
    return decompress_batch_any(c, p, streams, stream_stride, stream_len, n, frames, ICER_B200_PIX_U16, dims, rc, mem, stream);
}

int icer_b200_decompress_batch_pix(icer_b200_ctx *c, const icer_b200_params *p, const uint8_t *streams, size_t stream_stride,
                                   const uint64_t *stream_len, size_t n, void *frames, int pixfmt, uint32_t *dims, int32_t *rc, int mem,
                                   void *stream) {
    return decompress_batch_any(c, p, streams, stream_stride, stream_len, n, frames, pixfmt, dims, rc, mem, stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// one batch over several contexts (several GPUs of one box): frames are independent, so context g gets the
// frames [g*n/G, (g+1)*n/G) and there is no exchange between the devices.  One host thread per context.
//   mem = HOST   : every context moves its own slice over its own PCIe link
//   mem = DEVICE : the buffers live on the device of ctxs[0]; the other devices fetch their slice and return their
//                  results with peer copies (NVLink through the NVSwitch) -- the single scatter / gather of SURVEY 8(e)
// ---------------------------------------------------------------------------------------------------------
namespace {
struct MultiSlice { size_t f0, nb; };
MultiSlice slice_of(size_t n, int G, int g) {
    const size_t a = n * (size_t) g / (size_t) G, b = n * (size_t) (g + 1) / (size_t) G;
    return MultiSlice{a, b - a};
}
int own_stream(icer_b200_ctx *c, cudaStream_t *st) {
    CUDA_TRY(cudaSetDevice(c->device));
    if (!c->cs_own) CUDA_TRY(cudaStreamCreateWithFlags(&c->cs_own, cudaStreamNonBlocking));
    *st = c->cs_own;
    return ICER_RESULT_OK;
}
void enable_peer(int dev, int peer) {
    if (dev == peer) return;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, dev, peer) == cudaSuccess && can) {
        cudaSetDevice(dev);
        if (cudaDeviceEnablePeerAccess(peer, 0) != cudaSuccess) (void) cudaGetLastError();    // already enabled is fine
    }
}
template <typename F> int run_per_context(int nctx, F f) {
    std::vector<int> res((size_t) nctx, ICER_RESULT_OK);
    std::vector<std::thread> th;
    for (int g = 1; g < nctx; g++) th.emplace_back([&, g]() { res[(size_t) g] = f(g); });
    res[0] = f(0);
    for (auto &t : th) t.join();
    for (int r : res) if (r != ICER_RESULT_OK) return r;
    return ICER_RESULT_OK;
}
bool multi_args_ok(icer_b200_ctx *const *ctxs, int nctx) {
    if (!ctxs || nctx < 1 || nctx > 64) return false;
    for (int g = 0; g < nctx; g++) {
        if (!ctxs[g]) return false;
        for (int k = 0; k < g; k++) if (ctxs[k] == ctxs[g]) return false;      // a context serves one host thread at a time
    }
    return true;
}
}  // namespace

extern "C" int icer_b200_compress_batch_multi(icer_b200_ctx *const *ctxs, int nctx, const icer_b200_params *p, const void *frames,
                                              int pixfmt, size_t n, uint8_t *out, size_t out_stride, uint64_t *size_used, int32_t *rc,
                                              int mem) {
    if (!multi_args_ok(ctxs, nctx) || !p || !frames || !out || !size_used || !rc) return ICER_INVALID_INPUT;
    const size_t frame_bytes = (size_t) p->width * p->height * p->channels * (pixfmt == ICER_B200_PIX_U16 ? 2 : 1);
    const int dev0 = ctxs[0]->device;
    return run_per_context(nctx, [&](int g) -> int {
        icer_b200_ctx *c = ctxs[g];
        const MultiSlice sl = slice_of(n, nctx, g);
        if (sl.nb == 0) return ICER_RESULT_OK;
        cudaStream_t st;
        int e = own_stream(c, &st);
        if (e) return e;
        const uint8_t *src = (const uint8_t *) frames + sl.f0 * frame_bytes;
        uint8_t *dst = out + sl.f0 * out_stride;
        if (mem == ICER_B200_MEM_HOST || c->device == dev0) {
            e = compress_batch_any(c, p, src, pixfmt, sl.nb, dst, out_stride, nullptr, nullptr, size_used + sl.f0, rc + sl.f0, mem, st);
            if (e) return e;
            CUDA_TRY(cudaStreamSynchronize(st));
            return ICER_RESULT_OK;
        }
        // device buffers on dev0, this context on another GPU: scatter the slice in, gather the results back (peer copies)
        enable_peer(c->device, dev0);
        CUDA_TRY(cudaSetDevice(c->device));
        if ((e = c->p_in[0].ensure(sl.nb * frame_bytes))) return e;
        if ((e = c->p_out[0].ensure(sl.nb * out_stride))) return e;
        if ((e = c->p_used[0].ensure(sl.nb * 8))) return e;
        if ((e = c->p_rc[0].ensure(sl.nb * 4))) return e;
        CUDA_TRY(cudaMemcpyPeerAsync(c->p_in[0].p, c->device, src, dev0, sl.nb * frame_bytes, st));
        e = compress_batch_any(c, p, c->p_in[0].p, pixfmt, sl.nb, c->p_out[0].as<uint8_t>(), out_stride, nullptr, nullptr,
                               c->p_used[0].as<uint64_t>(), c->p_rc[0].as<int32_t>(), ICER_B200_MEM_DEVICE, st);
        if (e) return e;
        CUDA_TRY(cudaSetDevice(c->device));
        CUDA_TRY(cudaMemcpyPeerAsync(dst, dev0, c->p_out[0].p, c->device, sl.nb * out_stride, st));
        CUDA_TRY(cudaMemcpyPeerAsync(size_used + sl.f0, dev0, c->p_used[0].p, c->device, sl.nb * 8, st));
        CUDA_TRY(cudaMemcpyPeerAsync(rc + sl.f0, dev0, c->p_rc[0].p, c->device, sl.nb * 4, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return ICER_RESULT_OK;
    });
}

extern "C" int icer_b200_decompress_batch_multi(icer_b200_ctx *const *ctxs, int nctx, const icer_b200_params *p, const uint8_t *streams,
                                                size_t stream_stride, const uint64_t *stream_len, size_t n, void *frames, int pixfmt,
                                                uint32_t *dims, int32_t *rc, int mem) {
    if (!multi_args_ok(ctxs, nctx) || !p || !streams || !stream_len || !frames || !rc) return ICER_INVALID_INPUT;
    const size_t frame_bytes = (size_t) p->width * p->height * p->channels * (pixfmt == ICER_B200_PIX_U16 ? 2 : 1);
    const int dev0 = ctxs[0]->device;
    return run_per_context(nctx, [&](int g) -> int {
        icer_b200_ctx *c = ctxs[g];
        const MultiSlice sl = slice_of(n, nctx, g);
        if (sl.nb == 0) return ICER_RESULT_OK;
        cudaStream_t st;
        int e = own_stream(c, &st);
        if (e) return e;
        const uint8_t *src = streams + sl.f0 * stream_stride;
        uint8_t *dst = (uint8_t *) frames + sl.f0 * frame_bytes;
        uint32_t *dd = dims ? dims + 2 * sl.f0 : nullptr;
        if (mem == ICER_B200_MEM_HOST || c->device == dev0) {
            e = decompress_batch_any(c, p, src, stream_stride, stream_len + sl.f0, sl.nb, dst, pixfmt, dd, rc + sl.f0, mem, st);
            if (e) return e;
            CUDA_TRY(cudaStreamSynchronize(st));
            return ICER_RESULT_OK;
        }
        enable_peer(c->device, dev0);
        CUDA_TRY(cudaSetDevice(c->device));
        if ((e = c->p_in[0].ensure(sl.nb * stream_stride + 16))) return e;
        if ((e = c->p_out[0].ensure(sl.nb * frame_bytes))) return e;
        if ((e = c->p_used[0].ensure(sl.nb * 8))) return e;
        if ((e = c->p_rc[0].ensure(sl.nb * 4))) return e;
        if ((e = c->p_dims[0].ensure(sl.nb * 8))) return e;
        CUDA_TRY(cudaMemcpyPeerAsync(c->p_in[0].p, c->device, src, dev0, sl.nb * stream_stride, st));
        CUDA_TRY(cudaMemcpyPeerAsync(c->p_used[0].p, c->device, stream_len + sl.f0, dev0, sl.nb * 8, st));
        e = decompress_batch_any(c, p, c->p_in[0].as<uint8_t>(), stream_stride, c->p_used[0].as<uint64_t>(), sl.nb, c->p_out[0].p, pixfmt,
                                 c->p_dims[0].as<uint32_t>(), c->p_rc[0].as<int32_t>(), ICER_B200_MEM_DEVICE, st);
        if (e) return e;
        CUDA_TRY(cudaSetDevice(c->device));
        CUDA_TRY(cudaMemcpyPeerAsync(dst, dev0, c->p_out[0].p, c->device, sl.nb * frame_bytes, st));
        if (dd) CUDA_TRY(cudaMemcpyPeerAsync(dd, dev0, c->p_dims[0].p, c->device, sl.nb * 8, st));
        CUDA_TRY(cudaMemcpyPeerAsync(rc + sl.f0, dev0, c->p_rc[0].p, c->device, sl.nb * 4, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return ICER_RESULT_OK;
    });
}

extern "C" int icer_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void) cudaGetLastError(); return 0; }
    return n;
}

extern "C" {

// ---------------------------------------------------------------------------------------------------------
// wavelet-only hook
// ---------------------------------------------------------------------------------------------------------
int icer_b200_dwt_batch_u16(icer_b200_ctx *c, uint16_t *planes, size_t n, uint32_t width, uint32_t height, uint32_t stages,
                            uint32_t filter, int inverse, int32_t *rc, int mem, void *stream) {
    if (!c || !planes || filter > ICER_FILTER_Q) return ICER_INVALID_INPUT;
    if (n == 0) return ICER_RESULT_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t) stream;
    const bool host = mem == ICER_B200_MEM_HOST;
    const size_t ps = (size_t) width * height;
    if (dim_low(width, stages) < 3 || dim_low(height, stages) < 3) {   // icer_wavelet.c:66-68
        if (rc) {
            std::vector<int32_t> r(n, ICER_TOO_MANY_STAGES);
            CUDA_TRY(cudaMemcpy(rc, r.data(), n * 4, host ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice));
        }
        return ICER_RESULT_OK;
    }
    int e;
    if ((e = c->tmp.ensure(n * ps * 2))) return e;
    if ((e = c->ovf.ensure(n * 4))) return e;
    uint16_t *d = planes;
    if (host) {
        if ((e = c->fin.ensure(n * ps * 2))) return e;
        d = c->fin.as<uint16_t>();
        CUDA_TRY(cudaMemcpyAsync(d, planes, n * ps * 2, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemsetAsync(c->ovf.p, 0, n * 4, st));
    static const bool no_tma = getenv("ICER_B200_NO_TMA") != nullptr;
    if (!inverse && !no_tma && dwt_fused_applicable(d, width, height, ps)) {
        // same fused TMA path as the encoder; the result replaces the input planes
        if ((e = c->coefs.ensure(n * ps * 2))) return e;
        if ((e = c->tmp2.ensure(n * ps * 2))) return e;
        if ((e = launch_dwt_forward_fused(d, ps, c->coefs.as<uint16_t>(), c->tmp.as<uint16_t>(), c->tmp2.as<uint16_t>(), n, width, height, stages,
                                          filter, c->ovf.as<int>(), st, c->ls))) return e;
        CUDA_TRY(cudaMemcpyAsync(d, c->coefs.p, n * ps * 2, cudaMemcpyDeviceToDevice, st));
        e = ICER_RESULT_OK;
    } else if (inverse && !no_tma && !c->coefs.ensure(n * ps * 2) && !c->tmp2.ensure(n * ps * 2) &&
               dwt_inverse_fused_applicable(d, c->coefs.p, width, height, stages, filter) && ((uintptr_t) c->tmp.p % 16) == 0 && ((uintptr_t) c->tmp2.p % 16) == 0) {
        // same fused TMA path as the decoder; the result replaces the input planes
        if ((e = launch_dwt_inverse_fused(d, c->tmp.as<uint16_t>(), c->tmp2.as<uint16_t>(), c->coefs.as<uint16_t>(), n, width, height, stages, filter,
                                          false, st, c->ls))) return e;
        CUDA_TRY(cudaMemcpyAsync(d, c->coefs.p, n * ps * 2, cudaMemcpyDeviceToDevice, st));
        e = ICER_RESULT_OK;
    } else if (inverse) e = launch_dwt_inverse(d, c->tmp.as<uint16_t>(), n, width, height, stages, filter, false, st, c->ls);
    else e = launch_dwt_forward(d, c->tmp.as<uint16_t>(), n, width, height, stages, filter, c->ovf.as<int>(), st, c->ls);
    if (e) return e;
    std::vector<int> ovf(n, 0);
    CUDA_TRY(cudaMemcpyAsync(ovf.data(), c->ovf.p, n * 4, cudaMemcpyDeviceToHost, st));
    if (host) CUDA_TRY(cudaMemcpyAsync(planes, d, n * ps * 2, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (rc) {
        std::vector<int32_t> r(n);
        for (size_t i = 0; i < n; i++) r[i] = ovf[i] ? ICER_INTEGER_OVERFLOW : ICER_RESULT_OK;
        CUDA_TRY(cudaMemcpy(rc, r.data(), n * 4, host ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice));
    }
    return ICER_RESULT_OK;
}

// native uint8 transform (icer_wavelet_transform_stages_uint8 / inverse, icer_wavelet.c:8-28,32-53): int8 samples; the inverse
// reproduces icer_interleave_uint8's placement for odd line lengths
int icer_b200_dwt_batch_u8(icer_b200_ctx *c, uint8_t *planes, size_t n, uint32_t width, uint32_t height, uint32_t stages,
                           uint32_t filter, int inverse, int32_t *rc, int mem, void *stream) {
    if (!c || !planes || filter > ICER_FILTER_Q) return ICER_INVALID_INPUT;
    if (n == 0) return ICER_RESULT_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t) stream;
    const bool host = mem == ICER_B200_MEM_HOST;
    const size_t ps = (size_t) width * height;
    if (dim_low(width, stages) < 3 || dim_low(height, stages) < 3) {
        if (rc) {
            std::vector<int32_t> r(n, ICER_TOO_MANY_STAGES);
            CUDA_TRY(cudaMemcpy(rc, r.data(), n * 4, host ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice));
        }
        return ICER_RESULT_OK;
    }
    int e;
    if ((e = c->tmp.ensure(n * ps * 2))) return e;
    if ((e = c->fin.ensure(n * ps * 2))) return e;
    if ((e = c->ovf.ensure(n * 4))) return e;
    uint8_t *d8 = planes;
    if (host) {
        if ((e = c->pix.ensure(n * ps))) return e;
        d8 = c->pix.as<uint8_t>();
        CUDA_TRY(cudaMemcpyAsync(d8, planes, n * ps, cudaMemcpyHostToDevice, st));
    }
    uint16_t *d = c->fin.as<uint16_t>();
    if ((e = launch_s8_to_i16(d8, d, n * ps, st, c->ls))) return e;
    CUDA_TRY(cudaMemsetAsync(c->ovf.p, 0, n * 4, st));
    if (inverse) {
        if ((e = ensure_perms8(c, width, height, stages))) return e;
        e = launch_dwt_inverse(d, c->tmp.as<uint16_t>(), n, width, height, stages, filter, false, st, c->ls, true, &c->perms8);
    } else e = launch_dwt_forward(d, c->tmp.as<uint16_t>(), n, width, height, stages, filter, c->ovf.as<int>(), st, c->ls, true);
    if (e) return e;
    if ((e = launch_i16_to_b8(d, d8, n * ps, st, c->ls))) return e;
    std::vector<int> ovf(n, 0);
    CUDA_TRY(cudaMemcpyAsync(ovf.data(), c->ovf.p, n * 4, cudaMemcpyDeviceToHost, st));
    if (host) CUDA_TRY(cudaMemcpyAsync(planes, d8, n * ps, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (rc) {
        std::vector<int32_t> r(n);
        for (size_t i = 0; i < n; i++) r[i] = ovf[i] ? ICER_INTEGER_OVERFLOW : ICER_RESULT_OK;
        CUDA_TRY(cudaMemcpy(rc, r.data(), n * 4, host ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice));
    }
    return ICER_RESULT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// lib_icer-compatible single-image API (include/icer.h)
// ---------------------------------------------------------------------------------------------------------
static icer_b200_ctx *g_ctx = nullptr;
static std::mutex g_mu;
// lib_icer is non-re-entrant (global work buffers); the compat entry points share ONE default context, so they are
// serialised here instead of corrupting each other's workspace when called from several host threads
static std::mutex g_call_mu;

static icer_b200_ctx *default_ctx() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) {
        if (icer_b200_create(&g_ctx, -1) != ICER_RESULT_OK) g_ctx = nullptr;
        else {
            const char *pl = getenv("ICER_B200_PLANES"), *ms = getenv("ICER_B200_MAX_STAGES");
            icer_b200_set_limits(g_ctx, pl ? atoi(pl) : 0, ms ? atoi(ms) : 0, 0);
        }
    }
    return g_ctx;
}

int icer_init(void) {   // icer_init.c:24-35: table set-up; here it also binds the default device context
    icer_host_tables();
    return default_ctx() ? ICER_RESULT_OK : ICER_FATAL_ERROR;
}

int icer_init_output_struct(icer_output_data_buf_typedef *out, uint8_t *data, size_t buf_len, size_t byte_quota) {
    if (byte_quota * 2 > buf_len) return ICER_OUTPUT_BUF_TOO_SMALL;   // icer_util.c:38-45
    out->size_used = 0;
    out->data_start = data;
    out->size_allocated = byte_quota;
    out->rearrange_start = data + byte_quota;
    return ICER_RESULT_OK;
}

size_t icer_get_dim_n_low_stages(size_t dim, uint8_t stages) { return dim_low(dim, stages); }
size_t icer_get_dim_n_high_stages(size_t dim, uint8_t stages) { return dim_high(dim, stages); }

int icer_generate_partition_parameters(partition_param_typdef *params, size_t ll_w, size_t ll_h, uint16_t segments) {
    int n = partition_rects(ll_w, ll_h, segments, nullptr, nullptr, nullptr, nullptr, params);
    return n < 0 ? n : ICER_RESULT_OK;
}

}  // extern "C"

// T = uint16_t: the *_uint16 entry points; T = uint8_t: the native *_uint8 codec (int8 arithmetic, 7 planes)
template <typename T>
static int compress_common(T *const ch[3], uint32_t chan, size_t w, size_t h, uint8_t stages, int filt, uint8_t segments,
                           icer_output_data_buf_typedef *od) {
    icer_b200_ctx *c = default_ctx();
    if (!c) return ICER_FATAL_ERROR;
    if (!od || !ch[0]) return ICER_INVALID_INPUT;
    std::lock_guard<std::mutex> call_lock(g_call_mu);
    icer_b200_params p;
    p.width = (uint32_t) w; p.height = (uint32_t) h; p.channels = chan; p.stages = stages; p.filter = (uint32_t) filt;
    p.segments = segments; p.byte_quota = od->size_allocated;
    const size_t px = w * h;
    std::vector<T> stage;
    const T *src = ch[0];
    if (chan == 3) {
        stage.resize(px * 3);
        for (int k = 0; k < 3; k++) memcpy(stage.data() + k * px, ch[k], px * sizeof(T));
        src = stage.data();
    }
    std::vector<T> coefs(px * chan);
    uint64_t used = 0;
    int32_t rc = ICER_FATAL_ERROR;
    int e = compress_batch_any(c, &p, src, sizeof(T) == 2 ? ICER_B200_PIX_U16 : ICER_B200_PIX_NATIVE_U8, 1, od->rearrange_start,
                               od->size_allocated, od->data_start, (uint16_t *) coefs.data(), &used, &rc, ICER_B200_MEM_HOST, nullptr);
    if (e != ICER_RESULT_OK) return e;
    if (rc == ICER_RESULT_OK || rc == ICER_BYTE_QUOTA_EXCEEDED) {
        // lib_icer leaves the LL-mean-removed sign-magnitude coefficients in the caller's image buffer
        for (uint32_t k = 0; k < chan; k++) memcpy(ch[k], coefs.data() + k * px, px * sizeof(T));
        od->size_used = (size_t) used;
    }
    return rc;
}

template <typename T>
static int decompress_common(T *const ch[3], uint32_t chan, size_t *image_w, size_t *image_h, size_t bufsize,
                             const uint8_t *data, size_t len, uint8_t stages, int filt, uint8_t segments) {
    icer_b200_ctx *c = default_ctx();
    if (!c) return ICER_FATAL_ERROR;
    if (!data || !image_w || !image_h) return ICER_INVALID_INPUT;
    std::lock_guard<std::mutex> call_lock(g_call_mu);
    // dimensions of the LAST valid header decide the geometry (icer_compress.c:454-455)
    Parsed ps;
    parse_stream(data, len, chan, ps);
    if (!ps.any) return ICER_DECODER_OUT_OF_DATA;
    *image_w = ps.w; *image_h = ps.h;
    if (bufsize < (size_t) ps.w * ps.h) return ICER_BYTE_QUOTA_EXCEEDED;
    icer_b200_params p;
    p.width = ps.w; p.height = ps.h; p.channels = chan; p.stages = stages; p.filter = (uint32_t) filt; p.segments = segments;
    p.byte_quota = 0;
    const size_t px = (size_t) ps.w * ps.h;
    std::vector<T> outbuf(px * chan);
    uint64_t l = len;
    uint32_t dims[2];
    int32_t rc = ICER_FATAL_ERROR;
    int e = decompress_batch_any(c, &p, data, len, &l, 1, outbuf.data(), sizeof(T) == 2 ? ICER_B200_PIX_U16 : ICER_B200_PIX_NATIVE_U8, dims,
                                 &rc, ICER_B200_MEM_HOST, nullptr);
    if (e != ICER_RESULT_OK) return e;
    if (rc == ICER_RESULT_OK)
        for (uint32_t k = 0; k < chan; k++) memcpy(ch[k], outbuf.data() + k * px, px * sizeof(T));
    return rc;
}

extern "C" {

int icer_compress_image_uint16(uint16_t *image, size_t image_w, size_t image_h, uint8_t stages, enum icer_filter_types filt,
                               uint8_t segments, icer_output_data_buf_typedef *output_data) {
    uint16_t *ch[3] = {image, nullptr, nullptr};
    return compress_common<uint16_t>(ch, 1, image_w, image_h, stages, filt, segments, output_data);
}

int icer_compress_image_yuv_uint16(uint16_t *y, uint16_t *u, uint16_t *v, size_t image_w, size_t image_h, uint8_t stages,
                                   enum icer_filter_types filt, uint8_t segments, icer_output_data_buf_typedef *output_data) {
    uint16_t *ch[3] = {y, u, v};
    if (!u || !v) return ICER_INVALID_INPUT;
    return compress_common<uint16_t>(ch, 3, image_w, image_h, stages, filt, segments, output_data);
}

int icer_decompress_image_uint16(uint16_t *image, size_t *image_w, size_t *image_h, size_t image_bufsize, const uint8_t *datastream,
                                 size_t data_length, uint8_t stages, enum icer_filter_types filt, uint8_t segments) {
    uint16_t *ch[3] = {image, nullptr, nullptr};
    if (!image) return ICER_INVALID_INPUT;
    return decompress_common<uint16_t>(ch, 1, image_w, image_h, image_bufsize, datastream, data_length, stages, filt, segments);
}

int icer_decompress_image_yuv_uint16(uint16_t *y, uint16_t *u, uint16_t *v, size_t *image_w, size_t *image_h, size_t image_bufsize,
                                     const uint8_t *datastream, size_t data_length, uint8_t stages, enum icer_filter_types filt,
                                     uint8_t segments) {
    uint16_t *ch[3] = {y, u, v};
    if (!y || !u || !v) return ICER_INVALID_INPUT;
    return decompress_common<uint16_t>(ch, 3, image_w, image_h, image_bufsize, datastream, data_length, stages, filt, segments);
}

// ---- the native uint8 codec (lib_icer/inc/icer.h:386-390, :407-412): int8 wavelet arithmetic, 7 bit planes, 300 packets ----
int icer_compress_image_uint8(uint8_t *image, size_t image_w, size_t image_h, uint8_t stages, enum icer_filter_types filt,
                              uint8_t segments, icer_output_data_buf_typedef *output_data) {
    uint8_t *ch[3] = {image, nullptr, nullptr};
    return compress_common<uint8_t>(ch, 1, image_w, image_h, stages, filt, segments, output_data);
}

int icer_compress_image_yuv_uint8(uint8_t *y, uint8_t *u, uint8_t *v, size_t image_w, size_t image_h, uint8_t stages,
                                  enum icer_filter_types filt, uint8_t segments, icer_output_data_buf_typedef *output_data) {
    uint8_t *ch[3] = {y, u, v};
    if (!u || !v) return ICER_INVALID_INPUT;
    return compress_common<uint8_t>(ch, 3, image_w, image_h, stages, filt, segments, output_data);
}

int icer_decompress_image_uint8(uint8_t *image, size_t *image_w, size_t *image_h, size_t image_bufsize, const uint8_t *datastream,
                                size_t data_length, uint8_t stages, enum icer_filter_types filt, uint8_t segments) {
    uint8_t *ch[3] = {image, nullptr, nullptr};
    if (!image) return ICER_INVALID_INPUT;
    return decompress_common<uint8_t>(ch, 1, image_w, image_h, image_bufsize, datastream, data_length, stages, filt, segments);
}

int icer_decompress_image_yuv_uint8(uint8_t *y, uint8_t *u, uint8_t *v, size_t *image_w, size_t *image_h, size_t image_bufsize,
                                    const uint8_t *datastream, size_t data_length, uint8_t stages, enum icer_filter_types filt,
                                    uint8_t segments) {
    uint8_t *ch[3] = {y, u, v};
    if (!y || !u || !v) return ICER_INVALID_INPUT;
    return decompress_common<uint8_t>(ch, 3, image_w, image_h, image_bufsize, datastream, data_length, stages, filt, segments);
}

int icer_wavelet_transform_stages_uint8(uint8_t *image, size_t image_w, size_t image_h, uint8_t stages, enum icer_filter_types filt) {
    icer_b200_ctx *c = default_ctx();
    if (!c) return ICER_FATAL_ERROR;
    std::lock_guard<std::mutex> call_lock(g_call_mu);
    int32_t rc = ICER_FATAL_ERROR;
    int e = icer_b200_dwt_batch_u8(c, image, 1, (uint32_t) image_w, (uint32_t) image_h, stages, (uint32_t) filt, 0, &rc, ICER_B200_MEM_HOST, nullptr);
    return e != ICER_RESULT_OK ? e : rc;
}

int icer_inverse_wavelet_transform_stages_uint8(uint8_t *image, size_t image_w, size_t image_h, uint8_t stages,
                                                enum icer_filter_types filt) {
    icer_b200_ctx *c = default_ctx();
    if (!c) return ICER_FATAL_ERROR;
    std::lock_guard<std::mutex> call_lock(g_call_mu);
    int32_t rc = ICER_FATAL_ERROR;
    int e = icer_b200_dwt_batch_u8(c, image, 1, (uint32_t) image_w, (uint32_t) image_h, stages, (uint32_t) filt, 1, &rc, ICER_B200_MEM_HOST, nullptr);
    return e != ICER_RESULT_OK ? e : rc;
}

int icer_wavelet_transform_stages_uint16(uint16_t *image, size_t image_w, size_t image_h, uint8_t stages, enum icer_filter_types filt) {
    icer_b200_ctx *c = default_ctx();
    if (!c) return ICER_FATAL_ERROR;
    std::lock_guard<std::mutex> call_lock(g_call_mu);
    int32_t rc = ICER_FATAL_ERROR;
    int e = icer_b200_dwt_batch_u16(c, image, 1, (uint32_t) image_w, (uint32_t) image_h, stages, (uint32_t) filt, 0, &rc, ICER_B200_MEM_HOST, nullptr);
    return e != ICER_RESULT_OK ? e : rc;
}

int icer_inverse_wavelet_transform_stages_uint16(uint16_t *image, size_t image_w, size_t image_h, uint8_t stages,
                                                 enum icer_filter_types filt) {
    icer_b200_ctx *c = default_ctx();
    if (!c) return ICER_FATAL_ERROR;
    std::lock_guard<std::mutex> call_lock(g_call_mu);
    int32_t rc = ICER_FATAL_ERROR;
    int e = icer_b200_dwt_batch_u16(c, image, 1, (uint32_t) image_w, (uint32_t) image_h, stages, (uint32_t) filt, 1, &rc, ICER_B200_MEM_HOST, nullptr);
    // lib_icer reports inverse overflow too (icer_wavelet.c:102) but no caller looks at it; we return OK / TOO_MANY_STAGES
    return e != ICER_RESULT_OK ? e : rc;
}

}  // extern "C"
