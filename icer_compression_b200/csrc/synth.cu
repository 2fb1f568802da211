// synth.cu -- deterministic synthetic frame generator as a device kernel, so that bench.py can fill
// HBM with the benchmark workload without touching the test oracle.  One thread per frame (the per-frame xorshift32
// chain is sequential by definition); tests check it against oracle/icer_oracle.c:orc_synth_frame.
#include "common.h"

namespace {
__device__ __forceinline__ int tri(long long t, int p) {
    long long m = t % (2LL * p);
    if (m < 0) m += 2LL * p;
    m -= p;
    return (int) (m < 0 ? -m : m);
}

__global__ void k_synth(uint16_t *dst, size_t n, uint32_t w, uint32_t h, uint32_t first_frame, uint32_t seed, int kind) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t frame = first_frame + (uint32_t) i;
    uint32_t st = seed ^ (frame * 0x9E3779B9u);
    if (st == 0) st = 0x1234567u;
    uint16_t *o = dst + i * (size_t) w * h;
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            st ^= st << 13; st ^= st >> 17; st ^= st << 5;
            const uint32_t d = st;
            uint32_t v8;
            if (kind & 1) v8 = d & 0xFF;
            else v8 = 16u + (uint32_t) ((tri((long long) x + 2LL * y + 13LL * frame, 256) * 160) >> 8)
                          + (uint32_t) tri(3LL * x - (long long) y + 4096, 32) + d % 13u;
            if (kind & 2) {
                st ^= st << 13; st ^= st >> 17; st ^= st << 5;
                o[(size_t) y * w + x] = (uint16_t) (v8 * 16u + (st & 15u));
            } else {
                o[(size_t) y * w + x] = (uint16_t) v8;
            }
        }
}
}  // namespace

extern "C" int icer_b200_synth_frames_u16(icer_b200_ctx *ctx, uint16_t *dst_device, size_t n, uint32_t w, uint32_t h,
                                          uint32_t first_frame, uint32_t seed, int kind, void *stream) {
    if (!dst_device || n == 0) return ICER_INVALID_INPUT;
    if (ctx) CUDA_TRY(cudaSetDevice(ctx_device(ctx)));       // the stream belongs to the context's device (a process may hold several)
    k_synth<<<(unsigned) ((n + 31) / 32), 32, 0, (cudaStream_t) stream>>>(dst_device, n, w, h, first_frame, seed, kind);
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
