// wavelet_tma.cu -- K1 fused: one forward DWT stage (row lifting pass + column lifting pass) per launch, the input
// tile (with the 4-sample halo the lifting predictor needs on every side) staged into shared memory by TMA
// (cp.async.bulk.tensor.3d, mbarrier completion); out-of-range tile elements are zero-filled by the TMA unit and never
// used (the 1-D edge rules only read samples inside the region).
//
// Compared with the two separate passes of wavelet.cu this reads and writes every sample of the shrinking LL region
// once per stage instead of twice: 2·S·(region) bytes per stage -> (8/3)·P·S for all stages, and stage 1 reads the
// caller's frames directly (no working copy).  Each stage writes its three high subbands straight to their final
// place in the coefficient buffer and its LL quadrant to a ping-pong plane that the next stage reads.
//
// Arithmetic: identical to k_fwd_pass (icer_wavelet.c:385-465): stored values wrap to int16, the overflow flag is raised
// under the same conditions; the intermediate row-pass result is kept as int16 in shared memory exactly as the
// reference keeps it in the image between its row and column loops (:155-171).
//
// Requirements of the TMA path (else the caller falls back to the two-pass kernels): frame width a multiple of 8
// samples (16-byte global row stride) and 16-byte aligned plane bases.
#include "common.h"
#include <cuda.h>

namespace {

constexpr int TK = 32, TR = 32;                 // output tile: TK x TR positions of each of the four subbands
// input tile incl. halo: 4 samples are needed on every side; the innermost TMA coordinate has to be 16-byte aligned
// (8 samples), so the tile starts 8 samples left of the first pair and is 80 samples (160 bytes) wide
constexpr int HALO_X = 8, IN_W = 2 * TK + 2 * HALO_X, IN_H = 2 * TR + 8;

__constant__ int c_filt2[7][4] = {{0, 4, 4, 0}, {0, 4, 6, 4}, {-1, 4, 8, 6}, {0, 4, 5, 2},
                                  {0, 3, 8, 6}, {0, 3, 9, 8}, {0, 4, 4, 4}};

// one lifting output pair of a line of n samples: low l[k] (if k < ceil(n/2)) and high h[k] (if k < n/2).
// X(i) returns sample i of the line (int16 value); only in-range samples are requested.
template <class Acc>
__device__ __forceinline__ void lift_pair(const Acc &X, int n, int k, int f0, int f1, int f2, int f3, int &lo, int &hi, bool &ovf) {
    const int nl = (n + 1) >> 1, nh = n >> 1;
    const bool odd = n & 1;
    auto LO = [&](int j) -> int { return j < nh ? (int) (int16_t) ((X(2 * j) + X(2 * j + 1)) >> 1) : X(n - 1); };
    auto HI = [&](int j) -> int { return (int) (int16_t) (X(2 * j) - X(2 * j + 1)); };
    auto R = [&](int j) -> int { return (int) (int16_t) (LO(j - 1) - LO(j)); };
    lo = 0; hi = 0;
    if (k < nl) lo = LO(k);
    if (k >= nh) return;
    const int d = X(2 * k) - X(2 * k + 1);
    ovf |= d > 32767 || d < -32768;
    const int dk = (int) (int16_t) d;
    int sub;
    if (k == 0) sub = R(1) >> 2;
    else if (k == 1 && f0 != 0) {
        const int x = (odd && nl - 1 == 2) ? 0 : HI(1);
        sub = (2 * R(1) + 3 * R(2) - 2 * x + 4) >> 3;
    } else if (!odd && k == nh - 1) sub = R(k) >> 2;
    else {
        const int rm1 = (f0 != 0) ? ((k - 1 > 0) ? R(k - 1) : 1) : 0;
        const int dn = (odd && k + 1 == nl - 1) ? 0 : HI(k + 1);
        sub = (f0 * rm1 + f1 * R(k) + f2 * R(k + 1) - f3 * dn + 8) >> 4;
    }
    const int hv = dk - sub;
    ovf |= hv > 32767 || hv < -32768;
    hi = (int) (int16_t) hv;
}

struct StageArgs {
    int lw, lh, W;                 // region size, frame row stride (samples)
    size_t plane_stride;           // samples between planes in the destination buffers
    int16_t *hi_dst;               // final coefficient buffer (HL, LH, HH of this stage)
    int16_t *ll_dst;               // where this stage's LL quadrant goes (ping-pong plane, or the final buffer at the last stage)
    int filt;
    int *ovf;                      // per plane
};

__global__ void __launch_bounds__(256) k_fwd_stage_tma(const __grid_constant__ CUtensorMap tmap, StageArgs a) {
    __shared__ __align__(128) int16_t in[IN_H][IN_W];          // TMA destination
    __shared__ __align__(16) int16_t rowl[IN_H][TK], rowh[IN_H][TK];   // row-pass result of every tile row
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * TK, r0 = blockIdx.y * TR, plane = blockIdx.z;
    const int x0 = 2 * k0 - HALO_X, y0 = 2 * r0 - 4;
    const uint32_t mbar_a = (uint32_t) __cvta_generic_to_shared(&mbar);
    const uint32_t in_a = (uint32_t) __cvta_generic_to_shared(&in[0][0]);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"((uint32_t) sizeof(in)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(in_a), "l"(&tmap), "r"(x0), "r"(y0), "r"(plane), "r"(mbar_a)
                     : "memory");
    }
    {   // wait for the tile
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mbar_a) : "memory");
    }
    const int f0 = c_filt2[a.filt][0], f1 = c_filt2[a.filt][1], f2 = c_filt2[a.filt][2], f3 = c_filt2[a.filt][3];
    bool ovf = false;
    // ---- row pass over every tile row that lies inside the region --------------------------------------------
    for (int i = tid; i < IN_H * TK; i += 256) {
        const int yy = i / TK, kk = i % TK;
        const int gy = y0 + yy, k = k0 + kk;
        int lo = 0, hi = 0;
        if (gy >= 0 && gy < a.lh) {
            auto X = [&](int xi) -> int { return (int) in[yy][xi - x0]; };
            lift_pair(X, a.lw, k, f0, f1, f2, f3, lo, hi, ovf);
        }
        rowl[yy][kk] = (int16_t) lo;
        rowh[yy][kk] = (int16_t) hi;
    }
    __syncthreads();
    // ---- column pass: 2*TK columns (TK of the low half, TK of the high half) x TR output rows -----------------
    const int nlx = (a.lw + 1) >> 1, nhx = a.lw >> 1, nly = (a.lh + 1) >> 1, nhy = a.lh >> 1;
    int16_t *hi_p = a.hi_dst + (size_t) plane * a.plane_stride;
    int16_t *ll_p = a.ll_dst + (size_t) plane * a.plane_stride;
    for (int i = tid; i < 2 * TK * TR; i += 256) {
        const int kk = i % TK, half = (i / TK) & 1, rr = i / (2 * TK);
        const int k = k0 + kk, r = r0 + rr;
        if (half == 0 ? k >= nlx : k >= nhx) continue;
        if (r >= nly) continue;
        const int16_t (*col)[TK] = half == 0 ? rowl : rowh;
        auto Y = [&](int yi) -> int { return (int) col[yi - y0][kk]; };
        int lo, hi;
        lift_pair(Y, a.lh, r, f0, f1, f2, f3, lo, hi, ovf);
        const int xo = half == 0 ? k : nlx + k;
        if (half == 0) ll_p[(size_t) r * a.W + xo] = (int16_t) lo;          // LL
        else hi_p[(size_t) r * a.W + xo] = (int16_t) lo;                     // HL
        if (r < nhy) hi_p[(size_t) (nly + r) * a.W + xo] = (int16_t) hi;     // LH / HH
    }
    if (ovf) atomicOr(&a.ovf[plane], 1);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn) p;
    }
    return fn;
}

}  // namespace

bool dwt_fused_applicable(const void *src, uint32_t w, uint32_t h, size_t src_plane_stride) {
    return get_encode_fn() != nullptr && (w % 8) == 0 && ((uintptr_t) src % 16) == 0 && ((src_plane_stride * 2) % 16) == 0 && w >= 8 && h >= 1;
}

// src: the frames to transform ([nplanes] planes of w*h, plane stride src_plane_stride samples) -- NOT modified.
// dst: final coefficient buffer (same geometry, plane stride w*h); pp0 / pp1: ping-pong planes for the LL quadrants.
int launch_dwt_forward_fused(const uint16_t *src, size_t src_plane_stride, uint16_t *dst, uint16_t *pp0, uint16_t *pp1, size_t nplanes,
                             uint32_t w, uint32_t h, uint32_t stages, uint32_t filt, int *d_ovf, cudaStream_t st, LaunchStats &ls) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return ICER_FATAL_ERROR;
    const size_t ps = (size_t) w * h;
    uint32_t lw = w, lh = h;
    const uint16_t *cur = src;
    size_t cur_stride = src_plane_stride;
    for (uint32_t s = 0; s < stages; s++) {
        const bool last = s + 1 == stages;
        uint16_t *ll = last ? dst : ((s & 1) ? pp1 : pp0);
        for (size_t p0 = 0; p0 < nplanes; p0 += 32768) {
            const size_t cnt = nplanes - p0 < 32768 ? nplanes - p0 : 32768;
            alignas(64) CUtensorMap tm;
            cuuint64_t gdim[3] = {lw, lh, cnt};
            cuuint64_t gstr[2] = {(cuuint64_t) w * 2, (cuuint64_t) cur_stride * 2};
            cuuint32_t box[3] = {IN_W, IN_H, 1};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, (void *) (cur + p0 * cur_stride), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                fprintf(stderr, "[icer_b200] cuTensorMapEncodeTiled failed (%d)\n", (int) r);
                return ICER_FATAL_ERROR;
            }
            StageArgs a;
            a.lw = (int) lw; a.lh = (int) lh; a.W = (int) w; a.plane_stride = ps;
            a.hi_dst = reinterpret_cast<int16_t *>(dst) + p0 * ps;
            a.ll_dst = reinterpret_cast<int16_t *>(ll) + p0 * ps;
            a.filt = (int) filt; a.ovf = d_ovf + p0;
            dim3 grid(((lw + 1) / 2 + TK - 1) / TK, ((lh + 1) / 2 + TR - 1) / TR, (unsigned) cnt);
            k_fwd_stage_tma<<<grid, 256, 0, st>>>(tm, a);
            ls.launches++;
        }
        cur = ll; cur_stride = ps;
        lw = (lw + 1) / 2; lh = (lh + 1) / 2;
    }
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
