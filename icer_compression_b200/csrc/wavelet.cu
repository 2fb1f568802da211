// wavelet.cu -- K1/K2/K3/K8: multi-level reversible integer lifting DWT and the pre/post passes around it.
//
// Reference behaviour (lib_icer/src/icer_wavelet.c): one 1-D pass over a line of n samples produces
// ceil(n/2) lows followed by floor(n/2) highs (:385-465); a stage is every row then every column of the
// current LL region (:155-171); stages shrink the region by ceil(/2) (:57-77).  The reference works in place
// with a cycle-leader unshuffle; here every pass is out of place (src -> dst) between the coefficient buffer and
// a scratch buffer of the same size, so a thread that produces l[k] and h[k] writes them straight to their
// final positions.  All stored values wrap to int16 exactly like the reference's casts; the overflow flag is
// raised under the same conditions (:412-415, :460-461).
#include "common.h"

__constant__ int c_filt[7][4] = {{0, 4, 4, 0}, {0, 4, 6, 4}, {-1, 4, 8, 6}, {0, 4, 5, 2},
                                 {0, 3, 8, 6}, {0, 3, 9, 8}, {0, 4, 4, 4}};   // icer_config.c:18-24

namespace {

// Sample width of the stored values: the uint16 API wraps to int16 (icer_wavelet.c:412-415,460-461), the native uint8 API
// (:216-381) wraps to int8 -- here with the int8 values sign-extended in the same int16 containers.
template <bool B8> __device__ __forceinline__ int wrapv(int v) { return B8 ? (int) (int8_t) v : (int) (int16_t) v; }
template <bool B8> __device__ __forceinline__ bool ovfv(int v) { return B8 ? (v > 127 || v < -128) : (v > 32767 || v < -32768); }

// A line view: element i of line `line` of plane `p`.
template <bool COL>
struct LineView {
    const int16_t *s;
    size_t es;
    __device__ LineView(const int16_t *base, size_t plane_stride, int plane, int line, int W)
        : s(base + (size_t) plane * plane_stride + (COL ? (size_t) line : (size_t) line * W)), es(COL ? (size_t) W : 1) {}
    __device__ int operator()(int i) const { return (int) s[(size_t) i * es]; }
};

// ---- forward pass: thread (line, k) produces l[k] and h[k] --------------------------------------------
template <bool COL, bool B8>
__global__ void __launch_bounds__(256) k_fwd_pass(const int16_t *__restrict__ src, int16_t *__restrict__ dst, int n,
                                                  int nlines, int W, size_t plane_stride, int filt, int plane0,
                                                  int *__restrict__ ovf) {
    const int tx = blockIdx.x * blockDim.x + threadIdx.x, ty = blockIdx.y * blockDim.y + threadIdx.y;
    const int line = COL ? tx : ty, k = COL ? ty : tx;
    const int nl = (n + 1) >> 1, nh = n >> 1;
    if (line >= nlines || k >= nl) return;
    const int plane = plane0 + blockIdx.z;
    const bool odd = n & 1;
    LineView<COL> X(src, plane_stride, plane, line, W);
    int16_t *d = dst + (size_t) plane * plane_stride + (COL ? (size_t) line : (size_t) line * W);
    const size_t es = COL ? (size_t) W : 1;

    // stored (int16-wrapped) low / high of pair j
    auto LO = [&](int j) -> int { return j < nh ? wrapv<B8>((X(2 * j) + X(2 * j + 1)) >> 1) : X(n - 1); };
    auto HI = [&](int j) -> int { return wrapv<B8>(X(2 * j) - X(2 * j + 1)); };
    auto R = [&](int j) -> int { return (int) (int16_t) (LO(j - 1) - LO(j)); };   // get_r_int16 :206-208

    d[(size_t) k * es] = (int16_t) LO(k);
    if (k >= nh) return;
    const int a = X(2 * k), b = X(2 * k + 1);
    const int hi = a - b;
    bool o = ovfv<B8>(hi);
    const int dk = wrapv<B8>(hi);
    const int f0 = c_filt[filt][0], f1 = c_filt[filt][1], f2 = c_filt[filt][2], f3 = c_filt[filt][3];
    int sub;
    if (k == 0) {
        sub = R(1) >> 2;
    } else if (k == 1 && f0 != 0) {                       // filter C quirk :436-440
        const int x = (odd && nl - 1 == 2) ? 0 : HI(1);
        sub = (2 * R(1) + 3 * R(2) - 2 * x + 4) >> 3;
    } else if (!odd && k == nh - 1) {
        sub = R(k) >> 2;
    } else {
        const int rm1 = (f0 != 0) ? ((k - 1 > 0) ? R(k - 1) : 1) : 0;
        const int dn = (odd && k + 1 == nl - 1) ? 0 : HI(k + 1);
        sub = (f0 * rm1 + f1 * R(k) + f2 * R(k + 1) - f3 * dn + 8) >> 4;
    }
    const int hv = dk - sub;
    o |= ovfv<B8>(hv);
    d[(size_t) (nl + k) * es] = (int16_t) wrapv<B8>(hv);
    if (o) atomicOr(&ovf[plane], 1);
}

// ---- inverse pass, data-parallel form (beta == 0 and alpha_-1 == 0: filter A) -----------------------------
// perm (native uint8 API, odd n only): where lib_icer's icer_interleave_uint8 really puts element i of the lows|highs
// array (:570-628; it is NOT the perfect shuffle for odd n); nullptr = the true interleave.
template <bool COL, bool CLAMP, bool B8>
__global__ void __launch_bounds__(256) k_inv_pass_par(const int16_t *__restrict__ src, int16_t *__restrict__ dst, int n,
                                                      int nlines, int W, size_t plane_stride, int filt, int plane0,
                                                      const uint32_t *__restrict__ perm) {
    const int tx = blockIdx.x * blockDim.x + threadIdx.x, ty = blockIdx.y * blockDim.y + threadIdx.y;
    const int line = COL ? tx : ty, k = COL ? ty : tx;
    const int nl = (n + 1) >> 1, nh = n >> 1;
    if (line >= nlines || k >= nl) return;
    const int plane = plane0 + blockIdx.z;
    const bool odd = n & 1;
    LineView<COL> S(src, plane_stride, plane, line, W);
    int16_t *d = dst + (size_t) plane * plane_stride + (COL ? (size_t) line : (size_t) line * W);
    const size_t es = COL ? (size_t) W : 1;
    auto R = [&](int j) -> int { return (int) (int16_t) (S(j - 1) - S(j)); };
    if (k >= nh) {   // odd tail: low + floor(1/2) = low (:537-545)
        int v = S(nl - 1);
        if (CLAMP && v < 0) v = 0;
        d[(size_t) (perm ? (int) perm[nl - 1] : n - 1) * es] = (int16_t) v;
        return;
    }
    const int f1 = c_filt[filt][1], f2 = c_filt[filt][2];
    int add;
    if (k == 0) add = R(1) >> 2;
    else if (!odd && k == nh - 1) add = R(k) >> 2;
    else add = (f1 * R(k) + f2 * R(k + 1) + 8) >> 4;
    const int dk = wrapv<B8>(S(nl + k) + add);
    const int t = S(k) + ((dk + 1) >> 1);
    int v0 = wrapv<B8>(t), v1 = wrapv<B8>(t - dk);
    if (CLAMP) { v0 = v0 < 0 ? 0 : v0; v1 = v1 < 0 ? 0 : v1; }
    d[(size_t) (perm ? (int) perm[k] : 2 * k) * es] = (int16_t) v0;
    d[(size_t) (perm ? (int) perm[nl + k] : 2 * k + 1) * es] = (int16_t) v1;
}

// ---- inverse pass, serial along the line (beta != 0: d[k] needs the restored d[k+1]; :478-508) --------------
template <bool COL, bool CLAMP, bool B8>
__global__ void __launch_bounds__(128) k_inv_pass_ser(const int16_t *__restrict__ src, int16_t *__restrict__ dst, int n,
                                                      int nlines, int W, size_t plane_stride, int filt, int plane0,
                                                      const uint32_t *__restrict__ perm) {
    const int line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= nlines) return;
    const int plane = plane0 + blockIdx.z;
    const int nl = (n + 1) >> 1, nh = n >> 1;
    const bool odd = n & 1;
    LineView<COL> S(src, plane_stride, plane, line, W);
    int16_t *d = dst + (size_t) plane * plane_stride + (COL ? (size_t) line : (size_t) line * W);
    const size_t es = COL ? (size_t) W : 1;
    auto R = [&](int j) -> int { return (int) (int16_t) (S(j - 1) - S(j)); };
    const int f0 = c_filt[filt][0], f1 = c_filt[filt][1], f2 = c_filt[filt][2], f3 = c_filt[filt][3];
    if (odd) {
        int v = S(nl - 1);
        if (CLAMP && v < 0) v = 0;
        d[(size_t) (perm ? (int) perm[nl - 1] : n - 1) * es] = (int16_t) v;
    }
    int dnext = 0;
    for (int k = nh - 1; k >= 0; k--) {
        int add;
        if (k == 0) {
            add = R(1) >> 2;
        } else if (k == 1 && f0 != 0) {
            const int x = (odd && nl - 1 == 2) ? 0 : S(nl + 1);   // still the un-restored h[1] (:488-492)
            add = (2 * R(1) + 3 * R(2) - 2 * x + 4) >> 3;
        } else if (!odd && k == nh - 1) {
            add = R(k) >> 2;
        } else {
            const int rm1 = (f0 != 0) ? ((k - 1 > 0) ? R(k - 1) : 1) : 0;
            const int dn = (odd && k + 1 == nl - 1) ? 0 : dnext;
            add = (f0 * rm1 + f1 * R(k) + f2 * R(k + 1) - f3 * dn + 8) >> 4;
        }
        const int dk = wrapv<B8>(S(nl + k) + add);
        dnext = dk;
        const int t = S(k) + ((dk + 1) >> 1);
        int v0 = wrapv<B8>(t), v1 = wrapv<B8>(t - dk);
        if (CLAMP) { v0 = v0 < 0 ? 0 : v0; v1 = v1 < 0 ? 0 : v1; }
        d[(size_t) (perm ? (int) perm[k] : 2 * k) * es] = (int16_t) v0;
        d[(size_t) (perm ? (int) perm[nl + k] : 2 * k + 1) * es] = (int16_t) v1;
    }
}

// ---- the same recurrence for the uint16 codec with a sliding window of lows in registers (two loads per restored pair instead of
// eight) and, for the ROW pass, the restored samples collected in a shared-memory tile so that they leave as whole row segments (a
// thread per row otherwise stores 2 bytes into 32 different sectors per instruction).  Used for filters B..Q (beta != 0). ----------
template <bool COL, bool CLAMP>
__global__ void __launch_bounds__(64) k_inv_pass_ser2(const int16_t *__restrict__ src, int16_t *__restrict__ dst, int n, int nlines, int W,
                                                     size_t plane_stride, int filt, int plane0, int tile_stride /* samples, row pass only */) {
    extern __shared__ __align__(16) int16_t tile[];          // row pass: [64][tile_stride]
    const int line0 = blockIdx.x * 64, line = line0 + threadIdx.x;
    const int plane = plane0 + blockIdx.z;
    const int nl = (n + 1) >> 1, nh = n >> 1;
    const bool odd = n & 1;
    const bool live = line < nlines;
    const int16_t *s = src + (size_t) plane * plane_stride + (COL ? (size_t) line : (size_t) line * W);
    const size_t es = COL ? (size_t) W : 1;
    int16_t *d = dst + (size_t) plane * plane_stride + (COL ? (size_t) line : (size_t) line * W);
    int16_t *trow = tile + (size_t) threadIdx.x * tile_stride;
    auto put = [&](int pos, int v) {
        if (CLAMP && v < 0) v = 0;
        if (COL) d[(size_t) pos * es] = (int16_t) v; else trow[pos] = (int16_t) v;
    };
    if (live) {
        const int f0 = c_filt[filt][0], f1 = c_filt[filt][1], f2 = c_filt[filt][2], f3 = c_filt[filt][3];
        auto L = [&](int j) -> int { return (j >= 0 && j < nl) ? (int) s[(size_t) j * es] : 0; };
        if (odd) put(n - 1, L(nl - 1));
        int lk1 = L(nh), lk = L(nh - 1), lkm1 = L(nh - 2), lkm2 = L(nh - 3);      // l[k+1], l[k], l[k-1], l[k-2] for k = nh - 1
        int dnext = 0;
        for (int k = nh - 1; k >= 0; k--) {
            const int hk = (int) s[(size_t) (nl + k) * es];
            const int rk = (int) (int16_t) (lkm1 - lk), rk1 = (int) (int16_t) (lk - lk1);       // R(k), R(k+1)
            int add;
            if (k == 0) add = rk1 >> 2;
            else if (k == 1 && f0 != 0) {
                const int x = (odd && nl - 1 == 2) ? 0 : hk;              // still the un-restored h[1] (:488-492)
                add = (2 * rk + 3 * rk1 - 2 * x + 4) >> 3;
            } else if (!odd && k == nh - 1) add = rk >> 2;
            else {
                const int rm1 = (f0 != 0) ? ((k - 1 > 0) ? (int) (int16_t) (lkm2 - lkm1) : 1) : 0;
                const int dn = (odd && k + 1 == nl - 1) ? 0 : dnext;
                add = (f0 * rm1 + f1 * rk + f2 * rk1 - f3 * dn + 8) >> 4;
            }
            const int dk = (int) (int16_t) (hk + add);
            dnext = dk;
            const int t = lk + ((dk + 1) >> 1);
            put(2 * k, (int) (int16_t) t);
            put(2 * k + 1, (int) (int16_t) (t - dk));
            lk1 = lk; lk = lkm1; lkm1 = lkm2; lkm2 = L(k - 3);
        }
    }
    if (!COL) {
        __syncthreads();
        // 64 rows x n samples leave as row segments: consecutive threads write consecutive samples of one row
        const int rows = nlines - line0 < 64 ? nlines - line0 : 64;
        for (int r = 0; r < rows; r++) {
            int16_t *o = dst + (size_t) plane * plane_stride + (size_t) (line0 + r) * W;
            const int16_t *tr = tile + (size_t) r * tile_stride;
            for (int c = threadIdx.x; c < n; c += 64) o[c] = tr[c];
        }
    }
}

// ---- K3: LL mean (unsigned sum, icer_compress.c:289-302), subtract on LL, then sign-magnitude everywhere ----
template <bool B8>
__global__ void __launch_bounds__(256) k_ll_mean(uint16_t *__restrict__ planes, size_t plane_stride, int W, int llw,
                                                 int llh, uint16_t *__restrict__ mean_out, int *__restrict__ ovf) {
    __shared__ unsigned long long red[256];
    __shared__ unsigned s_mean;
    uint16_t *p = planes + (size_t) blockIdx.x * plane_stride;
    const int total = llw * llh;
    unsigned long long sum = 0;
    // the sum runs over the UNSIGNED reinterpretation of the samples (uint16 / uint8: icer_compress.c:289-296, :28-36)
    for (int i = threadIdx.x; i < total; i += blockDim.x) sum += p[(size_t) (i / llw) * W + (i % llw)] & (B8 ? 0xFFu : 0xFFFFu);
    red[threadIdx.x] = sum;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int) threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned m = (unsigned) ((red[0] / (unsigned long long) total) & (B8 ? 0xFFu : 0xFFFFu));   // uintN_t ll_mean = sum / n
        s_mean = m;
        mean_out[blockIdx.x] = (uint16_t) m;
        if (m > (B8 ? 127u : 32767u)) atomicOr(&ovf[blockIdx.x], 2);
    }
    __syncthreads();
    const int m = wrapv<B8>((int) s_mean);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        uint16_t *q = &p[(size_t) (i / llw) * W + (i % llw)];
        *q = (uint16_t) (int16_t) wrapv<B8>((int) (int16_t) *q - m);
    }
}

// B8: the int8 value lives sign-extended in the container; its 7-bit magnitude goes to bits 0-6 and the sign to bit 15, so
// that the coder sees the uint8 codec's samples (sign bit 7, icer_wavelet.c:851-857) in its own uint16 convention.
template <bool B8>
__device__ __forceinline__ uint32_t to_signmag(uint32_t v) {          // icer_wavelet.c:871-877
    const int s = (int) (int16_t) v;
    return s < 0 ? (uint32_t) ((-s) & (B8 ? 0x7F : 0x7FFF)) | 0x8000u : (uint32_t) s;
}
template <bool B8>
__global__ void __launch_bounds__(256) k_to_signmag(uint16_t *__restrict__ data, size_t count) {
    // 8 samples (128 bits) per thread per iteration
    size_t nvec = count / 8;
    uint4 *v = reinterpret_cast<uint4 *>(data);
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t) gridDim.x * blockDim.x) {
        uint4 q = v[i];
        uint32_t *w = reinterpret_cast<uint32_t *>(&q);
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] = to_signmag<B8>(w[j] & 0xFFFF) | (to_signmag<B8>(w[j] >> 16) << 16);
        v[i] = q;
    }
    if (blockIdx.x == 0)
        for (size_t i = nvec * 8 + threadIdx.x; i < count; i += blockDim.x) data[i] = (uint16_t) to_signmag<B8>(data[i]);
}

template <typename F>
int for_plane_chunks(size_t nplanes, F f) {
    for (size_t p0 = 0; p0 < nplanes; p0 += 32768) {
        size_t cnt = nplanes - p0 < 32768 ? nplanes - p0 : 32768;
        f((int) p0, (unsigned) cnt);
    }
    return 0;
}

}  // namespace

int launch_dwt_forward(uint16_t *planes, uint16_t *tmp, size_t nplanes, uint32_t w, uint32_t h, uint32_t stages,
                       uint32_t filt, int *d_ovf, cudaStream_t st, LaunchStats &ls, bool bits8) {
    uint32_t lw = w, lh = h;
    const size_t ps = (size_t) w * h;
    int16_t *A = reinterpret_cast<int16_t *>(planes), *B = reinterpret_cast<int16_t *>(tmp);
    for (uint32_t s = 0; s < stages; s++) {
        for_plane_chunks(nplanes, [&](int p0, unsigned cnt) {
            dim3 blk(64, 4);
            dim3 g1(((lw + 1) / 2 + 63) / 64, (lh + 3) / 4, cnt);
            dim3 g2((lw + 63) / 64, ((lh + 1) / 2 + 3) / 4, cnt);
            if (bits8) {
                k_fwd_pass<false, true><<<g1, blk, 0, st>>>(A, B, (int) lw, (int) lh, (int) w, ps, (int) filt, p0, d_ovf);
                k_fwd_pass<true, true><<<g2, blk, 0, st>>>(B, A, (int) lh, (int) lw, (int) w, ps, (int) filt, p0, d_ovf);
            } else {
                k_fwd_pass<false, false><<<g1, blk, 0, st>>>(A, B, (int) lw, (int) lh, (int) w, ps, (int) filt, p0, d_ovf);
                k_fwd_pass<true, false><<<g2, blk, 0, st>>>(B, A, (int) lh, (int) lw, (int) w, ps, (int) filt, p0, d_ovf);
            }
            ls.launches += 2;
        });
        lw = (lw + 1) / 2; lh = (lh + 1) / 2;
    }
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}

namespace {
template <bool B8>
void inverse_stage(int16_t *A, int16_t *B, int p0, unsigned cnt, uint32_t lw, uint32_t lh, uint32_t w, size_t ps, uint32_t filt,
                   bool parallel, bool clamp, const uint32_t *perm_col, const uint32_t *perm_row, cudaStream_t st) {
    if (parallel) {
        dim3 blk(64, 4);
        dim3 g1((lw + 63) / 64, ((lh + 1) / 2 + 3) / 4, cnt);
        k_inv_pass_par<true, false, B8><<<g1, blk, 0, st>>>(A, B, (int) lh, (int) lw, (int) w, ps, (int) filt, p0, perm_col);
        dim3 g2(((lw + 1) / 2 + 63) / 64, (lh + 3) / 4, cnt);
        if (clamp) k_inv_pass_par<false, true, B8><<<g2, blk, 0, st>>>(B, A, (int) lw, (int) lh, (int) w, ps, (int) filt, p0, perm_row);
        else k_inv_pass_par<false, false, B8><<<g2, blk, 0, st>>>(B, A, (int) lw, (int) lh, (int) w, ps, (int) filt, p0, perm_row);
    } else if (!B8 && (size_t) 64 * ((((lw + 1) / 2) | 1) * 2) * 2 <= 200 * 1024) {
        // uint16 codec: sliding-window recurrence; the row pass stages its output rows in shared memory (odd word stride: no bank conflicts)
        dim3 g1((lw + 63) / 64, 1, cnt);
        k_inv_pass_ser2<true, false><<<g1, 64, 0, st>>>(A, B, (int) lh, (int) lw, (int) w, ps, (int) filt, p0, 0);
        const int ts = (int) ((((lw + 1) / 2) | 1) * 2);               // samples per tile row: an odd number of 32-bit words (>= lw)
        const size_t smem = (size_t) 64 * ts * 2;
        dim3 g2((lh + 63) / 64, 1, cnt);
        if (clamp) {
            cudaFuncSetAttribute(k_inv_pass_ser2<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            k_inv_pass_ser2<false, true><<<g2, 64, smem, st>>>(B, A, (int) lw, (int) lh, (int) w, ps, (int) filt, p0, ts);
        } else {
            cudaFuncSetAttribute(k_inv_pass_ser2<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            k_inv_pass_ser2<false, false><<<g2, 64, smem, st>>>(B, A, (int) lw, (int) lh, (int) w, ps, (int) filt, p0, ts);
        }
    } else {
        dim3 g1((lw + 127) / 128, 1, cnt);
        k_inv_pass_ser<true, false, B8><<<g1, 128, 0, st>>>(A, B, (int) lh, (int) lw, (int) w, ps, (int) filt, p0, perm_col);
        dim3 g2((lh + 127) / 128, 1, cnt);
        if (clamp) k_inv_pass_ser<false, true, B8><<<g2, 128, 0, st>>>(B, A, (int) lw, (int) lh, (int) w, ps, (int) filt, p0, perm_row);
        else k_inv_pass_ser<false, false, B8><<<g2, 128, 0, st>>>(B, A, (int) lw, (int) lh, (int) w, ps, (int) filt, p0, perm_row);
    }
}
}  // namespace

// perms (bits8 only): device tables of icer_interleave_uint8's real placement for the ODD line lengths that occur
int launch_dwt_inverse(uint16_t *planes, uint16_t *tmp, size_t nplanes, uint32_t w, uint32_t h, uint32_t stages,
                       uint32_t filt, bool clamp_negative, cudaStream_t st, LaunchStats &ls, bool bits8,
                       const std::map<uint32_t, const uint32_t *> *perms) {
    const size_t ps = (size_t) w * h;
    int16_t *A = reinterpret_cast<int16_t *>(planes), *B = reinterpret_cast<int16_t *>(tmp);
    const bool parallel = (filt == ICER_FILTER_A);
    auto perm_of = [&](uint32_t n) -> const uint32_t * {
        if (!bits8 || !perms || !(n & 1)) return nullptr;
        auto it = perms->find(n);
        return it == perms->end() ? nullptr : it->second;
    };
    for (int s = (int) stages - 1; s >= 0; s--) {
        const uint32_t lw = (uint32_t) dim_low(w, s), lh = (uint32_t) dim_low(h, s);
        const bool clamp = clamp_negative && s == 0;
        for_plane_chunks(nplanes, [&](int p0, unsigned cnt) {
            if (bits8) inverse_stage<true>(A, B, p0, cnt, lw, lh, w, ps, filt, parallel, clamp, perm_of(lh), perm_of(lw), st);
            else inverse_stage<false>(A, B, p0, cnt, lw, lh, w, ps, filt, parallel, clamp, nullptr, nullptr, st);
            ls.launches += 2;
        });
    }
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}

int launch_ll_mean_signmag(uint16_t *planes, size_t nplanes, uint32_t w, uint32_t h, uint32_t llw, uint32_t llh,
                           uint16_t *d_mean, int *d_ovf, cudaStream_t st, LaunchStats &ls, bool bits8, bool skip_signmag) {
    const size_t ps = (size_t) w * h;
    const size_t count = ps * nplanes;
    unsigned blocks = (unsigned) ((count / 8 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks == 0) blocks = 1;
    if (bits8) {
        k_ll_mean<true><<<(unsigned) nplanes, 256, 0, st>>>(planes, ps, (int) w, (int) llw, (int) llh, d_mean, d_ovf);
        k_to_signmag<true><<<blocks, 256, 0, st>>>(planes, count);
    } else {
        k_ll_mean<false><<<(unsigned) nplanes, 256, 0, st>>>(planes, ps, (int) w, (int) llw, (int) llh, d_mean, d_ovf);
        if (!skip_signmag) k_to_signmag<false><<<blocks, 256, 0, st>>>(planes, count);
    }
    ls.launches += skip_signmag ? 1 : 2;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
