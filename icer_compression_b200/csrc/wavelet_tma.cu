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

// ---- the lifting step on values that are already in registers ---------------------------------------------------------
// One 1-D forward pass over a line of n samples yields lows l[j] = floor((x[2j] + x[2j+1]) / 2) (odd n: the last low is the last
// sample) and highs h[k] = d[k] - predictor(l[k-2..k+1], d[k+1]) with d[k] = x[2k] - x[2k+1] (icer_wavelet.c:385-465).  Every thread
// computes the low / difference of ONE pair exactly once; the predictor's neighbours travel by warp shuffle (row pass: lanes =
// consecutive pairs of a row) or sit in a sliding register window (column pass: a thread walks down its column).
struct Filt { int f0, f1, f2, f3; };

// stored low and difference of pair j from its two samples (a = x[2j], b = x[2j+1]); `last_odd`: j == n/2 of an odd line
__device__ __forceinline__ void pair_ld(int a, int b, bool last_odd, int &lo, int &d) {
    lo = last_odd ? a : (int) (int16_t) ((a + b) >> 1);
    d = a - b;
}

// high output of pair k.  lm2, lm1, l0, lp1 = stored lows l[k-2], l[k-1], l[k], l[k+1]; d0 = unwrapped difference of pair k;
// dp1 = stored (int16) difference of pair k+1.  Returns the stored high; ORs the overflow condition of the reference into ovf.
// INTERIOR: the caller guarantees 2 <= k < nh - 1 (and k + 1 < nl - 1), so none of the edge rules applies; F is a compile-time
// filter (the multiplications fold into shifts and the unused taps disappear).
__device__ __forceinline__ bool out_of_i16(int v) { return (unsigned) (v + 32768) > 65535u; }
template <bool INTERIOR, typename FT>
__device__ __forceinline__ int lift_high(const FT &F, int n, int k, int lm2, int lm1, int l0, int lp1, int d0, int dp1, bool &ovf) {
    ovf |= out_of_i16(d0);
    const int dk = (int) (int16_t) d0;
    const int rk = (int) (int16_t) (lm1 - l0), rk1 = (int) (int16_t) (l0 - lp1);       // r[k], r[k+1] (get_r_int16 :206-208)
    int sub;
    if (INTERIOR) {
        const int rm1 = (F.f0 != 0) ? (int) (int16_t) (lm2 - lm1) : 0;
        sub = (F.f0 * rm1 + F.f1 * rk + F.f2 * rk1 - F.f3 * dp1 + 8) >> 4;
    } else {
        const int nl = (n + 1) >> 1, nh = n >> 1;
        const bool odd = n & 1;
        if (k == 0) sub = rk1 >> 2;
        else if (k == 1 && F.f0 != 0) {                                                       // filter C quirk :436-440
            const int x = (odd && nl - 1 == 2) ? 0 : dk;
            sub = (2 * rk + 3 * rk1 - 2 * x + 4) >> 3;
        } else if (!odd && k == nh - 1) sub = rk >> 2;
        else {
            const int rm1 = (F.f0 != 0) ? ((k - 1 > 0) ? (int) (int16_t) (lm2 - lm1) : 1) : 0;
            const int dn = (odd && k + 1 == nl - 1) ? 0 : dp1;
            sub = (F.f0 * rm1 + F.f1 * rk + F.f2 * rk1 - F.f3 * dn + 8) >> 4;
        }
    }
    const int hv = dk - sub;
    ovf |= out_of_i16(hv);
    return (int) (int16_t) hv;
}
// the same with the filter as compile-time constants: filter A (0, 4, 4, 0) is the default of every tool and benchmark
struct FiltA { static constexpr int f0 = 0, f1 = 4, f2 = 4, f3 = 0; };

struct StageArgs {
    int lw, lh, W;                 // region size, frame row stride (samples)
    size_t plane_stride;           // samples between planes in the destination buffers
    int16_t *hi_dst;               // final coefficient buffer (HL, LH, HH of this stage)
    int16_t *ll_dst;               // where this stage's LL quadrant goes (ping-pong plane, or the final buffer at the last stage)
    int filt;
    int *ovf;                      // per plane
};

// ---- row pass: one warp per tile row, lane = pair.  Each lane reads its two samples with ONE 32-bit load; three lanes also
// evaluate the halo pairs k0-2, k0-1 and k0+TK; all neighbours then come from shuffles.
template <bool INTERIOR, typename FT>
__device__ __forceinline__ void fwd_row_pass(const FT &F, const int16_t (*in)[IN_W], int16_t (*rowl)[TK], int16_t (*rowh)[TK], int lw, int lh,
                                             int k0, int y0, int lane, int warp, bool &ovf) {
    const int nhx = lw >> 1;
    const int k = k0 + lane;
    const int ek = lane == 0 ? k0 - 2 : (lane == 1 ? k0 - 1 : k0 + TK);      // the halo pair of lanes 0, 1, 2
    for (int yy = warp; yy < IN_H; yy += 8) {
        const int gy = y0 + yy;
        if (gy < 0 || gy >= lh) { rowl[yy][lane] = 0; rowh[yy][lane] = 0; continue; }    // warp-uniform; outside the region
        const uint32_t *rw = reinterpret_cast<const uint32_t *>(&in[yy][0]);
        const uint32_t v = rw[(HALO_X >> 1) + lane];
        int l0, d0;
        pair_ld((int) (int16_t) (v & 0xFFFF), (int) (int16_t) (v >> 16), !INTERIOR && (lw & 1) && k == nhx, l0, d0);
        int le = 0, de = 0;
        if (lane < 3) {
            const uint32_t u = rw[(HALO_X >> 1) + (ek - k0)];
            pair_ld((int) (int16_t) (u & 0xFFFF), (int) (int16_t) (u >> 16), !INTERIOR && (lw & 1) && ek == nhx, le, de);
        }
        const int lm2h = __shfl_sync(0xffffffffu, le, 0), lm1h = __shfl_sync(0xffffffffu, le, 1);
        const int lph = __shfl_sync(0xffffffffu, le, 2), dph = __shfl_sync(0xffffffffu, de, 2);
        int lm1 = __shfl_up_sync(0xffffffffu, l0, 1), lm2 = __shfl_up_sync(0xffffffffu, l0, 2);
        int lp1 = __shfl_down_sync(0xffffffffu, l0, 1), dp1 = __shfl_down_sync(0xffffffffu, d0, 1);
        if (lane == 0) { lm1 = lm1h; lm2 = lm2h; }
        if (lane == 1) lm2 = lm1h;
        if (lane == 31) { lp1 = lph; dp1 = dph; }
        int hi = 0;
        if (INTERIOR || k < nhx) hi = lift_high<INTERIOR>(F, lw, k, lm2, lm1, l0, lp1, d0, (int) (int16_t) dp1, ovf);
        rowl[yy][lane] = (int16_t) l0;      // only read back for k < nlx
        rowh[yy][lane] = (int16_t) hi;
    }
}

// ---- column pass: warp = (half, strip of 8 output rows), lane = column; a thread walks down its column with the lows / differences
// of four consecutive vertical pairs in registers
template <bool INTERIOR, typename FT>
__device__ __forceinline__ void fwd_col_pass(const FT &F, const int16_t (*rowl)[TK], const int16_t (*rowh)[TK], const StageArgs &a, int k0, int r0, int y0,
                                             int plane, int lane, int warp, bool &ovf) {
    const int nlx = (a.lw + 1) >> 1, nhx = a.lw >> 1, nly = (a.lh + 1) >> 1, nhy = a.lh >> 1;
    const int half = warp & 1, strip = warp >> 1;
    const int k = k0 + lane;
    const bool col_ok = half == 0 ? k < nlx : k < nhx;
    const int16_t (*col)[TK] = half == 0 ? rowl : rowh;
    int16_t *hi_p = a.hi_dst + (size_t) plane * a.plane_stride;
    int16_t *ll_p = a.ll_dst + (size_t) plane * a.plane_stride;
    const int xo = half == 0 ? k : nlx + k;
    const int rs = r0 + strip * 8;
    if (!(col_ok && (INTERIOR || rs < nly))) return;
    auto ld_pair = [&](int r, int &lo, int &d) {
        // pairs outside [0, nly) are never used by lift_high (its edge rules only read in-range pairs)
        const int yi = 2 * r - y0;
        if (!INTERIOR && (r < 0 || r >= nly)) { lo = 0; d = 0; return; }
        pair_ld((int) col[yi][lane], (int) col[yi + 1][lane], !INTERIOR && (a.lh & 1) && r == nhy, lo, d);
    };
    int lm2, lm1, l0, lp1, dm, d0, dp1;
    ld_pair(rs - 2, lm2, dm);
    ld_pair(rs - 1, lm1, dm);
    ld_pair(rs, l0, d0);
    int16_t *dst_lo = (half == 0 ? ll_p : hi_p) + (size_t) rs * a.W + xo;                       // LL  /  HL
    int16_t *dst_hi = hi_p + (size_t) (nly + rs) * a.W + xo;                                    // LH  /  HH
#pragma unroll
    for (int i = 0; i < 8; i++, dst_lo += a.W, dst_hi += a.W) {
        const int r = rs + i;
        ld_pair(r + 1, lp1, dp1);
        if (INTERIOR || r < nly) {
            *dst_lo = (int16_t) l0;
            if (INTERIOR || r < nhy) *dst_hi = (int16_t) lift_high<INTERIOR>(F, a.lh, r, lm2, lm1, l0, lp1, d0, (int) (int16_t) dp1, ovf);
        }
        lm2 = lm1; lm1 = l0; l0 = lp1; d0 = dp1;
    }
}

template <bool FA>
__global__ void __launch_bounds__(256) k_fwd_stage_tma(const __grid_constant__ CUtensorMap tmap, StageArgs a) {
    __shared__ __align__(128) int16_t in[IN_H][IN_W];          // TMA destination
    __shared__ __align__(16) int16_t rowl[IN_H][TK], rowh[IN_H][TK];   // row-pass result of every tile row
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    Filt F;
    F.f0 = c_filt2[a.filt][0]; F.f1 = c_filt2[a.filt][1]; F.f2 = c_filt2[a.filt][2]; F.f3 = c_filt2[a.filt][3];
    const FiltA FAc;
    bool ovf = false;
    const int nhx = a.lw >> 1, nhy = a.lh >> 1;
    // a tile none of whose pairs meets an edge rule of the lifting step (first two / last pair of a line, the odd tail) takes the
    // branch-free interior form of the predictor: 6 of 8 tile columns / rows of a 512-wide region
    const bool int_x = k0 >= 2 && k0 + TK - 1 <= nhx - 2;
    const bool int_y = r0 >= 2 && r0 + TR - 1 <= nhy - 2;
    if (FA) {
        if (int_x) fwd_row_pass<true>(FAc, in, rowl, rowh, a.lw, a.lh, k0, y0, lane, warp, ovf);
        else fwd_row_pass<false>(FAc, in, rowl, rowh, a.lw, a.lh, k0, y0, lane, warp, ovf);
    } else {
        if (int_x) fwd_row_pass<true>(F, in, rowl, rowh, a.lw, a.lh, k0, y0, lane, warp, ovf);
        else fwd_row_pass<false>(F, in, rowl, rowh, a.lw, a.lh, k0, y0, lane, warp, ovf);
    }
    __syncthreads();
    if (FA) {
        if (int_y) fwd_col_pass<true>(FAc, rowl, rowh, a, k0, r0, y0, plane, lane, warp, ovf);
        else fwd_col_pass<false>(FAc, rowl, rowh, a, k0, r0, y0, plane, lane, warp, ovf);
    } else {
        if (int_y) fwd_col_pass<true>(F, rowl, rowh, a, k0, r0, y0, plane, lane, warp, ovf);
        else fwd_col_pass<false>(F, rowl, rowh, a, k0, r0, y0, plane, lane, warp, ovf);
    }
    if (ovf) atomicOr(&a.ovf[plane], 1);
}

// =========================================================================================================================
// K2 fused: one INVERSE stage per launch for filters without the beta term (filter A): column inverse then row inverse
// (icer_wavelet.c:175-191, :467-550) of a 64 x 64 output tile.  The four sub-band tiles the tile depends on (LL from the previous,
// coarser stage's output; HL, LH, HH from the coefficient buffer) arrive by four TMA box loads on one mbarrier; the column pass
// runs smem -> smem, the row pass writes 128-bit vectors of 8 reconstructed samples.
//   d[k] = h[k] + predictor(l[k-1], l[k], l[k+1]);  t = l[k] + floor((d[k] + 1) / 2);  x[2k] = t;  x[2k+1] = t - d[k]
// =========================================================================================================================
constexpr int IA_W = 48, IA_H = TR + 2;        // low-x boxes: columns k0-8 .. k0+39 (needs k0-1 .. k0+32); top boxes: rows r0-1 .. r0+32
constexpr int CR_W = 2 * TK + 2;               // column-pass result per output row: 34 low-x columns (k0-1 .. k0+32) then 32 high-x columns

struct InvArgs {
    int lw, lh, W;
    size_t plane_stride;
    int16_t *dst;                  // reconstructed region (row stride W)
    int f1, f2;
    int clamp;                     // last stage of the decoder: negative samples become 0 (icer_util.c:82-90)
};

__device__ __forceinline__ int inv_d(int f1, int f2, int n, int k, int lm1, int l0, int lp1, int h) {
    const int nh = n >> 1;
    const int rk = (int) (int16_t) (lm1 - l0), rk1 = (int) (int16_t) (l0 - lp1);
    int add;
    if (k == 0) add = rk1 >> 2;
    else if (!(n & 1) && k == nh - 1) add = rk >> 2;
    else add = (f1 * rk + f2 * rk1 + 8) >> 4;
    return (int) (int16_t) (h + add);
}

__global__ void __launch_bounds__(288) k_inv_stage_tma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                       const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapD,
                                                       InvArgs a) {
    __shared__ __align__(128) int16_t sA[IA_H][IA_W];      // LL   rows r0-1.., columns k0-8..
    __shared__ __align__(128) int16_t sB[IA_H][TK];        // HL   rows r0-1.., columns nlx+k0..
    __shared__ __align__(128) int16_t sC[TR][IA_W];        // LH   rows nly+r0.., columns k0-8..
    __shared__ __align__(128) int16_t sD[TR][TK];          // HH
    __shared__ __align__(16) int16_t cres[2 * TR][CR_W];   // after the column pass: 64 rows x (34 low-x | 32 high-x) columns
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k0 = blockIdx.x * TK, r0 = blockIdx.y * TR, plane = blockIdx.z;
    const int nlx = (a.lw + 1) >> 1, nhx = a.lw >> 1, nly = (a.lh + 1) >> 1, nhy = a.lh >> 1;
    const uint32_t mbar_a = (uint32_t) __cvta_generic_to_shared(&mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a),
                     "r"((uint32_t) (sizeof(sA) + sizeof(sB) + sizeof(sC) + sizeof(sD))) : "memory");
        auto tma = [&](const CUtensorMap *m, void *dst, int x, int y) {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"((uint32_t) __cvta_generic_to_shared(dst)), "l"(m), "r"(x), "r"(y), "r"(plane), "r"(mbar_a) : "memory");
        };
        tma(&mapA, &sA[0][0], k0 - 8, r0 - 1);
        tma(&mapB, &sB[0][0], nlx + k0, r0 - 1);
        tma(&mapC, &sC[0][0], k0 - 8, nly + r0);
        tma(&mapD, &sD[0][0], nlx + k0, nly + r0);
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mbar_a) : "memory");
    }
    const int f1 = a.f1, f2 = a.f2;
    // ---- column pass.  Vertical pair r of a column: lows l[r-1], l[r], l[r+1] from the top box (tile row r - r0 + 1), high from the
    // bottom box (tile row r - r0); results go to cres rows 2(r - r0), 2(r - r0) + 1.  Odd lh: the last low is copied (:537-545).
    auto col_pair = [&](int r, int lm1, int l0, int lp1, int h, int cx) {
        const int ry = 2 * (r - r0);
        if (r < nhy) {
            const int d = inv_d(f1, f2, a.lh, r, lm1, l0, lp1, h);
            const int t = l0 + ((d + 1) >> 1);
            cres[ry][cx] = (int16_t) t;
            cres[ry + 1][cx] = (int16_t) (t - d);
        } else if (r < nly) {
            cres[ry][cx] = (int16_t) l0;
        }
    };
    if (warp < 8) {
        // warp = (x half, strip of 8 vertical pairs), lane = column; sliding window of lows down the column
        const int half = warp & 1, rs = r0 + (warp >> 1) * 8;
        const int cx = half == 0 ? 1 + lane : 2 + TK + lane;                  // cres column: low-x k0+lane / high-x k0+lane
        const bool col_ok = half == 0 ? k0 + lane < nlx : k0 + lane < nhx;
        if (col_ok && rs < nly) {
            auto top = [&](int r) -> int { const int y = r - r0 + 1; return half == 0 ? (int) sA[y][8 + lane] : (int) sB[y][lane]; };
            auto bot = [&](int r) -> int { const int y = r - r0; return half == 0 ? (int) sC[y][8 + lane] : (int) sD[y][lane]; };
            int lm1 = top(rs - 1), l0 = top(rs);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int r = rs + i;
                const int lp1 = top(r + 1);
                col_pair(r, lm1, l0, lp1, bot(r), cx);
                lm1 = l0; l0 = lp1;
            }
        }
    } else {
        // the two halo columns of the low-x half (k0 - 1 and k0 + TK): lane = vertical pair
        const int r = r0 + lane;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int kx = e == 0 ? k0 - 1 : k0 + TK;
            if (kx < 0 || kx >= nlx || r >= nly) continue;
            const int sx = e == 0 ? 7 : 8 + TK;
            col_pair(r, (int) sA[lane][sx], (int) sA[lane + 1][sx], (int) sA[lane + 2][sx], (int) sC[lane][sx], e == 0 ? 0 : 1 + TK);
        }
    }
    __syncthreads();
    // ---- row pass: a thread reconstructs 4 horizontal pairs (8 samples) of one row and stores them as one 128-bit vector -----------
    if (warp < 8) {
        int16_t *dst_p = a.dst + (size_t) plane * a.plane_stride;
        const int q = lane & 7;
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int ry = warp * 8 + it * 4 + (lane >> 3);
            const int gy = 2 * r0 + ry;
            if (gy >= a.lh) continue;
            const int kb = k0 + 4 * q;                      // first of this thread's 4 pairs
            if (kb >= nlx) continue;
            const int16_t *cr = cres[ry];
            int lo[6];
#pragma unroll
            for (int j = 0; j < 6; j++) lo[j] = (int) cr[4 * q + j];          // l[kb-1 .. kb+4]
            int v[8];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int k = kb + j;
                int x0v = lo[j + 1], x1v = 0;
                if (k < nhx) {
                    const int d = inv_d(f1, f2, a.lw, k, lo[j], lo[j + 1], lo[j + 2], (int) cr[2 + TK + 4 * q + j]);
                    const int t = lo[j + 1] + ((d + 1) >> 1);
                    x0v = (int) (int16_t) t; x1v = (int) (int16_t) (t - d);
                }
                if (a.clamp) { x0v = x0v < 0 ? 0 : x0v; x1v = x1v < 0 ? 0 : x1v; }
                v[2 * j] = x0v; v[2 * j + 1] = x1v;
            }
            int16_t *o = dst_p + (size_t) gy * a.W + 2 * kb;
            if (2 * kb + 8 <= a.lw) {
                uint4 pk;
                pk.x = (uint32_t) (uint16_t) v[0] | ((uint32_t) (uint16_t) v[1] << 16);
                pk.y = (uint32_t) (uint16_t) v[2] | ((uint32_t) (uint16_t) v[3] << 16);
                pk.z = (uint32_t) (uint16_t) v[4] | ((uint32_t) (uint16_t) v[5] << 16);
                pk.w = (uint32_t) (uint16_t) v[6] | ((uint32_t) (uint16_t) v[7] << 16);
                *reinterpret_cast<uint4 *>(o) = pk;
            } else {
                for (int j = 0; j < 8 && 2 * kb + j < a.lw; j++) o[j] = (int16_t) v[j];
            }
        }
    }
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
            if (filt == ICER_FILTER_A) k_fwd_stage_tma<true><<<grid, 256, 0, st>>>(tm, a);
            else k_fwd_stage_tma<false><<<grid, 256, 0, st>>>(tm, a);
            ls.launches++;
        }
        cur = ll; cur_stride = ps;
        lw = (lw + 1) / 2; lh = (lh + 1) / 2;
    }
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}

// Fused inverse transform (filter A only: no beta term, so every line is data-parallel).  coefs: [nplanes] planes of w*h with the
// coefficients (NOT modified); pp0 / pp1: ping-pong planes for the intermediate LL regions; dst: the reconstructed planes.
bool dwt_inverse_fused_applicable(const void *coefs, const void *dst, uint32_t w, uint32_t h, uint32_t stages, uint32_t filt) {
    if (get_encode_fn() == nullptr || filt != ICER_FILTER_A || (w % 8) != 0 || ((uintptr_t) coefs % 16) != 0 || ((uintptr_t) dst % 16) != 0 ||
        (((size_t) w * h * 2) % 16) != 0)
        return false;
    for (uint32_t s = 0; s < stages; s++)          // the high-x boxes start at column nlx + k0: keep that a multiple of 8 samples (16 bytes)
        if ((dim_low(w, s + 1) % 8) != 0) return false;
    return stages >= 1;
}

int launch_dwt_inverse_fused(const uint16_t *coefs, uint16_t *pp0, uint16_t *pp1, uint16_t *dst, size_t nplanes, uint32_t w, uint32_t h,
                             uint32_t stages, uint32_t filt, bool clamp_negative, cudaStream_t st, LaunchStats &ls) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return ICER_FATAL_ERROR;
    const size_t ps = (size_t) w * h;
    static const int filt_tab[7][4] = {{0, 4, 4, 0}, {0, 4, 6, 4}, {-1, 4, 8, 6}, {0, 4, 5, 2}, {0, 3, 8, 6}, {0, 3, 9, 8}, {0, 4, 4, 4}};
    const uint16_t *ll_src = coefs;                 // the coarsest stage finds its LL quadrant in the coefficient buffer itself
    for (int s = (int) stages - 1; s >= 0; s--) {
        const uint32_t lw = (uint32_t) dim_low(w, s), lh = (uint32_t) dim_low(h, s);
        const uint32_t nlx = (lw + 1) / 2, nly = (lh + 1) / 2;
        uint16_t *out = s == 0 ? dst : (((stages - 1 - s) & 1) ? pp1 : pp0);
        for (size_t p0 = 0; p0 < nplanes; p0 += 32768) {
            const size_t cnt = nplanes - p0 < 32768 ? nplanes - p0 : 32768;
            alignas(64) CUtensorMap tm[4];
            const cuuint64_t gstr[2] = {(cuuint64_t) w * 2, (cuuint64_t) ps * 2};
            const cuuint32_t estr[3] = {1, 1, 1};
            const cuuint64_t dimLL[3] = {nlx, nly, cnt}, dimR[3] = {lw, lh, cnt};
            const cuuint32_t boxes[4][3] = {{IA_W, IA_H, 1}, {TK, IA_H, 1}, {IA_W, TR, 1}, {TK, TR, 1}};
            for (int b = 0; b < 4; b++) {
                const void *base = b == 0 ? (const void *) (ll_src + p0 * ps) : (const void *) (coefs + p0 * ps);
                CUresult r = enc(&tm[b], CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void *>(base), b == 0 ? dimLL : dimR, gstr, boxes[b], estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    fprintf(stderr, "[icer_b200] cuTensorMapEncodeTiled (inverse, box %d) failed (%d)\n", b, (int) r);
                    return ICER_FATAL_ERROR;
                }
            }
            InvArgs a;
            a.lw = (int) lw; a.lh = (int) lh; a.W = (int) w; a.plane_stride = ps;
            a.dst = reinterpret_cast<int16_t *>(out) + p0 * ps;
            a.f1 = filt_tab[filt][1]; a.f2 = filt_tab[filt][2];
            a.clamp = (clamp_negative && s == 0) ? 1 : 0;
            dim3 grid((nlx + TK - 1) / TK, (nly + TR - 1) / TR, (unsigned) cnt);
            k_inv_stage_tma<<<grid, 288, 0, st>>>(tm[0], tm[1], tm[2], tm[3], a);
            ls.launches++;
        }
        ll_src = out;
    }
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
