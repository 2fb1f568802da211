// coder.cu -- K4 (bit-plane context model + interleaved entropy ENCODER), K7 (the DECODER) and K5 (segment assembly).
//
// Granularity.  lib_icer resets its context model and entropy coder for every (channel, level, subband, bit plane,
// segment) (icer_partition.c:315-322), so each such STREAM is independent on the encode side.  On the decode side
// the planes of one segment form a chain (MSB -> LSB, icer_partition.c:427-443), but plane p only needs plane p+1
// to be a row + 1 pixel ahead (its right/down neighbours are taken at plane p+1, icer_context_modeller.c:510-521).
//
// Mapping (both kernels): one LANE per stream; the lanes of a warp are the bit planes of the same few segments, so
// they walk the same pixels (shared sectors) -- in lockstep for the encoder, as an asynchronous pipeline for the
// decoder (every tick a lane advances one pixel if the plane above is far enough ahead; positions travel by shfl).
// Warps fetch work dynamically, largest segments first.
//
// Two things keep the SIMT lanes together:
//   * every pixel is turned into at most three micro-ops (flush-fast-run, magnitude decision, sign decision) that
//     run through ONE code site for bin selection, ONE for codeword completion and ONE for bit emission;
//   * the "quiet" fast path: when the 3x3 neighbourhood has no bit at or above the current plane the decision is
//     (context 0, bit 0); runs of those only bump counters (adaptive counts, Golomb run length) until the next event
//     (run completes a codeword, probability bin changes, counts rescale), computed in closed form.
//
// Reference behaviour restated (lib_icer/src): icer_context_modeller.c:312-602 (contexts, adaptive counts with the
// discarded-ceil_div rescale), icer_encoding.c:37-189 and icer_decoding.c:46-194 (bin choice, Golomb /
// variable-to-variable / uncoded bins, 2048-codeword interleave window with forced flush and its decoder mirror),
// icer_util.c:48-56 (bin cut-offs), icer_partition.c:328-334 + icer_encoding.c:210-234 (segment header, quota).
#include "common.h"
#include <mutex>

namespace {

constexpr int NT = 128;   // threads per block (both kernels)
// Codewords that complete behind an older open one wait in the interleave window (lib_icer's 2048-slot circular buffer).  The
// window of every resident thread lives in global memory (2048 x 2 B x ~150 k threads); most codewords, however, complete within
// a few sequence numbers of the emission point: those are parked in a small per-thread shared-memory window instead and never
// touch global memory.  Must be a power of two (0 = everything goes through the global ring).
#ifndef ICER_ENC_MINI
#define ICER_ENC_MINI 8
#endif

struct DevTables {
    uint32_t cutoff[16];
    uint32_t crc[256];
    uint16_t enc[8][32];       // in_bits | out_bits << 4 | out << 8        (keyed by the gathered input word)
    uint16_t dec[8][32];       // in_bits | out_bits << 4 | out << 8        (keyed by the received word)
    uint16_t dlut[8][32];      // keyed by the NEXT 5 stream bits: codeword length | out_bits << 4 | out << 8 (0 = none)
    uint16_t gol_m[17], gol_l[17], gol_i[17];
    uint8_t flush[8][9][6];    // pad_word | pad_len << 4
    uint8_t ctx_hv[45], ctx_hh[25], ctx_sign[25], sign_pred[25];   // icer_config.c:26-67
};
__constant__ DevTables c_tab;
// __constant__ memory has one instance PER DEVICE: the upload is tracked per device ordinal (a process may hold contexts on
// several GPUs, e.g. icer_b200_*_batch_multi)
std::mutex g_tab_mu;
bool g_tables_uploaded[64] = {false};

const uint8_t h_ctx_hv[45] = {0, 1, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 6, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7,
                              8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8};
const uint8_t h_ctx_hh[25] = {0, 3, 6, 8, 8, 1, 4, 7, 8, 8, 2, 5, 7, 8, 8, 2, 5, 7, 8, 8, 2, 5, 7, 8, 8};
const uint8_t h_ctx_sign[25] = {14, 14, 15, 16, 16, 14, 14, 15, 16, 16, 13, 13, 12, 13, 13, 16, 16, 15, 14, 14, 16, 16, 15, 14, 14};
const uint8_t h_sign_pred[25] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

int upload_tables() {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return ICER_FATAL_ERROR;
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (g_tables_uploaded[dev]) return ICER_RESULT_OK;
    const HostTables &h = icer_host_tables();
    static DevTables d;
    for (int b = 0; b < 8; b++)
        for (int v = 0; v < 32; v++) {
            d.enc[b][v] = (uint16_t) (h.enc[b][v].in_bits | (h.enc[b][v].out_bits << 4) | (h.enc[b][v].out << 8));
            d.dec[b][v] = (uint16_t) (h.dec[b][v].in_bits | (h.dec[b][v].out_bits << 4) | (h.dec[b][v].out << 8));
        }
    memcpy(d.flush, h.flush, sizeof d.flush);
    // the variable-to-variable codes are complete prefix codes of at most 5 bits: the next 5 stream bits select the
    // codeword directly (same first-match order as the bit-by-bit loop of icer_decoding.c:155-177)
    for (int b = 0; b < 8; b++)
        for (int x = 0; x < 32; x++) {
            d.dlut[b][x] = 0;
            for (int L = 1; L <= 5; L++) {
                const V2VEntry &e = h.dec[b][x & ((1 << L) - 1)];
                if (e.in_bits == L) { d.dlut[b][x] = (uint16_t) (L | (e.out_bits << 4) | (e.out << 8)); break; }
            }
        }
    for (int b = 0; b < 17; b++) { d.gol_m[b] = h.gol_m[b]; d.gol_l[b] = h.gol_l[b]; d.gol_i[b] = h.gol_i[b]; }
    for (int b = 0; b < 16; b++) d.cutoff[b] = h.cutoff[b];
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        d.crc[n] = c;
    }
    memcpy(d.ctx_hv, h_ctx_hv, 45); memcpy(d.ctx_hh, h_ctx_hh, 25);
    memcpy(d.ctx_sign, h_ctx_sign, 25); memcpy(d.sign_pred, h_sign_pred, 25);
    CUDA_TRY(cudaMemcpyToSymbol(c_tab, &d, sizeof d));      // synchronous: complete before any later launch on any stream
    g_tables_uploaded[dev] = true;
    return ICER_RESULT_OK;
}

// tables every lane needs with divergent indices: staged in shared memory once per block
struct SmemTables {
    uint2 cutp[17];            // per bin: (cut-off below it, cut-off above it) with sentinels 0 / 65537 -- one 64-bit load validates a cached bin
    uint32_t cutoff[16];
    uint16_t v2v[8][32];
    uint16_t gol_m[17], gol_l[17], gol_i[17];
    uint8_t flush[8][9][6];
    uint8_t ctx_hv[45], ctx_hh[25], ctx_sign[25], sign_pred[25];
};

__device__ __forceinline__ void load_tables(SmemTables &T, bool decoder) {
    const int t = threadIdx.x;
    for (int i = t; i < 8 * 32; i += NT) (&T.v2v[0][0])[i] = decoder ? (&c_tab.dlut[0][0])[i] : (&c_tab.enc[0][0])[i];
    for (int i = t; i < 8 * 9 * 6; i += NT) (&T.flush[0][0][0])[i] = (&c_tab.flush[0][0][0])[i];
    for (int i = t; i < 45; i += NT) T.ctx_hv[i] = c_tab.ctx_hv[i];
    for (int i = t; i < 25; i += NT) { T.ctx_hh[i] = c_tab.ctx_hh[i]; T.ctx_sign[i] = c_tab.ctx_sign[i]; T.sign_pred[i] = c_tab.sign_pred[i]; }
    for (int i = t; i < 17; i += NT) { T.gol_m[i] = c_tab.gol_m[i]; T.gol_l[i] = c_tab.gol_l[i]; T.gol_i[i] = c_tab.gol_i[i]; }
    for (int i = t; i < 16; i += NT) T.cutoff[i] = c_tab.cutoff[i];
    for (int i = t; i < 17; i += NT) T.cutp[i] = make_uint2(i ? c_tab.cutoff[i - 1] : 0u, i < 16 ? c_tab.cutoff[i] : 65537u);
}

// probability bin: number of cut-offs c with zz * 65536 >= tot * c (icer_util.c:48-56); cut-offs ascend
__device__ __forceinline__ int bin_of(const SmemTables &T, uint32_t zz, uint32_t tot) {
    const uint32_t lhs = zz << 16;
    int bin = 0;
    if (lhs >= tot * T.cutoff[7]) bin = 8;
    if (lhs >= tot * T.cutoff[bin + 3]) bin += 4;
    if (lhs >= tot * T.cutoff[bin + 1]) bin += 2;
    if (lhs >= tot * T.cutoff[bin]) bin += 1;
    if (bin == 15 && lhs >= tot * T.cutoff[15]) bin = 16;
    return bin;
}

// The bin of a context moves slowly, so each context remembers its last bin (bits 27..31 of its count word); the guess is
// validated with two products (it holds iff cut[guess-1] * tot <= zz * 65536 < cut[guess] * tot) and the 5-step dependent
// search only runs when the context has crossed into another bin.
__device__ __forceinline__ int bin_cached(const SmemTables &T, uint32_t zz, uint32_t tot, int guess) {
    const uint32_t lhs = zz << 16;
    const uint2 c = T.cutp[guess];
    if (lhs >= tot * c.x && lhs < tot * c.y) return guess;
    return bin_of(T, zz, tot);
}
#define CNT_Z(zt) (((zt) >> 16) & 0x7FFu)
#define CNT_TOT(zt) ((zt) & 0xFFFFu)
#define CNT_BIN(zt) ((int) ((zt) >> 27))
#define CNT_PACK(bin, z, tot) (((uint32_t) (bin) << 27) | ((z) << 16) | (tot))

// number of consecutive all-zero decisions, starting at counts (z, tot), that stay in Golomb bin `bin`
__device__ __forceinline__ uint32_t zeros_in_same_bin(const SmemTables &T, int bin, uint32_t z, uint32_t tot) {
    if (bin >= 16) return 0x7FFFFFFFu;
    const uint32_t cut = T.cutoff[bin];
    const uint32_t num = tot * cut - (z << 16);       // > 0 because the cut-off test for `bin` failed
    return (num - 1) / (65536u - cut) + 1;
}

__device__ __forceinline__ int sig(uint32_t v, int plane) { return ((v & 0x7FFF) >> plane) != 0; }

// magnitude context of a pixel of category 0/1/2 (icer_context_modeller.c:355-391)
__device__ __forceinline__ int mag_context(const SmemTables &T, int cat, int lsb, int sb, uint32_t a0, uint32_t a1, uint32_t a2,
                                           uint32_t b0, uint32_t b2, uint32_t c0, uint32_t c1, uint32_t c2) {
    if (cat == 2) return 11;
    const int prev = lsb + 1;
    int hh = sig(b0, lsb) + sig(b2, prev);
    int vv = sig(a1, lsb) + sig(c1, prev);
    if (cat == 1) return (hh + vv == 0) ? 9 : 10;
    const int dd = sig(a0, lsb) + sig(a2, lsb) + sig(c0, prev) + sig(c2, prev);
    if (sb == 1) { const int tmp = hh; hh = vv; vv = tmp; }
    return (sb == 3) ? T.ctx_hh[(hh + vv) * 5 + dd] : T.ctx_hv[(hh * 3 + vv) * 5 + dd];
}
// sign context index into the 5x5 tables (icer_context_modeller.c:404-428)
__device__ __forceinline__ int sign_index(int lsb, int sb, uint32_t a1, uint32_t b0, uint32_t b2, uint32_t c1) {
    const int prev = lsb + 1;
    int shh = 2 - (int) ((b0 >> 15) & sig(b0, lsb)) - (int) ((b2 >> 15) & sig(b2, prev));
    int svv = 2 - (int) ((a1 >> 15) & sig(a1, lsb)) - (int) ((c1 >> 15) & sig(c1, prev));
    if (sb == 1) { const int tmp = shh; shh = svv; svv = tmp; }
    return shh * 5 + svv;
}

// =========================================================================================================
// Frame-interleaved coefficient layout ("T" layout)
// =========================================================================================================
// The coder kernels give every LANE of a warp the same stream of a different FRAME (32 consecutive frames = one
// group): identical geometry, so all loop bounds and edge tests are warp-uniform, and similar statistics, so the
// lanes mostly take the same branches.  To make their pixel accesses one coalesced 64-byte line, the sign-magnitude
// coefficients are stored as  T[group][channel][y][x][lane]  (uint16).  These two kernels convert between the planar
// frame layout the wavelet kernels use and T; the decode-side one also undoes sign-magnitude and adds the LL mean
// (icer_compress.c:520-531).

__global__ void __launch_bounds__(256) k_frames_to_T(const uint16_t *__restrict__ src, uint16_t *__restrict__ T, uint32_t nframes,
                                                     uint32_t chan, uint32_t plane_px) {
    // block: 32 frames (one group) x 32 consecutive pixels of one channel
    __shared__ uint16_t tile[32][34];
    const uint32_t g = blockIdx.z, c = blockIdx.y;
    const size_t px0 = (size_t) blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows of 32
    for (int f = ty; f < 32; f += 8) {
        const uint32_t frame = g * 32 + f;
        const size_t px = px0 + tx;
        tile[f][tx] = (frame < nframes && px < plane_px) ? src[((size_t) frame * chan + c) * plane_px + px] : (uint16_t) 0;
    }
    __syncthreads();
    for (int q = ty; q < 32; q += 8) {
        const size_t px = px0 + q;
        if (px < plane_px) T[(((size_t) g * chan + c) * plane_px + px) * 32 + tx] = tile[tx][q];
    }
}

__global__ void __launch_bounds__(256) k_T_to_frames(const uint16_t *__restrict__ T, uint16_t *__restrict__ dst,
                                                     const int32_t *__restrict__ frame_of, const uint16_t *__restrict__ mean,
                                                     uint32_t chan, uint32_t plane_px, uint64_t cap_px, int W, int llw, int llh,
                                                     int bits8) {
    __shared__ uint16_t tile[32][34];
    const uint32_t g = blockIdx.z, c = blockIdx.y;
    const size_t px0 = (size_t) blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int q = ty; q < 32; q += 8) {
        const size_t px = px0 + q;
        tile[q][tx] = px < plane_px ? T[(((size_t) g * chan + c) * plane_px + px) * 32 + tx] : (uint16_t) 0;
    }
    __syncthreads();
    for (int f = ty; f < 32; f += 8) {
        const int32_t frame = frame_of[g * 32 + f];
        const size_t px = px0 + tx;
        if (frame < 0 || px >= plane_px) continue;
        uint32_t v = tile[tx][f];
        v = (v & 0x8000u) ? ((0x8000u - v) & 0xFFFFu) : v;                      // icer_wavelet.c:880-886
        const int r = (int) (px / (size_t) W), cc = (int) (px % (size_t) W);
        if (r < llh && cc < llw) {
            const int m = (int) mean[(size_t) frame * chan + c];
            // uint16 API: int16 add (icer_compress.c:524-531); native uint8 API: int8 add of the header value's low byte (:258-265)
            v = bits8 ? (uint32_t) (uint16_t) (int16_t) (int8_t) ((int) (int16_t) v + (int) (int8_t) m)
                      : (uint32_t) (uint16_t) (int16_t) ((int) (int16_t) v + (int) (int16_t) m);
        }
        dst[((size_t) frame * chan + c) * cap_px + px] = (uint16_t) v;
    }
}

// 128-bit versions: a block moves a tile of 32 frames x 64 pixels; global reads and writes are 16-byte vectors (8 consecutive pixels
// of one frame on the planar side, 8 frames of one pixel on the T side -- a warp then covers 512 contiguous bytes of T).  The
// encode-side kernel also applies the sign-magnitude mapping (icer_wavelet.c:871-877), so the planar coefficients are read once.
// Requirements (else the scalar kernels above): plane_px, cap_px and W multiples of 8, 16-byte aligned bases.
constexpr int TT_PX = 64, TT_LD = TT_PX + 8;

template <bool SIGNMAG>
__global__ void __launch_bounds__(256) k_frames_to_T_v(const uint16_t *__restrict__ src, uint16_t *__restrict__ T, uint32_t nframes,
                                                       uint32_t chan, uint32_t plane_px) {
    __shared__ __align__(16) uint16_t tile[32][TT_LD];
    const uint32_t g = blockIdx.z, c = blockIdx.y;
    const size_t px0 = (size_t) blockIdx.x * TT_PX;
    const int t = threadIdx.x;
    {
        const int f = t >> 3, k = t & 7;
        const uint32_t frame = g * 32 + f;
        const size_t px = px0 + 8 * k;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (frame < nframes && px < plane_px) v = *reinterpret_cast<const uint4 *>(src + ((size_t) frame * chan + c) * plane_px + px);
        if (SIGNMAG) {
            uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int lo = (int) (int16_t) (w[j] & 0xFFFF), hi = (int) (int16_t) (w[j] >> 16);
                const uint32_t a = lo < 0 ? ((uint32_t) (-lo) & 0x7FFFu) | 0x8000u : (uint32_t) lo;
                const uint32_t b = hi < 0 ? ((uint32_t) (-hi) & 0x7FFFu) | 0x8000u : (uint32_t) hi;
                w[j] = a | (b << 16);
            }
        }
        *reinterpret_cast<uint4 *>(&tile[f][8 * k]) = v;
    }
    __syncthreads();
    {
        const int px = t >> 2, fq = t & 3;
        if (px0 + px < plane_px) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; j++) w[j] = (uint32_t) tile[8 * fq + 2 * j][px] | ((uint32_t) tile[8 * fq + 2 * j + 1][px] << 16);
            *reinterpret_cast<uint4 *>(T + (((size_t) g * chan + c) * plane_px + px0 + px) * 32 + 8 * fq) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

__global__ void __launch_bounds__(256) k_T_to_frames_v(const uint16_t *__restrict__ T, uint16_t *__restrict__ dst,
                                                       const int32_t *__restrict__ frame_of, const uint16_t *__restrict__ mean,
                                                       uint32_t chan, uint32_t plane_px, uint64_t cap_px, int W, int llw, int llh) {
    __shared__ __align__(16) uint16_t tile[32][TT_LD];
    const uint32_t g = blockIdx.z, c = blockIdx.y;
    const size_t px0 = (size_t) blockIdx.x * TT_PX;
    const int t = threadIdx.x;
    {
        const int px = t >> 2, fq = t & 3;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (px0 + px < plane_px) v = *reinterpret_cast<const uint4 *>(T + (((size_t) g * chan + c) * plane_px + px0 + px) * 32 + 8 * fq);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(&v);
#pragma unroll
        for (int j = 0; j < 4; j++) { tile[8 * fq + 2 * j][px] = (uint16_t) (w[j] & 0xFFFF); tile[8 * fq + 2 * j + 1][px] = (uint16_t) (w[j] >> 16); }
    }
    __syncthreads();
    {
        const int f = t >> 3, k = t & 7;
        const int32_t frame = frame_of[g * 32 + f];
        const size_t px = px0 + 8 * k;
        if (frame < 0 || px >= plane_px) return;
        uint4 v = *reinterpret_cast<const uint4 *>(&tile[f][8 * k]);
        uint32_t *w = reinterpret_cast<uint32_t *>(&v);
        const int r = (int) (px / (size_t) W), cc = (int) (px % (size_t) W);       // the 8 pixels lie in one row (W % 8 == 0)
        const int m = (int) (int16_t) mean[(size_t) frame * chan + c];
        const bool ll_row = r < llh;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t a = w[j] & 0xFFFF, b = w[j] >> 16;
            a = (a & 0x8000u) ? ((0x8000u - a) & 0xFFFFu) : a;                      // icer_wavelet.c:880-886
            b = (b & 0x8000u) ? ((0x8000u - b) & 0xFFFFu) : b;
            if (ll_row && cc + 2 * j < llw) a = (uint32_t) (uint16_t) (int16_t) ((int) (int16_t) a + m);        // icer_compress.c:524-531
            if (ll_row && cc + 2 * j + 1 < llw) b = (uint32_t) (uint16_t) (int16_t) ((int) (int16_t) b + m);
            w[j] = a | (b << 16);
        }
        *reinterpret_cast<uint4 *>(dst + ((size_t) frame * chan + c) * cap_px + px) = v;
    }
}

// =========================================================================================================
// K4: encoder
// =========================================================================================================
struct EncSmem {
    SmemTables T;
    uint32_t cnt[17][NT];      // zero << 16 | total, per context
    uint16_t bseq[17][NT];     // sequence number (low 16 bits; the window spans 2048) of the codeword each bin is building
    uint32_t bst[17][NT];      // Golomb: zeros counted; custom: input word (first bit = bit 0) | bits gathered << 16
#if ICER_ENC_MINI > 0
    uint16_t mini[ICER_ENC_MINI][NT];   // parked codewords close to the emission point (see Enc::complete)
#endif
};

// Interleave-window bookkeeping.  A codeword gets a sequence number when its first bit arrives (lib_icer: a ring
// slot, icer_encoding.c:59-67); `head` is the next sequence number to emit.  Everything in [head, oldest open) is
// complete by construction, so only codewords that complete BEHIND an older open one are parked in the ring.
struct Enc {
    EncSmem &S;
    const int t;
    uint16_t *ring;
    uint32_t *out;
    uint32_t cap, nwords = 0;
    unsigned long long acc = 0;
    int nacc = 0;
    uint32_t head = 0, tail = 0, open_mask = 0;   // head: first sequence number that is not complete yet (lib_icer's head)
    uint32_t emit_pos = 0;                        // next sequence number to pack into the payload (<= head)
    // bits 24.. of open_mask: which slots of S.mini hold a parked codeword (one register for both masks)
    int oldest = 0;
    bool overflow = false;

    __device__ Enc(EncSmem &s, int tid, uint16_t *r, uint32_t *o, uint32_t c) : S(s), t(tid), ring(r), out(o), cap(c) {}

    __device__ __forceinline__ uint32_t golomb_word(int bin, uint32_t k, uint32_t &nb) {   // icer_encoding.c:72-77
        const uint32_t i = S.T.gol_i[bin], l = S.T.gol_l[bin];
        const uint32_t longer = k >= i;
        nb = l + longer;
        return __brev(k + (longer ? i : 0)) >> (32 - nb);
    }

    // codeword of `bin` is complete (single call site in the pixel loop): park it in the ring; if it was the oldest
    // one, everything up to the next open codeword is now ready for emission
    __device__ __forceinline__ void complete(int bin, uint32_t code, uint32_t nb) {
        const uint32_t seq = S.bseq[bin][t];
        open_mask &= ~(1u << bin);
#if ICER_ENC_MINI > 0
        // a parked entry is live until emitted, and emit_pos only grows: every live entry of the small window has a sequence
        // number in [emit_pos, emit_pos + ICER_ENC_MINI), so slot = seq mod ICER_ENC_MINI is unambiguous (checked in pump())
        if (((seq - emit_pos) & 0xFFFFu) < (uint32_t) ICER_ENC_MINI) {
            S.mini[seq & (ICER_ENC_MINI - 1)][t] = (uint16_t) ((nb << 11) | code);
            open_mask |= 0x1000000u << (seq & (ICER_ENC_MINI - 1));
        } else
#endif
        ring[seq & (ICER_RING - 1)] = (uint16_t) ((nb << 11) | code);
        if (seq != (head & 0xFFFFu)) return;
        const uint32_t h1 = head + 1;
        uint32_t dist = tail - h1;                       // every other open codeword lies in [head + 1, tail)
        for (uint32_t m = open_mask & 0x1FFFFu; m; m &= m - 1) {
            const int b = __ffs(m) - 1;
            const uint32_t d = ((uint32_t) S.bseq[b][t] - h1) & 0xFFFFu;
            if (d < dist) { dist = d; oldest = b; }
        }
        head = h1 + dist;
    }

    // pack one ready codeword into the payload, LSB first (icer_encoding.c:114-139).  Called from ONE convergent site
    // per pixel so that the lanes of a warp emit together instead of each draining its backlog alone.
    __device__ __forceinline__ void pump() {
        if (emit_pos == head) return;
#if ICER_ENC_MINI > 0
        uint32_t v;
        const uint32_t slot = emit_pos & (ICER_ENC_MINI - 1);
        if ((open_mask >> (24 + slot)) & 1u) { v = S.mini[slot][t]; open_mask &= ~(0x1000000u << slot); }
        else v = ring[emit_pos & (ICER_RING - 1)];
#else
        const uint32_t v = ring[emit_pos & (ICER_RING - 1)];
#endif
        emit_pos++;
        acc |= (unsigned long long) (v & 0x7FF) << nacc;
        nacc += (int) (v >> 11);
        if (nacc >= 32) {
            if (nwords < cap) out[nwords] = (uint32_t) acc; else overflow = true;
            nwords++;
            acc >>= 32;
            nacc -= 32;
        }
    }

    // the word a forced completion gives the oldest codeword (icer_encoding.c:141-189): window full, or end of stream
    __device__ __forceinline__ uint32_t forced_word(int bin, uint32_t &nb) {
        const uint32_t st = S.bst[bin][t];
        if (bin > 7) {
            if (st == (uint32_t) S.T.gol_m[bin] - 1) { nb = 1; return 1; }
            return golomb_word(bin, st, nb);
        }
        const uint32_t pre = st & 0xFFFF, fill = st >> 16;
        const uint32_t fl = S.T.flush[bin][pre][fill];
        const uint32_t word = (pre | ((fl & 0xF) << fill)) & 0x3FF;
        const uint32_t e = S.T.v2v[bin][word & 31];
        nb = (e >> 4) & 0xF;
        return e >> 8;
    }
    // true when `bin` needs a new codeword but 2048 are already outstanding (icer_encoding.c:59-64)
    __device__ __forceinline__ bool window_full_for(int bin) const {
        return !((open_mask >> bin) & 1u) && tail - head >= ICER_RING;
    }
    // make sure `bin` has an open codeword (icer_encoding.c:56-67); the caller has dealt with a full window
    __device__ __forceinline__ void open_bin(int bin) {
        if (open_mask & (1u << bin)) return;
        while (tail - emit_pos >= ICER_RING) pump();       // physical ring space (emission may lag behind `head`)
        if ((open_mask & 0x1FFFFu) == 0) oldest = bin;
        S.bseq[bin][t] = (uint16_t) tail++;
        S.bst[bin][t] = 0;
        open_mask |= 1u << bin;
    }

    __device__ void finish(uint32_t &bits_out) {
        while (open_mask & 0x1FFFFu) {                         // icer_context_modeller.c:452-455
            uint32_t nb;
            const int bin = oldest;
            const uint32_t code = forced_word(bin, nb);
            complete(bin, code, nb);
        }
        while (emit_pos != head) pump();
        bits_out = nwords * 32u + (uint32_t) nacc;
        if (nacc > 0) { if (nwords < cap) out[nwords] = (uint32_t) acc; else overflow = true; }
    }
};

// One warp per (frame group, stream); lane = frame of the group.
__global__ void __launch_bounds__(NT, 8) k_encode(EncodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EncSmem &S = *reinterpret_cast<EncSmem *>(smem_raw);
    const int t = threadIdx.x;
    load_tables(S.T, false);
    __syncthreads();

    const uint32_t nstreams = a.nunits * a.planes;
    const unsigned long long total = (unsigned long long) a.ngroups * (a.list ? a.nlist : nstreams);
    const int lane = t & 31;
    uint16_t *ring = a.rings + (size_t) (blockIdx.x * NT + t) * ICER_RING;
    const size_t plane_px = (size_t) a.w * a.h;
    const int W = (int) a.w;
    const int RS = W * 32;     // row stride of the T layout, in elements

    for (;;) {
        unsigned long long wid = 0;
        if (lane == 0) wid = atomicAdd(a.work_counter, 1ull);
        wid = __shfl_sync(0xffffffffu, wid, 0);
        if (wid >= total) break;
        // stream-major order: the largest streams of ALL groups first, so the tail of the kernel consists of small streams
        const uint32_t g = (uint32_t) (wid % a.ngroups);
        const uint32_t s = a.list ? a.list[wid / a.ngroups] : (uint32_t) (wid / a.ngroups);
        const Unit u = a.units[s / a.planes];
        const int lsb = (int) (s % a.planes);
        const uint32_t frame = g * 32 + lane;
        const bool valid = frame < a.nframes;
        const uint16_t *seg = a.coefsT + (((size_t) g * a.chan + u.chan) * plane_px + (size_t) u.y * W + u.x) * 32 + lane;
        Enc cd(S, t, ring, valid ? a.scratch + (size_t) frame * a.scratch_words + a.soff[s] : nullptr, valid ? a.scap[s] : 0);
#pragma unroll
        for (int c = 0; c < 17; c++) S.cnt[c][t] = (2u << 16) | 4u;    // icer.h:147-148
        const int sw = u.w, sh = u.h, sb = u.subband;

        uint32_t fz_left = 0, fz_run = 0;    // quiet fast path: zeros still absorbable / absorbed but not yet applied
        int fz_bin = 0;

        for (int row = 0; row < sh; row++) {
            const bool up = row > 0, dn = row + 1 < sh;
            const uint16_t *p = seg + (size_t) row * RS;
            uint32_t a0 = 0, a1 = up ? p[-RS] : 0, a2, b0 = 0, b1 = p[0], b2, c0 = 0, c1 = dn ? p[RS] : 0, c2;
            // the column after next is fetched one pixel early so that the loads overlap the coding of this pixel
            uint32_t a2n = (up && sw > 1) ? p[32 - RS] : 0, b2n = sw > 1 ? p[32] : 0, c2n = (dn && sw > 1) ? p[32 + RS] : 0;
            for (int col = 0; col < sw; col++, p += 32) {
                a2 = a2n; b2 = b2n; c2 = c2n;
                const bool rt2 = col + 2 < sw;
                a2n = (up && rt2) ? p[64 - RS] : 0;
                b2n = rt2 ? p[64] : 0;
                c2n = (dn && rt2) ? p[64 + RS] : 0;
                const uint32_t v = b1;
                const bool quiet = ((((a0 | a1 | a2 | b0 | b1 | b2 | c0 | c1 | c2) & 0x7FFF) >> lsb) == 0);
                // micro-ops: bit0 = apply the pending fast run, bit1 = magnitude decision, bit2 = sign decision
                uint32_t todo;
                if (quiet && fz_left) { fz_run++; fz_left--; todo = 0; }
                else todo = (fz_run ? 1u : 0u) | 2u;
                int cat = 0;
                uint32_t mbit = 0;
                if (todo & 2u) {
                    const int msb = 31 - __clz((v & 0x7FFF) | 1);
                    cat = msb - lsb;                                   // icer_context_modeller.c:631-634
                    cat = cat < 0 ? 0 : (cat > 3 ? 3 : cat);
                    mbit = (v >> lsb) & 1;
                    if (cat == 0 && mbit) todo |= 4u;                  // first 1 of this coefficient: its sign follows
                }
                while (todo) {
                    const int st = __ffs(todo) - 1;
                    todo &= todo - 1;
                    bool done = false;
                    uint32_t code = 0, nb = 0;
                    int bin = 0;
                    if (st == 0) {
                        // apply `fz_run` (context 0, bit 0) decisions at once
                        bin = fz_bin;
                        const uint32_t zt = S.cnt[0][t];
                        uint32_t z = CNT_Z(zt) + fz_run, tot = CNT_TOT(zt) + fz_run;
                        if (tot >= 500) { tot >>= 1; if (z > tot) z >>= 1; }
                        S.cnt[0][t] = CNT_PACK(bin, z, tot);
                        const uint32_t k = S.bst[bin][t] + fz_run;
                        if (k >= S.T.gol_m[bin]) { done = true; code = 1; nb = 1; }
                        else S.bst[bin][t] = k;
                        fz_run = 0; fz_left = 0;
                    } else {
                        uint32_t bit;
                        if (st == 1 && cat == 3) {
                            bit = mbit; bin = 0;                              // p(0) = 1/2 -> uncoded bin
                            if (cd.window_full_for(0)) {
                                bin = cd.oldest; code = cd.forced_word(bin, nb); done = true; todo |= 1u << st;
                                goto emit_site;
                            }
                        } else {
                            int ctx;
                            if (st == 1) {
                                ctx = mag_context(S.T, cat, lsb, sb, a0, a1, a2, b0, b2, c0, c1, c2);
                                bit = mbit;
                            } else {
                                const int si = sign_index(lsb, sb, a1, b0, b2, c1);
                                ctx = S.T.ctx_sign[si];
                                bit = (S.T.sign_pred[si] ^ (v >> 15)) & 1;
                            }
                            const uint32_t zt = S.cnt[ctx][t];
                            uint32_t z = CNT_Z(zt), tot = CNT_TOT(zt);
                            uint32_t zz = z;
                            const bool inv = zz < (tot >> 1);
                            if (inv) zz = tot - zz;
                            bin = bin_cached(S.T, zz, tot, CNT_BIN(zt));
                            const int cbin = bin;
                            if (cd.window_full_for(bin)) {
                                // window full: force-complete the oldest codeword first, then redo this decision
                                bin = cd.oldest; code = cd.forced_word(bin, nb); done = true; todo |= 1u << st;
                                goto emit_site;
                            }
                            if (st == 1 && quiet && !inv && bin > 7) {
                                // enter the quiet fast path: this and the following quiet pixels only bump counters
                                cd.open_bin(bin);
                                const uint32_t n1 = (uint32_t) S.T.gol_m[bin] - S.bst[bin][t];
                                const uint32_t n2 = zeros_in_same_bin(S.T, bin, z, tot);
                                const uint32_t n3 = 500u - tot;
                                fz_left = min(n1, min(n2, n3)) - 1;
                                fz_run = 1;
                                fz_bin = bin;
                                continue;
                            }
                            tot++;
                            z += bit ^ 1u;
                            if (tot >= 500) { tot >>= 1; if (z > tot) z >>= 1; }   // icer_context_modeller.c:396-402
                            S.cnt[ctx][t] = CNT_PACK(cbin, z, tot);
                            bit ^= inv ? 1u : 0u;
                        }
                        cd.open_bin(bin);
                        if (bin > 7) {                                        // Golomb run-length bins
                            if (bit) {
                                code = cd.golomb_word(bin, S.bst[bin][t], nb);
                                done = true;
                            } else {
                                const uint32_t k = S.bst[bin][t] + 1;
                                S.bst[bin][t] = k;
                                if (k >= S.T.gol_m[bin]) { code = 1; nb = 1; done = true; }
                            }
                        } else if (bin != 0) {                                // variable-to-variable bins
                            const uint32_t stt = S.bst[bin][t];
                            const uint32_t fill = stt >> 16;
                            const uint32_t word = (stt & 0xFFFF) | (bit << fill);
                            const uint32_t e = S.T.v2v[bin][word & 31];
                            if ((e & 0xF) == fill + 1) { code = e >> 8; nb = (e >> 4) & 0xF; done = true; }
                            else S.bst[bin][t] = word | ((fill + 1) << 16);
                        } else {                                              // uncoded bin
                            code = bit & 1; nb = 1; done = true;
                        }
                    }
                emit_site:
                    if (done) cd.complete(bin, code, nb);
                }
                cd.pump();       // up to two codewords can start per pixel, so two emissions keep pace
                cd.pump();
                a0 = a1; a1 = a2; b0 = b1; b1 = b2; c0 = c1; c1 = c2;
            }
        }
        if (fz_run) {   // tail of a fast run
            const int bin = fz_bin;
            const uint32_t k = S.bst[bin][t] + fz_run;
            if (k >= S.T.gol_m[bin]) cd.complete(bin, 1, 1);
            else S.bst[bin][t] = k;
        }
        uint32_t bits;
        cd.finish(bits);
        if (valid) {
            a.bits[(size_t) frame * nstreams + s] = bits;
            if (cd.overflow) atomicOr(a.overflow, 1);
        }
        __syncwarp();
    }
}

// =========================================================================================================
// K7: decoder
// =========================================================================================================
// One BLOCK per (frame group, chain): warp j decodes bit plane (planes-1-j) of the chain for the 32 frames of the
// group.  The warps form a pipeline: warp j may process pixel n once warp j-1 has finished pixel n + w + 1 (its
// right/down neighbours are taken one plane up).  Progress is published in shared memory; a warp advances only when
// all of its live lanes can, so its position -- and therefore every address and edge test -- stays warp-uniform.
// Work items of the decoder are TASKS (chain, bit plane).  A block is a pool of DEC_WARPS warps that take tasks in
// order (chain-major, most significant plane first); task (c, p) may process pixel n of frame f once task (c, p+1)
// has finished pixel n + w + 1 of that frame (its right/down neighbours are taken one plane up).  A task with a
// smaller number has always been taken by some warp already, so waiting on it cannot deadlock; there is no block
// barrier anywhere, so the quick (quiet) upper planes run ahead into the next chains while the dense lower planes are
// still being decoded.  Chains in flight occupy one of DEC_SLOTS slots of progress / flag state in shared memory.
// Pool size (measured on config 2, k_decode ms): 8 warps x 4 blocks/SM 42.1, 16 x 2 38.0, 24 x 1 40.7, 32 x 1 (12 slots) 36.1 --
// one SM-wide pool balances the very unequal planes of a chain best (profiles/README.md).
#ifndef ICER_DEC_WARPS
#define ICER_DEC_WARPS 32
#endif
#ifndef ICER_DEC_SLOTS
#define ICER_DEC_SLOTS 12
#endif
#ifndef ICER_DEC_MINBLOCKS
#define ICER_DEC_MINBLOCKS 1
#endif
#ifndef ICER_DEC_RECHECK
#define ICER_DEC_RECHECK 32          /* pixels a plane decodes between two looks at the progress of the plane above */
#endif
#ifndef ICER_DEC_BACKOFF_NS
#define ICER_DEC_BACKOFF_NS 5000     /* sleep when the plane above is not far enough ahead */
#endif
#ifndef ICER_DEC_SLOTWAIT_NS
#define ICER_DEC_SLOTWAIT_NS 20000    /* poll interval while a chain waits for a free slot (chains last milliseconds) */
#endif
#ifndef ICER_DEC_CHAINWAIT_NS
#define ICER_DEC_CHAINWAIT_NS 5000   /* poll interval while a plane waits for its chain's slot to be published */
#endif
constexpr int DEC_WARPS = ICER_DEC_WARPS;
constexpr int DEC_BT = DEC_WARPS * 32;
constexpr int DEC_SLOTS = ICER_DEC_SLOTS;

struct DecSlot {
    volatile uint32_t chain;             // chain ordinal (within this block) that owns the slot; ~0 = never used
    volatile uint32_t freed;             // all planes done and cleaned up: the slot may be re-used
    volatile uint32_t done;              // planes finished
    volatile uint32_t anyerr;            // some lane hit a decode error (rare): only then are deadf / errf consulted
    volatile unsigned long long wid;     // global work id = group * nunits + unit
    volatile uint32_t prog[15];          // pixels completed per plane (warp-uniform position)
    volatile uint32_t posd[15][32];      // per lane: pixels it has (possibly) written -- needed to undo a voided plane
    volatile uint8_t deadf[15][32];      // lane's plane is void because a plane above it hit a decode error
    volatile uint8_t errf[15][32];       // lane hit a decode error in this plane
};
struct DecShared {
    SmemTables T;
    DecSlot slot[DEC_SLOTS];
    uint32_t next_task;
    volatile uint32_t fetch_turn;        // chain ordinal whose first task may fetch its global work id next
    volatile uint32_t k_end;             // first chain ordinal that does not exist (work exhausted)
};
// followed in dynamic shared memory by: cnt[17][DEC_BT] u32, stlow[17][DEC_BT] u32, depth[17][DEC_BT] i16

struct Dec {
    const SmemTables &T;
    const uint32_t *wp = nullptr;   // next aligned stream word to load
    int left = 0;                   // bytes from wp to the end of THIS FRAME's stream (bytes at or beyond it read as zero, like the oracle and the legacy calls)
    unsigned long long bb = 0;      // bit buffer, LSB = next stream bit
    int bbn = 0;                    // bits in bb that came from loads (zeros beyond)
    uint32_t total_bits = 0;
    uint32_t consumed = 0;          // bits consumed
    uint32_t words = 0;             // codewords consumed

    __device__ Dec(const SmemTables &tt) : T(tt) {}

    __device__ __forceinline__ uint32_t load_word() {
        uint32_t w = 0;
        if (left >= 4) w = *wp;
        else if (left > 0) {           // tail: byte loads, nothing behind `end` is touched (the caller's buffer may stop there)
            const uint8_t *bp = reinterpret_cast<const uint8_t *>(wp);
            w = bp[0];
            if (left > 1) w |= (uint32_t) bp[1] << 8;
            if (left > 2) w |= (uint32_t) bp[2] << 16;
        }
        wp++;
        left -= 4;
        return w;
    }
    __device__ __forceinline__ void start(const uint8_t *base, const uint8_t *e, uint32_t nbits) {
        total_bits = nbits;
        const uintptr_t ad = reinterpret_cast<uintptr_t>(base);
        wp = reinterpret_cast<const uint32_t *>(ad & ~(uintptr_t) 3);
        const ptrdiff_t room = e - reinterpret_cast<const uint8_t *>(wp);
        left = room > 0x7FFFFFF0 ? 0x7FFFFFF0 : (int) room;
        const int sh = (int) (ad & 3) * 8;
        bb = (unsigned long long) (load_word() >> sh);
        bbn = 32 - sh;
    }
    __device__ __forceinline__ void fill() {           // afterwards at least 33 valid bits are buffered
        if (bbn <= 32) { bb |= (unsigned long long) load_word() << bbn; bbn += 32; }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t) bb & ((1u << n) - 1u); }
    __device__ __forceinline__ void skip(uint32_t n) { bb >>= n; bbn -= (int) n; consumed += n; }

    // icer_get_bits_from_codeword / icer_pop_bits_from_codeword (icer_decoding.c:59-105): byte-chunked read that
    // fails only when a chunk is larger than the whole segment (`decoded_bits_total` is never advanced there)
    __device__ __forceinline__ int rd(uint32_t n, bool consume) {
        if (total_bits < 8) {
            uint32_t off = consumed & 7, left = n;
            while (left) {
                uint32_t take = 8 - off; if (take > left) take = left;
                if (take > total_bits) return ICER_DECODER_OUT_OF_DATA;
                left -= take; off = (off + take) & 7;
            }
        }
        const int v = (int) peek(n);
        if (consume) skip(n);
        return v;
    }

    // refill `bin` with the next codeword (icer_decoding.c:127-181); returns an error code or OK
    __device__ __forceinline__ int refill(int bin, int &depth, uint32_t &low) {
        depth = 0; low = 0;
        fill();
        if (bin > 7) {
            const uint32_t l = T.gol_l[bin], gi = T.gol_i[bin];
            if (peek(1)) {
                rd(1, true);
                depth = T.gol_m[bin];
            } else {
                uint32_t k = (__brev((uint32_t) rd(l, false) & 0xFFFFu) >> (32 - l)) & 0xFFFFu;
                if (k < gi) {
                    rd(l, true);
                } else {
                    k = (__brev((uint32_t) rd(l + 1, true) & 0xFFFFu) >> (31 - l)) & 0xFFFFu;
                    k = (k - gi) & 0xFFFFu;
                }
                depth = 1 + (int) k;
                low = 1;
            }
        } else if (bin != 0) {
            const uint32_t e = T.v2v[bin][peek(5)];
            const uint32_t L = e & 0xF;
            if (L == 0) return ICER_DECODED_INVALID_DATA;          // cannot happen: the codes are complete
            if (L >= total_bits) return ICER_DECODER_OUT_OF_DATA;  // the reference checks n + 1 >= total for every bit it reads
            low = e >> 8;
            depth = (int) ((e >> 4) & 0xF);
            const int word = (int) peek(L);
            if (word != rd(L, true)) return ICER_DECODED_INVALID_DATA;
        } else {
            const int v = rd(1, true);
            if (v == ICER_DECODER_OUT_OF_DATA) return v;
            low = v != 0;
            depth = 1;
        }
        words++;
        return ICER_RESULT_OK;
    }
};

__global__ void __launch_bounds__(DEC_BT, ICER_DEC_MINBLOCKS) k_decode(DecodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecShared &S = *reinterpret_cast<DecShared *>(smem_raw);
    const int t = threadIdx.x;
    constexpr int BT = DEC_BT;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem_raw + ((sizeof(DecShared) + 15) & ~(size_t) 15));
    uint32_t *stlow = cnt + 17 * BT;
    int16_t *depthA = reinterpret_cast<int16_t *>(stlow + 17 * BT);     // pending bits of a bin's stack: <= the largest Golomb m (512)
    {   // tables + pool state
        for (int i = t; i < 8 * 32; i += BT) (&S.T.v2v[0][0])[i] = (&c_tab.dlut[0][0])[i];
        for (int i = t; i < 45; i += BT) S.T.ctx_hv[i] = c_tab.ctx_hv[i];
        for (int i = t; i < 25; i += BT) { S.T.ctx_hh[i] = c_tab.ctx_hh[i]; S.T.ctx_sign[i] = c_tab.ctx_sign[i]; S.T.sign_pred[i] = c_tab.sign_pred[i]; }
        for (int i = t; i < 17; i += BT) { S.T.gol_m[i] = c_tab.gol_m[i]; S.T.gol_l[i] = c_tab.gol_l[i]; S.T.gol_i[i] = c_tab.gol_i[i]; }
        for (int i = t; i < 16; i += BT) S.T.cutoff[i] = c_tab.cutoff[i];
        for (int i = t; i < 17; i += BT) S.T.cutp[i] = make_uint2(i ? c_tab.cutoff[i - 1] : 0u, i < 16 ? c_tab.cutoff[i] : 65537u);
        if (t < DEC_SLOTS) { S.slot[t].chain = 0xFFFFFFFFu; S.slot[t].freed = 1; S.slot[t].done = 0; S.slot[t].anyerr = 0; }
        if (t == 0) { S.next_task = 0; S.fetch_turn = 0; S.k_end = 0xFFFFFFFFu; }
    }
    __syncthreads();
    const int lane = t & 31;
    const int P = (int) a.planes;
    const unsigned long long total = (unsigned long long) a.ngroups * a.nunits;
    const uint32_t nstreams = a.nunits * a.planes;
    const int W = (int) a.plan_w;
    const int RS = W * 32;
    const size_t plane_px = (size_t) a.plan_w * a.plan_h;
    const bool fast_ok = !(a.debug_flags & 1u);

    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(&S.next_task, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        const uint32_t k = task / (uint32_t) P;          // chain ordinal within this block
        const int j = (int) (task % (uint32_t) P);       // plane index from the top
        const int lsb = P - 1 - j;
        DecSlot &sl = S.slot[k % DEC_SLOTS];
        // ---- claim / find the slot of chain k ------------------------------------------------------------
        if (j == 0) {
            // The first task of a chain fetches the global work item and recycles the slot of chain k - DEC_SLOTS.
            // The fetch happens in CHAIN ORDER (ticket S.fetch_turn): work ids then grow with k inside a block, so "chain k
            // got no work" implies the same for every later chain -- the invariant k_end and the early exits rely on.
            // (Fetching after the slot wait let a later chain overtake an earlier one; an earlier chain that then drew
            // the end-of-work id made the warps of the later, valid chain give up: a stall seen with the 32-warp pool.)
            unsigned long long wid = 0;
            if (lane == 0) {
                while (S.fetch_turn != k) __nanosleep(40);
                wid = atomicAdd(a.work_counter, 1ull);
                if (wid >= total && S.k_end > k) S.k_end = k;
                __threadfence_block();
                S.fetch_turn = k + 1;
            }
            wid = __shfl_sync(0xffffffffu, wid, 0);
            if (wid >= total) break;
            if (lane == 0 && k >= (uint32_t) DEC_SLOTS)
                while (!(sl.chain == k - DEC_SLOTS && sl.freed)) __nanosleep(ICER_DEC_SLOTWAIT_NS);     // that chain is live: it will finish
            __syncwarp();
            __threadfence_block();
            for (int q = 0; q < 15; q++) { sl.deadf[q][lane] = 0; sl.errf[q][lane] = 0; sl.posd[q][lane] = 0; }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) {
                for (int q = 0; q < 15; q++) sl.prog[q] = 0;
                sl.done = 0; sl.anyerr = 0; sl.freed = 0; sl.wid = wid;
                __threadfence_block();
                sl.chain = k;
            }
            __syncwarp();
        } else {
            bool gone = false;
            if (lane == 0) while (sl.chain != k) { if (S.k_end <= k) { gone = true; break; } __nanosleep(ICER_DEC_CHAINWAIT_NS); }
            gone = __shfl_sync(0xffffffffu, gone ? 1 : 0, 0) != 0;
            if (gone) break;
        }
        __threadfence_block();
        const unsigned long long wid = sl.wid;
        if (wid >= total) break;                         // no more chains: every later task is void as well
        const uint32_t g = (uint32_t) (wid % a.ngroups), ui = (uint32_t) (wid / a.ngroups);      // largest chains of all groups first
        const Unit u = a.units[ui];
        const int32_t frame = a.frame_of[g * 32 + lane];
        bool alive = false;
        Dec dc(S.T);
        if (frame >= 0) {
            // a plane is decoded only if it and every plane above it are present (icer_partition.c:431)
            const uint32_t *row = a.seg_off + (size_t) frame * nstreams + ui * a.planes;
            alive = true;
            for (int q = P - 1; q > lsb; q--) alive = alive && row[q] != 0u;      // table entries are offset + 1, 0 = absent
            const uint32_t off1 = row[lsb];
            if (off1 == 0u) alive = false;
            if (alive) {
                // lib_icer's bit reader never stops at a segment's bit count (icer_decoding.c:59-105): a desynchronised stream
                // keeps reading.  Those reads end at the frame's own stream length; behind it every bit is zero.
                const uint8_t *fbase = a.streams + (size_t) frame * a.stream_stride;
                const uint64_t flen = a.stream_len[frame] < a.stream_stride ? a.stream_len[frame] : a.stream_stride;
                const uint8_t *hd = fbase + (off1 - 1);
                dc.start(hd + ICER_HDR_BYTES, fbase + flen, hd[16] | (hd[17] << 8) | (hd[18] << 16) | ((uint32_t) hd[19] << 24));
            }
        }
        uint16_t *seg = a.coefsT + (((size_t) g * a.chan + u.chan) * plane_px + (size_t) u.y * W + u.x) * 32 + lane;
        const int sw = u.w, sh = u.h, sb = u.subband;
        const uint32_t px = (uint32_t) sw * (uint32_t) sh;
        const uint32_t lag = (uint32_t) sw + 2;      // plane p stays this far behind plane p+1 (needs pixel n + w + 1, prefetches n + w + 2)
#pragma unroll
        for (int c = 0; c < 17; c++) { cnt[c * BT + t] = (2u << 16) | 4u; depthA[c * BT + t] = 0; stlow[c * BT + t] = 0; }

        uint32_t pos = 0;                  // warp-uniform
        int row = 0, col = 0;
        uint32_t a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0, c0 = 0, c1 = 0, c2 = 0, a2n = 0, b2n = 0, c2n = 0;
        uint32_t fz_left = 0, fz_run = 0;
        int fz_bin = 0;
        bool run = alive;                  // this lane still decodes
        bool dead = false;                 // voided by an error in a plane above
        uint32_t done_px = 0;              // pixels this lane has (possibly) written

        while (pos < px && __any_sync(0xffffffffu, run)) {
            uint32_t limit = px;
            if (j > 0) {
                const uint32_t up = sl.prog[j - 1];
                __threadfence_block();
                if (sl.anyerr) {     // an error above voids this plane for that frame
                    if (run && (sl.deadf[j - 1][lane] | sl.errf[j - 1][lane])) { run = false; dead = true; sl.deadf[j][lane] = 1; }
                }
                if (up < px) limit = (a.debug_flags & 2u) ? 0u : (up > lag ? up - lag : 0u);
            }
            if (pos >= limit) { __nanosleep(ICER_DEC_BACKOFF_NS); continue; }   // the plane above is not far enough ahead yet
            uint32_t stop = pos + ICER_DEC_RECHECK < limit ? pos + ICER_DEC_RECHECK : limit;      // re-check progress every ICER_DEC_RECHECK pixels
            for (; pos < stop; pos++) {
                uint16_t *p = seg + (size_t) row * RS + (size_t) col * 32;
                const bool up_ = row > 0, dn = row + 1 < sh, rt = col + 1 < sw;
                if (col == 0) {
                    a0 = b0 = c0 = 0; a1 = up_ ? p[-RS] : 0; b1 = p[0]; c1 = dn ? p[RS] : 0;
                    a2n = (up_ && rt) ? p[32 - RS] : 0; b2n = rt ? p[32] : 0; c2n = (dn && rt) ? p[32 + RS] : 0;
                }
                a2 = a2n; b2 = b2n; c2 = c2n;
                {   // fetch the column after next one pixel early (overlaps the coding of this pixel)
                    const bool rt2 = col + 2 < sw;
                    a2n = (up_ && rt2) ? p[64 - RS] : 0;
                    b2n = rt2 ? p[64] : 0;
                    c2n = (dn && rt2) ? p[64 + RS] : 0;
                }
                if (run) {
                    uint32_t v = b1;
                    const bool quiet = ((((a0 | a1 | a2 | b0 | b1 | b2 | c0 | c1 | c2) & 0x7FFF) >> lsb) == 0);
                    uint32_t todo;
                    if (quiet && fz_left) { fz_run++; fz_left--; todo = 0; }
                    else todo = (fz_run ? 1u : 0u) | 2u;
                    int cat = 0;
                    if (todo & 2u) {
                        const int msb = 31 - __clz((v & 0x7FFF) | 1);
                        cat = msb - lsb;
                        cat = cat < 0 ? 0 : (cat > 3 ? 3 : cat);
                    }
                    bool changed = false;
                    while (todo) {
                        const int st = __ffs(todo) - 1;
                        todo &= todo - 1;
                        if (st == 0) {
                            const int bin = fz_bin;
                            const uint32_t zt = cnt[t];
                            uint32_t z = CNT_Z(zt) + fz_run, tot = CNT_TOT(zt) + fz_run;
                            if (tot >= 500) { tot >>= 1; if (z > tot) z >>= 1; }
                            cnt[t] = CNT_PACK(bin, z, tot);
                            depthA[bin * BT + t] -= (int) fz_run;
                            fz_run = 0; fz_left = 0;
                            continue;
                        }
                        int bin, ctx = -1, si = 0;
                        bool inv = false;
                        uint32_t z = 0, tot = 0;
                        if (st == 1 && cat == 3) {
                            bin = 0;
                        } else {
                            if (st == 1) ctx = mag_context(S.T, cat, lsb, sb, a0, a1, a2, b0, b2, c0, c1, c2);
                            else { si = sign_index(lsb, sb, a1, b0, b2, c1); ctx = S.T.ctx_sign[si]; }
                            const uint32_t zt = cnt[ctx * BT + t];
                            z = CNT_Z(zt); tot = CNT_TOT(zt);
                            uint32_t zz = z;
                            inv = zz < (tot >> 1);
                            if (inv) zz = tot - zz;
                            bin = bin_cached(S.T, zz, tot, CNT_BIN(zt));
                        }
                        // next decoded bit of `bin` (icer_decoding.c:127, :186-192)
                        int depth = depthA[bin * BT + t];
                        const uint32_t slw = stlow[bin * BT + t];
                        uint32_t low = slw >> 24;
                        bool refilled = false;
                        if (depth <= 0 || dc.words - (slw & 0xFFFFFFu) >= ICER_RING) {
                            const int rc = dc.refill(bin, depth, low);
                            if (rc != ICER_RESULT_OK) { run = false; sl.errf[j][lane] = 1; sl.anyerr = 1; todo = 0; break; }
                            refilled = true;
                        }
                        const uint32_t stamp = refilled ? (dc.words & 0xFFFFFFu) : (slw & 0xFFFFFFu);
                        if (st == 1 && quiet && !inv && bin > 7 && fast_ok) {
                            // quiet fast path: the run of pending zeros of this Golomb bin is taken without touching the stack
                            const uint32_t zeros = (uint32_t) depth - (low & 1u);
                            if (zeros > 0) {
                                const uint32_t n2 = zeros_in_same_bin(S.T, bin, z, tot);
                                const uint32_t n3 = 500u - tot;
                                depthA[bin * BT + t] = (int16_t) depth;
                                stlow[bin * BT + t] = stamp | (low << 24);
                                fz_left = min(zeros, min(n2, n3)) - 1;
                                fz_run = 1;
                                fz_bin = bin;
                                continue;
                            }
                        }
                        const int top = depth - 1;
                        uint32_t b = 0;
                        if (top >= 0 && top < 8) { b = (low >> top) & 1u; low &= ~(1u << top); }
                        depthA[bin * BT + t] = (int16_t) (depth - 1);
                        stlow[bin * BT + t] = stamp | (low << 24);
                        const uint32_t bit = inv ? (b ^ 1u) : b;
                        if (ctx >= 0) {
                            tot++;
                            z += bit ^ 1u;
                            if (tot >= 500) { tot >>= 1; if (z > tot) z >>= 1; }
                            cnt[ctx * BT + t] = CNT_PACK(bin, z, tot);
                        }
                        if (st == 1) {
                            if (bit) { v |= 1u << lsb; changed = true; if (cat == 0) todo |= 4u; }
                        } else {
                            if ((bit ^ S.T.sign_pred[si]) & 1u) { v |= 0x8000u; changed = true; }
                        }
                    }
                    if (changed) { *p = (uint16_t) v; b1 = v; }      // also stores the magnitude bit when the sign read failed
                    done_px = pos + 1;
                }
                a0 = a1; a1 = a2; b0 = b1; b1 = b2; c0 = c1; c1 = c2;
                if (++col == sw) { col = 0; row++; }
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) sl.prog[j] = pos;
        }
        // ---- task finished -------------------------------------------------------------------------------
        sl.posd[j][lane] = done_px;
        __threadfence_block();
        __syncwarp();
        uint32_t old = 0;
        if (lane == 0) {
            sl.prog[j] = 0xFFFFFFFFu;                 // never hold the planes below
            __threadfence_block();
            old = atomicAdd((uint32_t *) &sl.done, 1u);
        }
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == (uint32_t) P - 1) {
            // last plane of the chain to finish.  A decode error voids every plane below it (lib_icer stops the chain
            // there and never decodes them, icer_partition.c:441); those planes ran ahead in the pipeline: take their
            // bits back, then clear the sign of coefficients whose magnitude is now zero.  Lane = frame.
            __threadfence_block();
            if (sl.anyerr) {
                bool voided = false, any_void = false;
                for (int q = 0; q < P; q++) {
                    if (q > 0 && (voided || sl.errf[q - 1][lane])) voided = true;     // everything below an error
                    if (!voided) continue;
                    const uint32_t npx = sl.posd[q][lane];
                    if (npx) any_void = true;
                    const uint32_t bitm = ~(1u << (P - 1 - q));
                    for (uint32_t n = 0; n < npx; n++) {
                        uint16_t *qq = seg + (size_t) (n / (uint32_t) sw) * RS + (size_t) (n % (uint32_t) sw) * 32;
                        *qq = (uint16_t) (*qq & bitm);
                    }
                }
                if (any_void)
                    for (uint32_t n = 0; n < px; n++) {
                        uint16_t *qq = seg + (size_t) (n / (uint32_t) sw) * RS + (size_t) (n % (uint32_t) sw) * 32;
                        if ((*qq & 0x7FFF) == 0 && *qq != 0) *qq = 0;
                    }
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) sl.freed = 1;
        }
    }
}

// =========================================================================================================
// K5: quota scan, then gather of header + payload into the output streams
// =========================================================================================================

// CRC-32 of every payload (icer_partition.c:330, icer_util.c:64-66): one thread per stream, word loads
__global__ void __launch_bounds__(128) k_payload_crc(AssembleArgs a) {
    __shared__ uint32_t crc[256];
    for (int i = threadIdx.x; i < 256; i += 128) crc[i] = c_tab.crc[i];
    __syncthreads();
    const unsigned long long id = (unsigned long long) blockIdx.x * 128 + threadIdx.x;
    if (id >= (unsigned long long) a.nframes * a.nstreams) return;
    const uint32_t f = (uint32_t) (id / a.nstreams), s = (uint32_t) (id % a.nstreams);
    if (a.bits[id] == ICER_BITS_NOT_CODED) return;
    const uint32_t nbytes = (a.bits[id] + 7) >> 3;
    const uint32_t *w = a.scratch + (size_t) f * a.scratch_words + a.soff[s];
    uint32_t c = 0xFFFFFFFFu, i = 0;
    for (; i + 4 <= nbytes; i += 4) {
        const uint32_t v = w[i >> 2];
        c = crc[(c ^ v) & 0xFF] ^ (c >> 8);
        c = crc[(c ^ (v >> 8)) & 0xFF] ^ (c >> 8);
        c = crc[(c ^ (v >> 16)) & 0xFF] ^ (c >> 8);
        c = crc[(c ^ (v >> 24)) & 0xFF] ^ (c >> 8);
    }
    if (i < nbytes) {
        const uint32_t v = w[i >> 2];
        for (uint32_t k = 0; i < nbytes; i++, k++) c = crc[(c ^ (v >> (8 * k))) & 0xFF] ^ (c >> 8);
    }
    a.crc[id] = ~c;
}

// One warp per frame walks the streams in priority order with lib_icer's quota rule (SURVEY A.14;
// icer_encoding.c:211-214, :133-135; icer_partition.c:323-326): 32 streams per step, a warp prefix sum gives the
// bytes used before each stream, the first lane whose stream does not fit ends the walk.  A second walk in
// rearranged order assigns the offsets of the final stream.
__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(128) k_quota_scan(AssembleArgs a) {
    const uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= a.nframes) return;
    int rc = ICER_RESULT_OK;
    for (uint32_t c = 0; c < a.chan && rc == ICER_RESULT_OK; c++)
        if (a.ovf[(size_t) f * a.chan + c]) rc = ICER_INTEGER_OVERFLOW;
    if (rc == ICER_RESULT_OK) rc = a.plan_status;      // ICER_PACKET_COUNT_EXCEEDED: found after the transform
    const uint32_t *bits = a.bits + (size_t) f * a.nstreams;
    uint64_t *offp = a.off_prio + (size_t) f * a.nstreams;     // indexed by STREAM ID; ~0 = dropped
    uint64_t used = 0;
    uint32_t kept = 0;
    bool stop = rc != ICER_RESULT_OK;
    for (uint32_t base = 0; base < a.nstreams; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < a.nstreams;
        const uint32_t s = valid ? a.prio[i] : 0;
        const uint32_t nb = valid ? bits[s] : 0;
        const uint64_t sz = valid ? ICER_HDR_BYTES + ((nb + 7) >> 3) : 0;
        const uint64_t inc = warp_incl_scan(sz, lane);
        const uint64_t before = used + inc - sz;
        uint32_t nkeep = 0;
        if (!stop) {
            bool fail = false;
            if (valid) {
                if (before > a.quota) fail = true;
                else {
                    const uint64_t room = a.quota - before;
                    fail = room < ICER_HDR_BYTES || (nb > 0 && (uint64_t) (nb >> 3) >= room - ICER_HDR_BYTES);
                }
            }
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, fail);
            if (m) {
                const int first = __ffs(m) - 1;
                nkeep = first;
                used = __shfl_sync(0xFFFFFFFFu, before, first);
                stop = true;
                rc = ICER_BYTE_QUOTA_EXCEEDED;
            } else {
                nkeep = a.nstreams - base < 32 ? a.nstreams - base : 32;
                used += __shfl_sync(0xFFFFFFFFu, inc, 31);
            }
            kept += nkeep;
        }
        if (valid) offp[s] = (uint32_t) lane < nkeep ? before : ~0ull;
    }
    __syncwarp();
    uint64_t *offr = a.off_rearr + (size_t) f * a.nstreams;
    uint64_t ro = 0;
    for (uint32_t base = 0; base < a.nstreams; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < a.nstreams;
        const uint32_t s = valid ? a.rearr[i] : 0;
        const bool keep = valid && offp[s] != ~0ull;
        const uint64_t sz = keep ? ICER_HDR_BYTES + ((bits[s] + 7) >> 3) : 0;
        const uint64_t inc = warp_incl_scan(sz, lane);
        if (valid) offr[s] = keep ? ro + inc - sz : ~0ull;
        ro += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (lane == 0) {
        a.size_used[f] = (rc == ICER_RESULT_OK || rc == ICER_BYTE_QUOTA_EXCEEDED) ? used : 0;
        a.rc[f] = rc;
        a.keep[f] = kept;
        atomicMax(&a.stat[0], kept);
        // priority-prefix encoding: the walk reached the first stream the encoder has not produced yet -> code more, walk again
        if (kept >= a.limit && a.limit < a.nstreams && (rc == ICER_RESULT_OK || rc == ICER_BYTE_QUOTA_EXCEEDED)) atomicOr(&a.stat[1], 1u);
    }
}

__device__ __forceinline__ uint32_t crc_bytes(const uint8_t *p, int n) {
    uint32_t c = 0xFFFFFFFFu;
    for (int i = 0; i < n; i++) c = c_tab.crc[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return ~c;
}

// One warp per (frame, stream): builds the 28-byte header (icer_encoding.c:210-234, icer_partition.c:328-332)
// and copies header + payload to the rearranged stream (and the priority-ordered one when requested).
__global__ void __launch_bounds__(256) k_gather(AssembleArgs a) {
    const unsigned long long wid = ((unsigned long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= (unsigned long long) a.nframes * a.nstreams) return;
    const uint32_t f = (uint32_t) (wid / a.nstreams), s = (uint32_t) (wid % a.nstreams);
    const uint64_t orr = a.off_rearr[(size_t) f * a.nstreams + s];
    if (orr == ~0ull) return;
    const uint64_t opr = a.off_prio[(size_t) f * a.nstreams + s];
    const uint32_t nb = a.bits[(size_t) f * a.nstreams + s];
    const uint32_t nbytes = (nb + 7) >> 3;
    const Unit u = a.units[s / a.planes];
    const uint32_t lsb = s % a.planes;
    __shared__ uint8_t hdr_s[8][ICER_HDR_BYTES + 4];
    uint8_t *hdr = hdr_s[threadIdx.x >> 5];
    if (lane == 0) {
        const uint32_t mean = a.mean[(size_t) f * a.chan + u.chan] & 0xFF;    // passes through a uint8_t (icer.h:270)
        hdr[0] = 0x5B; hdr[1] = 0x60; hdr[2] = (uint8_t) mean; hdr[3] = 0;
        hdr[4] = u.level; hdr[5] = u.subband; hdr[6] = u.seg; hdr[7] = (uint8_t) (lsb | (u.chan << 4));
        const uint32_t w = a.w, h = a.h, dc = a.crc[(size_t) f * a.nstreams + s];
        for (int i = 0; i < 4; i++) {
            hdr[8 + i] = (uint8_t) (w >> (8 * i)); hdr[12 + i] = (uint8_t) (h >> (8 * i));
            hdr[16 + i] = (uint8_t) (nb >> (8 * i)); hdr[20 + i] = (uint8_t) (dc >> (8 * i));
        }
        const uint32_t hc = crc_bytes(hdr, 24);
        for (int i = 0; i < 4; i++) hdr[24 + i] = (uint8_t) (hc >> (8 * i));
    }
    __syncwarp();
    const uint8_t *pay = reinterpret_cast<const uint8_t *>(a.scratch + (size_t) f * a.scratch_words + a.soff[s]);
    uint8_t *dst = a.out + (size_t) f * a.out_stride + orr;
    if (lane < ICER_HDR_BYTES) dst[lane] = hdr[lane];
    for (uint32_t i = lane; i < nbytes; i += 32) dst[ICER_HDR_BYTES + i] = pay[i];
    if (a.out_prio) {
        uint8_t *d2 = a.out_prio + (size_t) f * a.out_stride + opr;
        if (lane < ICER_HDR_BYTES) d2[lane] = hdr[lane];
        for (uint32_t i = lane; i < nbytes; i += 32) d2[ICER_HDR_BYTES + i] = pay[i];
    }
}

}  // namespace

bool frames_to_T_fuses_signmag(const uint16_t *frames, const uint16_t *T, uint32_t w, uint32_t h) {
    return (((size_t) w * h) % 8) == 0 && ((uintptr_t) frames % 16) == 0 && ((uintptr_t) T % 16) == 0;
}

uint32_t encode_resident_threads(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int bps = 0;
    cudaFuncSetAttribute(k_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(EncSmem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_encode, NT, sizeof(EncSmem));
    if (bps < 1) bps = 1;
    return (uint32_t) (sms * bps * NT);
}

int launch_frames_to_T(const uint16_t *frames, uint16_t *T, uint32_t nframes, uint32_t ngroups, uint32_t chan, uint32_t w, uint32_t h,
                       cudaStream_t st, LaunchStats &ls, bool signmag) {
    const size_t plane_px = (size_t) w * h;
    if ((plane_px % 8) == 0 && ((uintptr_t) frames % 16) == 0 && ((uintptr_t) T % 16) == 0) {
        for (uint32_t g0 = 0; g0 < ngroups; g0 += 32768) {
            const uint32_t cnt = ngroups - g0 < 32768 ? ngroups - g0 : 32768;
            dim3 grid((unsigned) ((plane_px + TT_PX - 1) / TT_PX), chan, cnt);
            const uint32_t f0 = g0 * 32;
            const uint16_t *src = frames + (size_t) f0 * chan * plane_px;
            uint16_t *dstT = T + (size_t) g0 * chan * plane_px * 32;
            if (signmag) k_frames_to_T_v<true><<<grid, 256, 0, st>>>(src, dstT, nframes - f0, chan, (uint32_t) plane_px);
            else k_frames_to_T_v<false><<<grid, 256, 0, st>>>(src, dstT, nframes - f0, chan, (uint32_t) plane_px);
            ls.launches++;
        }
        CUDA_TRY(cudaGetLastError());
        return ICER_RESULT_OK;
    }
    if (signmag) return ICER_INVALID_INPUT;      // the scalar kernel expects sign-magnitude input (callers check frames_to_T_fuses_signmag)
    for (uint32_t g0 = 0; g0 < ngroups; g0 += 32768) {
        const uint32_t cnt = ngroups - g0 < 32768 ? ngroups - g0 : 32768;
        dim3 grid((unsigned) ((plane_px + 31) / 32), chan, cnt);
        const uint32_t f0 = g0 * 32;
        k_frames_to_T<<<grid, 256, 0, st>>>(frames + (size_t) f0 * chan * plane_px, T + (size_t) g0 * chan * plane_px * 32,
                                            nframes - f0, chan, (uint32_t) plane_px);
        ls.launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}

int launch_T_to_frames(const uint16_t *T, uint16_t *frames, const int32_t *frame_of, const uint16_t *mean, uint32_t ngroups,
                       uint32_t chan, uint32_t w, uint32_t h, uint64_t cap_px, uint32_t llw, uint32_t llh, cudaStream_t st,
                       LaunchStats &ls, bool bits8) {
    const size_t plane_px = (size_t) w * h;
    if (!bits8 && (plane_px % 8) == 0 && (cap_px % 8) == 0 && (w % 8) == 0 && ((uintptr_t) frames % 16) == 0 && ((uintptr_t) T % 16) == 0) {
        for (uint32_t g0 = 0; g0 < ngroups; g0 += 32768) {
            const uint32_t cnt = ngroups - g0 < 32768 ? ngroups - g0 : 32768;
            dim3 grid((unsigned) ((plane_px + TT_PX - 1) / TT_PX), chan, cnt);
            k_T_to_frames_v<<<grid, 256, 0, st>>>(T + (size_t) g0 * chan * plane_px * 32, frames, frame_of + (size_t) g0 * 32, mean, chan,
                                                  (uint32_t) plane_px, cap_px, (int) w, (int) llw, (int) llh);
            ls.launches++;
        }
        CUDA_TRY(cudaGetLastError());
        return ICER_RESULT_OK;
    }
    for (uint32_t g0 = 0; g0 < ngroups; g0 += 32768) {
        const uint32_t cnt = ngroups - g0 < 32768 ? ngroups - g0 : 32768;
        dim3 grid((unsigned) ((plane_px + 31) / 32), chan, cnt);
        k_T_to_frames<<<grid, 256, 0, st>>>(T + (size_t) g0 * chan * plane_px * 32, frames, frame_of + (size_t) g0 * 32, mean, chan,
                                            (uint32_t) plane_px, cap_px, (int) w, (int) llw, (int) llh, bits8 ? 1 : 0);
        ls.launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}

int launch_encode(const EncodeArgs &a, cudaStream_t st, LaunchStats &ls) {
    int rc = upload_tables();
    if (rc) return rc;
    CUDA_TRY(cudaFuncSetAttribute(k_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(EncSmem)));
    const unsigned long long total = (unsigned long long) a.ngroups * (a.list ? a.nlist : a.nunits * a.planes);   // warps of work
    unsigned blocks = a.nring_threads / NT;
    const unsigned long long need = (total + (NT / 32) - 1) / (NT / 32);
    if (need < blocks) blocks = (unsigned) need;
    if (blocks == 0) return ICER_RESULT_OK;
    k_encode<<<blocks, NT, sizeof(EncSmem), st>>>(a);
    ls.launches++;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}

int launch_assemble(const AssembleArgs &a, cudaStream_t st, LaunchStats &ls) {
    int rc = upload_tables();
    if (rc) return rc;
    const unsigned long long nst = (unsigned long long) a.nframes * a.nstreams;
    const bool coded = a.plan_status == ICER_RESULT_OK;     // otherwise the coder did not run: only the per-frame status is produced
    if (a.phase != 2) {
        if (nst && coded) { k_payload_crc<<<(unsigned) ((nst + 127) / 128), 128, 0, st>>>(a); ls.launches++; }
        k_quota_scan<<<(a.nframes + 3) / 4, 128, 0, st>>>(a);
        ls.launches++;
    }
    const unsigned long long warps = (unsigned long long) a.nframes * a.nstreams;
    const unsigned long long blocks = (warps * 32 + 255) / 256;
    if (blocks > 0x7FFFFFFFull) return ICER_FATAL_ERROR;
    if (a.phase != 1 && blocks && coded) { k_gather<<<(unsigned) blocks, 256, 0, st>>>(a); ls.launches++; }
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}

int launch_decode(const DecodeArgs &a, cudaStream_t st, LaunchStats &ls) {
    int rc = upload_tables();
    if (rc) return rc;
    if (a.planes < 1 || a.planes > 15) return ICER_BITPLANE_OUT_OF_RANGE;
    const size_t smem = ((sizeof(DecShared) + 15) & ~(size_t) 15) + (size_t) 17 * DEC_BT * 10;
    CUDA_TRY(cudaFuncSetAttribute(k_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int dev = 0, sms = 148, bps = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_decode, DEC_BT, smem);
    if (bps < 1) bps = 1;
    // a block of DEC_WARPS warps keeps about DEC_WARPS / planes chains in flight
    const unsigned long long total = (unsigned long long) a.ngroups * a.nunits;
    unsigned long long blocks = (total * a.planes + DEC_WARPS - 1) / DEC_WARPS;
    if (blocks > (unsigned long long) sms * bps) blocks = (unsigned long long) sms * bps;
    if (blocks == 0) return ICER_RESULT_OK;
    k_decode<<<(unsigned) blocks, DEC_BT, smem, st>>>(a);
    ls.launches++;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
