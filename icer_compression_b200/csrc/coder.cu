// coder.cu -- K3 (bit slicing), K4 (bit-plane context model + interleaved entropy ENCODER), K7 (the DECODER), K8 (un-slicing) and
// K5 (segment assembly).
//
// Granularity.  lib_icer resets its context model and entropy coder for every (channel, level, subband, bit plane, segment)
// (icer_partition.c:315-322), so each such STREAM is independent on the encode side.  On the decode side the planes of one
// segment form a chain (MSB -> LSB, icer_partition.c:427-443), but plane p only needs plane p+1 to be a row + 1 pixel ahead (its
// right/down neighbours are taken one plane up, icer_context_modeller.c:510-521).
//
// Mapping (both coder kernels): a WARP per (group of 32 frames, stream), lane = frame.  The lanes run the per-lane state machines
// of coder_core.h on BIT SLICES of their frame: per chunk of 30 pixels of a segment row a handful of 32-bit words (magnitude bit,
// "significant at plane p", sign), stored [group][slot][field][lane] so that a warp's load of one field is one 128-byte line.  A lane
// steps through its chunk at its own pace (a sign decision, a codeword completion or a quiet run cost that lane a step, not the
// whole warp); the warp only re-converges at chunk borders.  K3 produces the slices from the planar wavelet coefficients, K8 turns
// the slices the decoder wrote back into planar coefficients.
//
// Reference behaviour restated (lib_icer/src): see coder_core.h; icer_partition.c:328-334 + icer_encoding.c:210-234 (segment
// header, quota) for K5.
#include "common.h"
#include <mutex>

namespace {

using icer_core::CoreTables;

#ifndef ICER_ENC_NT
#define ICER_ENC_NT 1024      /* one block per SM: measured on config 2, k_encode ms with 32 / 64 / 128 / 256 / 512 / 1024 threads per block (same 32 warps per SM): 21.3 / 20.7 / 20.7 / 20.2 / 19.4 / 19.3 */
#endif
constexpr int NT = ICER_ENC_NT;   // threads per encoder block
#ifndef ICER_ENC_MINBLOCKS
#define ICER_ENC_MINBLOCKS 1
#endif

struct DevTables {
    CoreTables enc, dec;
    uint32_t crc[256];
};
__constant__ DevTables c_tab;
// CRC-32 of the 24 header bytes without a serial loop: the register update is affine over GF(2), so
//     crc = ~(K ^ XOR_i P[i][byte_i])        with P[i][b] = A^(23-i)(tab[b]),  A(r) = tab[r & 0xFF] ^ (r >> 8),  K = A^24(0xFFFFFFFF)
// -- one table look-up per byte position (24 lanes) and a warp XOR reduction.  Global memory: the index differs in every lane.
__device__ uint32_t g_hdr_crc_pos[24][256];
__device__ uint32_t g_hdr_crc_k;
// __constant__ memory has one instance PER DEVICE: the upload is tracked per device ordinal (a process may hold contexts on
// several GPUs, e.g. icer_b200_*_batch_multi)
std::mutex g_tab_mu;
bool g_tables_uploaded[64] = {false};

int upload_tables() {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return ICER_FATAL_ERROR;
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (g_tables_uploaded[dev]) return ICER_RESULT_OK;
    const HostTables &h = icer_host_tables();
    static DevTables d;
    icer_tables::fill_core_tables(d.enc, h, false);
    icer_tables::fill_core_tables(d.dec, h, true);
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        d.crc[n] = c;
    }
    CUDA_TRY(cudaMemcpyToSymbol(c_tab, &d, sizeof d));      // synchronous: complete before any later launch on any stream
    {
        static uint32_t pos[24][256];
        auto A = [&](uint32_t r) { return d.crc[r & 0xFF] ^ (r >> 8); };
        for (int i = 0; i < 24; i++)
            for (uint32_t b = 0; b < 256; b++) {
                uint32_t r = d.crc[b];
                for (int k = 0; k < 23 - i; k++) r = A(r);
                pos[i][b] = r;
            }
        uint32_t kk = 0xFFFFFFFFu;
        for (int k = 0; k < 24; k++) kk = A(kk);
        CUDA_TRY(cudaMemcpyToSymbol(g_hdr_crc_pos, pos, sizeof pos));
        CUDA_TRY(cudaMemcpyToSymbol(g_hdr_crc_k, &kk, sizeof kk));
    }
    g_tables_uploaded[dev] = true;
    return ICER_RESULT_OK;
}

// tables every lane needs with divergent indices: staged in shared memory once per block
__device__ __forceinline__ void stage_tables(CoreTables &T, const CoreTables &src, int nthreads) {
    static_assert(sizeof(CoreTables) % 4 == 0, "CoreTables is copied word by word");
    uint32_t *d = reinterpret_cast<uint32_t *>(&T);
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&src);
    for (int i = threadIdx.x; i < (int) (sizeof(CoreTables) / 4); i += nthreads) d[i] = s[i];
}

// =========================================================================================================
// Bit slices
// =========================================================================================================
// A SLOT is one chunk of ICER_CHUNK (30) pixels of one row of one unit (segment); slots of a unit are numbered row-major from the
// unit's slot0.  For every slot the slice buffer holds F fields of 32 words (one per frame of the group):
//     [group][slot][field][lane]      word bit b <-> plane column x0 - 1 + b  (bit 0 / bit 31: the neighbours left / right of the chunk,
//                                                                                zero outside the unit)
// Encoder fields (K3 writes, K4 reads), P = planes coded, PS = min(P + 2, 14):
//     B_p (p < P): p            S_p (1 <= p <= PS): P + p - 1            N (sign): P + PS                F = P + PS + 1
//     (S_0 = S_1 | B_0; S_p = 0 for p > 14: magnitudes have 15 bits; the coder of plane lsb needs S up to lsb + 3, and the bits ABOVE the
//      top coded plane matter: they make encoder and decoder disagree for such coefficients exactly like lib_icer's do)
// Decoder fields (K7 writes and reads, K8 reads):
//     B_p: p            S_p (0 <= p < P): P + p            N_p (signs known after plane p, 0 <= p < P): 2P + p            F = 3P
//     (nothing is known above the top coded plane: S_p = N_p = 0 for p >= P)
struct SliceGeom { int P, PS, F; };
__host__ __device__ inline SliceGeom enc_geom(int P) { SliceGeom g; g.P = P; g.PS = P + 2 < 14 ? P + 2 : 14; g.F = P + g.PS + 1; return g; }
__host__ __device__ inline int dec_fields(int P) { return 3 * P; }

template <bool B8>
__device__ __forceinline__ uint32_t to_signmag16(uint32_t v) {          // icer_wavelet.c:871-877 (uint8 codec: :851-857, sign kept in bit 15)
    const int s = (int) (int16_t) v;
    return s < 0 ? (uint32_t) ((-s) & (B8 ? 0x7F : 0x7FFF)) | 0x8000u : (uint32_t) s;
}

// 16 x 16 bit-matrix transpose, done in BOTH 16-bit halves of 16 registers at once: afterwards bit i of half h of r[p] is what
// bit p of half h of r[i] was.  With r[i] = pixel i | pixel (i + 16) << 16 of a 32-pixel window this turns 16-bit samples into
// bit planes (word p = bit p of the 32 pixels) and back -- the operation is its own inverse.
template <int J>
__device__ __forceinline__ void transpose_stage(uint32_t (&r)[16]) {
    constexpr uint32_t m = J == 8 ? 0x00FF00FFu : (J == 4 ? 0x0F0F0F0Fu : (J == 2 ? 0x33333333u : 0x55555555u));
#pragma unroll
    for (int idx = 0; idx < 8; idx++) {
        const int k = ((idx & ~(J - 1)) << 1) | (idx & (J - 1));       // the rows whose bit J is clear
        const uint32_t t = ((r[k] >> J) ^ r[k + J]) & m;
        r[k + J] ^= t;
        r[k] ^= t << J;
    }
}
__device__ __forceinline__ void transpose16x16_halves(uint32_t (&r)[16]) {
    transpose_stage<8>(r); transpose_stage<4>(r); transpose_stage<2>(r); transpose_stage<1>(r);
}

// K3: planar coefficients -> encoder slices.  Warp = (group, slot), LANE = FRAME: a lane reads the 32 window samples of its frame
// (17 aligned 32-bit words), turns them into bit planes with the register transpose above, derives the S_p words as running ORs and
// stores every field as one coalesced 128-byte line of the slice buffer.
template <int P>
__global__ void __launch_bounds__(128) k_bitslice(const uint16_t *__restrict__ coefs, uint32_t *__restrict__ bs, const SlotDesc *__restrict__ slots,
                                                  uint32_t nslots, uint32_t nframes, uint32_t chan, uint32_t W, uint32_t H, int twos) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = blockIdx.y;
    const uint32_t frame = g * 32 + lane;
    constexpr int PS = P + 2 < 14 ? P + 2 : 14, F = P + PS + 1;
    const size_t plane_px = (size_t) W * H;
    const long long total = (long long) nframes * chan * (long long) plane_px;       // samples in the buffer
    const uint32_t *src32 = reinterpret_cast<const uint32_t *>(coefs);
    uint32_t *dstg = bs + (size_t) g * nslots * F * 32 + lane;
    for (uint32_t slot = blockIdx.x * 4 + warp; slot < nslots; slot += gridDim.x * 4) {
        const SlotDesc d = slots[slot];
        const uint32_t ch = (d.flags >> 2) & 3u;
        uint32_t r[16];
        if (frame < nframes) {
            const long long e0 = ((long long) frame * chan + ch) * (long long) plane_px + (long long) d.y * W + d.x0 - 1;     // first window sample (may lie one before / after the buffer)
            const long long w0 = e0 >> 1;
            const uint32_t sh = (uint32_t) (e0 & 1) * 16u;
            uint32_t w[17];
#pragma unroll
            for (int k = 0; k < 17; k++) {
                const long long wi = w0 + k;
                uint32_t v = 0;
                if (wi >= 0 && 2 * wi + 1 < total) v = src32[wi];
                else if (wi >= 0 && 2 * wi < total) v = coefs[2 * wi];
                w[k] = v;
            }
#pragma unroll
            for (int i = 0; i < 16; i++) {
                // window sample i -> low half, sample i + 16 -> high half
                const uint32_t va = __funnelshift_r(w[i >> 1], w[(i >> 1) + 1], sh), vb = __funnelshift_r(w[(i >> 1) + 8], w[(i >> 1) + 9], sh);
                r[i] = __byte_perm(va, vb, (i & 1) ? 0x7632 : 0x5410);
            }
            if (twos) {
#pragma unroll
                for (int i = 0; i < 16; i++) {        // two's complement -> sign-magnitude in both halves (icer_wavelet.c:871-877)
                    const uint32_t sg = (r[i] >> 15) & 0x00010001u;
                    const uint32_t t = (r[i] ^ (sg * 0xFFFFu)) + sg;
                    r[i] = (t & 0x7FFF7FFFu) | (sg << 15);
                }
            }
            transpose16x16_halves(r);              // r[p] = bit p of the 32 window samples (p = 15: the sign)
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) r[i] = 0;
        }
        // window columns that belong to the unit: 1..ncols, the left / right neighbour when it exists
        uint32_t vmask = ((2u << d.ncols) - 2u) | ((d.flags & 1u) ? 1u : 0u) | ((d.flags & 2u) ? (2u << d.ncols) : 0u);
        uint32_t *dst = dstg + (size_t) slot * F * 32;
        uint32_t sig = 0;
#pragma unroll
        for (int p = 14; p >= 0; p--) {
            sig |= r[p];                            // S_p = S_{p+1} | B_p
            if (p < P) dst[p * 32] = r[p] & vmask;
            if (p >= 1 && p <= PS) dst[(P + p - 1) * 32] = sig & vmask;
        }
        dst[(P + PS) * 32] = r[15] & vmask;
    }
}

// K8a: which planes of a (frame, unit) count.  A plane is decoded only if it and every plane above it are present
// (icer_partition.c:431), and lib_icer stops a chain at the first plane that fails (:441) -- that plane keeps the bits it decoded
// before the error; the planes below were decoded ahead of time by the pipeline and are ignored.  vp = lowest plane that counts.
enum { PST_OK = 0, PST_ERROR = 1, PST_ABSENT = 2 };
__global__ void __launch_bounds__(256) k_valid_planes(const uint8_t *__restrict__ pstat, uint8_t *__restrict__ vp, size_t count, int P) {
    const size_t i = (size_t) blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    int v = 0;
    for (int p = P - 1; p >= 0; p--) {
        const uint32_t st = pstat[i * P + p];
        if (st == PST_ABSENT) { v = p + 1; break; }
        if (st == PST_ERROR) { v = p; break; }
    }
    vp[i] = (uint8_t) v;
}

// K8: decoder slices -> planar coefficients.  Warp = (group, slot).  Load phase, lane = frame: the lane loads its B words (coalesced)
// and the register transpose turns them back into 32 sign-magnitude samples.  Store phase, lane = window column: the samples cross a
// shared-memory tile so that every store instruction writes the 60 contiguous bytes of ONE frame's row segment; on the way they
// are converted (icer_wavelet.c:880-886) and get the header LL mean (icer_compress.c:520-531).
template <int P>
__global__ void __launch_bounds__(128) k_unslice(const uint32_t *__restrict__ bs, uint16_t *__restrict__ dst, const SlotDesc *__restrict__ slots,
                                                 const uint8_t *__restrict__ vpmap, const int32_t *__restrict__ frame_of,
                                                 const uint16_t *__restrict__ mean, uint32_t nslots, uint32_t nunits, uint32_t chan, uint32_t W, uint32_t H,
                                                 uint64_t cap_px, int llw, int llh, int bits8) {
    __shared__ uint32_t tile_s[4][32][17];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t (*tile)[17] = tile_s[warp];
    const uint32_t g = blockIdx.y;
    const int32_t frame = frame_of[g * 32 + lane];
    constexpr int F = 3 * P;
    const uint32_t *srcg = bs + (size_t) g * nslots * F * 32 + lane;
    for (uint32_t slot = blockIdx.x * 4 + warp; slot < nslots; slot += gridDim.x * 4) {
        const SlotDesc d = slots[slot];
        const uint32_t ch = (d.flags >> 2) & 3u;
        {
            uint32_t r[16];
            if (frame >= 0) {
                const int vp = vpmap[(size_t) frame * nunits + d.unit];
                const uint32_t *src = srcg + (size_t) slot * F * 32;
#pragma unroll
                for (int p = 0; p < 15; p++) r[p] = (p < P && p >= vp) ? src[p * 32] : 0u;
                r[15] = vp < P ? src[(2 * P + vp) * 32] : 0u;
                transpose16x16_halves(r);              // r[i] = sample i | sample (i + 16) << 16, sign-magnitude
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) r[i] = 0;
            }
#pragma unroll
            for (int i = 0; i < 16; i++) tile[lane][i] = r[i];
        }
        __syncwarp();
        const int y = (int) d.y, x = (int) d.x0 - 1 + lane;
        const bool col_ok = lane >= 1 && lane <= (int) d.ncols;
        const bool in_ll = y < llh && x < llw;
        for (int f = 0; f < 32; f++) {
            const int32_t fr = __shfl_sync(0xffffffffu, frame, f);
            if (fr < 0 || !col_ok) continue;
            uint32_t v = (tile[f][lane & 15] >> ((lane >> 4) * 16)) & 0xFFFFu;
            if (!(v & 0x7FFFu)) v = 0;                                              // a sign without magnitude cannot occur; be safe
            v = (v & 0x8000u) ? ((0x8000u - v) & 0xFFFFu) : v;                      // icer_wavelet.c:880-886
            if (in_ll) {
                const int m = (int) mean[(size_t) fr * chan + ch];
                // uint16 API: int16 add (icer_compress.c:524-531); native uint8 API: int8 add of the header value's low byte (:258-265)
                v = bits8 ? (uint32_t) (uint16_t) (int16_t) (int8_t) ((int) (int16_t) v + (int) (int8_t) m)
                          : (uint32_t) (uint16_t) (int16_t) ((int) (int16_t) v + (int) (int16_t) m);
            }
            dst[((size_t) fr * chan + ch) * cap_px + (size_t) y * W + x] = (uint16_t) v;
        }
        __syncwarp();
    }
}

// =========================================================================================================
// K4: encoder
// =========================================================================================================
struct EncSmem {
    CoreTables T;
    uint32_t cnt[17][NT];      // cached bin << 27 | zero << 16 | total, per context
    uint16_t bseq[17][NT];     // sequence number (low 16 bits; the window spans 2048) of the codeword each bin is building
    uint16_t bst[17][NT];      // Golomb: zeros counted; variable-to-variable: input word (first bit = bit 0) | bits gathered << 8
};

// One warp per (frame group, stream); lane = frame of the group.
__global__ void __launch_bounds__(NT, ICER_ENC_MINBLOCKS) k_encode(EncodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EncSmem &S = *reinterpret_cast<EncSmem *>(smem_raw);
    const int t = threadIdx.x;
    stage_tables(S.T, c_tab.enc, NT);
    __syncthreads();

    const int P = (int) a.planes;
    const SliceGeom G = enc_geom(P);
    const uint32_t nstreams = a.nunits * a.planes;
    const unsigned long long total = (unsigned long long) a.ngroups * (a.list ? a.nlist : nstreams);
    const int lane = t & 31;
    uint16_t *ring = a.rings + (size_t) (blockIdx.x * NT + t) * ICER_RING;

    // Work items in stream-major order: the largest streams of ALL groups first, so the tail of the kernel consists of small
    // streams.  The FIRST item of every warp is dealt round-robin over the blocks (item i -> block i mod gridDim.x): neighbouring
    // items are the bit planes of one segment -- dense low planes, nearly empty high planes -- and an SM that started with 32 dense
    // planes of the largest segments would still be running when the rest of the chip has gone idle (4096^2 frames: 1.7 s instead
    // of 0.9 s).  After that the warps fetch dynamically.
    const unsigned long long dealt = (unsigned long long) gridDim.x * (NT / 32);
    bool first = true;
    for (;;) {
        unsigned long long wid = 0;
        if (first) { wid = (unsigned long long) blockIdx.x + (unsigned long long) (t >> 5) * gridDim.x; first = false; }
        else {
            if (lane == 0) wid = atomicAdd(a.work_counter, 1ull) + dealt;
            wid = __shfl_sync(0xffffffffu, wid, 0);
        }
        if (wid >= total) break;
        const uint32_t g = (uint32_t) (wid % a.ngroups);
        const uint32_t s = a.list ? a.list[wid / a.ngroups] : (uint32_t) (wid / a.ngroups);
        const Unit u = a.units[s / a.planes];
        const int lsb = (int) (s % a.planes);
        const uint32_t frame = g * 32 + lane;
        const bool valid = frame < a.nframes;
        icer_core::EncLane<NT> ln(S.T, &S.cnt[0][t], &S.bseq[0][t], &S.bst[0][t], ring,
                                                 valid ? a.scratch + (size_t) frame * a.scratch_words + a.soff[s] : nullptr, valid ? a.scap[s] : 0);
        const int sw = u.w, sh = u.h, sb = u.subband, cw = (int) u.cw;
        // field offsets (in words) inside a slot; -1: the field is identically zero
        const int oB = lsb * 32, oN = (P + G.PS) * 32;
        const int oP = lsb + 1 <= G.PS ? (P + lsb) * 32 : -1, o2 = lsb + 2 <= G.PS ? (P + lsb + 1) * 32 : -1, o3 = lsb + 3 <= G.PS ? (P + lsb + 2) * 32 : -1;
        const size_t slot_words = (size_t) G.F * 32, row_words = slot_words * cw;
        const uint32_t *rowp = a.bs + ((size_t) g * a.nslots + u.slot0) * slot_words + lane;
        for (int row = 0; row < sh; row++, rowp += row_words) {
            const bool up = row > 0, dn = row + 1 < sh;
            const uint32_t *p = rowp;
            for (int j = 0; j < cw; j++, p += slot_words) {
                const int ncols = sw - j * ICER_CHUNK < ICER_CHUNK ? sw - j * ICER_CHUNK : ICER_CHUNK;
                icer_core::EncChunk k;
                {
                    // the slice words are used once per stream: they bypass L1 (ld.global.cg), which is left to the codeword ring
                    const uint32_t *pa = p - row_words, *pc = p + row_words;
                    const uint32_t bP = oP >= 0 ? __ldcg(p + oP) : 0u, bB = __ldcg(p + oB), bN = __ldcg(p + oN);
                    const uint32_t b2 = o2 >= 0 ? __ldcg(p + o2) : 0u, b3 = o3 >= 0 ? __ldcg(p + o3) : 0u;
                    const uint32_t aP = (up && oP >= 0) ? __ldcg(pa + oP) : 0u, aB = up ? __ldcg(pa + oB) : 0u, aN = up ? __ldcg(pa + oN) : 0u;
                    const uint32_t cP = (dn && oP >= 0) ? __ldcg(pc + oP) : 0u, cN = dn ? __ldcg(pc + oN) : 0u;
                    k.set(aP, aB, aN, bP, bB, bN, b2, b3, cP, cN, ncols);
                }
                // a chunk that is quiet in every frame of the group and lies inside every lane's current run of (context 0, bit 0)
                // decisions is absorbed without a single step (the empty upper planes consist of little else)
                if (__all_sync(0xffffffffu, ln.fz_left >= (uint32_t) ncols && (k.nq & ((2u << ncols) - 2u)) == 0u)) {
                    ln.fz_run += (uint32_t) ncols; ln.fz_left -= (uint32_t) ncols;
                    continue;
                }
                ln.begin_chunk();
                for (;;) {
                    const bool busy = !ln.chunk_done(ncols);
                    const unsigned lanes = __ballot_sync(0xffffffffu, busy);
                    if (!lanes) break;
                    if (busy) ln.step(k, sb, lanes);
                    ln.pump();       // one emission per step keeps pace: a step starts at most one codeword
                }
            }
        }
        uint32_t bits;
        ln.finish(bits);
        if (valid) {
            a.bits[(size_t) frame * nstreams + s] = bits;
            if (ln.overflow) atomicOr(a.overflow, 1);
        }
        __syncwarp();
    }
}

// =========================================================================================================
// K7: decoder
// =========================================================================================================
// Work items of the decoder are TASKS (chain, bit plane).  A block is a pool of DEC_WARPS warps that take tasks in order
// (chain-major, most significant plane first); task (c, p) may process slot n of the unit once task (c, p+1) has STORED slot
// n + cw + 1 (the words of the row below, incl. the halo bit that depends on the chunk after it).  A task with a smaller number has
// always been taken by some warp already, so waiting on it cannot deadlock; there is no block barrier anywhere, so the quick (quiet)
// upper planes run ahead into the next chains while the dense lower planes are still being decoded.  Chains in flight occupy one of
// DEC_SLOTS slots of progress state in shared memory.
// Pool size (round 1, config 2, k_decode ms): 8 warps x 4 blocks/SM 42.1, 16 x 2 38.0, 24 x 1 40.7, 32 x 1 (12 slots) 36.1 --
// one SM-wide pool balances the very unequal planes of a chain best (profiles/README.md).
#ifndef ICER_DEC_WARPS
#define ICER_DEC_WARPS 32
#endif
#ifndef ICER_DEC_SLOTS
#define ICER_DEC_SLOTS 48
#endif
#ifndef ICER_DEC_MINBLOCKS
#define ICER_DEC_MINBLOCKS 1
#endif
#ifndef ICER_DEC_BACKOFF_NS
#define ICER_DEC_BACKOFF_NS 500      /* sleep when the plane above is not far enough ahead */
#endif
#ifndef ICER_DEC_SLACK
#define ICER_DEC_SLACK 0            /* extra chunks a blocked plane lets the plane above get ahead before it continues */
#endif
#ifndef ICER_DEC_SLOTWAIT_NS
#define ICER_DEC_SLOTWAIT_NS 20000   /* poll interval while a chain waits for a free slot (chains last milliseconds) */
#endif
#ifndef ICER_DEC_CHAINWAIT_NS
#define ICER_DEC_CHAINWAIT_NS 4000   /* poll interval while a plane waits for its chain's slot to be published */
#endif
constexpr int DEC_WARPS = ICER_DEC_WARPS;
constexpr int DEC_BT = DEC_WARPS * 32;
constexpr int DEC_SLOTS = ICER_DEC_SLOTS;

struct DecSlot {
    volatile uint32_t chain;             // chain ordinal (within this block) that owns the slot; ~0 = never used
    volatile uint32_t freed;             // all planes done: the slot may be re-used
    volatile uint32_t done;              // planes finished
    volatile unsigned long long wid;     // global work id = unit * ngroups + group
    volatile uint32_t prog[15];          // slots STORED per plane (warp-uniform)
};
struct DecShared {
    CoreTables T;
    DecSlot slot[DEC_SLOTS];
    uint32_t next_task;
    volatile uint32_t fetch_turn;        // chain ordinal whose first task may fetch its global work id next
    volatile uint32_t k_end;             // first chain ordinal that does not exist (work exhausted)
};
// followed in dynamic shared memory by: cnt[17][DEC_BT] u32, stlow[17][DEC_BT] u32, depth[17][DEC_BT] i16

__global__ void __launch_bounds__(DEC_BT, ICER_DEC_MINBLOCKS) k_decode(DecodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecShared &S = *reinterpret_cast<DecShared *>(smem_raw);
    const int t = threadIdx.x;
    constexpr int BT = DEC_BT;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem_raw + ((sizeof(DecShared) + 15) & ~(size_t) 15));
    uint32_t *stlow = cnt + 17 * BT;
    int16_t *depthA = reinterpret_cast<int16_t *>(stlow + 17 * BT);
    stage_tables(S.T, c_tab.dec, BT);
    if (t < DEC_SLOTS) { S.slot[t].chain = 0xFFFFFFFFu; S.slot[t].freed = 1; S.slot[t].done = 0; }
    if (t == 0) { S.next_task = 0; S.fetch_turn = 0; S.k_end = 0xFFFFFFFFu; }
    __syncthreads();
    const int lane = t & 31;
    const int P = (int) a.planes;
    const int F = dec_fields(P);
    const size_t slot_words = (size_t) F * 32;
    const unsigned long long total = (unsigned long long) a.ngroups * a.nunits;
    const uint32_t nstreams = a.nunits * a.planes;

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
            // The first task of a chain fetches the global work item and recycles the slot of chain k - DEC_SLOTS.  The fetch happens
            // in CHAIN ORDER (ticket S.fetch_turn): work ids then grow with k inside a block, so "chain k got no work" implies the same
            // for every later chain -- the invariant k_end and the early exits rely on.
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
            if (lane == 0) {
                if (k >= (uint32_t) DEC_SLOTS) {
                    uint32_t ns = ICER_DEC_SLOTWAIT_NS;
                    while (!(sl.chain == k - DEC_SLOTS && sl.freed)) { __nanosleep(ns); if (ns < 8 * ICER_DEC_SLOTWAIT_NS) ns += ns; }     // that chain is live: it will finish
                }
                __threadfence_block();
                for (int q = 0; q < 15; q++) sl.prog[q] = 0;
                sl.done = 0; sl.freed = 0; sl.wid = wid;
                __threadfence_block();
                sl.chain = k;
            }
            __syncwarp();
        } else {
            bool gone = false;
            if (lane == 0) {
                uint32_t ns = ICER_DEC_CHAINWAIT_NS;
                while (sl.chain != k) { if (S.k_end <= k) { gone = true; break; } __nanosleep(ns); if (ns < 16 * ICER_DEC_CHAINWAIT_NS) ns += ns; }
            }
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
        const uint8_t *pay = nullptr, *fend = nullptr;
        uint32_t nbits = 0;
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
                pay = hd + ICER_HDR_BYTES; fend = fbase + flen;
                nbits = hd[16] | (hd[17] << 8) | (hd[18] << 16) | ((uint32_t) hd[19] << 24);
            }
        }
        icer_core::DecLane<BT> ln(S.T, cnt + t, stlow + t, depthA + t, alive);
        if (alive) ln.br.start(pay, fend, nbits);
        const int sw = u.w, sh = u.h, sb = u.subband, cw = (int) u.cw;
        const uint32_t nsl = (uint32_t) sh * (uint32_t) cw;
        // field offsets (in words); -1: identically zero (nothing is known above the top coded plane)
        const int oB = lsb * 32, oS = (P + lsb) * 32, oN = (2 * P + lsb) * 32;
        const int oP = lsb + 1 < P ? (P + lsb + 1) * 32 : -1, o2 = lsb + 2 < P ? (P + lsb + 2) * 32 : -1, o3 = lsb + 3 < P ? (P + lsb + 3) * 32 : -1;
        const int oNp = lsb + 1 < P ? (2 * P + lsb + 1) * 32 : -1;
        const size_t row_words = slot_words * cw;
        uint32_t *base = a.bs + ((size_t) g * a.nslots + u.slot0) * slot_words + lane;
        uint32_t stored = 0;               // slots of this plane whose words are final (warp-uniform)
        uint32_t avail = 0;                // slots the plane above is known to have stored
        uint32_t tslot = 0;
        for (int row = 0; row < sh; row++) {
            const bool up = row > 0, dn = row + 1 < sh;
            uint32_t carryB = 0, carryN = 0, pendS = 0, pendN = 0;
            for (int jc = 0; jc < cw; jc++, tslot++) {
                if (j > 0) {
                    // the plane above must have stored the words of the row below this chunk (incl. their right halo)
                    const uint32_t need = tslot + cw + 1 < nsl ? tslot + cw + 1 : nsl;
                    if (avail < need) {                      // warp-uniform: `avail` is what lane 0 saw last
                        // wait until the plane above is a few chunks further than needed: this plane then runs that many chunks without
                        // looking again (planes of similar density would otherwise block and poll at every chunk)
                        const uint32_t want = need + ICER_DEC_SLACK < nsl ? need + ICER_DEC_SLACK : nsl;
                        if (lane == 0) {
                            uint32_t ns = ICER_DEC_BACKOFF_NS;
                            while ((avail = sl.prog[j - 1]) < want) { __nanosleep(ns); if (ns < 8 * ICER_DEC_BACKOFF_NS) ns += ns; }
                        }
                        avail = __shfl_sync(0xffffffffu, avail, 0);
                        __threadfence_block();
                    }
                }
                uint32_t *p = base + (size_t) tslot * slot_words;
                const int ncols = sw - jc * ICER_CHUNK < ICER_CHUNK ? sw - jc * ICER_CHUNK : ICER_CHUNK;
                icer_core::DecChunk kc;
                {
                    // slice words come from L2 (ld.global.cg): written by another warp of this block (the plane above) or by this
                    // warp a row earlier, and used once
                    const uint32_t *pa = p - row_words, *pc = p + row_words;
                    const uint32_t bP = oP >= 0 ? __ldcg(p + oP) : 0u, bNp = oNp >= 0 ? __ldcg(p + oNp) : 0u;
                    const uint32_t b2 = o2 >= 0 ? __ldcg(p + o2) : 0u, b3 = o3 >= 0 ? __ldcg(p + o3) : 0u;
                    const uint32_t cP = (dn && oP >= 0) ? __ldcg(pc + oP) : 0u, cN = (dn && oNp >= 0) ? __ldcg(pc + oNp) : 0u;
                    // row above: this plane's own words
                    const uint32_t aS = up ? __ldcg(pa + oS) : 0u, aN = up ? __ldcg(pa + oN) : 0u;
                    kc.set(aS, aN, bP, bNp, b2, b3, cP, cN, carryB, carryN, ncols);
                }
                // a chunk that is quiet in every frame and inside every lane's pending run of zeros needs no step (see k_encode)
                const bool absorbed = __all_sync(0xffffffffu, !ln.run || (ln.fz_left >= (uint32_t) ncols && ((kc.nq | (kc.bB << 1)) & ((2u << ncols) - 2u)) == 0u));
                if (absorbed) {
                    if (ln.run) { ln.fz_run += (uint32_t) ncols; ln.fz_left -= (uint32_t) ncols; }
                } else {
                    ln.begin_chunk();
                    for (;;) {
                        const bool busy = !ln.chunk_done(ncols);
                        const unsigned lanes = __ballot_sync(0xffffffffu, busy);
                        if (!lanes) break;
                        if (busy) ln.step(kc, sb, lanes);
                    }
                }
                // this chunk's words; the right halo of the chunk before is patched now that this chunk's first column is known
                p[oB] = kc.bB;
                if (jc > 0) {
                    uint32_t *pp = p - slot_words;
                    pp[oS] = pendS | (((kc.bB >> 1) & 1u) << 31);
                    pp[oN] = pendN | (((kc.bNn >> 1) & 1u) << 31);
                }
                pendS = kc.bP | kc.bB; pendN = kc.bNp | kc.bNn;
                carryB = kc.bB >> ICER_CHUNK; carryN = kc.bNn >> ICER_CHUNK;
                if (jc == cw - 1) {
                    p[oS] = pendS;
                    p[oN] = pendN;
                    stored = tslot + 1;
                } else stored = tslot;
                __threadfence_block();
                __syncwarp();
                if (lane == 0) sl.prog[j] = stored;
            }
        }
        // ---- task finished -------------------------------------------------------------------------------
        if (frame >= 0) a.pstat[((size_t) frame * a.nunits + ui) * P + lsb] = (uint8_t) (ln.err != icer_core::ST_OK ? PST_ERROR : (alive ? PST_OK : PST_ABSENT));
        __threadfence_block();
        __syncwarp();
        uint32_t old = 0;
        if (lane == 0) {
            sl.prog[j] = 0xFFFFFFFFu;                 // never hold the planes below
            __threadfence_block();
            old = atomicAdd((uint32_t *) &sl.done, 1u);
            if (old == (uint32_t) P - 1) { __threadfence_block(); sl.freed = 1; }
        }
        __syncwarp();
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
    // every lane builds the header word its byte lives in (lane = byte index), the CRC of bytes 0..23 comes from the position tables
    const uint32_t mean = a.mean[(size_t) f * a.chan + u.chan] & 0xFF;    // passes through a uint8_t (icer.h:270)
    const uint32_t dc = a.crc[(size_t) f * a.nstreams + s];
    const int wi = lane >> 2;
    uint32_t word = wi == 0 ? (0x5Bu | (0x60u << 8) | (mean << 16))
                  : wi == 1 ? ((uint32_t) u.level | ((uint32_t) u.subband << 8) | ((uint32_t) u.seg << 16) | ((lsb | ((uint32_t) u.chan << 4)) << 24))
                  : wi == 2 ? a.w : wi == 3 ? a.h : wi == 4 ? nb : wi == 5 ? dc : 0u;
    uint32_t hb = (word >> (8 * (lane & 3))) & 0xFFu;
    uint32_t x = lane < 24 ? g_hdr_crc_pos[lane][hb] : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x ^= __shfl_xor_sync(0xffffffffu, x, d);
    const uint32_t hc = ~(x ^ g_hdr_crc_k);
    if (wi == 6) hb = (hc >> (8 * (lane & 3))) & 0xFFu;
    const uint8_t *pay = reinterpret_cast<const uint8_t *>(a.scratch + (size_t) f * a.scratch_words + a.soff[s]);
    uint8_t *dst = a.out + (size_t) f * a.out_stride + orr;
    if (lane < ICER_HDR_BYTES) dst[lane] = (uint8_t) hb;
    for (uint32_t i = lane; i < nbytes; i += 32) dst[ICER_HDR_BYTES + i] = pay[i];
    if (a.out_prio) {
        uint8_t *d2 = a.out_prio + (size_t) f * a.out_stride + opr;
        if (lane < ICER_HDR_BYTES) d2[lane] = (uint8_t) hb;
        for (uint32_t i = lane; i < nbytes; i += 32) d2[ICER_HDR_BYTES + i] = pay[i];
    }
}

}  // namespace

size_t enc_slice_words(uint32_t planes, uint32_t nslots) { return (size_t) enc_geom((int) planes).F * 32 * nslots; }
size_t dec_slice_words(uint32_t planes, uint32_t nslots) { return (size_t) dec_fields((int) planes) * 32 * nslots; }

uint32_t encode_resident_threads(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int bps = 0;
    cudaFuncSetAttribute(k_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(EncSmem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_encode, NT, sizeof(EncSmem));
    if (bps < 1) bps = 1;
    return (uint32_t) (sms * bps * NT);
}

static unsigned slice_grid_x(uint32_t nslots, uint32_t ngroups) {
    // blocks of 4 warps (one slot per warp and iteration); a few waves of 148 SMs x 8 blocks over all groups
    unsigned gx = (unsigned) ((8u * 1184u + ngroups - 1) / ngroups);
    const unsigned need = (nslots + 3) / 4;
    if (gx > need) gx = need;
    return gx < 1 ? 1 : gx;
}

template <int P> struct SliceLaunch {
    static void enc(dim3 grid, cudaStream_t st, const uint16_t *coefs, uint32_t *bs, const SlotDesc *slots, uint32_t nslots, uint32_t nframes,
                    uint32_t chan, uint32_t w, uint32_t h, int twos) {
        k_bitslice<P><<<grid, 128, 0, st>>>(coefs, bs, slots, nslots, nframes, chan, w, h, twos);
    }
    static void dec(dim3 grid, cudaStream_t st, const uint32_t *bs, uint16_t *frames, const SlotDesc *slots, const uint8_t *vp, const int32_t *frame_of,
                    const uint16_t *mean, uint32_t nslots, uint32_t nunits, uint32_t chan, uint32_t w, uint32_t h, uint64_t cap_px, int llw, int llh, int bits8) {
        k_unslice<P><<<grid, 128, 0, st>>>(bs, frames, slots, vp, frame_of, mean, nslots, nunits, chan, w, h, cap_px, llw, llh, bits8);
    }
};
#define ICER_FOR_PLANES(P, CALL)                                                                                              \
    switch (P) {                                                                                                              \
        case 1: SliceLaunch<1>::CALL; break; case 2: SliceLaunch<2>::CALL; break; case 3: SliceLaunch<3>::CALL; break;        \
        case 4: SliceLaunch<4>::CALL; break; case 5: SliceLaunch<5>::CALL; break; case 6: SliceLaunch<6>::CALL; break;        \
        case 7: SliceLaunch<7>::CALL; break; case 8: SliceLaunch<8>::CALL; break; case 9: SliceLaunch<9>::CALL; break;        \
        case 10: SliceLaunch<10>::CALL; break; case 11: SliceLaunch<11>::CALL; break; case 12: SliceLaunch<12>::CALL; break;  \
        case 13: SliceLaunch<13>::CALL; break; case 14: SliceLaunch<14>::CALL; break; case 15: SliceLaunch<15>::CALL; break;  \
        default: return ICER_BITPLANE_OUT_OF_RANGE;                                                                           \
    }

int launch_bitslice(const uint16_t *coefs, uint32_t *bs, const SlotDesc *slots, uint32_t nslots, uint32_t nframes, uint32_t ngroups, uint32_t chan,
                    uint32_t w, uint32_t h, uint32_t planes, bool twos_complement, cudaStream_t st, LaunchStats &ls) {
    if (!nslots || !ngroups) return ICER_RESULT_OK;
    for (uint32_t g0 = 0; g0 < ngroups; g0 += 32768) {
        const uint32_t cnt = ngroups - g0 < 32768 ? ngroups - g0 : 32768;
        dim3 grid(slice_grid_x(nslots, cnt), cnt);
        const uint16_t *src = coefs + (size_t) g0 * 32 * chan * w * h;
        uint32_t *dst = bs + (size_t) g0 * enc_slice_words(planes, nslots);
        ICER_FOR_PLANES((int) planes, enc(grid, st, src, dst, slots, nslots, nframes - g0 * 32, chan, w, h, twos_complement ? 1 : 0))
        ls.launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}

int launch_unslice(const uint32_t *bs, uint16_t *frames, const SlotDesc *slots, const uint8_t *pstat, uint8_t *vp, size_t nbatch, const int32_t *frame_of,
                   const uint16_t *mean, uint32_t nslots, uint32_t nunits, uint32_t ngroups, uint32_t chan, uint32_t w, uint32_t h, uint64_t cap_px,
                   uint32_t planes, uint32_t llw, uint32_t llh, cudaStream_t st, LaunchStats &ls, bool bits8) {
    if (!nslots || !ngroups) return ICER_RESULT_OK;
    const size_t count = nbatch * nunits;
    k_valid_planes<<<(unsigned) ((count + 255) / 256), 256, 0, st>>>(pstat, vp, count, (int) planes);
    ls.launches++;
    for (uint32_t g0 = 0; g0 < ngroups; g0 += 32768) {
        const uint32_t cnt = ngroups - g0 < 32768 ? ngroups - g0 : 32768;
        dim3 grid(slice_grid_x(nslots, cnt), cnt);
        const uint32_t *src = bs + (size_t) g0 * dec_slice_words(planes, nslots);
        ICER_FOR_PLANES((int) planes, dec(grid, st, src, frames, slots, vp, frame_of + (size_t) g0 * 32, mean, nslots, nunits, chan, w, h, cap_px,
                                          (int) llw, (int) llh, bits8 ? 1 : 0))
        ls.launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}

int launch_encode(const EncodeArgs &a, cudaStream_t st, LaunchStats &ls) {
    int rc = upload_tables();
    if (rc) return rc;
    if (a.planes < 1 || a.planes > 15) return ICER_BITPLANE_OUT_OF_RANGE;
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
