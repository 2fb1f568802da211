// encode.cu -- K4 (bit-plane context model + interleaved entropy coder) and K5 (segment assembly).
//
// K4, v1 mapping: ONE THREAD PER STREAM, where a stream is one bit plane of one error-containment segment of
// one subband of one channel of one frame -- exactly the granularity at which lib_icer resets its context model
// and entropy coder (icer_partition.c:315-322), so all streams are independent.  Consecutive lanes of a warp
// hold the `planes` bit planes of the SAME segment: they walk the same pixels in lockstep, so their neighbourhood
// loads coalesce into the same sectors.  Warps fetch work dynamically (biggest segments first).
//
// Reference behaviour restated (lib_icer/src): icer_context_modeller.c:312-457 (contexts, adaptive counts with
// the discarded-ceil_div rescale), icer_encoding.c:37-189 (bin choice, Golomb / variable-to-variable / uncoded
// bins, 2048-slot interleave window with forced flush), icer_util.c:48-56 (bin cut-offs).
//
// Window bookkeeping differs from the reference's DONE flags: a codeword gets a sequence number when its first
// bit arrives; `head` is the next sequence number to emit.  Everything in [head, oldest open sequence) is
// complete by construction, so a codeword that completes AT the head is emitted straight from registers and only
// codewords that complete behind an older open one are parked in the ring (global memory, 2048 x u16 / thread).
#include "common.h"

namespace {

constexpr int NT = 128;   // threads per block

struct DevTables {
    uint16_t enc[8][32];       // in_bits | out_bits << 4 | out << 8
    uint8_t flush[8][9][6];    // pad_word | pad_len << 4
    uint16_t gol_m[17], gol_l[17], gol_i[17];
    uint32_t cutoff[16];
    uint32_t crc[256];
};
__constant__ DevTables c_tab;
// context tables, icer_config.c:26-67
__constant__ uint8_t c_ctx_hv[3][3][5] = {{{0, 1, 2, 2, 2}, {3, 3, 3, 3, 3}, {4, 4, 4, 4, 4}},
                                          {{5, 6, 7, 7, 7}, {7, 7, 7, 7, 7}, {7, 7, 7, 7, 7}},
                                          {{8, 8, 8, 8, 8}, {8, 8, 8, 8, 8}, {8, 8, 8, 8, 8}}};
__constant__ uint8_t c_ctx_hh[5][5] = {{0, 3, 6, 8, 8}, {1, 4, 7, 8, 8}, {2, 5, 7, 8, 8}, {2, 5, 7, 8, 8}, {2, 5, 7, 8, 8}};
__constant__ uint8_t c_ctx_sign[5][5] = {{14, 14, 15, 16, 16}, {14, 14, 15, 16, 16}, {13, 13, 12, 13, 13},
                                         {16, 16, 15, 14, 14}, {16, 16, 15, 14, 14}};
__constant__ uint8_t c_sign_pred[5][5] = {{1, 1, 1, 1, 1}, {1, 1, 1, 1, 1}, {0, 0, 0, 1, 1}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};

bool g_tables_uploaded = false;

struct EncSmem {
    uint32_t crc[256];
    uint16_t enc[8][32];
    uint8_t flush[8][9][6];
    uint8_t ctx_hv[3][3][5], ctx_hh[5][5], ctx_sign[5][5], sign_pred[5][5];
    uint16_t gol_m[17], gol_l[17], gol_i[17];
    uint32_t cutoff[16];
    uint32_t cnt[17][NT];      // zero << 16 | total, per context
    uint32_t bseq[17][NT];     // sequence number of the codeword each bin is building
    uint16_t bstate[17][NT];   // Golomb: zeros counted; custom: input bits gathered (first bit = bit 0)
    uint8_t bfill[17][NT];     // custom: number of input bits gathered
};

struct Coder {
    EncSmem &S;
    const int t;               // threadIdx.x
    uint16_t *ring;
    uint32_t *out;             // payload words
    uint32_t cap;              // capacity in words
    uint32_t nwords = 0;
    unsigned long long acc = 0;
    int nacc = 0;
    uint32_t crc = 0xFFFFFFFFu;
    uint32_t head = 0, tail = 0;
    uint32_t open_mask = 0;
    int oldest = 0;            // bin holding sequence number `head` (valid when open_mask != 0)
    bool overflow = false;

    __device__ Coder(EncSmem &s, int tid, uint16_t *r, uint32_t *o, uint32_t c) : S(s), t(tid), ring(r), out(o), cap(c) {}

    __device__ __forceinline__ void emit(uint32_t code, uint32_t nb) {   // LSB first (icer_encoding.c:114-139)
        acc |= (unsigned long long) code << nacc;
        nacc += (int) nb;
        if (nacc >= 32) {
            const uint32_t wv = (uint32_t) acc;
            if (nwords < cap) out[nwords] = wv; else overflow = true;
            nwords++;
            uint32_t c = crc;
            c = S.crc[(c ^ wv) & 0xFF] ^ (c >> 8);
            c = S.crc[(c ^ (wv >> 8)) & 0xFF] ^ (c >> 8);
            c = S.crc[(c ^ (wv >> 16)) & 0xFF] ^ (c >> 8);
            c = S.crc[(c ^ (wv >> 24)) & 0xFF] ^ (c >> 8);
            crc = c;
            acc >>= 32;
            nacc -= 32;
        }
    }

    __device__ __forceinline__ uint32_t golomb_word(int bin, uint32_t k, uint32_t &nb) {   // icer_encoding.c:72-77
        const uint32_t i = S.gol_i[bin], l = S.gol_l[bin];
        const uint32_t longer = k >= i;
        nb = l + longer;
        return __brev(k + (longer ? i : 0)) >> (32 - nb);
    }

    // codeword of `bin` is complete: emit it (and whatever it was blocking) or park it in the ring
    __device__ __forceinline__ void complete(int bin, uint32_t code, uint32_t nb) {
        const uint32_t seq = S.bseq[bin][t];
        open_mask &= ~(1u << bin);
        if (seq != head) {
            ring[seq & (ICER_RING - 1)] = (uint16_t) ((nb << 11) | code);
            return;
        }
        emit(code, nb);
        head++;
        // new oldest open codeword
        uint32_t limit = tail;
        for (uint32_t m = open_mask; m; m &= m - 1) {
            const int b = __ffs(m) - 1;
            const uint32_t sq = S.bseq[b][t];
            if ((int32_t) (sq - head) < (int32_t) (limit - head)) { limit = sq; oldest = b; }
        }
        while (head != limit) {
            const uint32_t v = ring[head & (ICER_RING - 1)];
            emit(v & 0x7FF, v >> 11);
            head++;
        }
    }

    // forced completion of the oldest codeword (icer_encoding.c:141-189)
    __device__ void flush_head() {
        const int bin = oldest;
        uint32_t code, nb;
        if (bin > 7) {
            const uint32_t k = S.bstate[bin][t];
            if (k == (uint32_t) S.gol_m[bin] - 1) { code = 1; nb = 1; }
            else code = golomb_word(bin, k, nb);
        } else {
            const uint32_t pre = S.bstate[bin][t], fill = S.bfill[bin][t];
            const uint32_t fl = S.flush[bin][pre][fill];
            const uint32_t word = (pre | ((fl & 0xF) << fill)) & 0x3FF;
            const uint32_t e = S.enc[bin][word & 31];
            code = e >> 8; nb = (e >> 4) & 0xF;
        }
        complete(bin, code, nb);
    }

    __device__ __forceinline__ void put(int bin, uint32_t bit) {   // icer_encoding.c:52-108
        if (!(open_mask & (1u << bin))) {
            if (tail - head >= ICER_RING) flush_head();
            if (open_mask == 0) oldest = bin;
            S.bseq[bin][t] = tail++;
            S.bstate[bin][t] = 0;
            S.bfill[bin][t] = 0;
            open_mask |= 1u << bin;
        }
        if (bin > 7) {
            if (bit) {
                uint32_t nb;
                const uint32_t code = golomb_word(bin, S.bstate[bin][t], nb);
                complete(bin, code, nb);
            } else {
                const uint32_t k = (uint32_t) S.bstate[bin][t] + 1;
                S.bstate[bin][t] = (uint16_t) k;
                if (k >= S.gol_m[bin]) complete(bin, 1, 1);
            }
        } else if (bin != 0) {
            const uint32_t fill = S.bfill[bin][t];
            const uint32_t word = S.bstate[bin][t] | (bit << fill);
            const uint32_t e = S.enc[bin][word & 31];
            if ((e & 0xF) == fill + 1) {
                complete(bin, e >> 8, (e >> 4) & 0xF);
            } else {
                S.bstate[bin][t] = (uint16_t) word;
                S.bfill[bin][t] = (uint8_t) (fill + 1);
            }
        } else {
            complete(0, bit & 1, 1);
        }
    }

    // adaptive decision: probability from (zero, total) of `ctx`, then count update with the rescale quirk
    __device__ __forceinline__ void put_ctx(int ctx, uint32_t bit) {
        const uint32_t zt = S.cnt[ctx][t];
        uint32_t z = zt >> 16, tot = zt & 0xFFFF;
        uint32_t b = bit;
        uint32_t zz = z;
        if (zz < (tot >> 1)) { zz = tot - zz; b ^= 1; }
        // bin = number of cut-offs c with zz * 65536 >= tot * c (icer_util.c:48-56); cut-offs ascend
        const uint32_t lhs = zz << 16;
        int bin = 0;
        if (lhs >= tot * S.cutoff[7]) bin = 8;
        if (lhs >= tot * S.cutoff[bin + 3]) bin += 4;
        if (lhs >= tot * S.cutoff[bin + 1]) bin += 2;
        if (lhs >= tot * S.cutoff[bin]) bin += 1;
        if (bin == 15 && lhs >= tot * S.cutoff[15]) bin = 16;
        put(bin, b);
        tot++;
        z += (bit ^ 1);
        if (tot >= 500) { tot >>= 1; if (z > tot) z >>= 1; }
        S.cnt[ctx][t] = (z << 16) | tot;
    }

    __device__ void finish(uint32_t &bits_out, uint32_t &crc_out) {
        while (open_mask) flush_head();                        // icer_context_modeller.c:452-455
        bits_out = nwords * 32u + (uint32_t) nacc;
        const int nbytes = (nacc + 7) >> 3;
        if (nbytes) {
            const uint32_t wv = (uint32_t) acc;
            if (nwords < cap) out[nwords] = wv; else overflow = true;
            uint32_t c = crc;
            for (int i = 0; i < nbytes; i++) c = S.crc[(c ^ (wv >> (8 * i))) & 0xFF] ^ (c >> 8);
            crc = c;
        }
        crc_out = ~crc;
    }
};

__device__ __forceinline__ int sig(uint32_t v, int plane) { return ((v & 0x7FFF) >> plane) != 0; }

__global__ void __launch_bounds__(NT) k_encode(EncodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EncSmem &S = *reinterpret_cast<EncSmem *>(smem_raw);
    const int t = threadIdx.x;
    for (int i = t; i < 256; i += NT) S.crc[i] = c_tab.crc[i];
    for (int i = t; i < 8 * 32; i += NT) (&S.enc[0][0])[i] = (&c_tab.enc[0][0])[i];
    for (int i = t; i < 8 * 9 * 6; i += NT) (&S.flush[0][0][0])[i] = (&c_tab.flush[0][0][0])[i];
    for (int i = t; i < 45; i += NT) (&S.ctx_hv[0][0][0])[i] = (&c_ctx_hv[0][0][0])[i];
    for (int i = t; i < 25; i += NT) {
        (&S.ctx_hh[0][0])[i] = (&c_ctx_hh[0][0])[i];
        (&S.ctx_sign[0][0])[i] = (&c_ctx_sign[0][0])[i];
        (&S.sign_pred[0][0])[i] = (&c_sign_pred[0][0])[i];
    }
    for (int i = t; i < 17; i += NT) { S.gol_m[i] = c_tab.gol_m[i]; S.gol_l[i] = c_tab.gol_l[i]; S.gol_i[i] = c_tab.gol_i[i]; }
    for (int i = t; i < 16; i += NT) S.cutoff[i] = c_tab.cutoff[i];
    __syncthreads();

    const uint32_t nstreams = a.nunits * a.planes;
    const unsigned long long total = (unsigned long long) a.nframes * nstreams;
    const int lane = t & 31;
    uint16_t *ring = a.rings + (size_t) (blockIdx.x * NT + t) * ICER_RING;
    const size_t plane_px = (size_t) a.w * a.h;

    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.work_counter, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) break;
        const unsigned long long sid = base + lane;
        if (sid < total) {
            const uint32_t frame = (uint32_t) (sid / nstreams), s = (uint32_t) (sid % nstreams);
            const uint32_t ui = s / a.planes;
            const int lsb = (int) (s % a.planes);
            const Unit u = a.units[ui];
            const int W = (int) a.w;
            const uint16_t *seg = a.coefs + ((size_t) frame * a.chan + u.chan) * plane_px + (size_t) u.y * W + u.x;
            uint32_t *out = a.scratch + (size_t) frame * a.scratch_words + a.soff[s];
            Coder cd(S, t, ring, out, a.scap[s]);
#pragma unroll
            for (int c = 0; c < 17; c++) S.cnt[c][t] = (2u << 16) | 4u;    // icer.h:147-148
            const int sw = u.w, sh = u.h, sb = u.subband;
            const int prev = lsb + 1;
            for (int row = 0; row < sh; row++) {
                const uint16_t *pr = seg + (size_t) row * W;
                const bool up = row > 0, dn = row + 1 < sh;
                // 3x3 neighbourhood values; outside the segment = 0 (never significant)
                uint32_t a0 = 0, a1 = 0, a2, b0 = 0, b1 = 0, b2, c0 = 0, c1 = 0, c2;
                a2 = (up && sw > 0) ? pr[-W] : 0;
                b2 = sw > 0 ? pr[0] : 0;
                c2 = (dn && sw > 0) ? pr[W] : 0;
                for (int col = 0; col < sw; col++) {
                    a0 = a1; a1 = a2; b0 = b1; b1 = b2; c0 = c1; c1 = c2;
                    const bool rt = col + 1 < sw;
                    a2 = (up && rt) ? pr[col + 1 - W] : 0;
                    b2 = rt ? pr[col + 1] : 0;
                    c2 = (dn && rt) ? pr[col + 1 + W] : 0;
                    const uint32_t v = b1;
                    const int msb = 31 - __clz((v & 0x7FFF) | 1);
                    int cat = msb - lsb;                      // icer_context_modeller.c:631-634
                    cat = cat < 0 ? 0 : (cat > 3 ? 3 : cat);
                    const uint32_t bit = (v >> lsb) & 1;
                    if (cat == 3) { cd.put(0, bit); continue; }    // p(0) = 1/2 -> uncoded bin
                    int ctx;
                    if (cat == 2) {
                        ctx = 11;
                    } else {
                        int hh = sig(b0, lsb) + sig(b2, prev);
                        int vv = sig(a1, lsb) + sig(c1, prev);
                        if (cat == 1) {
                            ctx = (hh + vv == 0) ? 9 : 10;
                        } else {
                            const int dd = sig(a0, lsb) + sig(a2, lsb) + sig(c0, prev) + sig(c2, prev);
                            if (sb == 1) { const int tmp = hh; hh = vv; vv = tmp; }
                            ctx = (sb == 3) ? S.ctx_hh[hh + vv][dd] : S.ctx_hv[hh][vv][dd];
                        }
                    }
                    cd.put_ctx(ctx, bit);
                    if (cat == 0 && bit) {                         // first 1 of this coefficient: code its sign
                        int shh = 2 - ((b0 >> 15) & sig(b0, lsb)) - ((b2 >> 15) & sig(b2, prev));
                        int svv = 2 - ((a1 >> 15) & sig(a1, lsb)) - ((c1 >> 15) & sig(c1, prev));
                        if (sb == 1) { const int tmp = shh; shh = svv; svv = tmp; }
                        const int sctx = S.ctx_sign[shh][svv];
                        const uint32_t agree = (S.sign_pred[shh][svv] ^ (v >> 15)) & 1;
                        cd.put_ctx(sctx, agree);
                    }
                }
            }
            uint32_t bits, crc;
            cd.finish(bits, crc);
            a.bits[(size_t) frame * nstreams + s] = bits;
            a.crc[(size_t) frame * nstreams + s] = crc;
            if (cd.overflow) atomicOr(a.overflow, 1);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// K5: quota scan, then gather of header + payload into the output streams
// ---------------------------------------------------------------------------------------------------------

// One thread per frame walks the streams in priority order with lib_icer's quota rule (SURVEY A.14;
// icer_encoding.c:211-214, :133-135; icer_partition.c:323-326).  Then a second walk in rearranged order assigns
// the offsets of the final stream.
__global__ void k_quota_scan(AssembleArgs a) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.nframes) return;
    int rc = a.plan_status;
    for (uint32_t c = 0; c < a.chan && rc == ICER_RESULT_OK; c++)
        if (a.ovf[(size_t) f * a.chan + c]) rc = ICER_INTEGER_OVERFLOW;
    const uint32_t *bits = a.bits + (size_t) f * a.nstreams;
    uint64_t *offp = a.off_prio + (size_t) f * a.nstreams;     // indexed by STREAM ID; ~0 = dropped
    uint64_t used = 0;
    uint32_t kept = 0;
    if (rc == ICER_RESULT_OK) {
        for (uint32_t i = 0; i < a.nstreams; i++) {
            const uint32_t s = a.prio[i];
            const uint32_t nb = bits[s];
            const uint64_t room = a.quota - used;
            if (room < ICER_HDR_BYTES) { rc = ICER_BYTE_QUOTA_EXCEEDED; break; }
            if (nb > 0 && (uint64_t) (nb >> 3) >= room - ICER_HDR_BYTES) { rc = ICER_BYTE_QUOTA_EXCEEDED; break; }
            offp[s] = used;
            used += ICER_HDR_BYTES + ((nb + 7) >> 3);
            kept++;
        }
    }
    // mark the rest as dropped
    for (uint32_t i = kept; i < a.nstreams; i++) offp[a.prio[i]] = ~0ull;
    uint64_t *offr = a.off_rearr + (size_t) f * a.nstreams;
    uint64_t ro = 0;
    for (uint32_t i = 0; i < a.nstreams; i++) {
        const uint32_t s = a.rearr[i];
        if (offp[s] == ~0ull) { offr[s] = ~0ull; continue; }
        offr[s] = ro;
        ro += ICER_HDR_BYTES + ((bits[s] + 7) >> 3);
    }
    a.size_used[f] = (rc == ICER_RESULT_OK || rc == ICER_BYTE_QUOTA_EXCEEDED) ? used : 0;
    a.rc[f] = rc;
    a.keep[f] = kept;
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

int upload_tables() {
    if (g_tables_uploaded) return ICER_RESULT_OK;
    const HostTables &h = icer_host_tables();
    static DevTables d;
    for (int b = 0; b < 8; b++)
        for (int v = 0; v < 32; v++) d.enc[b][v] = (uint16_t) (h.enc[b][v].in_bits | (h.enc[b][v].out_bits << 4) | (h.enc[b][v].out << 8));
    memcpy(d.flush, h.flush, sizeof d.flush);
    for (int b = 0; b < 17; b++) { d.gol_m[b] = h.gol_m[b]; d.gol_l[b] = h.gol_l[b]; d.gol_i[b] = h.gol_i[b]; }
    for (int b = 0; b < 16; b++) d.cutoff[b] = h.cutoff[b];
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        d.crc[n] = c;
    }
    CUDA_TRY(cudaMemcpyToSymbol(c_tab, &d, sizeof d));
    g_tables_uploaded = true;
    return ICER_RESULT_OK;
}

}  // namespace

uint32_t encode_resident_threads(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int bps = 0;
    cudaFuncSetAttribute(k_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(EncSmem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_encode, NT, sizeof(EncSmem));
    if (bps < 1) bps = 1;
    return (uint32_t) (sms * bps * NT);
}

int launch_encode(const EncodeArgs &a, cudaStream_t st, LaunchStats &ls) {
    int rc = upload_tables();
    if (rc) return rc;
    CUDA_TRY(cudaFuncSetAttribute(k_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(EncSmem)));
    const unsigned long long total = (unsigned long long) a.nframes * a.nunits * a.planes;
    unsigned blocks = a.nring_threads / NT;
    const unsigned long long need = (total + NT - 1) / NT;
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
    k_quota_scan<<<(a.nframes + 63) / 64, 64, 0, st>>>(a);
    const unsigned long long warps = (unsigned long long) a.nframes * a.nstreams;
    const unsigned long long blocks = (warps * 32 + 255) / 256;
    if (blocks > 0x7FFFFFFFull) return ICER_FATAL_ERROR;
    if (blocks) k_gather<<<(unsigned) blocks, 256, 0, st>>>(a);
    ls.launches += 2;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
