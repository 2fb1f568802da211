// decode.cu -- K7: bit-plane context model + interleaved entropy DECODER.
//
// v1 mapping: one thread per chain, where a chain is one error-containment segment of one subband of one channel
// of one frame; its bit planes are decoded MSB -> LSB, stopping at the first missing plane or decode error
// (icer_partition.c:427-443).  Within a plane the scan is strictly serial (the context of a pixel depends on bits
// decoded just before it), so parallelism comes from chains x frames.
//
// Reference behaviour restated (lib_icer/src): icer_context_modeller.c:461-602 (contexts from partially decoded
// data), icer_decoding.c:108-194 (per-bin stacks of decoded bits, refill when empty or when 2048 codewords have
// gone by -- the mirror of the encoder's forced flush), :46-105 (bit readers; `decoded_bits_total` is never
// advanced there, so the only live range checks are the ones reproduced in rd() and the custom-code loop).
#include "common.h"

namespace {

constexpr int NTD = 128;

struct DecTables {
    uint16_t dec[8][32];       // in_bits | out_bits << 4 | out << 8   (keyed by the received word)
    uint16_t gol_m[17], gol_l[17], gol_i[17];
    uint32_t cutoff[16];
};
__constant__ DecTables c_dtab;
__constant__ uint8_t c_dctx_hv[3][3][5] = {{{0, 1, 2, 2, 2}, {3, 3, 3, 3, 3}, {4, 4, 4, 4, 4}},
                                           {{5, 6, 7, 7, 7}, {7, 7, 7, 7, 7}, {7, 7, 7, 7, 7}},
                                           {{8, 8, 8, 8, 8}, {8, 8, 8, 8, 8}, {8, 8, 8, 8, 8}}};
__constant__ uint8_t c_dctx_hh[5][5] = {{0, 3, 6, 8, 8}, {1, 4, 7, 8, 8}, {2, 5, 7, 8, 8}, {2, 5, 7, 8, 8}, {2, 5, 7, 8, 8}};
__constant__ uint8_t c_dctx_sign[5][5] = {{14, 14, 15, 16, 16}, {14, 14, 15, 16, 16}, {13, 13, 12, 13, 13},
                                          {16, 16, 15, 14, 14}, {16, 16, 15, 14, 14}};
__constant__ uint8_t c_dsign_pred[5][5] = {{1, 1, 1, 1, 1}, {1, 1, 1, 1, 1}, {0, 0, 0, 1, 1}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
bool g_dtables_uploaded = false;

struct DecSmem {
    uint16_t dec[8][32];
    uint8_t ctx_hv[3][3][5], ctx_hh[5][5], ctx_sign[5][5], sign_pred[5][5];
    uint16_t gol_m[17], gol_l[17], gol_i[17];
    uint32_t cutoff[16];
    uint32_t cnt[17][NTD];
    int32_t depth[17][NTD];     // pending decoded bits per bin
    uint32_t stamp[17][NTD];    // codeword count when the bin was last refilled
    uint8_t low[17][NTD];       // the low 8 stack positions (nothing above can hold a 1)
};

struct Decoder {
    DecSmem &S;
    const int t;
    const uint8_t *base;        // payload start
    const uint8_t *end;         // end of the whole stream buffer (reads beyond return 0)
    uint32_t total_bits;
    uint32_t consumed = 0;      // bits consumed
    uint32_t words = 0;         // codewords consumed

    __device__ Decoder(DecSmem &s, int tid, const uint8_t *b, const uint8_t *e, uint32_t nbits)
        : S(s), t(tid), base(b), end(e), total_bits(nbits) {}

    // n <= 16 bits starting `ahead` bits after the cursor, LSB first
    __device__ __forceinline__ uint32_t bits_at(uint32_t ahead, uint32_t n) const {
        const uint32_t pos = consumed + ahead;
        const uint8_t *p = base + (pos >> 3);
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) v |= (p + i < end ? (uint32_t) p[i] : 0u) << (8 * i);
        return (v >> (pos & 7)) & ((1u << n) - 1);
    }
    // icer_get_bits_from_codeword / icer_pop_bits_from_codeword (icer_decoding.c:59-105): byte-chunked read that
    // fails only when a chunk is larger than the whole segment
    __device__ __forceinline__ int rd(uint32_t n, bool consume) {
        if (total_bits < 8) {
            uint32_t off = consumed & 7, left = n;
            while (left) {
                uint32_t take = 8 - off; if (take > left) take = left;
                if (take > total_bits) return ICER_DECODER_OUT_OF_DATA;
                left -= take; off = (off + take) & 7;
            }
        }
        const int v = (int) bits_at(0, n);
        if (consume) consumed += n;
        return v;
    }

    __device__ int get(int bin, bool inv, uint32_t &bit) {
        int depth = S.depth[bin][t];
        uint32_t low = S.low[bin][t];
        if (depth <= 0 || words - S.stamp[bin][t] >= ICER_RING) {
            depth = 0; low = 0;
            if (bin > 7) {
                const uint32_t l = S.gol_l[bin], gi = S.gol_i[bin];
                if (bits_at(0, 1)) {
                    rd(1, true);
                    depth = S.gol_m[bin];
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
                uint32_t word = 0, n = 0;
                bool found = false;
                do {
                    if (n + 1 >= total_bits) return ICER_DECODER_OUT_OF_DATA;
                    word |= bits_at(n, 1) << n;
                    n++;
                    if (word >= 32) return ICER_DECODED_INVALID_DATA;
                    const uint32_t e = S.dec[bin][word];
                    if ((e & 0xF) == n) {
                        low = e >> 8;
                        depth = (int) ((e >> 4) & 0xF);
                        if ((int) word != rd(n, true)) return ICER_DECODED_INVALID_DATA;
                        found = true;
                    }
                } while (!found && n < 10);
            } else {
                const int v = rd(1, true);
                if (v == ICER_DECODER_OUT_OF_DATA) return v;
                low = v != 0;
                depth = 1;
            }
            words++;
            S.stamp[bin][t] = words;
        }
        const int top = depth - 1;
        uint32_t b = 0;
        if (top >= 0 && top < 8) { b = (low >> top) & 1u; low &= ~(1u << top); }
        S.depth[bin][t] = depth - 1;
        S.low[bin][t] = (uint8_t) low;
        bit = inv ? (b ^ 1u) : b;
        return ICER_RESULT_OK;
    }

    __device__ __forceinline__ int get_ctx(int ctx, uint32_t &bit) {
        const uint32_t zt = S.cnt[ctx][t];
        uint32_t z = zt >> 16, tot = zt & 0xFFFF;
        uint32_t zz = z;
        bool inv = false;
        if (zz < (tot >> 1)) { zz = tot - zz; inv = true; }
        const uint32_t lhs = zz << 16;
        int bin = 0;
        if (lhs >= tot * S.cutoff[7]) bin = 8;
        if (lhs >= tot * S.cutoff[bin + 3]) bin += 4;
        if (lhs >= tot * S.cutoff[bin + 1]) bin += 2;
        if (lhs >= tot * S.cutoff[bin]) bin += 1;
        if (bin == 15 && lhs >= tot * S.cutoff[15]) bin = 16;
        const int rc = get(bin, inv, bit);
        if (rc != ICER_RESULT_OK) return rc;
        tot++;
        z += (bit ^ 1);
        if (tot >= 500) { tot >>= 1; if (z > tot) z >>= 1; }
        S.cnt[ctx][t] = (z << 16) | tot;
        return ICER_RESULT_OK;
    }
};

__device__ __forceinline__ int dsig(uint32_t v, int plane) { return ((v & 0x7FFF) >> plane) != 0; }

__global__ void __launch_bounds__(NTD) k_decode(DecodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecSmem &S = *reinterpret_cast<DecSmem *>(smem_raw);
    const int t = threadIdx.x;
    for (int i = t; i < 8 * 32; i += NTD) (&S.dec[0][0])[i] = (&c_dtab.dec[0][0])[i];
    for (int i = t; i < 45; i += NTD) (&S.ctx_hv[0][0][0])[i] = (&c_dctx_hv[0][0][0])[i];
    for (int i = t; i < 25; i += NTD) {
        (&S.ctx_hh[0][0])[i] = (&c_dctx_hh[0][0])[i];
        (&S.ctx_sign[0][0])[i] = (&c_dctx_sign[0][0])[i];
        (&S.sign_pred[0][0])[i] = (&c_dsign_pred[0][0])[i];
    }
    for (int i = t; i < 17; i += NTD) { S.gol_m[i] = c_dtab.gol_m[i]; S.gol_l[i] = c_dtab.gol_l[i]; S.gol_i[i] = c_dtab.gol_i[i]; }
    for (int i = t; i < 16; i += NTD) S.cutoff[i] = c_dtab.cutoff[i];
    __syncthreads();

    const unsigned long long total = (unsigned long long) a.nframes * a.nunits;
    const int lane = t & 31;
    const uint32_t nstreams = a.nunits * a.planes;
    const uint8_t *buf_end = a.streams + a.total_bytes;

    for (;;) {
        unsigned long long basei = 0;
        if (lane == 0) basei = atomicAdd(a.work_counter, 32ull);
        basei = __shfl_sync(0xffffffffu, basei, 0);
        if (basei >= total) break;
        const unsigned long long cid = basei + lane;
        if (cid < total) {
            const uint32_t frame = (uint32_t) (cid / a.nunits), ui = (uint32_t) (cid % a.nunits);
            if (a.frame_ok[frame]) {
                const Unit u = a.units[ui];
                const int W = (int) a.plan_w;
                uint16_t *seg = a.coefs + ((size_t) frame * a.chan + u.chan) * a.cap_px + (size_t) u.y * W + u.x;
                const uint8_t *stream = a.streams + (size_t) frame * a.stream_stride;
                const int sw = u.w, sh = u.h, sb = u.subband;
                for (int lsb = (int) a.planes - 1; lsb >= 0; lsb--) {
                    const uint32_t off = a.seg_off[(size_t) frame * nstreams + ui * a.planes + lsb];
                    if (off == 0xFFFFFFFFu) break;
                    const uint8_t *hd = stream + off;
                    const uint32_t nbits = hd[16] | (hd[17] << 8) | (hd[18] << 16) | ((uint32_t) hd[19] << 24);
                    Decoder dc(S, t, hd + ICER_HDR_BYTES, buf_end, nbits);
#pragma unroll
                    for (int c = 0; c < 17; c++) { S.cnt[c][t] = (2u << 16) | 4u; S.depth[c][t] = 0; S.stamp[c][t] = 0; S.low[c][t] = 0; }
                    const int prev = lsb + 1;
                    int rc = ICER_RESULT_OK;
                    for (int row = 0; row < sh && rc == ICER_RESULT_OK; row++) {
                        uint16_t *pr = seg + (size_t) row * W;
                        const bool up = row > 0, dn = row + 1 < sh;
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
                            uint32_t v = b1;
                            const int msb = 31 - __clz((v & 0x7FFF) | 1);
                            int cat = msb - lsb;
                            cat = cat < 0 ? 0 : (cat > 3 ? 3 : cat);
                            uint32_t bit;
                            if (cat == 3) {
                                rc = dc.get(0, false, bit);
                                if (rc != ICER_RESULT_OK) break;
                                v |= bit << lsb;
                            } else {
                                int ctx;
                                if (cat == 2) {
                                    ctx = 11;
                                } else {
                                    int hh = dsig(b0, lsb) + dsig(b2, prev);
                                    int vv = dsig(a1, lsb) + dsig(c1, prev);
                                    if (cat == 1) {
                                        ctx = (hh + vv == 0) ? 9 : 10;
                                    } else {
                                        const int dd = dsig(a0, lsb) + dsig(a2, lsb) + dsig(c0, prev) + dsig(c2, prev);
                                        if (sb == 1) { const int tmp = hh; hh = vv; vv = tmp; }
                                        ctx = (sb == 3) ? S.ctx_hh[hh + vv][dd] : S.ctx_hv[hh][vv][dd];
                                    }
                                }
                                rc = dc.get_ctx(ctx, bit);
                                if (rc != ICER_RESULT_OK) break;
                                v |= bit << lsb;
                                if (cat == 0 && bit) {
                                    int shh = 2 - ((b0 >> 15) & dsig(b0, lsb)) - ((b2 >> 15) & dsig(b2, prev));
                                    int svv = 2 - ((a1 >> 15) & dsig(a1, lsb)) - ((c1 >> 15) & dsig(c1, prev));
                                    if (sb == 1) { const int tmp = shh; shh = svv; svv = tmp; }
                                    uint32_t agree;
                                    // the magnitude bit is already stored when the sign read fails (reference order :502,:583)
                                    pr[col] = (uint16_t) v;
                                    rc = dc.get_ctx(S.ctx_sign[shh][svv], agree);
                                    if (rc != ICER_RESULT_OK) break;
                                    v |= ((agree ^ S.sign_pred[shh][svv]) & 1u) << 15;
                                }
                            }
                            pr[col] = (uint16_t) v;
                            b1 = v;
                        }
                    }
                    if (rc != ICER_RESULT_OK) break;
                }
            }
        }
        __syncwarp();
    }
}

int upload_dtables() {
    if (g_dtables_uploaded) return ICER_RESULT_OK;
    const HostTables &h = icer_host_tables();
    static DecTables d;
    for (int b = 0; b < 8; b++)
        for (int v = 0; v < 32; v++) d.dec[b][v] = (uint16_t) (h.dec[b][v].in_bits | (h.dec[b][v].out_bits << 4) | (h.dec[b][v].out << 8));
    for (int b = 0; b < 17; b++) { d.gol_m[b] = h.gol_m[b]; d.gol_l[b] = h.gol_l[b]; d.gol_i[b] = h.gol_i[b]; }
    for (int b = 0; b < 16; b++) d.cutoff[b] = h.cutoff[b];
    CUDA_TRY(cudaMemcpyToSymbol(c_dtab, &d, sizeof d));
    g_dtables_uploaded = true;
    return ICER_RESULT_OK;
}

}  // namespace

int launch_decode(const DecodeArgs &a, cudaStream_t st, LaunchStats &ls) {
    int rc = upload_dtables();
    if (rc) return rc;
    CUDA_TRY(cudaFuncSetAttribute(k_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(DecSmem)));
    int dev = 0, sms = 148, bps = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_decode, NTD, sizeof(DecSmem));
    if (bps < 1) bps = 1;
    const unsigned long long total = (unsigned long long) a.nframes * a.nunits;
    unsigned long long blocks = (total + NTD - 1) / NTD;
    if (blocks > (unsigned long long) sms * bps) blocks = (unsigned long long) sms * bps;
    if (blocks == 0) return ICER_RESULT_OK;
    k_decode<<<(unsigned) blocks, NTD, sizeof(DecSmem), st>>>(a);
    ls.launches++;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
