// coder_core.h -- the per-lane state machines of K4 (bit-plane context model + interleaved entropy ENCODER) and K7 (the DECODER).
//
// Everything in here is lane-local (no warp intrinsics, no shared-memory declarations): the kernels in coder.cu give every lane
// one instance and call step() until the lane has finished its chunk.  Because the code is plain C++ it also compiles for the host:
// tests/emul builds it with g++ and checks it against the oracle without a GPU (test infrastructure; the product library only
// contains the device instantiation).
//
// Data representation: BIT SLICES.  A stream (one bit plane `lsb` of one segment) never looks at coefficient values, only at a few
// one-bit properties of each pixel and of its eight neighbours:
//     B_p  bit p of the magnitude              S_p  "significant at plane p": (magnitude >> p) != 0              N  the sign bit
// For 30 consecutive pixels of a segment row those properties are 32-bit words -- bit 0 the pixel left of the chunk, bits 1..30 the
// chunk's pixels, bit 31 the pixel right of it (0 outside the segment, which is how the reference treats segment borders,
// icer_context_modeller.c:361-372) -- so a neighbourhood is a shift and a mask, a run of quiet pixels is a count-trailing-zeros, and
// a lane can be at ANY column of its chunk: lanes of a warp no longer wait for each other at every pixel (a sign decision, a
// codeword completion), only at chunk borders.
//
// Reference behaviour restated (lib_icer/src): icer_context_modeller.c:312-602 (contexts, categories, adaptive counts with the
// discarded-ceil_div rescale), icer_encoding.c:37-189 and icer_decoding.c:46-194 (bin choice, Golomb / variable-to-variable /
// uncoded bins, 2048-codeword interleave window with forced flush and its decoder mirror), icer_util.c:48-56 (bin cut-offs).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ICER_HD __host__ __device__ __forceinline__
#else
#define ICER_HD inline
#endif
// Re-convergence point of the lanes `mask` of a warp that all call the same step(): without it the compiler threads the jumps of
// the set-up branches straight into the common tail and the tail runs once per branch (measured: 13 of 32 lanes active there).
#if defined(__CUDA_ARCH__)
#define ICER_CONVERGE(mask) __syncwarp(mask)
#else
#define ICER_CONVERGE(mask) ((void) (mask))
#endif

#ifndef ICER_RING
#define ICER_RING 2048
#endif
#ifndef ICER_BIN_PROBE
#define ICER_BIN_PROBE 1          /* one probe of the neighbouring bin before the full search: k_encode 19.4 -> 18.6, k_decode 18.6 -> 18.0 ms */
#endif
#define ICER_CHUNK 30          /* pixels per bit-slice word */

namespace icer_core {

enum { ST_OK = 0, ST_OUT_OF_DATA = -7, ST_INVALID_DATA = -8 };       // = ICER_DECODER_OUT_OF_DATA / ICER_DECODED_INVALID_DATA

ICER_HD int ctz32(uint32_t x) {          // x != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int) x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
ICER_HD int ffs32(uint32_t x) {          // 1-based, 0 for x == 0
#if defined(__CUDA_ARCH__)
    return __ffs((int) x);
#else
    return __builtin_ffs((int) x);
#endif
}
ICER_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
ICER_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

// Tables every lane needs with divergent indices (shared memory on the device).
struct CoreTables {
    struct alignas(8) Cut { uint32_t lo, hi; } cut[17];   // per bin: cut-off below it / above it, sentinels 0 / 65537 -- one 64-bit load validates a cached bin
    uint32_t cutoff[16];
    uint16_t v2v[8][32];             // encoder: keyed by the gathered input word: in_bits | out_bits << 4 | out << 8
                                     // decoder: keyed by the NEXT 5 stream bits: codeword length | out_bits << 4 | out << 8
    uint16_t gol_m[17], gol_l[17], gol_i[17];
    uint8_t flush[8][9][6];          // pad_word | pad_len << 4
    uint8_t lut[6][256];             // rows 0..3 [subband][neighbour mask] -> context of a category-0 pixel (icer_config.c:26-45);
                                     // rows 4, 5 [subband == HL][mask of the NEGATIVE significant neighbours] -> sign context | predicted sign << 7
                                     // (:47-67; only the four direct neighbours count, the diagonal bits of the mask are ignored)
};
// neighbour mask bits: 0,1,2 = row above (left, mid, right); 3 = left; 4 = right; 5,6,7 = row below (left, mid, right)
#define ICER_NM_HV 0x5Au           /* mid-above, left, right, mid-below */

// probability bin: number of cut-offs c with zz * 65536 >= tot * c (icer_util.c:48-56); cut-offs ascend
ICER_HD int bin_of(const CoreTables &T, uint32_t zz, uint32_t tot) {
    const uint32_t lhs = zz << 16;
    int bin = 0;
    if (lhs >= tot * T.cutoff[7]) bin = 8;
    if (lhs >= tot * T.cutoff[bin + 3]) bin += 4;
    if (lhs >= tot * T.cutoff[bin + 1]) bin += 2;
    if (lhs >= tot * T.cutoff[bin]) bin += 1;
    if (bin == 15 && lhs >= tot * T.cutoff[15]) bin = 16;
    return bin;
}
// The bin of a context moves slowly, so each context remembers its last bin (bits 27..31 of its count word); the guess holds iff
// cut[guess-1] * tot <= zz * 65536 < cut[guess] * tot.
ICER_HD int bin_cached(const CoreTables &T, uint32_t zz, uint32_t tot, int guess) {
    const uint32_t lhs = zz << 16;
    const CoreTables::Cut c = T.cut[guess];
    if (lhs >= tot * c.lo && lhs < tot * c.hi) return guess;
#if ICER_BIN_PROBE
    // a context that leaves its bin almost always lands in a neighbouring one: one more probe before the full search
    const int nbr = lhs >= tot * c.hi ? guess + 1 : guess - 1;     // in range: the sentinels 65537 / 0 never fail
    const CoreTables::Cut c2 = T.cut[nbr];
    if (lhs >= tot * c2.lo && lhs < tot * c2.hi) return nbr;
#endif
    return bin_of(T, zz, tot);
}
#define ICER_CNT_Z(zt) (((zt) >> 16) & 0x7FFu)
#define ICER_CNT_TOT(zt) ((zt) & 0xFFFFu)
#define ICER_CNT_BIN(zt) ((int) ((zt) >> 27))
#define ICER_CNT_PACK(bin, z, tot) (((uint32_t) (bin) << 27) | ((z) << 16) | (tot))
#define ICER_CNT_INIT ((2u << 16) | 4u) /* icer.h:147-148 */

// number of consecutive all-zero decisions, starting at counts (z, tot), that stay in Golomb bin `bin`
ICER_HD uint32_t zeros_in_same_bin(const CoreTables &T, int bin, uint32_t z, uint32_t tot) {
    if (bin >= 16) return 0x7FFFFFFFu;
    const uint32_t cut = T.cutoff[bin];
    const uint32_t num = tot * cut - (z << 16);       // > 0 because the cut-off test for `bin` failed
    return (num - 1) / (65536u - cut) + 1;
}

ICER_HD uint32_t dilate3(uint32_t x) { return x | (x << 1) | (x >> 1); }

// context of a decision from its table entry (category 0 and sign decisions), the neighbour mask (category 1) or nothing
// (category 2); icer_context_modeller.c:355-391
ICER_HD int ctx_of(int cat, uint32_t e, uint32_t nm) {
    if (cat == 0) return (int) (e & 0x7Fu);
    if (cat == 1) return (nm & ICER_NM_HV) ? 10 : 9;
    return 11;
}

// =========================================================================================================================
// Encoder
// =========================================================================================================================
// The bit-slice words of one chunk as the encoder of plane `lsb` sees them ("cur" = significant at lsb, "prev" = at lsb + 1).
struct EncChunk {
    uint32_t aS, cP;            // row above: S_cur; row below: S_prev
    uint32_t bSc, bP;           // this row: S_cur, S_prev
    uint32_t b2, b3, bB;        // this row: S_{lsb+2}, S_{lsb+3}, the magnitude bit B_lsb
    uint32_t nsA, nsC;          // negative & significant: row above (cur), row below (prev)
    uint32_t nsBc, nsBp, bN;    // this row: negative & S_cur, negative & S_prev, the sign bits
    uint32_t nq;                // columns whose decision is NOT (context 0, bit 0); every bit above the last valid column is set

    // aP/aB/aN: S_prev, B, N of the row above; b*: this row; cP/cN: row below.  ncols = valid columns (1..30).
    ICER_HD void set(uint32_t aP, uint32_t aB, uint32_t aN, uint32_t bP_, uint32_t bB_, uint32_t bN_, uint32_t b2_, uint32_t b3_,
                     uint32_t cP_, uint32_t cN, int ncols) {
        aS = aP | aB; cP = cP_;
        bP = bP_; bB = bB_; bSc = bP_ | bB_; b2 = b2_; b3 = b3_; bN = bN_;
        nsA = aN & aS; nsC = cN & cP_; nsBc = bN_ & bSc; nsBp = bN_ & bP_;
        nq = dilate3(aS) | dilate3(cP) | (bSc << 1) | (bP >> 1) | bSc | (0xFFFFFFFEu << ncols);
    }
};

template <int STRIDE>
struct EncLane {
    const CoreTables &T;
    uint32_t *cnt;      // [17] (stride STRIDE): cached bin << 27 | zero count << 16 | total count, per context
    uint16_t *bseq;     // [17] sequence number (low 16 bits; the window spans 2048) of the codeword each bin is building
    uint16_t *bst;      // [17] Golomb: zeros counted; variable-to-variable: input word (first bit = bit 0) | bits gathered << 8
    uint16_t *ring;     // [ICER_RING] the interleave window: codewords wait here for the older ones still open (global memory)
    uint32_t *out;
    uint32_t cap, nwords = 0;
    unsigned long long acc = 0;
    int nacc = 0;
    // Interleave-window bookkeeping.  A codeword gets a sequence number when its first bit arrives (lib_icer: a ring slot,
    // icer_encoding.c:59-67); `head` = first sequence number that is not complete; everything in [emit_pos, head) is ready.
    uint32_t head = 0, tail = 0, open_mask = 0;
    uint32_t emit_pos = 0;
    int oldest = 0;
    bool overflow = false;
    // quiet fast path: zeros still absorbable / absorbed but not yet applied
    uint32_t fz_left = 0, fz_run = 0;
    int fz_bin = 0;
    // position inside the chunk
    int col = 1;        // 1..ncols; > ncols: chunk finished
    int phase = 0;      // 0: magnitude decision next, 1: sign decision of this pixel next

    ICER_HD EncLane(const CoreTables &t, uint32_t *c, uint16_t *bs, uint16_t *st, uint16_t *r, uint32_t *o, uint32_t cp)
        : T(t), cnt(c), bseq(bs), bst(st), ring(r), out(o), cap(cp) {
        for (int k = 0; k < 17; k++) cnt[k * STRIDE] = ICER_CNT_INIT;
    }

    ICER_HD uint32_t golomb_word(int bin, uint32_t k, uint32_t &nb) {   // icer_encoding.c:72-77
        const uint32_t i = T.gol_i[bin], l = T.gol_l[bin];
        const uint32_t longer = k >= i;
        nb = l + longer;
        return brev32(k + (longer ? i : 0)) >> (32 - nb);
    }

    // codeword of `bin` is complete: park it; if it was the oldest one, everything up to the next open codeword is ready
    ICER_HD void complete(int bin, uint32_t code, uint32_t nb) {
        const uint32_t seq = bseq[bin * STRIDE];
        open_mask &= ~(1u << bin);
        ring[seq & (ICER_RING - 1)] = (uint16_t) ((nb << 11) | code);
        if (seq != (head & 0xFFFFu)) return;
        const uint32_t h1 = head + 1;
        uint32_t dist = tail - h1;                       // every other open codeword lies in [head + 1, tail)
        for (uint32_t m = open_mask; m; m &= m - 1) {
            const int b = ctz32(m);
            const uint32_t d = ((uint32_t) bseq[b * STRIDE] - h1) & 0xFFFFu;
            if (d < dist) { dist = d; oldest = b; }
        }
        head = h1 + dist;
    }

    // pack one ready codeword into the payload, LSB first (icer_encoding.c:114-139).  One emission per step keeps pace (a step starts
    // at most one codeword); emitting two per call was measured and changes nothing.
    ICER_HD void pump() {
        if (emit_pos == head) return;
        const uint32_t v = ring[emit_pos & (ICER_RING - 1)];
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
    ICER_HD uint32_t forced_word(int bin, uint32_t &nb) {
        const uint32_t st = bst[bin * STRIDE];
        if (bin > 7) {
            if (st == (uint32_t) T.gol_m[bin] - 1) { nb = 1; return 1; }
            return golomb_word(bin, st, nb);
        }
        const uint32_t pre = st & 0xFF, fill = st >> 8;
        const uint32_t fl = T.flush[bin][pre][fill];
        const uint32_t word = (pre | ((fl & 0xF) << fill)) & 0x3FF;
        const uint32_t e = T.v2v[bin][word & 31];
        nb = (e >> 4) & 0xF;
        return e >> 8;
    }
    // true when `bin` needs a new codeword but 2048 are already outstanding (icer_encoding.c:59-64)
    ICER_HD bool window_full_for(int bin) const { return !((open_mask >> bin) & 1u) && tail - head >= ICER_RING; }
    // make sure `bin` has an open codeword (icer_encoding.c:56-67); the caller has dealt with a full window
    ICER_HD void open_bin(int bin) {
        if (open_mask & (1u << bin)) return;
        while (tail - emit_pos >= ICER_RING) pump();       // physical ring space (emission may lag behind `head`)
        if ((open_mask) == 0) oldest = bin;
        bseq[bin * STRIDE] = (uint16_t) tail++;
        bst[bin * STRIDE] = 0;
        open_mask |= 1u << bin;
    }

    ICER_HD void begin_chunk() { col = 1; }
    ICER_HD bool chunk_done(int ncols) const { return col > ncols; }

    // One micro-operation: skip a run of quiet pixels, apply a pending run, or code one magnitude / sign decision.  The set-up
    // branches only derive (context, bit); the lanes meet again before the common part -- adaptive counts, bin, interleave window,
    // codeword building, completion -- which therefore runs once per step for the whole warp.
    ICER_HD void step(const EncChunk &k, int sb, unsigned lanes) {
        enum { OP_NONE = 0, OP_CODE = 1, OP_DONE = 3 };
        int op = OP_NONE;
        int kind = 0;                                       // 1: magnitude decision, 2: ... of a quiet pixel, 4: ... whose one starts the sign decision; 8: sign decision
        int bin = 0, ctx = -1;
        uint32_t bit = 0, code = 0, nb = 0;
        bool decide = false;
        if (phase == 0) {
            const uint32_t nqs = k.nq >> col;               // bit 0: this column; never zero (the bits above the chunk are set)
            const uint32_t r = fz_left ? (uint32_t) ctz32(nqs) : 0u;
            if (r) {                                        // r quiet pixels ahead: they only bump counters
                const uint32_t n = umin32(r, fz_left);
                fz_run += n; fz_left -= n; col += (int) n;
            } else if (fz_run) {
                // apply `fz_run` (context 0, bit 0) decisions at once
                bin = fz_bin;
                const uint32_t zt = cnt[0];
                uint32_t zz = ICER_CNT_Z(zt) + fz_run, tt = ICER_CNT_TOT(zt) + fz_run;
                if (tt >= 500) { tt >>= 1; if (zz > tt) zz >>= 1; }
                cnt[0] = ICER_CNT_PACK(bin, zz, tt);
                const uint32_t kk = bst[bin * STRIDE] + fz_run;
                fz_run = 0; fz_left = 0;
                if (kk >= T.gol_m[bin]) { op = OP_DONE; code = 1; nb = 1; }
                else bst[bin * STRIDE] = (uint16_t) kk;
            } else decide = true;
        } else decide = true;
        if (decide) {
            // Magnitude and sign decisions share ONE set-up: the same neighbour-mask extraction runs on the significance words (magnitude)
            // or on the "negative and significant" words (sign), and one table look-up gives the context (and the predicted sign).
            const bool sg = phase != 0;
            const uint32_t wA = sg ? k.nsA : k.aS, wB0 = sg ? k.nsBc : k.bSc, wB2 = sg ? k.nsBp : k.bP, wC = sg ? k.nsC : k.cP;
            const int s = col - 1;
            const uint32_t nm = ((wA >> s) & 7u) | (((wB0 >> s) & 1u) << 3) | (((wB2 >> s) & 4u) << 2) | (((wC >> s) & 7u) << 5);
            const uint32_t e = T.lut[sg ? 4 + (sb == 1) : sb][nm];
            const int selfp = sg ? 0 : (int) ((k.bP >> col) & 1u);
            const int cat = selfp ? 1 + (int) ((k.b2 >> col) & 1u) + (int) ((k.b3 >> col) & 1u) : 0;     // icer_context_modeller.c:631-634
            bit = (((sg ? k.bN : k.bB) >> col) ^ (e >> 7)) & 1u;       // sign: agreement with the prediction; magnitude entries have no bit 7
            if (cat != 3) ctx = ctx_of(cat, e, nm);                     // category 3: p(0) = 1/2 -> uncoded bin, no context
            kind = sg ? 8 : (1 | (((k.nq >> col) & 1u) ? 0 : 2) | ((cat == 0 && bit) ? 4 : 0));
        }
        ICER_CONVERGE(lanes);       // ---- all lanes meet here ----
        if (kind) {
            uint32_t z = 0, tot = 0;
            bool inv = false;
            if (ctx >= 0) {
                const uint32_t zt = cnt[ctx * STRIDE];
                z = ICER_CNT_Z(zt); tot = ICER_CNT_TOT(zt);
                uint32_t zz = z;
                inv = zz < (tot >> 1);
                if (inv) zz = tot - zz;
                bin = bin_cached(T, zz, tot, ICER_CNT_BIN(zt));
            }
            if (window_full_for(bin)) {
                // window full: force-complete the oldest codeword (icer_encoding.c:59-64); the decision is retried in the next step
                bin = oldest;
                code = forced_word(bin, nb);
                op = OP_DONE;
            } else if ((kind & 2) && ctx >= 0 && !inv && bin > 7) {
                // enter the quiet fast path: this and the following quiet pixels only bump counters
                open_bin(bin);
                const uint32_t n1 = (uint32_t) T.gol_m[bin] - bst[bin * STRIDE];
                const uint32_t n2 = zeros_in_same_bin(T, bin, z, tot);
                const uint32_t n3 = 500u - tot;
                fz_left = umin32(n1, umin32(n2, n3)) - 1;
                fz_run = 1;
                fz_bin = bin;
                col++;
            } else {
                if (kind & 4) phase = 1;                    // first 1 of this coefficient: its sign follows
                else { phase = 0; col++; }
                if (ctx >= 0) {
                    tot++;
                    z += bit ^ 1u;
                    if (tot >= 500) { tot >>= 1; if (z > tot) z >>= 1; }   // icer_context_modeller.c:396-402
                    cnt[ctx * STRIDE] = ICER_CNT_PACK(bin, z, tot);
                    bit ^= inv ? 1u : 0u;
                }
                open_bin(bin);      // (one call site for both branches was measured: k_encode 18.6 -> 19.1 ms)
                const uint32_t stt = bst[bin * STRIDE];
                if (bin > 7) {                                        // Golomb run-length bins
                    if (bit) {
                        code = golomb_word(bin, stt, nb);
                        op = OP_DONE;
                    } else {
                        bst[bin * STRIDE] = (uint16_t) (stt + 1);
                        if (stt + 1 >= T.gol_m[bin]) { code = 1; nb = 1; op = OP_DONE; }
                    }
                } else {                                              // variable-to-variable bins; bin 0 (uncoded) is the identity code
                    const uint32_t fill = stt >> 8;
                    const uint32_t word = (stt & 0xFF) | (bit << fill);
                    const uint32_t e = T.v2v[bin][word & 31];
                    if ((e & 0xF) == fill + 1) { code = e >> 8; nb = (e >> 4) & 0xF; op = OP_DONE; }
                    else bst[bin * STRIDE] = (uint16_t) (word | ((fill + 1) << 8));
                }
            }
        }
        ICER_CONVERGE(lanes);
        if (op == OP_DONE) complete(bin, code, nb);
    }

    // force-complete the oldest codeword (end of stream)
    ICER_HD void force_oldest() {
        uint32_t nb;
        const int b = oldest;
        const uint32_t code = forced_word(b, nb);
        complete(b, code, nb);
    }

    ICER_HD void finish(uint32_t &bits_out) {
        if (fz_run) {   // tail of a fast run
            const int bin = fz_bin;
            const uint32_t kk = bst[bin * STRIDE] + fz_run;
            if (kk >= T.gol_m[bin]) complete(bin, 1, 1);
            else bst[bin * STRIDE] = (uint16_t) kk;
            fz_run = 0;
        }
        while (open_mask) force_oldest();           // icer_context_modeller.c:452-455
        while (emit_pos != head) pump();
        bits_out = nwords * 32u + (uint32_t) nacc;
        if (nacc > 0) { if (nwords < cap) out[nwords] = (uint32_t) acc; else overflow = true; }
    }
};

// =========================================================================================================================
// Decoder
// =========================================================================================================================
struct BitReader {
    const uint32_t *wp = nullptr;   // next aligned stream word to load
    int left = 0;                   // bytes from wp to the end of THIS FRAME's stream (bytes at or beyond it read as zero, like the oracle and the legacy calls)
    unsigned long long bb = 0;      // bit buffer, LSB = next stream bit
    int bbn = 0;                    // bits in bb that came from loads (zeros beyond)
    uint32_t total_bits = 0;
    uint32_t consumed = 0;          // bits consumed
    uint32_t words = 0;             // codewords consumed

    ICER_HD uint32_t load_word() {
        uint32_t w = 0;
        if (left >= 4) w = *wp;
        else if (left > 0) {           // tail: byte loads, nothing behind the end is touched (the caller's buffer may stop there)
            const uint8_t *bp = reinterpret_cast<const uint8_t *>(wp);
            w = bp[0];
            if (left > 1) w |= (uint32_t) bp[1] << 8;
            if (left > 2) w |= (uint32_t) bp[2] << 16;
        }
        wp++;
        left -= 4;
        return w;
    }
    ICER_HD void start(const uint8_t *base, const uint8_t *e, uint32_t nbits) {
        total_bits = nbits;
        const uintptr_t ad = reinterpret_cast<uintptr_t>(base);
        wp = reinterpret_cast<const uint32_t *>(ad & ~(uintptr_t) 3);
        const long long room = (long long) (e - reinterpret_cast<const uint8_t *>(wp));
        left = room > 0x7FFFFFF0ll ? 0x7FFFFFF0 : (room < -0x7FFFFFF0ll ? -0x7FFFFFF0 : (int) room);
        const int sh = (int) (ad & 3) * 8;
        bb = (unsigned long long) (load_word() >> sh);
        bbn = 32 - sh;
    }
    ICER_HD void fill() {           // afterwards at least 33 valid bits are buffered
        if (bbn <= 32) { bb |= (unsigned long long) load_word() << bbn; bbn += 32; }
    }
    ICER_HD uint32_t peek(uint32_t n) const { return (uint32_t) bb & ((1u << n) - 1u); }
    ICER_HD void skip(uint32_t n) { bb >>= n; bbn -= (int) n; consumed += n; }

    // icer_get_bits_from_codeword / icer_pop_bits_from_codeword (icer_decoding.c:59-105): byte-chunked read that
    // fails only when a chunk is larger than the whole segment (`decoded_bits_total` is never advanced there)
    ICER_HD int rd(uint32_t n, bool consume) {
        if (total_bits < 8) {
            uint32_t off = consumed & 7, lft = n;
            while (lft) {
                uint32_t take = 8 - off; if (take > lft) take = lft;
                if (take > total_bits) return ST_OUT_OF_DATA;
                lft -= take; off = (off + take) & 7;
            }
        }
        const int v = (int) peek(n);
        if (consume) skip(n);
        return v;
    }

    // refill `bin` with the next codeword (icer_decoding.c:127-181); returns an error code or OK
    ICER_HD int refill(const CoreTables &T, int bin, int &depth, uint32_t &low) {
        depth = 0; low = 0;
        fill();
        if (total_bits >= 8) {
            // rd() can only fail when a byte-sized chunk is larger than the whole segment: not here -- plain peeks and skips
            if (bin > 7) {
                const uint32_t l = T.gol_l[bin], gi = T.gol_i[bin];
                if (peek(1)) {
                    skip(1);
                    depth = T.gol_m[bin];
                } else {
                    uint32_t k = (brev32(peek(l)) >> (32 - l)) & 0xFFFFu;
                    if (k < gi) skip(l);
                    else { k = (((brev32(peek(l + 1)) >> (31 - l)) & 0xFFFFu) - gi) & 0xFFFFu; skip(l + 1); }
                    depth = 1 + (int) k;
                    low = 1;
                }
            } else if (bin != 0) {
                const uint32_t e = T.v2v[bin][peek(5)];
                const uint32_t L = e & 0xF;
                if (L == 0) return ST_INVALID_DATA;                    // cannot happen: the codes are complete
                if (L >= total_bits) return ST_OUT_OF_DATA;            // the reference checks n + 1 >= total for every bit it reads
                low = e >> 8;
                depth = (int) ((e >> 4) & 0xF);
                skip(L);
            } else {
                low = peek(1);
                skip(1);
                depth = 1;
            }
            words++;
            return ST_OK;
        }
        if (bin > 7) {
            const uint32_t l = T.gol_l[bin], gi = T.gol_i[bin];
            if (peek(1)) {
                rd(1, true);
                depth = T.gol_m[bin];
            } else {
                uint32_t k = (brev32((uint32_t) rd(l, false) & 0xFFFFu) >> (32 - l)) & 0xFFFFu;
                if (k < gi) {
                    rd(l, true);
                } else {
                    k = (brev32((uint32_t) rd(l + 1, true) & 0xFFFFu) >> (31 - l)) & 0xFFFFu;
                    k = (k - gi) & 0xFFFFu;
                }
                depth = 1 + (int) k;
                low = 1;
            }
        } else if (bin != 0) {
            const uint32_t e = T.v2v[bin][peek(5)];
            const uint32_t L = e & 0xF;
            if (L == 0) return ST_INVALID_DATA;                    // cannot happen: the codes are complete
            if (L >= total_bits) return ST_OUT_OF_DATA;            // the reference checks n + 1 >= total for every bit it reads
            low = e >> 8;
            depth = (int) ((e >> 4) & 0xF);
            const int word = (int) peek(L);
            if (word != rd(L, true)) return ST_INVALID_DATA;
        } else {
            const int v = rd(1, true);
            if (v == ST_OUT_OF_DATA) return v;
            low = v != 0;
            depth = 1;
        }
        words++;
        return ST_OK;
    }
};

// The bit-slice words of one chunk as the decoder of plane `lsb` sees them.  Pixels this plane has not visited yet are significant
// "at lsb" exactly when they are "at lsb + 1", so one rule serves both sides of the scan position.
struct DecChunk {
    uint32_t aS, cP, bP, b2, b3;    // row above: S_cur (this plane's own output); row below / this row: S_prev; this row: S_{lsb+2}, S_{lsb+3}
    uint32_t nsA, nsC;              // negative & significant: row above (cur), row below (prev)
    uint32_t bNp;                   // this row: sign bits known from the planes above
    uint32_t bB, bNn;               // this row, growing: magnitude bits / sign bits decoded by this plane (bit 0: carried over from the chunk before)
    uint32_t nq;                    // static part of the non-quiet mask

    ICER_HD void set(uint32_t aS_, uint32_t aN, uint32_t bP_, uint32_t bNp_, uint32_t b2_, uint32_t b3_, uint32_t cP_, uint32_t cN,
                     uint32_t carryB, uint32_t carryN, int ncols) {
        aS = aS_; cP = cP_; bP = bP_; b2 = b2_; b3 = b3_;
        nsA = aN & aS_; nsC = cN & cP_; bNp = bNp_;
        bB = carryB & 1u; bNn = carryN & 1u;
        nq = dilate3(aS) | dilate3(cP) | dilate3(bP) | (0xFFFFFFFEu << ncols);
    }
};

template <int STRIDE>
struct DecLane {
    const CoreTables &T;
    uint32_t *cnt;      // [17] contexts
    uint32_t *stlow;    // [17] per bin: codeword stamp (24 bits) | low bits of the pending stack << 24
    int16_t *depth;     // [17] per bin: pending bits of the stack (<= the largest Golomb m, 512)
    BitReader br;
    uint32_t fz_left = 0, fz_run = 0;
    int fz_bin = 0;
    int col = 1, phase = 0;
    bool run;           // still decoding (false: absent plane, or stopped by a decode error)
    int err = ST_OK;

    ICER_HD DecLane(const CoreTables &t, uint32_t *c, uint32_t *sl, int16_t *d, bool alive) : T(t), cnt(c), stlow(sl), depth(d), run(alive) {
        for (int k = 0; k < 17; k++) { cnt[k * STRIDE] = ICER_CNT_INIT; depth[k * STRIDE] = 0; stlow[k * STRIDE] = 0; }
    }
    ICER_HD void begin_chunk() { col = 1; }
    ICER_HD bool chunk_done(int ncols) const { return !run || col > ncols; }

    // One micro-operation (see EncLane::step): the set-up branches meet before the common decoding tail.
    ICER_HD void step(DecChunk &k, int sb, unsigned lanes) {
        int bin = 0, ctx = -1;
        int kind = -1;                                          // -1: nothing to decode in this step, 0: sign decision, 2: magnitude decision of a category-0 pixel (a one starts the sign decision), 3: other magnitude decision
        uint32_t z = 0, tot = 0, e = 0;
        bool inv = false, quiet = false, decide = false;
        if (phase == 0) {
            const uint32_t nqs = (k.nq | (k.bB << 1)) >> col;
            const uint32_t r = fz_left ? (uint32_t) ctz32(nqs) : 0u;
            if (r) {
                const uint32_t n = umin32(r, fz_left);
                fz_run += n; fz_left -= n; col += (int) n;
            } else {
                if (fz_run) {                                   // apply the pending run (no stream access), then go on with this pixel
                    const int fb = fz_bin;
                    const uint32_t zt = cnt[0];
                    uint32_t zz = ICER_CNT_Z(zt) + fz_run, tt = ICER_CNT_TOT(zt) + fz_run;
                    if (tt >= 500) { tt >>= 1; if (zz > tt) zz >>= 1; }
                    cnt[0] = ICER_CNT_PACK(fb, zz, tt);
                    depth[fb * STRIDE] = (int16_t) (depth[fb * STRIDE] - (int) fz_run);
                    fz_run = 0; fz_left = 0;
                }
                quiet = !(nqs & 1u);
                decide = true;
            }
        } else decide = true;
        if (decide) {
            // one set-up for magnitude and sign decisions (see EncLane::step)
            const bool sg = phase != 0;
            const uint32_t bSc = k.bP | k.bB;
            const uint32_t wA = sg ? k.nsA : k.aS, wB0 = sg ? ((k.bNp | k.bNn) & bSc) : bSc, wB2 = sg ? (k.bNp & k.bP) : k.bP, wC = sg ? k.nsC : k.cP;
            const int s = col - 1;
            const uint32_t nm = ((wA >> s) & 7u) | (((wB0 >> s) & 1u) << 3) | (((wB2 >> s) & 4u) << 2) | (((wC >> s) & 7u) << 5);
            e = T.lut[sg ? 4 + (sb == 1) : sb][nm];
            const int selfp = sg ? 0 : (int) ((k.bP >> col) & 1u);
            const int cat = selfp ? 1 + (int) ((k.b2 >> col) & 1u) + (int) ((k.b3 >> col) & 1u) : 0;
            if (cat != 3) ctx = ctx_of(cat, e, nm);
            kind = sg ? 0 : ((cat == 0) ? 2 : 3);
        }
        ICER_CONVERGE(lanes);       // ---- all lanes meet here ----
        int dep = 0;
        uint32_t low = 0, stamp = 0;
        if (kind >= 0) {
            if (ctx >= 0) {
                const uint32_t zt = cnt[ctx * STRIDE];
                z = ICER_CNT_Z(zt); tot = ICER_CNT_TOT(zt);
                uint32_t zz = z;
                inv = zz < (tot >> 1);
                if (inv) zz = tot - zz;
                bin = bin_cached(T, zz, tot, ICER_CNT_BIN(zt));
            }
            // next decoded bit of `bin` (icer_decoding.c:127, :186-192)
            dep = depth[bin * STRIDE];
            const uint32_t slw = stlow[bin * STRIDE];
            low = slw >> 24;
            stamp = slw & 0xFFFFFFu;
            if (dep <= 0 || br.words - stamp >= ICER_RING) {
                const int rc = br.refill(T, bin, dep, low);
                if (rc != ST_OK) { run = false; err = rc; kind = -1; }
                stamp = br.words & 0xFFFFFFu;
            }
        }
        ICER_CONVERGE(lanes);       // the lanes that had to fetch a codeword are back
        if (kind < 0) return;
        if (kind && quiet && !inv && bin > 7 && (uint32_t) dep > (low & 1u)) {
            // quiet fast path: the run of pending zeros of this Golomb bin is taken without touching the stack
            const uint32_t zeros = (uint32_t) dep - (low & 1u);
            const uint32_t n2 = zeros_in_same_bin(T, bin, z, tot);
            const uint32_t n3 = 500u - tot;
            depth[bin * STRIDE] = (int16_t) dep;
            stlow[bin * STRIDE] = stamp | (low << 24);
            fz_left = umin32(zeros, umin32(n2, n3)) - 1;
            fz_run = 1;
            fz_bin = bin;
            col++;
            return;
        }
        const int top = dep - 1;
        uint32_t b = 0;
        if (top >= 0 && top < 8) { b = (low >> top) & 1u; low &= ~(1u << top); }
        depth[bin * STRIDE] = (int16_t) (dep - 1);
        stlow[bin * STRIDE] = stamp | (low << 24);
        const uint32_t bit = inv ? (b ^ 1u) : b;
        if (ctx >= 0) {
            tot++;
            z += bit ^ 1u;
            if (tot >= 500) { tot >>= 1; if (z > tot) z >>= 1; }
            cnt[ctx * STRIDE] = ICER_CNT_PACK(bin, z, tot);
        }
        if (kind) {
            k.bB |= bit << col;
            if (bit && kind == 2) phase = 1; else col++;
        } else {
            k.bNn |= ((bit ^ (e >> 7)) & 1u) << col;
            phase = 0; col++;
        }
    }
};

}  // namespace icer_core
