// tests/emul/emul_coder.cpp -- TEST INFRASTRUCTURE: host build of the lane state machines of icer_compression_b200/csrc/coder_core.h.
// One "lane" codes one segment plane exactly as a lane of k_encode / k_decode does (same chunk loop, same halo rules for the words the
// decoder produces); tests/test_emul_core.py compares the result with the oracle.  Not part of libicer_b200.so.
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "../../icer_compression_b200/csrc/coder_tables.h"

using namespace icer_core;

namespace {
struct Seg {
    const uint16_t *p; size_t w, h, rs;
    // window word of row r, chunk j: bit b <-> column 30 j + b - 1
    template <typename F> uint32_t word(long r, size_t j, F f) const {
        if (r < 0 || r >= (long) h) return 0;
        uint32_t v = 0;
        for (int b = 0; b < 32; b++) {
            const long c = (long) (ICER_CHUNK * j) + b - 1;
            if (c < 0 || c >= (long) w) continue;
            if (f(p[r * rs + c])) v |= 1u << b;
        }
        return v;
    }
    uint32_t S(long r, size_t j, int pl) const { return pl > 14 ? 0u : word(r, j, [pl](uint16_t x) { return ((x & 0x7FFF) >> pl) != 0; }); }
    uint32_t B(long r, size_t j, int pl) const { return word(r, j, [pl](uint16_t x) { return ((x >> pl) & 1) != 0; }); }
    uint32_t N(long r, size_t j) const { return word(r, j, [](uint16_t x) { return (x & 0x8000) != 0; }); }
};
CoreTables g_enc, g_dec;
bool g_ready = false;
void ready() {
    if (g_ready) return;
    const HostTables h = icer_tables::make_tables();
    icer_tables::fill_core_tables(g_enc, h, false);
    icer_tables::fill_core_tables(g_dec, h, true);
    g_ready = true;
}
}  // namespace

extern "C" long emul_encode_segment_plane(const uint16_t *seg, size_t w, size_t h, size_t rowstride, int subband, int lsb, uint8_t *out,
                                          size_t max_bytes, uint32_t *steps_out) {
    ready();
    Seg s{seg, w, h, rowstride};
    uint32_t cnt[17]; uint16_t bseq[17]; uint16_t bst[17];
    std::vector<uint16_t> ring(ICER_RING);
    std::vector<uint32_t> words(max_bytes / 4 + 2);
    EncLane<1> ln(g_enc, cnt, bseq, bst, ring.data(), words.data(), (uint32_t) (max_bytes / 4));
    const size_t cw = (w + ICER_CHUNK - 1) / ICER_CHUNK;
    uint32_t steps = 0;
    for (size_t r = 0; r < h; r++)
        for (size_t j = 0; j < cw; j++) {
            const int ncols = (int) ((w - j * ICER_CHUNK) < ICER_CHUNK ? (w - j * ICER_CHUNK) : ICER_CHUNK);
            EncChunk k;
            const long ra = (long) r - 1, rc = (long) r + 1;
            k.set(s.S(ra, j, lsb + 1), s.B(ra, j, lsb), s.N(ra, j), s.S(r, j, lsb + 1), s.B(r, j, lsb), s.N(r, j), s.S(r, j, lsb + 2), s.S(r, j, lsb + 3),
                  s.S(rc, j, lsb + 1), s.N(rc, j), ncols);
            ln.begin_chunk();
            while (!ln.chunk_done(ncols)) { ln.step(k, subband, 1u); ln.pump(); steps++; }
        }
    uint32_t bits = 0;
    ln.finish(bits);
    if (steps_out) *steps_out = steps;
    if (ln.overflow) return -2;
    const size_t nbytes = (bits + 7) / 8;
    for (size_t i = 0; i < nbytes; i++) out[i] = (uint8_t) (words[i / 4] >> (8 * (i % 4)));
    return (long) bits;
}

// decodes plane `lsb` into seg (which holds the planes above); in / in_bytes: the payload and the bytes readable behind its start
extern "C" int emul_decode_segment_plane(uint16_t *seg, size_t w, size_t h, size_t rowstride, int subband, int lsb, const uint8_t *in,
                                         size_t in_bytes, uint32_t nbits) {
    ready();
    Seg s{seg, w, h, rowstride};
    uint32_t cnt[17]; uint32_t stlow[17]; int16_t depth[17];
    DecLane<1> ln(g_dec, cnt, stlow, depth, true);
    ln.br.start(in, in + in_bytes, nbits);
    const size_t cw = (w + ICER_CHUNK - 1) / ICER_CHUNK;
    // the words this plane produces, as the kernel stores them (with both halo bits)
    std::vector<uint32_t> Sw(h * cw, 0), Nw(h * cw, 0), Bw(h * cw, 0);
    for (size_t r = 0; r < h; r++) {
        uint32_t carryB = 0, carryN = 0;
        uint32_t pendS = 0, pendN = 0;
        for (size_t j = 0; j < cw; j++) {
            const int ncols = (int) ((w - j * ICER_CHUNK) < ICER_CHUNK ? (w - j * ICER_CHUNK) : ICER_CHUNK);
            DecChunk k;
            const uint32_t aS = r ? Sw[(r - 1) * cw + j] : 0u, aN = r ? Nw[(r - 1) * cw + j] : 0u;
            const long rc = (long) r + 1;
            k.set(aS, aN, s.S(r, j, lsb + 1), s.N(r, j), s.S(r, j, lsb + 2), s.S(r, j, lsb + 3), s.S(rc, j, lsb + 1), s.N(rc, j), carryB, carryN, ncols);
            ln.begin_chunk();
            while (!ln.chunk_done(ncols)) ln.step(k, subband, 1u);
            // this chunk's words; the right halo is patched when the next chunk's first column is known
            const uint32_t sw = k.bP | k.bB, nw = k.bNp | k.bNn;
            if (j > 0) {
                pendS |= ((k.bB >> 1) & 1u) << 31; pendN |= ((k.bNn >> 1) & 1u) << 31;
                Sw[r * cw + j - 1] = pendS; Nw[r * cw + j - 1] = pendN;
            }
            pendS = sw; pendN = nw;
            Bw[r * cw + j] = k.bB;
            carryB = k.bB >> ICER_CHUNK; carryN = k.bNn >> ICER_CHUNK;
        }
        Sw[r * cw + cw - 1] = pendS; Nw[r * cw + cw - 1] = pendN;
    }
    for (size_t r = 0; r < h; r++)
        for (size_t c = 0; c < w; c++) {
            const size_t j = c / ICER_CHUNK; const int b = (int) (c % ICER_CHUNK) + 1;
            uint16_t v = seg[r * rowstride + c];
            if ((Bw[r * cw + j] >> b) & 1u) v |= (uint16_t) (1u << lsb);
            if ((Nw[r * cw + j] >> b) & 1u) v |= 0x8000u;
            seg[r * rowstride + c] = v;
        }
    return ln.err;
}
