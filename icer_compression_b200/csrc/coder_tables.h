// coder_tables.h -- the ICER code tables (lib_icer builds the same at run time in src/icer_init.c:24-255; context tables
// icer_config.c:18-107) and their packed forms for the coder kernels.  Header-only plain C++ so that the host emulation test of
// coder_core.h (tests/emul) uses exactly the tables the device gets.
#pragma once
#include <stdint.h>
#include <string.h>
#include "coder_core.h"

struct V2VEntry { uint8_t in_bits, out_bits, out; };
struct HostTables {
    V2VEntry enc[8][32];       // [bin 0..7][input word]   (bin 0 unused)
    V2VEntry dec[8][32];       // [bin][emitted word] -> input word, bit-reversed
    uint8_t flush[8][9][6];    // [bin][prefix][len] -> pad_word | pad_len << 4
    uint16_t gol_m[17], gol_l[17], gol_i[17];
    uint32_t cutoff[17];
};

namespace icer_tables {
// {bin, input word (first bit = bit 0), input bits, output word (first emitted bit = bit 0), output bits}
static const uint8_t kV2V[][5] = {
    {1, 0x01, 2, 0x02, 2}, {1, 0x03, 3, 0x03, 3}, {1, 0x07, 4, 0x0F, 4}, {1, 0x0F, 4, 0x10, 5}, {1, 0x02, 2, 0x01, 2},
    {1, 0x04, 3, 0x04, 3}, {1, 0x08, 4, 0x08, 4}, {1, 0x10, 5, 0x00, 5}, {1, 0x00, 5, 0x07, 4},
    {2, 0x02, 2, 0x01, 2}, {2, 0x04, 3, 0x00, 2}, {2, 0x00, 4, 0x03, 3}, {2, 0x18, 5, 0x12, 5}, {2, 0x08, 5, 0x0F, 4},
    {2, 0x01, 2, 0x06, 3}, {2, 0x03, 4, 0x07, 4}, {2, 0x0B, 4, 0x02, 5}, {2, 0x07, 3, 0x0A, 4},
    {3, 0x02, 2, 0x02, 2}, {3, 0x04, 3, 0x03, 3}, {3, 0x00, 3, 0x00, 2}, {3, 0x01, 2, 0x01, 2}, {3, 0x03, 2, 0x07, 3},
    {4, 0x00, 2, 0x01, 1}, {4, 0x02, 3, 0x00, 3}, {4, 0x06, 3, 0x0A, 4}, {4, 0x05, 3, 0x02, 4}, {4, 0x09, 4, 0x0E, 4},
    {4, 0x01, 5, 0x04, 4}, {4, 0x11, 5, 0x06, 5}, {4, 0x03, 3, 0x0C, 4}, {4, 0x07, 3, 0x16, 5},
    {5, 0x01, 1, 0x02, 2}, {5, 0x02, 3, 0x03, 3}, {5, 0x06, 3, 0x0F, 4}, {5, 0x04, 3, 0x05, 3}, {5, 0x08, 4, 0x01, 3},
    {5, 0x10, 5, 0x07, 4}, {5, 0x00, 5, 0x00, 2},
    {6, 0x00, 3, 0x00, 1}, {6, 0x04, 3, 0x01, 3}, {6, 0x02, 3, 0x05, 3}, {6, 0x06, 3, 0x0F, 5}, {6, 0x03, 2, 0x07, 4},
    {6, 0x01, 3, 0x03, 3}, {6, 0x05, 3, 0x1F, 5},
    {7, 0x02, 2, 0x05, 3}, {7, 0x04, 3, 0x01, 3}, {7, 0x00, 4, 0x00, 1}, {7, 0x08, 5, 0x07, 4}, {7, 0x18, 5, 0x0F, 5},
    {7, 0x01, 2, 0x03, 3}, {7, 0x03, 2, 0x1F, 5}};
// {bin, incomplete prefix, prefix bits, pad word, pad bits}: what a forced flush appends
static const uint8_t kFlush[][5] = {
    {1, 0x1, 1, 0, 1}, {1, 0x3, 2, 0, 1}, {1, 0x7, 3, 0, 1}, {1, 0x0, 1, 1, 1}, {1, 0x0, 2, 1, 1}, {1, 0x0, 3, 1, 1},
    {1, 0x0, 4, 0, 1},
    {2, 0x0, 1, 1, 1}, {2, 0x0, 2, 1, 1}, {2, 0x0, 3, 0, 1}, {2, 0x8, 4, 0, 1}, {2, 0x1, 1, 0, 1}, {2, 0x3, 2, 1, 1},
    {2, 0x3, 3, 0, 1},
    {3, 0x0, 1, 1, 1}, {3, 0x0, 2, 0, 1}, {3, 0x1, 1, 0, 1},
    {4, 0x0, 1, 0, 1}, {4, 0x2, 2, 0, 1}, {4, 0x1, 2, 1, 1}, {4, 0x1, 3, 1, 1}, {4, 0x1, 4, 0, 1}, {4, 0x1, 1, 1, 2},
    {4, 0x3, 2, 0, 1},
    {5, 0x0, 1, 1, 2}, {5, 0x1, 2, 0, 1}, {5, 0x0, 2, 1, 1}, {5, 0x0, 3, 1, 1}, {5, 0x0, 4, 0, 1},
    {6, 0x0, 1, 0, 2}, {6, 0x0, 2, 0, 1}, {6, 0x2, 2, 0, 1}, {6, 0x1, 1, 1, 1}, {6, 0x1, 2, 0, 1},
    {7, 0x0, 1, 1, 1}, {7, 0x0, 2, 1, 1}, {7, 0x0, 3, 0, 1}, {7, 0x8, 4, 0, 1}, {7, 0x1, 1, 0, 1}};
static const int kGolombM[17] = {0, 0, 0, 0, 0, 0, 0, 0, 5, 6, 7, 11, 17, 31, 70, 200, 512};
static const uint32_t kCutoff[17] = {35298, 37345, 40503, 43591, 47480, 50133, 53645, 55902, 57755,
                              58894, 60437, 62267, 63613, 64557, 65134, 65392, 65536};

inline unsigned rev(unsigned v, unsigned n) { unsigned r = 0; while (n--) { r = (r << 1) | (v & 1); v >>= 1; } return r; }

inline HostTables make_tables() {
    HostTables t;
    memset(&t, 0, sizeof t);
    for (auto &e : kV2V) {
        t.enc[e[0]][e[1]] = V2VEntry{e[2], e[4], e[3]};
        t.dec[e[0]][e[3]] = V2VEntry{e[4], e[2], (uint8_t) rev(e[1], e[2])};
    }
    for (auto &e : kFlush) t.flush[e[0]][e[1]][e[2]] = (uint8_t) (e[3] | (e[4] << 4));
    for (int b = 0; b < 17; b++) {
        t.cutoff[b] = kCutoff[b];
        if (kGolombM[b] > 0) {
            unsigned m = (unsigned) kGolombM[b], l = 0;
            while ((1u << l) < m) l++;
            t.gol_m[b] = (uint16_t) m; t.gol_l[b] = (uint16_t) l; t.gol_i[b] = (uint16_t) ((1u << l) - m);
        }
    }
    return t;
}

// icer_config.c:26-67
static const uint8_t kCtxHV[45] = {0, 1, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 6, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7,
                                   8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8};
static const uint8_t kCtxHH[25] = {0, 3, 6, 8, 8, 1, 4, 7, 8, 8, 2, 5, 7, 8, 8, 2, 5, 7, 8, 8, 2, 5, 7, 8, 8};
static const uint8_t kCtxSign[25] = {14, 14, 15, 16, 16, 14, 14, 15, 16, 16, 13, 13, 12, 13, 13, 16, 16, 15, 14, 14, 16, 16, 15, 14, 14};
static const uint8_t kSignPred[25] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// the tables in the form the lane state machines use (coder_core.h); `decoder` selects which variable-to-variable table is staged
inline void fill_core_tables(icer_core::CoreTables &d, const HostTables &h, bool decoder) {
    memset(&d, 0, sizeof d);
    for (int b = 0; b < 16; b++) d.cutoff[b] = h.cutoff[b];
    for (int b = 0; b < 17; b++) {
        d.cut[b].lo = b ? h.cutoff[b - 1] : 0u;
        d.cut[b].hi = b < 16 ? h.cutoff[b] : 65537u;
        d.gol_m[b] = h.gol_m[b]; d.gol_l[b] = h.gol_l[b]; d.gol_i[b] = h.gol_i[b];
    }
    for (int b = 0; b < 8; b++)
        for (int x = 0; x < 32; x++) {
            if (!decoder) {
                d.v2v[b][x] = (uint16_t) (h.enc[b][x].in_bits | (h.enc[b][x].out_bits << 4) | (h.enc[b][x].out << 8));
                if (b == 0) d.v2v[b][x] = (uint16_t) (1 | (1 << 4) | ((x & 1) << 8));      // the uncoded bin as a one-bit identity code
            } else {
                // the variable-to-variable codes are complete prefix codes of at most 5 bits: the next 5 stream bits select the
                // codeword directly (same first-match order as the bit-by-bit loop of icer_decoding.c:155-177)
                d.v2v[b][x] = 0;
                for (int L = 1; L <= 5; L++) {
                    const V2VEntry &e = h.dec[b][x & ((1 << L) - 1)];
                    if (e.in_bits == L) { d.v2v[b][x] = (uint16_t) (L | (e.out_bits << 4) | (e.out << 8)); break; }
                }
            }
        }
    memcpy(d.flush, h.flush, sizeof d.flush);
    // neighbour mask -> context of a category-0 pixel (icer_context_modeller.c:355-385): bits 0,1,2 row above, 3 left, 4 right, 5,6,7 row below
    for (int sb = 0; sb < 4; sb++)
        for (int nm = 0; nm < 256; nm++) {
            int hh = ((nm >> 3) & 1) + ((nm >> 4) & 1), vv = ((nm >> 1) & 1) + ((nm >> 6) & 1);
            const int dd = (nm & 1) + ((nm >> 2) & 1) + ((nm >> 5) & 1) + ((nm >> 7) & 1);
            if (sb == 1) { const int t = hh; hh = vv; vv = t; }
            d.lut[sb][nm] = sb == 3 ? kCtxHH[(hh + vv) * 5 + dd] : kCtxHV[(hh * 3 + vv) * 5 + dd];
        }
    // mask of the negative significant neighbours (same bit positions; only above = bit 1, left = 3, right = 4, below = 6 count) -> sign
    // context | prediction << 7 (:404-428); get_sign is -1 for a negative significant neighbour and 0 otherwise, so only the indices
    // 0..2 of the 5x5 tables are reachable
    for (int sw = 0; sw < 2; sw++)
        for (int m = 0; m < 256; m++) {
            int shh = 2 - ((m >> 3) & 1) - ((m >> 4) & 1), svv = 2 - ((m >> 1) & 1) - ((m >> 6) & 1);
            if (sw) { const int t = shh; shh = svv; svv = t; }
            d.lut[4 + sw][m] = (uint8_t) (kCtxSign[shh * 5 + svv] | (kSignPred[shh * 5 + svv] << 7));
        }
}
}  // namespace icer_tables
