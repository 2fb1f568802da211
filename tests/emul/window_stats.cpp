// tests/emul/window_stats.cpp -- TEST INFRASTRUCTURE (CPU): how deep is ICER's interleave window on real coefficients?  Reads a raw uint16
// sign-magnitude segment from stdin, runs the encoder lane state machine (coder_core.h) over every bit plane and prints, per plane, the
// steps per pixel, the maximum of tail - emit_pos and the fraction of steps at which at least 32 / 128 / 256 codewords are outstanding.
//   g++ -O2 -std=c++17 -o /tmp/window_stats tests/emul/window_stats.cpp && /tmp/window_stats W H SUBBAND PLANES < segment.u16
#include <stdio.h>
#include <stdlib.h>
#include "emul_coder.cpp"
int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: window_stats W H SUBBAND PLANES < segment.u16\n"); return 2; }
    const size_t w = (size_t) atoi(argv[1]), h = (size_t) atoi(argv[2]);
    const int sb = atoi(argv[3]), planes = atoi(argv[4]);
    std::vector<uint16_t> seg(w * h);
    if (fread(seg.data(), 2, w * h, stdin) != w * h) return 1;
    ready();
    for (int lsb = planes - 1; lsb >= 0; lsb--) {
        Seg s{seg.data(), w, h, w};
        uint32_t cnt[17]; uint16_t bseq[17]; uint16_t bst[17];
        std::vector<uint16_t> ring(ICER_RING);
        std::vector<uint32_t> words(w * h + 64);
        EncLane<1> ln(g_enc, cnt, bseq, bst, ring.data(), words.data(), (uint32_t) (w * h));
        const size_t cw = (w + ICER_CHUNK - 1) / ICER_CHUNK;
        uint32_t maxwin = 0;
        unsigned long long steps = 0, o32 = 0, o128 = 0, o256 = 0;
        for (size_t r = 0; r < h; r++)
            for (size_t j = 0; j < cw; j++) {
                const int ncols = (int) ((w - j * ICER_CHUNK) < ICER_CHUNK ? (w - j * ICER_CHUNK) : ICER_CHUNK);
                EncChunk k;
                const long ra = (long) r - 1, rc = (long) r + 1;
                k.set(s.S(ra, j, lsb + 1), s.B(ra, j, lsb), s.N(ra, j), s.S(r, j, lsb + 1), s.B(r, j, lsb), s.N(r, j), s.S(r, j, lsb + 2), s.S(r, j, lsb + 3),
                      s.S(rc, j, lsb + 1), s.N(rc, j), ncols);
                ln.begin_chunk();
                while (!ln.chunk_done(ncols)) {
                    ln.step(k, sb, 1u); ln.pump(); steps++;
                    const uint32_t win = ln.tail - ln.emit_pos;
                    if (win > maxwin) maxwin = win;
                    o32 += win >= 32; o128 += win >= 128; o256 += win >= 256;
                }
            }
        uint32_t bits; ln.finish(bits);
        printf("plane %2d  steps/px %.3f  max window %4u  >=32 %.3f  >=128 %.3f  >=256 %.3f  bits/px %.3f\n", lsb, (double) steps / (double) (w * h), maxwin,
               (double) o32 / (double) steps, (double) o128 / (double) steps, (double) o256 / (double) steps, (double) bits / (double) (w * h));
    }
    return 0;
}
