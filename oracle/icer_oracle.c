/*
 * icer_oracle.c -- TEST INFRASTRUCTURE ONLY (see icer_oracle.h).
 *
 * Plain-C restatement of lib_icer's uint16 codec path.  Every function cites the reference lines it
 * follows (paths relative to /root/reference/lib_icer).  Written independently: data structures and
 * control flow are ours, the arithmetic (including every quirk in SURVEY.md Appendix A) is lib_icer's.
 * Parity is pinned against the compiled reference in tests/test_oracle_vs_ref.py.
 */
#include "icer_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* constant tables                                                                                   */
/* ------------------------------------------------------------------------------------------------ */

/* {alpha_-1, alpha_0, alpha_1, beta} x16 : src/icer_config.c:18-24 */
static const int FILT[7][4] = {{0, 4, 4, 0}, {0, 4, 6, 4}, {-1, 4, 8, 6}, {0, 4, 5, 2},
                               {0, 3, 8, 6}, {0, 3, 9, 8}, {0, 4, 4, 4}};

/* src/icer_config.c:26-42 : context from (h, v, d) for LL/LH/HL (HL swaps h and v first) */
static const uint8_t CTX_HV[3][3][5] = {
    {{0, 1, 2, 2, 2}, {3, 3, 3, 3, 3}, {4, 4, 4, 4, 4}},
    {{5, 6, 7, 7, 7}, {7, 7, 7, 7, 7}, {7, 7, 7, 7, 7}},
    {{8, 8, 8, 8, 8}, {8, 8, 8, 8, 8}, {8, 8, 8, 8, 8}}};
/* src/icer_config.c:44-50 : context from (h+v, d) for HH */
static const uint8_t CTX_HH[5][5] = {
    {0, 3, 6, 8, 8}, {1, 4, 7, 8, 8}, {2, 5, 7, 8, 8}, {2, 5, 7, 8, 8}, {2, 5, 7, 8, 8}};
/* src/icer_config.c:52-58 */
static const uint8_t CTX_SIGN[5][5] = {{14, 14, 15, 16, 16}, {14, 14, 15, 16, 16}, {13, 13, 12, 13, 13},
                                       {16, 16, 15, 14, 14}, {16, 16, 15, 14, 14}};
/* src/icer_config.c:61-67 : 1 = predict negative */
static const uint8_t SIGN_PRED[5][5] = {
    {1, 1, 1, 1, 1}, {1, 1, 1, 1, 1}, {0, 0, 0, 1, 1}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
/* src/icer_config.c:69-87 */
static const uint32_t BIN_CUTOFF[17] = {35298, 37345, 40503, 43591, 47480, 50133, 53645, 55902, 57755,
                                        58894, 60437, 62267, 63613, 64557, 65134, 65392, 65536};
/* src/icer_config.c:89-107 : 0 = uncoded, -1 = custom variable-to-variable code, >0 = Golomb m */
static const int BIN_SCHEME[17] = {0, -1, -1, -1, -1, -1, -1, -1, 5, 6, 7, 11, 17, 31, 70, 200, 512};

/* custom variable-to-variable codes, src/icer_init.c:124-188.  Each row: bin index, input word (first
 * input bit in bit 0), input length, output word (first emitted bit in bit 0), output length. */
static const uint8_t V2V[][5] = {
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
/* forced-flush completion bits, src/icer_init.c:191-237.  Row: bin, prefix, prefix length, pad word, pad length */
static const uint8_t V2V_FLUSH[][5] = {
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

typedef struct { uint8_t in_bits, out_bits, out; } v2v_entry;
static v2v_entry ENC_TAB[17][32];        /* indexed by input word  : icer_custom_coding_scheme  */
static v2v_entry DEC_TAB[17][32];        /* indexed by output word : icer_custom_decode_scheme  */
static uint8_t FLUSH_TAB[17][9][6][2];   /* [bin][prefix][len] -> {pad word, pad len}            */
static struct { unsigned m, l, i; } GOLOMB[17];
static int tables_ready = 0;

static unsigned rev_bits(unsigned v, unsigned n) { /* inc/icer.h:600-608 */
    unsigned r = 0;
    for (unsigned k = 0; k < n; k++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

/* src/icer_init.c:24-255 */
void orc_init(void) {
    if (tables_ready) return;
    memset(ENC_TAB, 0, sizeof ENC_TAB);
    memset(DEC_TAB, 0, sizeof DEC_TAB);
    memset(FLUSH_TAB, 0, sizeof FLUSH_TAB);
    for (size_t k = 0; k < sizeof V2V / sizeof V2V[0]; k++) {
        const uint8_t *e = V2V[k];
        ENC_TAB[e[0]][e[1]].in_bits = e[2];
        ENC_TAB[e[0]][e[1]].out = e[3];
        ENC_TAB[e[0]][e[1]].out_bits = e[4];
        /* decode table is keyed by the emitted word; its payload is the input word, bit-reversed so that
         * popping from the top of the bin stack returns the first input bit first (icer_init.c:107-119) */
        DEC_TAB[e[0]][e[3]].in_bits = e[4];
        DEC_TAB[e[0]][e[3]].out = (uint8_t) rev_bits(e[1], e[2]);
        DEC_TAB[e[0]][e[3]].out_bits = e[2];
    }
    for (size_t k = 0; k < sizeof V2V_FLUSH / sizeof V2V_FLUSH[0]; k++) {
        const uint8_t *e = V2V_FLUSH[k];
        FLUSH_TAB[e[0]][e[1]][e[2]][0] = e[3];
        FLUSH_TAB[e[0]][e[1]][e[2]][1] = e[4];
    }
    for (int b = 0; b < 17; b++) { /* icer_init.c:239-255 */
        GOLOMB[b].m = GOLOMB[b].l = GOLOMB[b].i = 0;
        if (BIN_SCHEME[b] > 0) {
            unsigned m = (unsigned) BIN_SCHEME[b], l = 0;
            while ((1u << l) < m) l++;
            GOLOMB[b].m = m; GOLOMB[b].l = l; GOLOMB[b].i = (1u << l) - m;
        }
    }
    tables_ready = 1;
}

void orc_get_code_tables(uint8_t enc[17][32][3], uint8_t dec[17][32][3], uint8_t flush[17][9][6][2],
                         uint16_t golomb[17][3]) {
    orc_init();
    for (int b = 0; b < 17; b++) {
        for (int v = 0; v < 32; v++) {
            enc[b][v][0] = ENC_TAB[b][v].in_bits; enc[b][v][1] = ENC_TAB[b][v].out_bits; enc[b][v][2] = ENC_TAB[b][v].out;
            dec[b][v][0] = DEC_TAB[b][v].in_bits; dec[b][v][1] = DEC_TAB[b][v].out_bits; dec[b][v][2] = DEC_TAB[b][v].out;
        }
        golomb[b][0] = (uint16_t) GOLOMB[b].m; golomb[b][1] = (uint16_t) GOLOMB[b].l; golomb[b][2] = (uint16_t) GOLOMB[b].i;
    }
    memcpy(flush, FLUSH_TAB, sizeof FLUSH_TAB);
}

/* src/icer_util.c:48-56 */
int orc_bin_of(uint32_t zero, uint32_t total) {
    uint32_t lhs = zero * 65536u;
    for (int b = 16; b > 0; b--)
        if (lhs >= total * BIN_CUTOFF[b - 1]) return b;
    return 0;
}

/* src/crc32.c:157-169 : reflected CRC-32, poly 0xEDB88320, init/xorout 0xFFFFFFFF */
uint32_t orc_crc32(const uint8_t *buf, size_t len) {
    static uint32_t tab[256];
    static int have = 0;
    if (!have) {
        for (uint32_t n = 0; n < 256; n++) {
            uint32_t c = n;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            tab[n] = c;
        }
        have = 1;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t k = 0; k < len; k++) c = tab[(c ^ buf[k]) & 0xFF] ^ (c >> 8);
    return ~c;
}

static int32_t fdiv(int32_t a, int32_t b) { /* inc/icer.h:562-566 (b > 0 here) */
    int32_t q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

/* ------------------------------------------------------------------------------------------------ */
/* wavelet                                                                                           */
/* ------------------------------------------------------------------------------------------------ */

static size_t dim_low(size_t d, int s)  { size_t p = (size_t) 1 << s; return (d + p - 1) / p; }       /* icer_wavelet.c:107-109 */
static size_t dim_high(size_t d, int s) { return dim_low(d, s - 1) / 2; }                             /* icer_wavelet.c:111-113 */

#define AT(k) line[(size_t)(k) * stride]

/*
 * src/icer_wavelet.c:385-465.  Layout after the call: ceil(n/2) lows then floor(n/2) highs.
 * The reference does pair-split in place, then an in-place cycle-leader unshuffle
 * (icer_deinterleave_uint16 :765-820) whose net effect is the plain deinterleave done here through a
 * temporary, then the lifting update on the highs.  r[k] = l[k-1]-l[k] is truncated to int16 (:206-208);
 * stores wrap to int16 and raise the overflow flag (:412-415,:460-461).
 */
/* b8 != 0: the native uint8 twin (src/icer_wavelet.c:216-305): identical steps, but stores wrap to int8 and the overflow
 * flag is raised at the int8 range; the int8 samples live sign-extended in the int16 line */
#define WRAP(v) (b8 ? (int16_t) (int8_t) (v) : (int16_t) (v))
#define OVF(v) (b8 ? ((v) > 127 || (v) < -128) : ((v) > 32767 || (v) < -32768))
static int fwd_1d(uint16_t *data, size_t n, size_t stride, int filt, int b8) {
    int16_t *line = (int16_t *) data;
    size_t nl = (n + 1) / 2, nh = n / 2;
    int odd = (int) (n & 1), ovf = 0;
    int16_t *tmp = (int16_t *) malloc(n * sizeof(int16_t));
    for (size_t k = 0; k < nh; k++) {
        int32_t a = AT(2 * k), b = AT(2 * k + 1);
        int32_t lo = fdiv(a + b, 2), hi = a - b;
        if (OVF(lo) || OVF(hi)) ovf = 1;
        tmp[k] = WRAP(lo);
        tmp[nl + k] = WRAP(hi);
    }
    if (odd) tmp[nl - 1] = AT(n - 1);
    for (size_t k = 0; k < n; k++) AT(k) = tmp[k];
    free(tmp);

#define R(k)  ((int32_t)(int16_t)(AT((k) - 1) - AT(k)))   /* get_r_int16, k >= 1 */
    /* NOTE: like the reference, l[] and d[] are read back from the line itself, so d[k] for k < current n has
     * already been replaced by h[k]; only d[n] and d[n+1] (still untouched) are ever read. */
    const int *f = FILT[filt];
    for (size_t k = 0; k < nh; k++) {
        int32_t sub;
        if (k == 0) {
            sub = fdiv(R(1), 4);
        } else if (k == 1 && f[0] != 0) {
            /* filter C special case (:436-440): the "d" term reads data[low_N + 2] = slot of d[1] itself,
             * and is 0 when n is odd and low_N == 2 (n == 5) */
            size_t low_N = nl - 1;
            int32_t x = (odd && low_N == 2) ? 0 : (int32_t) AT(low_N + 2);
            sub = fdiv(2 * R(1) + 3 * R(2) - 2 * x + 4, 8);
        } else if (!odd && k == n / 2 - 1) {
            sub = fdiv(R(n / 2 - 1), 4);
        } else {
            int32_t rm1 = (k - 1 > 0) ? R(k - 1) : 1;  /* get_r(0) == 1 (:207) */
            int32_t dn = (odd && k + 1 == nl - 1) ? 0 : (int32_t) AT(nl + k + 1);
            sub = fdiv(f[0] * rm1 + f[1] * R(k) + f[2] * R(k + 1) - f[3] * dn + 8, 16);
        }
        int32_t h = (int32_t) AT(nl + k) - sub;
        if (OVF(h)) ovf = 1;
        AT(nl + k) = WRAP(h);
    }
    return ovf ? ORC_INTEGER_OVERFLOW : ORC_OK;
}
int orc_dwt_forward_1d_u16(uint16_t *data, size_t n, size_t stride, int filt) { return fwd_1d(data, n, stride, filt, 0); }

/*
 * Where icer_interleave_uint8 (src/icer_wavelet.c:570-628) puts element i of a length-n lows|highs array.  Even n: the
 * perfect in-shuffle.  Odd n: `halfleft = left / 2 - (is_odd ? 0 : 1)` (:599) is one too large and the three reversals
 * (:601-605) therefore rotate the wrong span, so the result is a different (fixed) permutation -- SURVEY A.5.  The
 * reference's uint8 decoder applies it after every inverse 1-D pass, so the oracle does too.  Derived by running the
 * function's index operations on an identity array: odd-tail rotation (:586-592), then per slice of length 3^k + 1
 * (k from icer_find_k's binary search :823-847 -- which returns the previous probe, not the match, on equality) the
 * three reversals and the cycle-leader walk (:607-624).
 */
static void rev_span(uint32_t *a, size_t s, size_t e) { while (s < e) { uint32_t t = a[s]; a[s] = a[e]; a[e] = t; s++; e--; } }
static void interleave8_dst(size_t len, uint32_t *dst) {
    uint32_t *a = (uint32_t *) malloc(len * sizeof(uint32_t));
    for (size_t i = 0; i < len; i++) a[i] = (uint32_t) i;
    size_t slice[12], p3 = 1;
    for (int k = 0; k < 12; k++) { slice[k] = p3 + 1; p3 *= 3; }
    int odd = (int) (len & 1);
    size_t n = odd ? len - 1 : len;
    if (odd) {
        uint32_t keep = a[n / 2];
        for (size_t i = n / 2; i < n; i++) a[i] = a[i + 1];
        a[len - 1] = keep;
    }
    size_t done = 0;
    while (done < n) {
        size_t want = n - done;
        int lo = 0, hi = 11, k = 0;
        while (lo < hi) {
            int mid = (lo + hi) / 2;
            if (want > slice[mid]) { lo = mid + 1; k = mid; }
            else if (want < slice[mid]) hi = mid - 1;
            else break;
        }
        size_t seg = slice[k], half = seg / 2, halfleft = want / 2 - (odd ? 0 : 1);
        rev_span(a, done + half, done + halfleft + half);
        rev_span(a, done + half, done + seg - 1);
        rev_span(a, done + seg, done + halfleft + half);
        for (size_t i = 1; i < seg; i *= 3) {
            size_t j = i;
            uint32_t carry = a[j + done];
            do {
                j = j < half ? 2 * j : (j - half) * 2 + 1;
                uint32_t t = a[j + done]; a[j + done] = carry; carry = t;
            } while (j != i);
        }
        done += seg;
    }
    for (size_t pos = 0; pos < len; pos++) dst[a[pos]] = (uint32_t) pos;
    free(a);
}
void orc_interleave8_map(size_t len, uint32_t *dst) { interleave8_dst(len, dst); }

/* src/icer_wavelet.c:467-550 : highs restored from the last one down (each uses the already restored
 * d[k+1] when beta != 0), then pairs rebuilt, then plain interleave (icer_interleave_uint16 :705-763) */
static int inv_1d(uint16_t *data, size_t n, size_t stride, int filt, int b8) {
    int16_t *line = (int16_t *) data;
    size_t nl = (n + 1) / 2, nh = n / 2;
    int odd = (int) (n & 1), ovf = 0;
    const int *f = FILT[filt];
    for (size_t it = 0; it < nh; it++) {
        size_t k = nh - 1 - it;
        int32_t add;
        if (k == 0) {
            add = fdiv(R(1), 4);
        } else if (k == 1 && f[0] != 0) {
            size_t low_N = nl - 1;
            int32_t x = (odd && low_N == 2) ? 0 : (int32_t) AT(low_N + 2);   /* still holds h[1] here (:488-492) */
            add = fdiv(2 * R(1) + 3 * R(2) - 2 * x + 4, 8);
        } else if (!odd && k == n / 2 - 1) {
            add = fdiv(R(n / 2 - 1), 4);
        } else {
            int32_t rm1 = (k - 1 > 0) ? R(k - 1) : 1;
            int32_t dn = (odd && k + 1 == nl - 1) ? 0 : (int32_t) AT(nl + k + 1);
            add = fdiv(f[0] * rm1 + f[1] * R(k) + f[2] * R(k + 1) - f[3] * dn + 8, 16);
        }
        int32_t d = (int32_t) AT(nl + k) + add;
        if (OVF(d)) ovf = 1;
        AT(nl + k) = WRAP(d);
    }
#undef R
    int16_t *tmp = (int16_t *) malloc(n * sizeof(int16_t));
    /* uint16: plain interleave.  uint8 with odd n: the reference's icer_interleave_uint8 placement (see above) */
    uint32_t *dst = NULL;
    if (b8 && odd) { dst = (uint32_t *) malloc(n * sizeof(uint32_t)); interleave8_dst(n, dst); }
    for (size_t k = 0; k < nh; k++) {
        int32_t lo = AT(k), hi = AT(nl + k);
        int32_t t = lo + fdiv(hi + 1, 2);
        if (OVF(t) || OVF(t - hi)) ovf = 1;
        tmp[dst ? dst[k] : 2 * k] = WRAP(t);
        tmp[dst ? dst[nl + k] : 2 * k + 1] = WRAP(t - hi);
    }
    if (odd) tmp[dst ? dst[nl - 1] : n - 1] = AT(nl - 1); /* low + floor(1/2) = low (:537-545) */
    for (size_t k = 0; k < n; k++) AT(k) = tmp[k];
    free(tmp);
    free(dst);
    return ovf ? ORC_INTEGER_OVERFLOW : ORC_OK;
}
int orc_dwt_inverse_1d_u16(uint16_t *data, size_t n, size_t stride, int filt) { return inv_1d(data, n, stride, filt, 0); }
#undef AT

/* src/icer_wavelet.c:57-77 + :155-171 : rows then columns of the shrinking LL region */
static int dwt_forward(uint16_t *img, size_t w, size_t h, int stages, int filt, int b8) {
    if (dim_low(w, stages) < 3 || dim_low(h, stages) < 3) return ORC_TOO_MANY_STAGES;
    int ovf = 0;
    size_t lw = w, lh = h;
    for (int s = 0; s < stages; s++) {
        for (size_t r = 0; r < lh; r++) ovf |= fwd_1d(img + r * w, lw, 1, filt, b8) != 0;
        for (size_t c = 0; c < lw; c++) ovf |= fwd_1d(img + c, lh, w, filt, b8) != 0;
        lw = (lw + 1) / 2; lh = (lh + 1) / 2;
    }
    return ovf ? ORC_INTEGER_OVERFLOW : ORC_OK;
}
int orc_dwt_forward_u16(uint16_t *img, size_t w, size_t h, int stages, int filt) { return dwt_forward(img, w, h, stages, filt, 0); }

/* src/icer_wavelet.c:81-103 + :175-191 : columns then rows, coarse to fine */
static int dwt_inverse(uint16_t *img, size_t w, size_t h, int stages, int filt, int b8) {
    if (dim_low(w, stages) < 3 || dim_low(h, stages) < 3) return ORC_TOO_MANY_STAGES;
    int ovf = 0;
    for (int s = stages - 1; s >= 0; s--) {
        size_t lw = dim_low(w, s), lh = dim_low(h, s);
        for (size_t c = 0; c < lw; c++) ovf |= inv_1d(img + c, lh, w, filt, b8) != 0;
        for (size_t r = 0; r < lh; r++) ovf |= inv_1d(img + r * w, lw, 1, filt, b8) != 0;
    }
    return ovf ? ORC_INTEGER_OVERFLOW : ORC_OK;
}
int orc_dwt_inverse_u16(uint16_t *img, size_t w, size_t h, int stages, int filt) { return dwt_inverse(img, w, h, stages, filt, 0); }

/* the native uint8 transform (src/icer_wavelet.c:8-53): samples widened to sign-extended int16, transformed with int8 wrap */
static uint16_t *widen8(const uint8_t *img, size_t count) {
    uint16_t *t = (uint16_t *) malloc(count * sizeof(uint16_t));
    for (size_t k = 0; k < count; k++) t[k] = (uint16_t) (int16_t) (int8_t) img[k];
    return t;
}
int orc_dwt_forward_u8(uint8_t *img, size_t w, size_t h, int stages, int filt) {
    uint16_t *t = widen8(img, w * h);
    int rc = dwt_forward(t, w, h, stages, filt, 1);
    if (rc != ORC_TOO_MANY_STAGES) for (size_t k = 0; k < w * h; k++) img[k] = (uint8_t) t[k];
    free(t);
    return rc;
}
int orc_dwt_inverse_u8(uint8_t *img, size_t w, size_t h, int stages, int filt) {
    uint16_t *t = widen8(img, w * h);
    int rc = dwt_inverse(t, w, h, stages, filt, 1);
    if (rc != ORC_TOO_MANY_STAGES) for (size_t k = 0; k < w * h; k++) img[k] = (uint8_t) t[k];
    free(t);
    return rc;
}

/* src/icer_wavelet.c:871-877 */
static uint16_t to_signmag(uint16_t v) {
    int16_t s = (int16_t) v;
    uint16_t mask = (uint16_t) (s >> 15);
    return (uint16_t) ((((int16_t) (s + (int16_t) mask)) ^ mask) | (v & 0x8000));
}
/* src/icer_wavelet.c:880-886 */
static uint16_t from_signmag(uint16_t v) {
    if (!(v & 0x8000)) return v;
    return (uint16_t) (0x8000 - v);  /* (0x8000 - v) mod 2^16 == -(v & 0x7fff) */
}

/* ------------------------------------------------------------------------------------------------ */
/* partition geometry                                                                                */
/* ------------------------------------------------------------------------------------------------ */

/* src/icer_partition.c:7-54 followed by the row/column walk of :299-385 (same for decode :411-491) */
int orc_partition_rects(size_t w, size_t h, int segments, orc_rect *out) {
    size_t s = (size_t) segments;
    if (s > w * h || segments > 32) return ORC_TOO_MANY_SEGMENTS;
    size_t r, c, r_t, h_t, x_t, c_t0, y_t, r_t0, x_b = 0, c_b0 = 0, y_b = 0, r_b0 = 0;
    if (h > (s - 1) * w) r = s;
    else for (r = 1; r < s && (r + 1) * r * w < h * s; r++) {}
    c = s / r;
    r_t = (c + 1) * r - s;
    h_t = ((2 * h * c * r_t + s) / 2) / s;
    if (h_t < r_t) h_t = r_t;
    x_t = w / c;
    c_t0 = (x_t + 1) * c - w;
    y_t = h_t / r_t;
    r_t0 = (y_t + 1) * r_t - h_t;
    if (r_t < r) {
        x_b = w / (c + 1);
        c_b0 = (x_b + 1) * (c + 1) - w;
        y_b = (h - h_t) / (r - r_t);
        r_b0 = (y_b + 1) * (r - r_t) - (h - h_t);
    }
    int n = 0;
    size_t y = 0;
    for (size_t row = 0; row < r_t; row++) {
        size_t sh = y_t + (row >= r_t0), x = 0;
        for (size_t col = 0; col < c; col++) {
            size_t sw = x_t + (col >= c_t0);
            out[n].x = (int) x; out[n].y = (int) y; out[n].w = (int) sw; out[n].h = (int) sh; n++;
            x += sw;
        }
        y += sh;
    }
    for (size_t row = 0; row < r - r_t; row++) {
        size_t sh = y_b + (row >= r_b0), x = 0;
        for (size_t col = 0; col < c + 1; col++) {
            size_t sw = x_b + (col >= c_b0);
            out[n].x = (int) x; out[n].y = (int) y; out[n].w = (int) sw; out[n].h = (int) sh; n++;
            x += sw;
        }
        y += sh;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------ */
/* interleaved entropy coder -- encoder side                                                         */
/* ------------------------------------------------------------------------------------------------ */

#define RING 2048
#define W_DONE 0x0400u
#define W_DATA 0x03FFu

typedef struct {
    uint16_t slot[RING];       /* [15:11] bits-or-bin  [10] done  [9:0] payload : icer.h:220-224 */
    size_t head, tail, used;
    int open[17];              /* slot index of the codeword each bin is building, -1 = none */
    int fill[17];              /* input bits gathered so far (custom bins) */
    uint8_t *out;
    size_t out_ind, out_max;
    unsigned bit_off;
} enc_state;

static void enc_reset(enc_state *e, uint8_t *out, size_t out_max) { /* icer_encoding.c:15-34 */
    e->head = e->tail = e->used = 0;
    for (int b = 0; b < 17; b++) { e->open[b] = -1; e->fill[b] = 0; }
    e->out = out; e->out_max = out_max; e->out_ind = 0; e->bit_off = 0;
    e->out[0] = 0;
}

/* icer_encoding.c:114-139 : emit finished codewords from the head, LSB first */
static int enc_drain(enc_state *e) {
    while (e->used > 0 && (e->slot[e->head] & W_DONE)) {
        unsigned v = e->slot[e->head];
        e->head = (e->head + 1) % RING;
        e->used--;
        unsigned nb = v >> 11;
        while (nb) {
            unsigned take = 8 - e->bit_off;
            if (take > nb) take = nb;
            e->out[e->out_ind] |= (uint8_t) ((v & ((1u << take) - 1)) << e->bit_off);
            v >>= take; nb -= take;
            unsigned pos = e->bit_off + take;
            e->bit_off = pos % 8;
            if (pos / 8) { e->out_ind++; e->out[e->out_ind] = 0; }
            if (e->out_ind == e->out_max) return ORC_BYTE_QUOTA_EXCEEDED;
        }
    }
    return ORC_OK;
}

static uint16_t golomb_word(int bin, unsigned k) { /* icer_encoding.c:72-77 */
    unsigned longer = k >= GOLOMB[bin].i;
    unsigned nb = GOLOMB[bin].l + longer;
    unsigned code = rev_bits(k + (longer ? GOLOMB[bin].i : 0), nb);
    return (uint16_t) ((nb << 11) | W_DONE | (code & W_DATA));
}

/* icer_encoding.c:141-189 : force-complete the oldest codeword */
static int enc_flush_head(enc_state *e) {
    uint16_t *w = &e->slot[e->head];
    if (!(*w & W_DONE)) {
        int bin = *w >> 11;
        if (bin > 7) {
            unsigned k = *w & W_DATA;
            *w = (k == GOLOMB[bin].m - 1) ? (uint16_t) ((1u << 11) | W_DONE | 1u) : golomb_word(bin, k);
            e->open[bin] = -1;
        } else if (bin != 0) {
            const uint8_t *fl = FLUSH_TAB[bin][*w & W_DATA][e->fill[bin]];
            unsigned word = (*w | (fl[0] << e->fill[bin])) & W_DATA;
            const v2v_entry *t = &ENC_TAB[bin][word];
            *w = (uint16_t) ((t->out_bits << 11) | W_DONE | t->out);
            e->open[bin] = -1;
            e->fill[bin] = 0;
        }
    }
    return enc_drain(e);
}

/* icer_encoding.c:37-112 */
static int enc_put(enc_state *e, unsigned bit, uint32_t zero, uint32_t total) {
    if (zero < (total >> 1)) { zero = total - zero; bit ^= 1; }
    int bin = orc_bin_of(zero, total);
    if (e->open[bin] < 0) {
        if (e->used >= RING) {                      /* window full: :59-64 */
            if (enc_flush_head(e) == ORC_BYTE_QUOTA_EXCEEDED) return ORC_BYTE_QUOTA_EXCEEDED;
        }
        e->open[bin] = (int) e->tail;
        e->tail = (e->tail + 1) % RING;
        e->used++;
        e->slot[e->open[bin]] = (uint16_t) (bin << 11);
    }
    uint16_t *w = &e->slot[e->open[bin]];
    if (bin > 7) {
        if (bit) {
            *w = golomb_word(bin, *w & W_DATA);
            e->open[bin] = -1;
        } else {
            (*w)++;
            if ((*w & W_DATA) >= GOLOMB[bin].m) { *w = (uint16_t) ((1u << 11) | W_DONE | 1u); e->open[bin] = -1; }
        }
    } else if (bin != 0) {
        *w |= (uint16_t) (bit << e->fill[bin]);
        e->fill[bin]++;
        const v2v_entry *t = &ENC_TAB[bin][*w & W_DATA];
        if (t->in_bits == e->fill[bin]) {
            *w = (uint16_t) ((t->out_bits << 11) | W_DONE | t->out);
            e->open[bin] = -1;
            e->fill[bin] = 0;
        }
    } else {
        *w = (uint16_t) ((1u << 11) | W_DONE | (bit & 1));
        e->open[bin] = -1;
    }
    return enc_drain(e);
}

/* ------------------------------------------------------------------------------------------------ */
/* context model                                                                                     */
/* ------------------------------------------------------------------------------------------------ */

typedef struct { uint32_t zero[17], total[17]; } ctx_model;

static void model_reset(ctx_model *m) { /* icer_context_modeller.c:607-613, icer.h:147-148 */
    for (int k = 0; k < 17; k++) { m->zero[k] = 2; m->total[k] = 4; }
}
static void model_update(ctx_model *m, int c, unsigned bit) { /* icer_context_modeller.c:396-402 */
    m->total[c]++;
    m->zero[c] += !bit;
    if (m->total[c] >= 500) {
        m->total[c] >>= 1;
        if (m->zero[c] > m->total[c]) m->zero[c] >>= 1;   /* else: value of the ceil_div is discarded */
    }
}

static int px_category(uint16_t v, int lsb) { /* icer_context_modeller.c:631-634 */
    unsigned mag = (v & 0x7FFFu) | 1u;
    int msb = 0;
    while (mag >> (msb + 1)) msb++;
    int cat = msb < lsb ? 0 : msb - lsb;
    return cat > 3 ? 3 : cat;
}
static int px_sig(uint16_t v, int plane) { return ((v & 0x7FFFu) >> plane) != 0; }              /* :636-638 */
static int px_sgn(uint16_t v, int plane) { return (v & 0x8000u) ? -px_sig(v, plane) : 0; }      /* :640-642 */

/* neighbourhood -> magnitude context / sign context; shared by encoder and decoder
 * (icer_context_modeller.c:355-391 and :404-428) */
static int mag_context(const uint16_t *p, size_t stride, size_t row, size_t col, size_t w, size_t h, int cat,
                       int lsb, int subband) {
    if (cat == 2) return 11;
    int prev = lsb + 1, hh = 0, vv = 0, dd = 0;
    int L = col > 0, Rr = col + 1 < w, U = row > 0, D = row + 1 < h;
    if (L) hh += px_sig(p[-1], lsb);
    if (Rr) hh += px_sig(p[1], prev);
    if (U) vv += px_sig(p[-(ptrdiff_t) stride], lsb);
    if (D) vv += px_sig(p[stride], prev);
    if (L && U) dd += px_sig(p[-(ptrdiff_t) stride - 1], lsb);
    if (L && D) dd += px_sig(p[stride - 1], prev);
    if (Rr && U) dd += px_sig(p[-(ptrdiff_t) stride + 1], lsb);
    if (Rr && D) dd += px_sig(p[stride + 1], prev);
    if (cat == 1) return (hh + vv == 0) ? 9 : 10;
    if (subband == 1) { int t = hh; hh = vv; vv = t; }       /* HL */
    return subband == 3 ? CTX_HH[hh + vv][dd] : CTX_HV[hh][vv][dd];
}
static void sign_context(const uint16_t *p, size_t stride, size_t row, size_t col, size_t w, size_t h, int lsb,
                         int subband, int *ctx, int *pred) {
    int prev = lsb + 1, sh = 2, sv = 2;
    if (col > 0) sh += px_sgn(p[-1], lsb);
    if (col + 1 < w) sh += px_sgn(p[1], prev);
    if (row > 0) sv += px_sgn(p[-(ptrdiff_t) stride], lsb);
    if (row + 1 < h) sv += px_sgn(p[stride], prev);
    if (subband == 1) { int t = sh; sh = sv; sv = t; }
    *ctx = CTX_SIGN[sh][sv];
    *pred = SIGN_PRED[sh][sv];
}

/* icer_context_modeller.c:312-457 */
long orc_encode_segment_plane(const uint16_t *seg, size_t w, size_t h, size_t rowstride, int subband, int lsb,
                              uint8_t *out, size_t max_bytes) {
    orc_init();
    if (lsb + 1 >= 16) return ORC_BITPLANE_OUT_OF_RANGE;
    enc_state *e = (enc_state *) malloc(sizeof *e);
    ctx_model m;
    model_reset(&m);
    enc_reset(e, out, max_bytes);
    int rc = ORC_OK;
    for (size_t row = 0; row < h && rc == ORC_OK; row++) {
        const uint16_t *p = seg + row * rowstride;
        for (size_t col = 0; col < w; col++, p++) {
            int cat = px_category(*p, lsb);
            unsigned bit = (*p >> lsb) & 1;
            if (cat == 3) {
                if ((rc = enc_put(e, bit, 1, 2)) != ORC_OK) break;
                continue;
            }
            int c = mag_context(p, rowstride, row, col, w, h, cat, lsb, subband);
            if ((rc = enc_put(e, bit, m.zero[c], m.total[c])) != ORC_OK) break;
            model_update(&m, c, bit);
            if (cat == 0 && bit) {
                int sc, pred;
                sign_context(p, rowstride, row, col, w, h, lsb, subband, &sc, &pred);
                unsigned agree = (unsigned) (pred ^ ((*p & 0x8000u) != 0)) & 1;
                if ((rc = enc_put(e, agree, m.zero[sc], m.total[sc])) != ORC_OK) break;
                model_update(&m, sc, agree);
            }
        }
    }
    while (rc == ORC_OK && e->used > 0) rc = enc_flush_head(e);   /* :452-455 */
    long bits = (long) (e->out_ind * 8 + e->bit_off);
    free(e);
    return rc == ORC_OK ? bits : rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* interleaved entropy coder -- decoder side (icer_decoding.c)                                       */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    const uint8_t *in;
    size_t byte, avail;            /* read cursor; avail = readable bytes (reads past it return 0) */
    unsigned bit_off;
    uint32_t total_bits;
    size_t words;                  /* codewords consumed so far */
    /* per-bin stack of pending decoded bits.  Only the low 32 positions can ever hold a 1 (custom words
     * are <= 5 bits; a Golomb run is a single 1 at position 0 under a pile of zeros). */
    int depth[17];
    uint32_t low[17];
    size_t stamp[17];
} dec_state;

static unsigned dec_peek(const dec_state *d, unsigned nth /* 1-based */) { /* icer_decoding.c:46-57 */
    size_t pos = d->bit_off + (nth - 1);
    size_t b = d->byte + pos / 8;
    return b < d->avail ? (d->in[b] >> (pos % 8)) & 1u : 0u;
}
/* icer_decoding.c:59-105 : `consume`=0 is icer_get_bits_from_codeword, 1 is icer_pop_bits_from_codeword.
 * decoded_bits_total stays 0 in the reference, so the only live range check is chunk > total_bits. */
static int dec_bits(dec_state *d, unsigned n, int consume) {
    int val = 0;
    unsigned got = 0, off = d->bit_off;
    size_t b = d->byte;
    while (n) {
        unsigned take = 8 - off;
        if (take > n) take = n;
        if (take > d->total_bits) return ORC_DECODER_OUT_OF_DATA;
        unsigned byte = b < d->avail ? d->in[b] : 0;
        val |= (int) (((byte >> off) & ((1u << take) - 1)) << got);
        n -= take; got += take; off += take;
        if (off >= 8) { off -= 8; b++; }
    }
    if (consume) { d->bit_off = off; d->byte = b; }
    return val;
}
static void dec_push(dec_state *d, int bin, unsigned bits, unsigned n) { /* icer_decoding.c:28-44 */
    if (d->depth[bin] < 32) d->low[bin] |= (n >= 32 ? bits : (bits & ((1u << n) - 1))) << d->depth[bin];
    d->depth[bin] += (int) n;
}

/* icer_decoding.c:108-194 */
static int dec_get(dec_state *d, unsigned *bit, uint32_t zero, uint32_t total) {
    int inv = 0;
    if (zero < (total >> 1)) { zero = total - zero; inv = 1; }
    int bin = orc_bin_of(zero, total);
    if (d->depth[bin] <= 0 || d->words - d->stamp[bin] >= RING) {
        d->depth[bin] = 0;
        d->low[bin] = 0;
        if (bin > 7) {
            if (dec_peek(d, 1)) {
                dec_bits(d, 1, 1);
                dec_push(d, bin, 0, GOLOMB[bin].m);
            } else {
                unsigned l = GOLOMB[bin].l;
                unsigned k = rev_bits((unsigned) dec_bits(d, l, 0) & 0xFFFFu, l) & 0xFFFFu;
                if (k < GOLOMB[bin].i) {
                    dec_bits(d, l, 1);
                } else {
                    k = rev_bits((unsigned) dec_bits(d, l + 1, 1) & 0xFFFFu, l + 1) & 0xFFFFu;
                    k = (k - GOLOMB[bin].i) & 0xFFFFu;
                }
                dec_push(d, bin, 1, 1);
                dec_push(d, bin, 0, k);
            }
        } else if (bin != 0) {
            unsigned word = 0, n = 0;
            int found = 0;
            do {
                if (n + 1 >= d->total_bits) return ORC_DECODER_OUT_OF_DATA;
                word |= dec_peek(d, n + 1) << n;
                n++;
                if (word >= 32) return ORC_DECODED_INVALID_DATA;
                if (DEC_TAB[bin][word].in_bits == n) {
                    dec_push(d, bin, DEC_TAB[bin][word].out, DEC_TAB[bin][word].out_bits);
                    if ((int) word != dec_bits(d, n, 1)) return ORC_DECODED_INVALID_DATA;
                    found = 1;
                }
            } while (!found && n < 10);
        } else {
            int v = dec_bits(d, 1, 1);
            if (v == ORC_DECODER_OUT_OF_DATA) return v;
            dec_push(d, bin, v != 0, 1);
        }
        d->words++;
        d->stamp[bin] = d->words;
    }
    int top = d->depth[bin] - 1;
    unsigned b = 0;
    if (top >= 0 && top < 32) { b = (d->low[bin] >> top) & 1u; d->low[bin] &= ~(1u << top); }
    d->depth[bin]--;
    *bit = inv ? !b : b;    /* (*bit) = inv == !b  (:192) */
    return ORC_OK;
}

/* icer_context_modeller.c:461-602 */
static int decode_plane_bounded(uint16_t *seg, size_t w, size_t h, size_t rowstride, int subband, int lsb,
                                const uint8_t *in, uint32_t nbits, size_t avail) {
    if (lsb + 1 >= 16) return ORC_BITPLANE_OUT_OF_RANGE;
    dec_state d;
    memset(&d, 0, sizeof d);
    d.in = in; d.avail = avail; d.total_bits = nbits;
    ctx_model m;
    model_reset(&m);
    for (size_t row = 0; row < h; row++) {
        uint16_t *p = seg + row * rowstride;
        for (size_t col = 0; col < w; col++, p++) {
            int cat = px_category(*p, lsb), rc;
            unsigned bit;
            if (cat == 3) {
                if ((rc = dec_get(&d, &bit, 1, 2)) != ORC_OK) return rc;
                *p |= (uint16_t) (bit << lsb);
                continue;
            }
            int c = mag_context(p, rowstride, row, col, w, h, cat, lsb, subband);
            if ((rc = dec_get(&d, &bit, m.zero[c], m.total[c])) != ORC_OK) return rc;
            *p |= (uint16_t) (bit << lsb);
            model_update(&m, c, bit);
            if (cat == 0 && bit) {
                int sc, pred;
                unsigned agree;
                sign_context(p, rowstride, row, col, w, h, lsb, subband, &sc, &pred);
                if ((rc = dec_get(&d, &agree, m.zero[sc], m.total[sc])) != ORC_OK) return rc;
                *p |= (uint16_t) (((agree ^ (unsigned) pred) & 1u) << 15);
                model_update(&m, sc, agree);
            }
        }
    }
    return ORC_OK;
}

int orc_decode_segment_plane(uint16_t *seg, size_t w, size_t h, size_t rowstride, int subband, int lsb,
                             const uint8_t *in, uint32_t nbits) {
    orc_init();
    return decode_plane_bounded(seg, w, h, rowstride, subband, lsb, in, nbits, (nbits + 7) / 8);
}

/* ------------------------------------------------------------------------------------------------ */
/* image level                                                                                       */
/* ------------------------------------------------------------------------------------------------ */

#define HDR 28 /* sizeof(icer_image_segment_typedef), icer.h:293-305 */

static void put16(uint8_t *p, unsigned v) { p[0] = (uint8_t) v; p[1] = (uint8_t) (v >> 8); }
static void put32(uint8_t *p, uint32_t v) { put16(p, v & 0xFFFF); put16(p + 2, v >> 16); }
static unsigned get16(const uint8_t *p) { return p[0] | (p[1] << 8); }
static uint32_t get32(const uint8_t *p) { return get16(p) | ((uint32_t) get16(p + 2) << 16); }

typedef struct {
    int level, subband, lsb, chan;
    uint64_t prio;
    int order;   /* insertion index: glibc qsort is a stable merge sort for these sizes */
} packet;

static int packet_cmp(const void *a, const void *b) { /* icer_compress.c:8-15 + stability */
    const packet *x = a, *y = b;
    if (x->prio != y->prio) return x->prio > y->prio ? -1 : 1;
    if (x->subband != y->subband) return x->subband > y->subband ? 1 : -1;
    return x->order - y->order;
}

static void subband_geom(size_t w, size_t h, int level, int sb, size_t *x0, size_t *y0, size_t *sw, size_t *sh) {
    /* icer_compress.c:380-399 */
    *x0 = (sb == 1 || sb == 3) ? dim_low(w, level) : 0;
    *y0 = (sb == 2 || sb == 3) ? dim_low(h, level) : 0;
    *sw = (sb == 1 || sb == 3) ? dim_high(w, level) : dim_low(w, level);
    *sh = (sb == 2 || sb == 3) ? dim_high(h, level) : dim_low(h, level);
}

/* icer_compress.c:279-426 (gray) and icer_color.c:343-530 (3 planar channels) */
/* native8: the uint8 twins icer_compress.c:18-159 / icer_color.c:19-206 -- int8 samples sign-extended in the containers,
 * LL mean over the uint8 reinterpretation, 7-bit magnitudes (the sign sits in bit 15 here, bit 7 in the reference's bytes),
 * and for colour the rearrangement loops run ASCENDING (icer_color.c:184-203) */
static int compress_core(uint16_t *const img[3], int chan, size_t w, size_t h, int stages, int filt, int segments,
                         const orc_limits *lim, uint8_t *out, size_t quota, size_t *size_used, int native8) {
    orc_init();
    const int planes = lim->planes;
    *size_used = 0;
    if (stages > lim->max_stages) return ORC_TOO_MANY_STAGES;   /* reference: out-of-bounds table index */
    for (int ch = 0; ch < chan; ch++) {
        int rc = dwt_forward(img[ch], w, h, stages, filt, native8);
        if (rc != ORC_OK) return rc;
    }
    size_t llw = dim_low(w, stages), llh = dim_low(h, stages);
    uint16_t mean[3] = {0, 0, 0};
    for (int ch = 0; ch < chan; ch++) {                    /* icer_compress.c:289-302 (:28-40 uint8) : unsigned sum */
        uint64_t sum = 0;
        for (size_t r = 0; r < llh; r++)
            for (size_t c = 0; c < llw; c++) sum += native8 ? (img[ch][r * w + c] & 0xFF) : img[ch][r * w + c];
        mean[ch] = native8 ? (uint16_t) (uint8_t) (sum / (llw * llh)) : (uint16_t) (sum / (llw * llh));
        if (mean[ch] > (native8 ? 127 : 32767)) return ORC_INTEGER_OVERFLOW;
    }
    for (int ch = 0; ch < chan; ch++) {
        for (size_t r = 0; r < llh; r++)
            for (size_t c = 0; c < llw; c++) {
                int32_t d = (int32_t) (int16_t) img[ch][r * w + c] - (native8 ? (int32_t) (int8_t) mean[ch] : (int32_t) (int16_t) mean[ch]);
                img[ch][r * w + c] = (uint16_t) (native8 ? (int16_t) (int8_t) d : (int16_t) d);
            }
        for (size_t k = 0; k < w * h; k++) {
            if (!native8) { img[ch][k] = to_signmag(img[ch][k]); continue; }
            int32_t v = (int16_t) img[ch][k];           /* icer_wavelet.c:851-857: -128 becomes "-0" */
            img[ch][k] = v < 0 ? (uint16_t) (0x8000u | ((uint32_t) (-v) & 0x7Fu)) : (uint16_t) v;
        }
    }

    /* packet list: icer_compress.c:315-363 / icer_color.c:398-456 */
    packet *pk = (packet *) malloc(sizeof(packet) * (size_t) lim->max_packets);
    int np = 0;
#define ADD(LV, SB, LSB, CH, PR) do { pk[np].level = (LV); pk[np].subband = (SB); pk[np].lsb = (LSB); pk[np].chan = (CH); \
        pk[np].prio = (PR); pk[np].order = np; np++; if (np >= lim->max_packets) { free(pk); return ORC_PACKET_COUNT_EXCEEDED; } } while (0)
    if (chan == 1) {
        for (int st = 1; st <= stages; st++) {
            uint64_t pr = (uint64_t) 1 << st;            /* icer_pow_uint returns unsigned */
            for (int lsb = 0; lsb < planes; lsb++) {
                ADD(st, 1, lsb, 0, pr << lsb);
                ADD(st, 2, lsb, 0, pr << lsb);
                ADD(st, 3, lsb, 0, ((pr / 2) << lsb) + 1);
            }
        }
        uint64_t pr = (uint64_t) 1 << stages;
        for (int lsb = 0; lsb < planes; lsb++) ADD(stages, 0, lsb, 0, (2 * pr) << lsb);
    } else {
        /* colour: `priority` is a uint32_t that is doubled at every Y channel, cumulatively over the lsb
         * loop (icer_color.c:401-404, :439-448) */
        for (int st = 1; st <= stages; st++) {
            uint32_t pr = 1u << st;
            for (int lsb = 0; lsb < planes; lsb++)
                for (int ch = 0; ch < 3; ch++) {
                    if (ch == 0) pr *= 2;
                    ADD(st, 1, lsb, ch, (uint32_t) (pr << lsb));
                    ADD(st, 2, lsb, ch, (uint32_t) (pr << lsb));
                    ADD(st, 3, lsb, ch, (uint32_t) (((pr / 2) << lsb) + 1));
                }
        }
        uint32_t pr = 1u << stages;
        for (int lsb = 0; lsb < planes; lsb++)
            for (int ch = 0; ch < 3; ch++) {
                if (ch == 0) pr *= 2;
                ADD(stages, 0, lsb, ch, (uint32_t) ((2 * pr) << lsb));
            }
    }
#undef ADD
    qsort(pk, (size_t) np, sizeof(packet), packet_cmp);

    /* where each finished segment sits in the priority-ordered stream, for the rearrangement pass */
    size_t tab_n = (size_t) 3 * (size_t) (lim->max_stages + 1) * 4 * 15 * 33;
    long *where = (long *) malloc(tab_n * sizeof(long));
    for (size_t k = 0; k < tab_n; k++) where[k] = -1;
#define WHERE(CH, LV, SB, LSB, SEG) where[((((size_t)(CH) * (lim->max_stages + 1) + (LV)) * 4 + (SB)) * 15 + (LSB)) * 33 + (SEG)]

    size_t used = 0;
    int rc = ORC_OK;
    orc_rect rects[40];
    for (int it = 0; it < np && rc == ORC_OK; it++) {
        size_t x0, y0, sw, sh;
        subband_geom(w, h, pk[it].level, pk[it].subband, &x0, &y0, &sw, &sh);
        int nseg = orc_partition_rects(sw, sh, segments, rects);
        if (nseg < 0) { rc = nseg; break; }           /* reference ignores this (UB); we report it */
        const uint16_t *base = img[pk[it].chan] + y0 * w + x0;
        for (int sg = 0; sg < nseg; sg++) {          /* icer_partition.c:299-385 */
            if (quota - used < HDR) { rc = ORC_BYTE_QUOTA_EXCEEDED; break; }   /* icer_encoding.c:211-214 */
            uint8_t *hd = out + used;
            size_t room = quota - used - HDR;
            long bits = orc_encode_segment_plane(base + (size_t) rects[sg].y * w + rects[sg].x, (size_t) rects[sg].w,
                                                 (size_t) rects[sg].h, w, pk[it].subband, pk[it].lsb, hd + HDR, room);
            if (bits < 0) { rc = (int) bits; break; }  /* header rolled back: icer_partition.c:323-326 */
            size_t nbytes = (size_t) (bits + 7) / 8;
            put16(hd + 0, 0x605B);
            put16(hd + 2, mean[pk[it].chan] & 0xFF);  /* ll_mean_val passes through a uint8_t: icer.h:270 */
            hd[4] = (uint8_t) pk[it].level;
            hd[5] = (uint8_t) pk[it].subband;
            hd[6] = (uint8_t) sg;
            hd[7] = (uint8_t) (pk[it].lsb | (pk[it].chan << 4));
            put32(hd + 8, (uint32_t) w);
            put32(hd + 12, (uint32_t) h);
            put32(hd + 16, (uint32_t) bits);
            put32(hd + 20, orc_crc32(hd + HDR, nbytes));
            put32(hd + 24, orc_crc32(hd, 24));
            WHERE(pk[it].chan, pk[it].level, pk[it].subband, pk[it].lsb, sg) = (long) used;
            used += HDR + nbytes;
        }
    }
    /* rearrangement: icer_compress.c:409-423 / icer_color.c:508-527; the uint8 colour encoder loops ascending (:184-203) */
    size_t off = 0;
    const int asc = native8 && chan == 3;
    for (int sg = 0; sg <= 32; sg++)
        for (int i1 = 0; i1 <= 3; i1++)
            for (int i2 = 0; i2 <= lim->max_stages; i2++)
                for (int i3 = 0; i3 < planes; i3++)
                    for (int ch = 0; ch < chan; ch++) {
                        const int sb = asc ? i1 : 3 - i1, lv = asc ? i2 : lim->max_stages - i2, lsb = asc ? i3 : planes - 1 - i3;
                        long at = WHERE(ch, lv, sb, lsb, sg);
                        if (at < 0) continue;
                        size_t len = HDR + (get32(out + at + 16) + 7) / 8;
                        memcpy(out + quota + off, out + at, len);
                        off += len;
                    }
#undef WHERE
    free(where);
    free(pk);
    *size_used = used;
    return rc;
}

int orc_compress_u16(uint16_t *const img[3], int chan, size_t w, size_t h, int stages, int filt, int segments,
                     const orc_limits *lim, uint8_t *out, size_t quota, size_t *size_used) {
    return compress_core(img, chan, w, h, stages, filt, segments, lim, out, quota, size_used, 0);
}

int orc_compress_u8(uint8_t *const img[3], int chan, size_t w, size_t h, int stages, int filt, int segments,
                    const orc_limits *lim, uint8_t *out, size_t quota, size_t *size_used) {
    uint16_t *t[3] = {NULL, NULL, NULL};
    for (int ch = 0; ch < chan; ch++) t[ch] = widen8(img[ch], w * h);
    int rc = compress_core(t, chan, w, h, stages, filt, segments, lim, out, quota, size_used, 1);
    for (int ch = 0; ch < chan; ch++) {
        /* what the reference leaves in the caller's buffers: transform output (on ICER_INTEGER_OVERFLOW after the
         * transform) or 8-bit sign-magnitude coefficients; the caller compares them only on success */
        for (size_t k = 0; k < w * h; k++) img[ch][k] = (uint8_t) ((t[ch][k] & 0x7F) | ((t[ch][k] >> 8) & 0x80));
        free(t[ch]);
    }
    return rc;
}

/* icer_compress.c:569-588 : returns 1 and the segment length when a CRC-valid segment starts at p */
static int segment_at(const uint8_t *p, size_t left, size_t *len) {
    if (left < HDR) {
        /* the reference reads the header fields without a length check (UB past the buffer); any such
         * tail can never carry a valid segment, so we simply reject it */
        return 0;
    }
    if (get16(p) != 0x605B) return 0;
    if (get32(p + 24) != orc_crc32(p, 24)) return 0;
    size_t nbytes = ((size_t) get32(p + 16) + 7) / 8;
    if (nbytes > left - HDR) return 0;
    if (get32(p + 20) != orc_crc32(p + HDR, nbytes)) return 0;
    *len = HDR + nbytes;
    return 1;
}

/* icer_compress.c:430-536 / icer_color.c:534-663 */
static int decompress_core(uint16_t *const img[3], int chan, size_t *w_out, size_t *h_out, size_t bufsize_px,
                           const uint8_t *stream, size_t len, int stages, int filt, int segments,
                           const orc_limits *lim, int native8) {
    orc_init();
    const int planes = lim->planes;
    size_t tab_n = (size_t) 3 * (size_t) (lim->max_stages + 1) * 4 * 33 * 15;
    long *seen = (long *) malloc(tab_n * sizeof(long));
    for (size_t k = 0; k < tab_n; k++) seen[k] = -1;
#define SEEN(CH, LV, SB, SEG, LSB) seen[((((size_t)(CH) * (lim->max_stages + 1) + (LV)) * 4 + (SB)) * 33 + (SEG)) * 15 + (LSB)]
    uint16_t mean[3] = {0, 0, 0};
    size_t w = 0, h = 0, off = 0;
    int any = 0;
    while (off < len) {
        size_t slen;
        if (segment_at(stream + off, len - off, &slen)) {
            const uint8_t *p = stream + off;
            int ch = chan == 1 ? 0 : (p[7] >> 4);
            int lv = p[4], sb = p[5], sg = p[6], lsb = p[7] & 0x0F;
            if (ch < 3 && lv <= lim->max_stages && sb < 4 && sg <= 32 && lsb < 15) {
                SEEN(ch, lv, sb, sg, lsb) = (long) off;
            }
            w = get32(p + 8); h = get32(p + 12);
            if (ch < 3) mean[ch] = (uint16_t) get16(p + 2);
            any = 1;
            off += slen;
        } else {
            off++;
        }
    }
    if (!any) { free(seen); return ORC_DECODER_OUT_OF_DATA; }   /* reference: *image_w/h left uninitialised */
    *w_out = w; *h_out = h;
    if (bufsize_px < w * h) { free(seen); return ORC_BYTE_QUOTA_EXCEEDED; }
    for (int ch = 0; ch < chan; ch++) memset(img[ch], 0, w * h * sizeof(uint16_t));

    orc_rect rects[40];
    int rc = ORC_OK;
    for (int st = 1; st <= stages && rc == ORC_OK; st++) {
        for (int ch = 0; ch < chan && rc == ORC_OK; ch++) {
            for (int sb = 0; sb < 4 && rc == ORC_OK; sb++) {
                if (sb == 0 && st != stages) continue;    /* LL only at the last stage, and first (:473-483) */
                size_t x0, y0, sw, sh;
                subband_geom(w, h, st, sb, &x0, &y0, &sw, &sh);
                int nseg = orc_partition_rects(sw, sh, segments, rects);
                if (nseg < 0) { rc = nseg; break; }
                for (int sg = 0; sg < nseg; sg++) {       /* icer_partition.c:427-443 */
                    uint16_t *base = img[ch] + (y0 + (size_t) rects[sg].y) * w + x0 + (size_t) rects[sg].x;
                    for (int lsb = planes - 1; lsb >= 0; lsb--) {
                        long at = st <= lim->max_stages ? SEEN(ch, st, sb, sg, lsb) : -1;
                        if (at < 0) break;
                        const uint8_t *p = stream + at;
                        /* subband type is taken from the top-plane header (:432-434); it always equals sb here */
                        int r2 = decode_plane_bounded(base, (size_t) rects[sg].w, (size_t) rects[sg].h, w, sb, lsb,
                                                      p + HDR, get32(p + 16), len - (size_t) at - HDR);
                        if (r2 != ORC_OK) break;
                    }
                }
            }
        }
    }
#undef SEEN
    free(seen);
    if (rc != ORC_OK) return rc;
    size_t llw = dim_low(w, stages), llh = dim_low(h, stages);
    for (int ch = 0; ch < chan; ch++) {
        for (size_t k = 0; k < w * h; k++) img[ch][k] = from_signmag(img[ch][k]);
        for (size_t r = 0; r < llh; r++)
            for (size_t c = 0; c < llw; c++) {
                /* uint8: `(int8_t)(*px + (int8_t) ll_mean)` with ll_mean the header's 16-bit field (icer_compress.c:258-265) */
                int32_t d = (int32_t) (int16_t) img[ch][r * w + c] + (native8 ? (int32_t) (int8_t) mean[ch] : (int32_t) (int16_t) mean[ch]);
                img[ch][r * w + c] = (uint16_t) (native8 ? (int16_t) (int8_t) d : (int16_t) d);
            }
        dwt_inverse(img[ch], w, h, stages, filt, native8);   /* result ignored: icer_compress.c:533 */
        for (size_t k = 0; k < w * h; k++)
            if ((int16_t) img[ch][k] < 0) img[ch][k] = 0;   /* icer_util.c:82-90 */
    }
    return ORC_OK;
}

int orc_decompress_u16(uint16_t *const img[3], int chan, size_t *w_out, size_t *h_out, size_t bufsize_px,
                       const uint8_t *stream, size_t len, int stages, int filt, int segments,
                       const orc_limits *lim) {
    return decompress_core(img, chan, w_out, h_out, bufsize_px, stream, len, stages, filt, segments, lim, 0);
}

int orc_decompress_u8(uint8_t *const img[3], int chan, size_t *w_out, size_t *h_out, size_t bufsize_px,
                      const uint8_t *stream, size_t len, int stages, int filt, int segments,
                      const orc_limits *lim) {
    uint16_t *t[3] = {NULL, NULL, NULL};
    for (int ch = 0; ch < chan; ch++) t[ch] = (uint16_t *) calloc(bufsize_px ? bufsize_px : 1, sizeof(uint16_t));
    int rc = decompress_core(t, chan, w_out, h_out, bufsize_px, stream, len, stages, filt, segments, lim, 1);
    for (int ch = 0; ch < chan; ch++) {
        if (rc == ORC_OK) for (size_t k = 0; k < *w_out * *h_out; k++) img[ch][k] = (uint8_t) t[ch][k];
        free(t[ch]);
    }
    return rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* deterministic synthetic frames; shared by tests, bench and the CPU baseline */
/* ------------------------------------------------------------------------------------------------ */
static uint32_t xs32(uint32_t *s) { uint32_t x = *s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; return *s = x; }
static int tri(long t, int p) { long m = t % (2L * p); if (m < 0) m += 2L * p; m -= p; return (int) (m < 0 ? -m : m); }

/* kind: 0 natural-like 8-bit, 1 noise 8-bit, 2 natural-like 12-bit, 3 noise 12-bit */
void orc_synth_frame(uint16_t *out, size_t w, size_t h, uint32_t frame, uint32_t seed, int kind) {
    uint32_t st = seed ^ (frame * 0x9E3779B9u);
    if (st == 0) st = 0x1234567u;
    for (size_t y = 0; y < h; y++)
        for (size_t x = 0; x < w; x++) {
            uint32_t d = xs32(&st);
            unsigned v8;
            if (kind & 1) v8 = d & 0xFF;
            else v8 = 16u + (unsigned) ((tri((long) x + 2 * (long) y + 13 * (long) frame, 256) * 160) >> 8)
                          + (unsigned) tri(3 * (long) x - (long) y + 4096, 32) + d % 13u;
            if (kind & 2) { uint32_t d2 = xs32(&st); out[y * w + x] = (uint16_t) (v8 * 16u + (d2 & 15u)); }
            else out[y * w + x] = (uint16_t) v8;
        }
}
