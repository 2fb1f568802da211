/*
 * icer_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the ICER codec as implemented by lib_icer
 * (TheRealOrange/icer_compression @ 0e2961a).  It exists to CHECK the CUDA product path; it is
 * never linked into, imported by, or called from the product (libicer_b200.so).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py compares every entry point below byte-for-byte
 * with the unmodified reference compiled into oracle/_ref/libicer_ref{,15}.so (the reference ships no
 * golden vectors of its own), and tests/golden/ holds vectors generated from that
 * reference by tests/golden/make_golden.py.  The native uint8 entry points (orc_*_u8) are pinned by
 * tests/test_oracle_u8_vs_ref.py and tests/golden/golden_u8.json (tests/golden/make_golden_u8.py).
 *
 * Differences from the reference that are deliberate:
 *   - bit planes / max stages / max packets are run-time parameters (reference: compile-time macros
 *     ICER_BITPLANES_TO_COMPRESS_16, ICER_MAX_DECOMP_STAGES, ICER_MAX_PACKETS_16; icer.h:29-46)
 *   - no global state; re-entrant
 *   - where the reference has undefined behaviour (compress with a subband smaller than `segments`,
 *     icer_compress.c:401 ignores the partition error) the oracle returns ORC_TOO_MANY_SEGMENTS.
 */
#ifndef ICER_ORACLE_H
#define ICER_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same numeric values as enum icer_status, lib_icer/inc/icer.h:92-105 */
enum {
    ORC_OK = 0, ORC_INTEGER_OVERFLOW = -1, ORC_OUTPUT_BUF_TOO_SMALL = -2, ORC_TOO_MANY_SEGMENTS = -3,
    ORC_TOO_MANY_STAGES = -4, ORC_BYTE_QUOTA_EXCEEDED = -5, ORC_BITPLANE_OUT_OF_RANGE = -6,
    ORC_DECODER_OUT_OF_DATA = -7, ORC_DECODED_INVALID_DATA = -8, ORC_PACKET_COUNT_EXCEEDED = -9,
    ORC_FATAL_ERROR = -10, ORC_INVALID_INPUT = -11
};

typedef struct {
    int planes;       /* bit planes coded, 1..15 (reference default 9) */
    int max_stages;   /* reference default 6 */
    int max_packets;  /* reference default 800 */
} orc_limits;

typedef struct { int x, y, w, h; } orc_rect;

void orc_init(void);

/* wavelet (icer_wavelet.c:57-103) */
int orc_dwt_forward_u16(uint16_t *img, size_t w, size_t h, int stages, int filt);
int orc_dwt_inverse_u16(uint16_t *img, size_t w, size_t h, int stages, int filt);
int orc_dwt_forward_1d_u16(uint16_t *data, size_t n, size_t stride, int filt);
int orc_dwt_inverse_1d_u16(uint16_t *data, size_t n, size_t stride, int filt);

/* the native uint8 transform (icer_wavelet.c:8-53, :216-381): int8 arithmetic; the inverse reproduces
 * icer_interleave_uint8's placement for odd line lengths (:570-628) */
int orc_dwt_forward_u8(uint8_t *img, size_t w, size_t h, int stages, int filt);
int orc_dwt_inverse_u8(uint8_t *img, size_t w, size_t h, int stages, int filt);
/* dst[i] = position icer_interleave_uint8 gives element i of a length-len array */
void orc_interleave8_map(size_t len, uint32_t *dst);

/* partition geometry (icer_partition.c:7-54 + the segment walk of :299-385); returns #rects or <0 */
int orc_partition_rects(size_t w, size_t h, int segments, orc_rect *out /* >= 33 entries */);

/* one segment of one bit plane (icer_context_modeller.c:312-457 + icer_encoding.c), no header.
 * returns bit count (>=0) or a negative status; out must hold max_bytes+1 bytes */
long orc_encode_segment_plane(const uint16_t *seg, size_t w, size_t h, size_t rowstride, int subband, int lsb,
                              uint8_t *out, size_t max_bytes);
/* mirror (icer_context_modeller.c:461-602 + icer_decoding.c) */
int orc_decode_segment_plane(uint16_t *seg, size_t w, size_t h, size_t rowstride, int subband, int lsb,
                             const uint8_t *in, uint32_t nbits);

uint32_t orc_crc32(const uint8_t *buf, size_t len);

/* whole image. chan = 1 (gray) or 3 (planar Y,U,V).  `out` must hold 2*quota bytes: [0,quota) receives
 * the priority-ordered stream (reference data_start), [quota,2*quota) the rearranged stream
 * (reference rearrange_start).  Images are clobbered exactly like the reference (sign-magnitude coefs). */
int orc_compress_u16(uint16_t *const img[3], int chan, size_t w, size_t h, int stages, int filt, int segments,
                     const orc_limits *lim, uint8_t *out, size_t quota, size_t *size_used);
int orc_decompress_u16(uint16_t *const img[3], int chan, size_t *w, size_t *h, size_t bufsize_px,
                       const uint8_t *stream, size_t len, int stages, int filt, int segments,
                       const orc_limits *lim);

/* the native uint8 codec (icer_compress.c:18-274, icer_color.c:19-338): call with lim = {7, 6, 300} for the reference's
 * defaults (ICER_BITPLANES_TO_COMPRESS_8, ICER_MAX_PACKETS) */
int orc_compress_u8(uint8_t *const img[3], int chan, size_t w, size_t h, int stages, int filt, int segments,
                    const orc_limits *lim, uint8_t *out, size_t quota, size_t *size_used);
int orc_decompress_u8(uint8_t *const img[3], int chan, size_t *w, size_t *h, size_t bufsize_px,
                      const uint8_t *stream, size_t len, int stages, int filt, int segments,
                      const orc_limits *lim);

/* table access for cross-checks with the reference's globals */
int orc_bin_of(uint32_t zero, uint32_t total);
void orc_get_code_tables(uint8_t enc[17][32][3], uint8_t dec[17][32][3], uint8_t flush[17][9][6][2],
                         uint16_t golomb[17][3]);

/* deterministic synthetic frames, SURVEY.md 8(d): kind 0 natural 8-bit, 1 noise 8-bit, 2 natural 12-bit, 3 noise 12-bit */
void orc_synth_frame(uint16_t *out, size_t w, size_t h, uint32_t frame, uint32_t seed, int kind);

#ifdef __cplusplus
}
#endif
#endif
