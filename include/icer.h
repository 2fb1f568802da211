/*
 * icer.h -- drop-in public header of libicer_b200 (Blackwell-native ICER codec).
 *
 * This header declares the SAME C entry points, status codes, filter enum and caller-visible structs as
 * lib_icer's lib_icer/inc/icer.h in TheRealOrange/icer_compression, so that the reference's example/
 * programs and the icer_util CLI compile and link against libicer_b200.so unchanged
 * (callers: example/src/example_encode.c:36,66,69, example_decode.c:31,54, example_encode_color.c:65,93,96,
 * example_decode_color.c:61,84, example/src/icer_util.c:186,201,205,283,308,318,466).
 *
 * It is written from scratch: only the interface is shared.  All work happens in hand-written sm_100a
 * CUDA kernels behind these calls; there is no CPU fallback -- without a CUDA device every codec call
 * returns ICER_FATAL_ERROR and prints the reason on stderr.
 *
 * Each declaration cites the reference declaration it replaces (lib_icer/inc/icer.h:LINE).
 */
#ifndef ICER_B200_COMPAT_ICER_H
#define ICER_B200_COMPAT_ICER_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* compile-time limits of the reference (icer.h:27-46).  In this library they are only DEFAULTS: the
 * run-time values are set with icer_b200_set_limits() (include/icer_b200.h). */
#define ICER_CIRC_BUF_SIZE 2048
#ifndef ICER_MAX_SEGMENTS
#define ICER_MAX_SEGMENTS 32
#endif
#ifndef ICER_MAX_DECOMP_STAGES
#define ICER_MAX_DECOMP_STAGES 6
#endif
#ifndef ICER_MAX_PACKETS
#define ICER_MAX_PACKETS 300
#endif
#ifndef ICER_MAX_PACKETS_16
#define ICER_MAX_PACKETS_16 800
#endif
#ifndef ICER_BITPLANES_TO_COMPRESS_8
#define ICER_BITPLANES_TO_COMPRESS_8 7
#endif
#ifndef ICER_BITPLANES_TO_COMPRESS_16
#define ICER_BITPLANES_TO_COMPRESS_16 9
#endif

/* icer.h:92-105 */
enum icer_status {
    ICER_RESULT_OK = 0,
    ICER_INTEGER_OVERFLOW = -1,
    ICER_OUTPUT_BUF_TOO_SMALL = -2,
    ICER_TOO_MANY_SEGMENTS = -3,
    ICER_TOO_MANY_STAGES = -4,
    ICER_BYTE_QUOTA_EXCEEDED = -5,
    ICER_BITPLANE_OUT_OF_RANGE = -6,
    ICER_DECODER_OUT_OF_DATA = -7,
    ICER_DECODED_INVALID_DATA = -8,
    ICER_PACKET_COUNT_EXCEEDED = -9,
    ICER_FATAL_ERROR = -10,
    ICER_INVALID_INPUT = -11,
};

/* icer.h:107-115 */
enum icer_filter_types {
    ICER_FILTER_A = 0,
    ICER_FILTER_B,
    ICER_FILTER_C,
    ICER_FILTER_D,
    ICER_FILTER_E,
    ICER_FILTER_F,
    ICER_FILTER_Q
};

/* icer.h:180-186 */
#define ICER_SUBBAND_MAX 3
enum icer_subband_types { ICER_SUBBAND_LL = 0, ICER_SUBBAND_HL, ICER_SUBBAND_LH, ICER_SUBBAND_HH = ICER_SUBBAND_MAX };

/* icer.h:342-348 */
#define ICER_CHANNEL_MIN 0
#define ICER_CHANNEL_MAX 2
enum icer_color_channels { ICER_CHANNEL_Y = ICER_CHANNEL_MIN, ICER_CHANNEL_U = 1, ICER_CHANNEL_V = ICER_CHANNEL_MAX };

/* segment geometry, icer.h:126-142 (same field names and order) */
typedef struct {
    uint16_t w, h, r, c, r_t, h_t, x_t, c_t0, y_t, r_t0, x_b, c_b0, y_b, r_b0, s;
} partition_param_typdef;

/* wire format of one segment header: 28 bytes, little endian, no padding (icer.h:293-305) */
#define ICER_PACKET_PREAMBLE 0x605B
#define ICER_SEGMENT_LSB_MASK 0x0f
#define ICER_SEGMENT_CHANNEL_MASK 0xf0
#define ICER_GET_LSB_MACRO(x) ((x) & ICER_SEGMENT_LSB_MASK)
#define ICER_GET_CHANNEL_MACRO(x) (((x) & ICER_SEGMENT_CHANNEL_MASK) >> 4)
#define ICER_SET_CHANNEL_MACRO(x) ((x) << 4)
typedef struct {
    uint16_t preamble;
    uint16_t ll_mean_val;
    uint8_t decomp_level;
    uint8_t subband_type;
    uint8_t segment_number;
    uint8_t lsb_chan;
    uint32_t image_w;
    uint32_t image_h;
    uint32_t data_length; /* payload length in BITS */
    uint32_t data_crc32;
    uint32_t crc32;
} icer_image_segment_typedef;

/* caller-owned output descriptor (icer.h:307-312): [data_start, +size_used) is the priority-ordered
 * stream, [rearrange_start, +size_used) the final segment-ordered stream callers write out */
typedef struct {
    size_t size_used;
    size_t size_allocated;
    uint8_t *data_start;
    uint8_t *rearrange_start;
} icer_output_data_buf_typedef;

/* ---- the seven entry points the reference's tools import ------------------------------------------ */

int icer_init(void);                                                                               /* icer.h:370 */
int icer_init_output_struct(icer_output_data_buf_typedef *out, uint8_t *data, size_t buf_len,
                            size_t byte_quota);                                                    /* icer.h:524 */
int icer_get_image_dimensions(const uint8_t *datastream, size_t data_length, size_t *image_w,
                              size_t *image_h);                                                    /* icer.h:374 */

int icer_compress_image_uint16(uint16_t *image, size_t image_w, size_t image_h, uint8_t stages,
                               enum icer_filter_types filt, uint8_t segments,
                               icer_output_data_buf_typedef *output_data);                         /* icer.h:440-441 */
int icer_compress_image_yuv_uint16(uint16_t *y_channel, uint16_t *u_channel, uint16_t *v_channel, size_t image_w,
                                   size_t image_h, uint8_t stages, enum icer_filter_types filt, uint8_t segments,
                                   icer_output_data_buf_typedef *output_data);                     /* icer.h:442-444 */
int icer_decompress_image_uint16(uint16_t *image, size_t *image_w, size_t *image_h, size_t image_bufsize,
                                 const uint8_t *datastream, size_t data_length, uint8_t stages,
                                 enum icer_filter_types filt, uint8_t segments);                   /* icer.h:461-462 */
int icer_decompress_image_yuv_uint16(uint16_t *y_channel, uint16_t *u_channel, uint16_t *v_channel, size_t *image_w,
                                     size_t *image_h, size_t image_bufsize, const uint8_t *datastream,
                                     size_t data_length, uint8_t stages, enum icer_filter_types filt,
                                     uint8_t segments);                                            /* icer.h:463-466 */

/* ---- the native uint8 codec (int8 wavelet arithmetic, ICER_BITPLANES_TO_COMPRESS_8 planes, ICER_MAX_PACKETS packets) ---- */

int icer_compress_image_uint8(uint8_t *image, size_t image_w, size_t image_h, uint8_t stages,
                              enum icer_filter_types filt, uint8_t segments,
                              icer_output_data_buf_typedef *output_data);                          /* icer.h:386-387 */
int icer_compress_image_yuv_uint8(uint8_t *y_channel, uint8_t *u_channel, uint8_t *v_channel, size_t image_w,
                                  size_t image_h, uint8_t stages, enum icer_filter_types filt, uint8_t segments,
                                  icer_output_data_buf_typedef *output_data);                      /* icer.h:388-390 */
int icer_decompress_image_uint8(uint8_t *image, size_t *image_w, size_t *image_h, size_t image_bufsize,
                                const uint8_t *datastream, size_t data_length, uint8_t stages,
                                enum icer_filter_types filt, uint8_t segments);                    /* icer.h:407-408 */
int icer_decompress_image_yuv_uint8(uint8_t *y_channel, uint8_t *u_channel, uint8_t *v_channel, size_t *image_w,
                                    size_t *image_h, size_t image_bufsize, const uint8_t *datastream,
                                    size_t data_length, uint8_t stages, enum icer_filter_types filt,
                                    uint8_t segments);                                             /* icer.h:409-412 */
int icer_wavelet_transform_stages_uint8(uint8_t *image, size_t image_w, size_t image_h, uint8_t stages,
                                        enum icer_filter_types filt);                              /* icer.h:392 */
int icer_inverse_wavelet_transform_stages_uint8(uint8_t *image, size_t image_w, size_t image_h, uint8_t stages,
                                                enum icer_filter_types filt);                      /* icer.h:414 */

/* ---- public sub-APIs kept as kernel-level parity hooks -------------------------------------------- */

int icer_wavelet_transform_stages_uint16(uint16_t *image, size_t image_w, size_t image_h, uint8_t stages,
                                         enum icer_filter_types filt);                             /* icer.h:446 */
int icer_inverse_wavelet_transform_stages_uint16(uint16_t *image, size_t image_w, size_t image_h, uint8_t stages,
                                                 enum icer_filter_types filt);                     /* icer.h:468 */
int icer_generate_partition_parameters(partition_param_typdef *params, size_t ll_w, size_t ll_h,
                                       uint16_t segments);                                         /* icer.h:519 */
int icer_find_packet_in_bytestream(const icer_image_segment_typedef **seg, const uint8_t *datastream,
                                   size_t data_length, size_t *offset);                            /* icer.h:512 */
uint32_t icer_calculate_packet_crc32(const icer_image_segment_typedef *pkt);                       /* icer.h:521 */
uint32_t icer_calculate_segment_crc32(const icer_image_segment_typedef *pkt);                      /* icer.h:522 */
size_t icer_get_dim_n_low_stages(size_t dim, uint8_t stages);                                      /* icer.h:515 */
size_t icer_get_dim_n_high_stages(size_t dim, uint8_t stages);                                     /* icer.h:516 */

/* header inlines the tools use (example_decode.c:63; icer.h:586-598) */
static inline int icer_max_int(int a, int b) { return (a > b) ? a : b; }
static inline int icer_min_int(int a, int b) { return (a < b) ? a : b; }
static inline unsigned icer_max_uint(unsigned a, unsigned b) { return (a > b) ? a : b; }

#ifdef __cplusplus
}
#endif
#endif /* ICER_B200_COMPAT_ICER_H */
