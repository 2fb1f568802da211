// common.h -- internal declarations shared by the host orchestration and the CUDA kernels of libicer_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <vector>
#include <map>
#include <tuple>

#include "../../include/icer.h"
#include "../../include/icer_b200.h"
#include "coder_tables.h"

#define ICER_HDR_BYTES 28
#define ICER_RING 2048

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            fprintf(stderr, "[icer_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), __FILE__,  \
                    __LINE__, cudaGetErrorString(_e));                                                   \
            return ICER_FATAL_ERROR;                                                                     \
        }                                                                                                \
    } while (0)

// ---- code tables (ICER standard; lib_icer builds the same at run time in src/icer_init.c) ---------------
const HostTables &icer_host_tables();

// ---- plan: everything that depends only on (w, h, channels, stages, segments, planes) ---------------------
// A "unit" is one error-containment segment of one subband of one channel; it carries `planes` bit-plane
// streams (stream id = unit * planes + lsb).  Units are ordered big-first (level 1 first) so that warps get
// similar amounts of work and the long streams start early.
struct Unit {
    int32_t x, y, w, h;        // rectangle inside the channel plane
    uint8_t chan, level, subband, seg;
    uint32_t px;
    uint32_t slot0, cw;        // bit-slice slots: first slot of the unit, chunks (slots) per row; the unit owns h * cw slots, row-major
};
// One bit-slice slot = one chunk of ICER_CHUNK pixels of one row of one unit (coder.cu).  flags bit 0 / 1: the pixel left / right of
// the chunk belongs to the same unit (else that halo bit is zero).
struct SlotDesc { uint16_t x0, y, unit; uint8_t ncols, flags; };

struct Plan {
    uint32_t w, h, chan, stages, segments, planes, max_stages;
    int status;                          // ICER_RESULT_OK or the error every frame of this shape gets
    int enc_status;                      // errors only the encoder reports (ICER_PACKET_COUNT_EXCEEDED)
    std::vector<Unit> units;
    std::vector<SlotDesc> slots;         // bit-slice slots of all units (unit after unit, row-major inside a unit)
    uint32_t nslots = 0;
    SlotDesc *d_slots = nullptr;
    uint32_t nstreams;                   // units.size() * planes
    std::vector<uint32_t> prio_order;    // stream ids in (priority packet, segment) order
    std::vector<uint32_t> rearr_order;   // stream ids in the rearranged output order
    std::vector<uint32_t> scratch_off;   // per stream: 32-bit word offset into the frame's payload scratch
    std::vector<uint32_t> scratch_cap;   // per stream: capacity in 32-bit words
    uint64_t scratch_words;              // per frame
    int scratch_bits;
    // device copies
    Unit *d_units = nullptr;
    uint32_t *d_prio = nullptr, *d_rearr = nullptr, *d_soff = nullptr, *d_scap = nullptr;
    std::vector<uint32_t> unit_lut;      // [chan][max_stages+1][4][ICER_MAX_SEGMENTS+1] -> unit index or ~0 (device scanner)
    uint32_t *d_unit_lut = nullptr;
    // lookup (chan, level, subband, seg) -> unit index, for the decoder's segment table
    std::map<std::tuple<int, int, int, int>, uint32_t> unit_of;
    // priority-prefix encoding (byte quotas far below the lossless size): how many streams of the priority order the encoder
    // produces first; learned from the previous batch of this shape (0 = all at once)
    uint32_t enc_prefix = 0;
    uint64_t enc_prefix_quota = 0;       // the quota the hint was learned for
    uint32_t *d_list = nullptr;          // [nstreams]: the first `list_prefix` streams of the priority order, then the rest; each part big streams first
    uint32_t list_prefix = 0;
};
int upload_prefix_lists(Plan &pl, uint32_t prefix, cudaStream_t st);

int build_plan(Plan &pl, uint32_t w, uint32_t h, uint32_t chan, uint32_t stages, uint32_t segments, uint32_t planes,
               uint32_t max_stages, uint32_t max_packets, int scratch_bits, bool native_u8 = false);
void u8_interleave_dst(uint32_t len, std::vector<uint32_t> &dst);
void free_plan_device(Plan &pl);
int upload_plan(Plan &pl);

static inline size_t dim_low(size_t d, unsigned s) { size_t p = (size_t) 1 << s; return (d + p - 1) / p; }
static inline size_t dim_high(size_t d, unsigned s) { return dim_low(d, s - 1) / 2; }
int partition_rects(size_t w, size_t h, int segments, int *xs, int *ys, int *ws, int *hs, partition_param_typdef *pp);
uint32_t host_crc32(const uint8_t *buf, size_t len);

int ctx_device(const icer_b200_ctx *c);      // api.cu

// ---- kernel launchers (wavelet.cu / wavelet_tma.cu / coder.cu) -------------------------------------------
struct LaunchStats { uint64_t launches = 0; };

// forward / inverse multi-stage DWT over `nplanes` planes of w*h int16, in place (tmp has the same size)
int launch_dwt_forward(uint16_t *planes, uint16_t *tmp, size_t nplanes, uint32_t w, uint32_t h, uint32_t stages,
                       uint32_t filt, int *d_ovf /* per plane, pre-zeroed */, cudaStream_t st, LaunchStats &ls, bool bits8 = false);
int launch_dwt_inverse(uint16_t *planes, uint16_t *tmp, size_t nplanes, uint32_t w, uint32_t h, uint32_t stages,
                       uint32_t filt, bool clamp_negative, cudaStream_t st, LaunchStats &ls, bool bits8 = false,
                       const std::map<uint32_t, const uint32_t *> *perms = nullptr);
// LL mean removal + sign-magnitude (encode; skip_signmag: the bit-slice kernel applies the mapping instead); the decode-side inverse is fused into launch_unslice
int launch_ll_mean_signmag(uint16_t *planes, size_t nplanes, uint32_t w, uint32_t h, uint32_t llw, uint32_t llh,
                           uint16_t *d_mean, int *d_ovf, cudaStream_t st, LaunchStats &ls, bool bits8 = false, bool skip_signmag = false);

struct EncodeArgs {
    const uint32_t *bs;        // encoder bit slices [ngroups][nslots][F][32] (coder.cu), 32 frames per group
    uint32_t nslots;
    uint32_t nframes, ngroups, planes;
    const Unit *units; uint32_t nunits;
    const uint32_t *soff, *scap;   // per stream
    uint64_t scratch_words;        // per frame
    uint32_t *scratch;             // [nframes][scratch_words]
    uint32_t *bits;                // [nframes][nstreams] payload length in bits
    uint16_t *rings;               // [resident threads][2048]
    uint32_t nring_threads;
    unsigned long long *work_counter;  // dynamic work fetch, pre-zeroed
    int *overflow;                 // set when any stream outgrew its scratch
    const uint32_t *list;          // optional: the streams to code (a prefix of the priority order, big ones first); nullptr = all
    uint32_t nlist;
};
// bit count of a stream the encoder has not produced (priority-prefix encoding): larger than any quota, so the quota walk stops there
#define ICER_BITS_NOT_CODED 0xF0F0F0F0u      /* a memset pattern */
int launch_encode(const EncodeArgs &a, cudaStream_t st, LaunchStats &ls);
uint32_t encode_resident_threads(int device);

struct AssembleArgs {
    uint32_t nframes, nstreams, planes, chan, w, h;
    const Unit *units;
    const uint32_t *prio, *rearr;  // stream id orders
    const uint32_t *soff;
    uint64_t scratch_words;
    const uint32_t *scratch, *bits;
    uint32_t *crc;                 // [nframes][nstreams] payload CRC-32, filled by the assemble step
    const uint16_t *mean;          // [nframes][chan]
    const int *ovf;                // [nframes][chan] wavelet/mean overflow
    int plan_status;
    uint64_t quota;
    uint8_t *out; uint64_t out_stride;        // rearranged streams
    uint8_t *out_prio;                        // optional
    uint64_t *size_used; int32_t *rc;         // [nframes]
    uint32_t *keep;                           // [nframes] number of streams (in priority order) kept
    uint64_t *off_prio, *off_rearr;           // [nframes][nstreams] byte offsets (scratch)
    uint32_t limit;                           // streams (in priority order) the encoder has produced so far; nstreams = all
    uint32_t *stat;                           // [0] max over frames of `keep`, [1] != 0: some frame kept all `limit` streams and wants more
    int phase;                                // 0: CRC + quota walk + gather; 1: CRC + quota walk only; 2: gather only
};
int launch_assemble(const AssembleArgs &a, cudaStream_t st, LaunchStats &ls);

struct DecodeArgs {
    uint32_t *bs;                  // decoder bit slices [ngroups][nslots][3 * planes][32]: written and read by the planes of a chain
    uint32_t nslots;
    uint8_t *pstat;                // [batch frames][nunits][planes] status of every plane (coder.cu PST_*), read by the un-slicing kernel
    uint32_t ngroups, planes;
    const Unit *units; uint32_t nunits;
    const int32_t *frame_of;       // [ngroups*32] batch index of each lane's frame, -1 = padding
    const uint8_t *streams; uint64_t stream_stride;
    const uint64_t *stream_len;    // [batch frames] bytes of each frame's stream; the bit readers of a frame see zeros behind it
    const uint32_t *seg_off;       // [batch frames][nstreams] 1 + byte offset of the segment header inside the frame's stream, 0 = absent
    unsigned long long *work_counter;
};
int launch_decode(const DecodeArgs &a, cudaStream_t st, LaunchStats &ls);
// planar coefficients [n][chan][h][w] -> encoder bit slices (twos_complement: apply the sign-magnitude mapping on the way), and the
// decoder's bit slices -> planar coefficients (undoes sign-magnitude, adds the header LL mean; frame_of maps (group, lane) to the
// batch index, -1 = none)
size_t enc_slice_words(uint32_t planes, uint32_t nslots);      // 32-bit words per group
size_t dec_slice_words(uint32_t planes, uint32_t nslots);
int launch_bitslice(const uint16_t *coefs, uint32_t *bs, const SlotDesc *slots, uint32_t nslots, uint32_t nframes, uint32_t ngroups, uint32_t chan,
                    uint32_t w, uint32_t h, uint32_t planes, bool twos_complement, cudaStream_t st, LaunchStats &ls);
// pstat: [nbatch][nunits][planes] plane status written by the decoder; vp: [nbatch][nunits] scratch (lowest plane that counts)
int launch_unslice(const uint32_t *bs, uint16_t *frames, const SlotDesc *slots, const uint8_t *pstat, uint8_t *vp, size_t nbatch, const int32_t *frame_of,
                   const uint16_t *mean, uint32_t nslots, uint32_t nunits, uint32_t ngroups, uint32_t chan, uint32_t w, uint32_t h, uint64_t cap_px,
                   uint32_t planes, uint32_t llw, uint32_t llh, cudaStream_t st, LaunchStats &ls, bool bits8 = false);

// ---- K6: device stream scanner (parse.cu) ------------------------------------------------------------
struct ParseInfo {               // per frame, what lib_icer's scan loop leaves behind (icer_compress.c:443-459)
    uint32_t w, h;               // from the LAST accepted header
    uint16_t mean[3];            // LL mean per channel, last accepted header of that channel
    uint8_t any, overflow;       // any segment found; too many candidates for the device path
    uint32_t naccepted;
};
struct ParseArgs {
    const uint8_t *streams; uint64_t stride; const uint64_t *lens; uint32_t n, chan;
    uint2 *cand;                 // [n][cap] (offset, payload bytes); after k_walk: the accepted segments in stream order
    uint32_t cap, cap_pow2;
    uint32_t *ncand;             // [n]
    ParseInfo *info;             // [n]
};
int launch_parse(const ParseArgs &a, cudaStream_t st, LaunchStats &ls);
struct TableArgs {
    const uint32_t *list; uint32_t nlist;      // batch indices of the frames of one dimension group
    const uint2 *cand; uint32_t cap;
    const uint8_t *streams; uint64_t stride;
    const ParseInfo *info;
    uint32_t *seg_off; uint32_t nstreams, planes, chan, max_level;
    const uint32_t *unit_lut;
};
int launch_table(const TableArgs &t, cudaStream_t st, LaunchStats &ls);

// ---- pixel formats at the boundary (pixfmt.cu) ---------------------------------------------------------
int launch_u8_to_u16(const uint8_t *src, uint16_t *dst, size_t count, cudaStream_t st, LaunchStats &ls);
int launch_u16_to_u8(const uint16_t *src, uint8_t *dst, size_t count, cudaStream_t st, LaunchStats &ls);
// native uint8 codec: int8 samples sign-extended into the int16 containers and back (low byte); sign-magnitude 16 -> 8
int launch_s8_to_i16(const uint8_t *src, uint16_t *dst, size_t count, cudaStream_t st, LaunchStats &ls);
int launch_i16_to_b8(const uint16_t *src, uint8_t *dst, size_t count, cudaStream_t st, LaunchStats &ls);
int launch_sm16_to_sm8(const uint16_t *src, uint8_t *dst, size_t count, cudaStream_t st, LaunchStats &ls);
int launch_rgb_to_yuv(const uint8_t *rgb, uint16_t *yuv, size_t n, size_t px, cudaStream_t st, LaunchStats &ls);
int launch_yuv_to_rgb(const uint16_t *yuv, uint8_t *rgb, size_t n, size_t px, size_t cap_px, cudaStream_t st, LaunchStats &ls);

// ---- fused forward DWT stage with TMA-staged tiles (wavelet_tma.cu) -------------------------------------
bool dwt_fused_applicable(const void *src, uint32_t w, uint32_t h, size_t src_plane_stride);
// fused inverse stages (filter A): coefs is read-only, pp0 / pp1 hold the intermediate LL regions, dst receives the planes
bool dwt_inverse_fused_applicable(const void *coefs, const void *dst, uint32_t w, uint32_t h, uint32_t stages, uint32_t filt);
int launch_dwt_inverse_fused(const uint16_t *coefs, uint16_t *pp0, uint16_t *pp1, uint16_t *dst, size_t nplanes, uint32_t w, uint32_t h,
                             uint32_t stages, uint32_t filt, bool clamp_negative, cudaStream_t st, LaunchStats &ls);
int launch_dwt_forward_fused(const uint16_t *src, size_t src_plane_stride, uint16_t *dst, uint16_t *pp0, uint16_t *pp1, size_t nplanes,
                             uint32_t w, uint32_t h, uint32_t stages, uint32_t filt, int *d_ovf, cudaStream_t st, LaunchStats &ls);
