// parse.cu -- K6: the byte-stream scanner on the device.
//
// lib_icer's decoder first walks the stream byte by byte looking for segments (icer_compress.c:443-459 calling
// icer_find_packet_in_bytestream :569-588): at offset o a segment is accepted iff the 16-bit preamble is 0x605B, the
// CRC-32 of the 24 header bytes matches, the payload length fits in what is left of the stream and the payload CRC-32
// matches; then o jumps past the segment, otherwise o advances by one byte.  The last accepted header supplies the
// image dimensions and (per channel) the LL mean; for duplicate (channel, level, subband, segment, plane) entries the
// last one wins.
//
// Device formulation (exactly equivalent):
//   k_scan    every byte offset is tested independently for "a valid segment COULD start here" (preamble, header CRC,
//             length) -> candidate list per frame (unordered, atomics)
//   k_verify  payload CRC of every candidate (one thread each)
//   k_walk    one block per frame: bitonic sort of the candidates by offset, then thread 0 replays the sequential
//             walk over them -- a candidate is accepted iff it starts at or after the end of the previously accepted
//             one; offsets that are not valid candidates only ever advance the reference's cursor by one byte
//   k_table   per dimension group: accepted segments -> [frame][stream] offset table through the plan's unit LUT
#include "common.h"
#include <mutex>

namespace {

__constant__ uint32_t c_crc[256];
// CRC-32 of a 24-byte header by position tables (see coder.cu g_hdr_crc_pos): crc = ~(K ^ XOR_i P[i][byte_i])
__device__ uint32_t g_pos_crc[24][256];
__device__ uint32_t g_pos_crc_k;
std::mutex g_crc_mu;
bool g_crc_uploaded[64] = {false};       // per device: __constant__ memory has one instance per GPU

int upload_crc() {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return ICER_FATAL_ERROR;
    std::lock_guard<std::mutex> lk(g_crc_mu);
    if (g_crc_uploaded[dev]) return ICER_RESULT_OK;
    uint32_t tab[256];
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        tab[n] = c;
    }
    CUDA_TRY(cudaMemcpyToSymbol(c_crc, tab, sizeof tab));
    {
        static uint32_t pos[24][256];
        auto A = [&](uint32_t r) { return tab[r & 0xFF] ^ (r >> 8); };
        for (int i = 0; i < 24; i++)
            for (uint32_t b = 0; b < 256; b++) {
                uint32_t r = tab[b];
                for (int k = 0; k < 23 - i; k++) r = A(r);
                pos[i][b] = r;
            }
        uint32_t kk = 0xFFFFFFFFu;
        for (int k = 0; k < 24; k++) kk = A(kk);
        CUDA_TRY(cudaMemcpyToSymbol(g_pos_crc, pos, sizeof pos));
        CUDA_TRY(cudaMemcpyToSymbol(g_pos_crc_k, &kk, sizeof kk));
    }
    g_crc_uploaded[dev] = true;
    return ICER_RESULT_OK;
}

__device__ __forceinline__ uint32_t ld32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t) p[3] << 24); }

__device__ __forceinline__ uint32_t byte_at(const uint8_t *s, uint64_t len, long long o) { return (o >= 0 && (uint64_t) o < len) ? s[o] : 0u; }

// A thread tests 16 byte offsets for the preamble (four aligned 32-bit words + one for the look-ahead, a SWAR test for "some byte is
// 0x5B" per word).  The hits of a warp are then checked ONE AFTER THE OTHER BY THE WHOLE WARP: 24 lanes take one header byte each, the
// position tables and an XOR reduction give the header CRC -- a candidate costs ~20 warp instructions instead of a 24-step serial CRC
// in a single lane (which was most of this kernel).
__global__ void __launch_bounds__(256) k_scan(ParseArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t frame = blockIdx.y;
    const uint64_t len = a.lens[frame] < a.stride ? a.lens[frame] : a.stride;
    const uint8_t *s = a.streams + (size_t) frame * a.stride;
    const int mis = (int) ((uintptr_t) s & 3);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(s - mis);          // aligned view; word j covers stream offsets 4j - mis .. 4j - mis + 3
    const uint64_t wj = ((uint64_t) blockIdx.x * 256 + threadIdx.x) * 4;       // this thread's four words
    const long long o0 = (long long) wj * 4 - mis;                             // stream offset of this thread's first byte
    uint32_t hits = 0;                                                          // bit i: a preamble starts at o0 + i and a header fits behind it
    if (o0 < (long long) len) {
        uint32_t w[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const long long o = o0 + 4 * k;
            if (o >= 0 && (uint64_t) o + 4 <= len) w[k] = sw[wj + k];
            else w[k] = byte_at(s, len, o) | (byte_at(s, len, o + 1) << 8) | (byte_at(s, len, o + 2) << 16) | (byte_at(s, len, o + 3) << 24);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t t = w[k] ^ 0x5B5B5B5Bu;
            if (((t - 0x01010101u) & ~t & 0x80808080u) == 0u) continue;       // no byte of this word is 0x5B
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const long long o = o0 + 4 * k + b;
                if ((__funnelshift_r(w[k], w[k + 1], 8 * b) & 0xFFFFu) == 0x605Bu && o >= 0 && (uint64_t) o + ICER_HDR_BYTES <= len) hits |= 1u << (4 * k + b);
            }
        }
    }
    for (;;) {
        const uint32_t who = __ballot_sync(0xffffffffu, hits != 0u);
        if (!who) break;
        const int src = __ffs((int) who) - 1;
        const uint32_t hm = __shfl_sync(0xffffffffu, hits, src);
        const int pos = __ffs((int) hm) - 1;
        const long long ob = __shfl_sync(0xffffffffu, o0, src);
        if (lane == src) hits &= hits - 1;
        const uint8_t *h = s + (ob + pos);
        uint32_t x = lane < 24 ? g_pos_crc[lane][h[lane]] : 0u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x ^= __shfl_xor_sync(0xffffffffu, x, d);
        if (lane == 0 && ~(x ^ g_pos_crc_k) == ld32(h + 24)) {
            const uint64_t o = (uint64_t) (ob + pos);
            const uint32_t nbits = ld32(h + 16);
            const uint64_t nbytes = ((uint64_t) nbits + 7) >> 3;
            if (nbytes <= len - o - ICER_HDR_BYTES) {
                const uint32_t slot = atomicAdd(&a.ncand[frame], 1u);
                if (slot < a.cap) a.cand[(size_t) frame * a.cap + slot] = make_uint2((uint32_t) o, (uint32_t) nbytes);
            }
        }
    }
}

__global__ void __launch_bounds__(128) k_verify(ParseArgs a) {
    __shared__ uint32_t crc[256];
    for (int i = threadIdx.x; i < 256; i += 128) crc[i] = c_crc[i];
    __syncthreads();
    const uint32_t frame = blockIdx.y;
    const uint32_t k = blockIdx.x * 128 + threadIdx.x;
    uint32_t cnt = a.ncand[frame];
    if (cnt > a.cap) cnt = a.cap;
    if (k >= cnt) return;
    uint2 cd = a.cand[(size_t) frame * a.cap + k];
    const uint8_t *h = a.streams + (size_t) frame * a.stride + cd.x;
    const uint8_t *p = h + ICER_HDR_BYTES;
    uint32_t c = 0xFFFFFFFFu;
    uint32_t n = cd.y, i = 0;
    // bytes up to a 4-byte boundary, then words
    while (i < n && ((uintptr_t) (p + i) & 3)) { c = crc[(c ^ p[i]) & 0xFF] ^ (c >> 8); i++; }
    for (; i + 4 <= n; i += 4) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(p + i);
        c = crc[(c ^ w) & 0xFF] ^ (c >> 8);
        c = crc[(c ^ (w >> 8)) & 0xFF] ^ (c >> 8);
        c = crc[(c ^ (w >> 16)) & 0xFF] ^ (c >> 8);
        c = crc[(c ^ (w >> 24)) & 0xFF] ^ (c >> 8);
    }
    for (; i < n; i++) c = crc[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    if (~c != ld32(h + 20)) a.cand[(size_t) frame * a.cap + k].y = 0xFFFFFFFFu;   // payload CRC mismatch: not a segment
}

// one block per frame; dynamic shared memory: keys[np2], vals[np2]
__global__ void __launch_bounds__(256) k_walk(ParseArgs a) {
    extern __shared__ uint32_t sm[];
    const uint32_t frame = blockIdx.x;
    uint32_t cnt = a.ncand[frame];
    ParseInfo info;
    info.w = info.h = 0; info.mean[0] = info.mean[1] = info.mean[2] = 0; info.any = 0; info.overflow = 0; info.naccepted = 0;
    if (cnt > a.cap) {           // too many candidates for the device path: the host scans this frame
        if (threadIdx.x == 0) { info.overflow = 1; a.info[frame] = info; }
        return;
    }
    uint32_t np2 = 1;
    while (np2 < cnt) np2 <<= 1;
    uint32_t *keys = sm, *vals = sm + a.cap_pow2;
    uint2 *cand = a.cand + (size_t) frame * a.cap;
    for (uint32_t i = threadIdx.x; i < np2; i += 256) {
        if (i < cnt) { const uint2 c = cand[i]; keys[i] = c.x; vals[i] = c.y; }
        else { keys[i] = 0xFFFFFFFFu; vals[i] = 0xFFFFFFFFu; }
    }
    __syncthreads();
    for (uint32_t k = 2; k <= np2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < np2; i += 256) {
                const uint32_t l = i ^ j;
                if (l > i) {
                    const bool asc = (i & k) == 0;
                    const uint32_t ki = keys[i], kl = keys[l];
                    if ((ki > kl) == asc) { keys[i] = kl; keys[l] = ki; const uint32_t t = vals[i]; vals[i] = vals[l]; vals[l] = t; }
                }
            }
            __syncthreads();
        }
    // the sequential part touches shared memory only: compact the accepted candidates to the front, in stream order
    __shared__ uint32_t s_nacc;
    __shared__ unsigned long long s_mean[3];
    if (threadIdx.x < 3) s_mean[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        uint64_t cur = 0;
        uint32_t nacc = 0;
        for (uint32_t i = 0; i < cnt; i++) {
            const uint32_t v = vals[i], off = keys[i];
            if (v == 0xFFFFFFFFu) continue;               // failed the payload CRC
            if (off < cur) continue;                      // inside a segment that was accepted before it
            keys[nacc] = off; vals[nacc] = v; nacc++;     // nacc <= i: in place
            cur = (uint64_t) off + ICER_HDR_BYTES + v;
        }
        s_nacc = nacc;
    }
    __syncthreads();
    const uint32_t nacc = s_nacc;
    const uint8_t *s = a.streams + (size_t) frame * a.stride;
    for (uint32_t j = threadIdx.x; j < nacc; j += 256) {
        const uint32_t off = keys[j];
        cand[j] = make_uint2(off, vals[j]);
        const uint8_t *h = s + off;
        const uint32_t ch = a.chan == 1 ? 0 : (h[7] >> 4);
        // per channel the mean of the LAST accepted header wins (icer_compress.c:454-456)
        if (ch < 3) atomicMax(&s_mean[ch], ((unsigned long long) (j + 1) << 16) | (uint32_t) (h[2] | (h[3] << 8)));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (nacc) {
            const uint8_t *h = s + keys[nacc - 1];
            info.w = ld32(h + 8); info.h = ld32(h + 12);
            info.any = 1;
        }
        for (int c = 0; c < 3; c++) info.mean[c] = (uint16_t) (s_mean[c] & 0xFFFF);
        info.naccepted = nacc;
        a.info[frame] = info;
    }
}

// accepted segments of the frames of one dimension group -> segment table (entries hold offset + 1, 0 = absent).  One
// thread per accepted segment; for duplicates lib_icer's table assignment lets the LAST one in stream order win
// (icer_compress.c:453) -- offsets grow in stream order, so that is an atomicMax.
__global__ void __launch_bounds__(128) k_table(TableArgs t) {
    const uint32_t frame = t.list[blockIdx.y];
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= t.info[frame].naccepted) return;
    const uint32_t off = t.cand[(size_t) frame * t.cap + i].x;
    const uint8_t *h = t.streams + (size_t) frame * t.stride + off;
    const uint32_t level = h[4], sb = h[5], seg = h[6], lsb = h[7] & 0x0F, ch = t.chan == 1 ? 0 : (h[7] >> 4);
    if (lsb >= t.planes || ch >= t.chan || level > t.max_level || sb > 3 || seg > ICER_MAX_SEGMENTS) return;
    const uint32_t unit = t.unit_lut[((ch * (t.max_level + 1) + level) * 4 + sb) * (ICER_MAX_SEGMENTS + 1) + seg];
    if (unit == 0xFFFFFFFFu) return;
    atomicMax(&t.seg_off[(size_t) frame * t.nstreams + unit * t.planes + lsb], off + 1);
}

}  // namespace

int launch_parse(const ParseArgs &a, cudaStream_t st, LaunchStats &ls) {
    int rc = upload_crc();
    if (rc) return rc;
    if (a.n == 0) return ICER_RESULT_OK;
    CUDA_TRY(cudaMemsetAsync(a.ncand, 0, (size_t) a.n * 4, st));
    for (uint32_t f0 = 0; f0 < a.n; f0 += 32768) {
        ParseArgs b = a;
        const uint32_t cnt = a.n - f0 < 32768 ? a.n - f0 : 32768;
        b.streams += (size_t) f0 * a.stride; b.lens += f0; b.cand += (size_t) f0 * a.cap; b.ncand += f0; b.info += f0; b.n = cnt;
        dim3 g1((unsigned) ((a.stride + 3 + 256 * 16 - 1) / (256 * 16)), cnt);      // 16 byte offsets per thread (+3: the aligned view may start up to 3 bytes early)
        k_scan<<<g1, 256, 0, st>>>(b);
        dim3 g2((a.cap + 127) / 128, cnt);
        k_verify<<<g2, 128, 0, st>>>(b);
        const size_t smem = (size_t) a.cap_pow2 * 8;
        CUDA_TRY(cudaFuncSetAttribute(k_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        k_walk<<<cnt, 256, smem, st>>>(b);
        ls.launches += 3;
    }
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}

int launch_table(const TableArgs &t, cudaStream_t st, LaunchStats &ls) {
    if (t.nlist == 0) return ICER_RESULT_OK;
    for (uint32_t l0 = 0; l0 < t.nlist; l0 += 32768) {
        TableArgs b = t;
        b.list += l0; b.nlist = t.nlist - l0 < 32768 ? t.nlist - l0 : 32768;
        k_table<<<dim3((t.cap + 127) / 128, b.nlist), 128, 0, st>>>(b);
    }
    ls.launches++;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
