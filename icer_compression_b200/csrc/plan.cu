// plan.cu -- host-side planning: code tables, segment geometry, packet priority order, stream orderings.
//
// Reference behaviour restated here (lib_icer/src): icer_init.c:24-255 (tables), icer_partition.c:7-54 (+ the
// segment walk :299-385), icer_compress.c:315-365 / icer_color.c:398-458 (packet list, priority, sort),
// icer_compress.c:409-423 / icer_color.c:508-527 (rearranged order).
#include "common.h"
#include <algorithm>
#include <string.h>

// ---------------------------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------------------------
#include "coder_tables.h"

const HostTables &icer_host_tables() {
    static const HostTables t = icer_tables::make_tables();
    return t;
}

uint32_t host_crc32(const uint8_t *buf, size_t len) {
    static uint32_t tab[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t n = 0; n < 256; n++) {
            uint32_t c = n;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            tab[n] = c;
        }
        ready = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t k = 0; k < len; k++) c = tab[(c ^ buf[k]) & 0xFF] ^ (c >> 8);
    return ~c;
}

// ---------------------------------------------------------------------------------------------------------
// segment geometry
// ---------------------------------------------------------------------------------------------------------
int partition_rects(size_t w, size_t h, int segments, int *xs, int *ys, int *ws, int *hs, partition_param_typdef *pp) {
    size_t s = (size_t) segments;
    if (segments < 1 || s > w * h || segments > ICER_MAX_SEGMENTS) return ICER_TOO_MANY_SEGMENTS;
    size_t r, c;
    if (h > (s - 1) * w) r = s;
    else for (r = 1; r < s && (r + 1) * r * w < h * s; r++) {}
    c = s / r;
    size_t r_t = (c + 1) * r - s;
    size_t h_t = ((2 * h * c * r_t + s) / 2) / s;
    if (h_t < r_t) h_t = r_t;
    size_t x_t = w / c, c_t0 = (x_t + 1) * c - w, y_t = h_t / r_t, r_t0 = (y_t + 1) * r_t - h_t;
    size_t x_b = 0, c_b0 = 0, y_b = 0, r_b0 = 0;
    if (r_t < r) {
        x_b = w / (c + 1);
        c_b0 = (x_b + 1) * (c + 1) - w;
        y_b = (h - h_t) / (r - r_t);
        r_b0 = (y_b + 1) * (r - r_t) - (h - h_t);
    }
    if (pp) {
        pp->w = (uint16_t) w; pp->h = (uint16_t) h; pp->s = (uint16_t) s; pp->r = (uint16_t) r; pp->c = (uint16_t) c;
        pp->r_t = (uint16_t) r_t; pp->h_t = (uint16_t) h_t; pp->x_t = (uint16_t) x_t; pp->c_t0 = (uint16_t) c_t0;
        pp->y_t = (uint16_t) y_t; pp->r_t0 = (uint16_t) r_t0; pp->x_b = (uint16_t) x_b; pp->c_b0 = (uint16_t) c_b0;
        pp->y_b = (uint16_t) y_b; pp->r_b0 = (uint16_t) r_b0;
    }
    if (!xs) return (int) s;
    int n = 0;
    size_t y = 0;
    for (size_t row = 0; row < r_t; row++) {
        size_t sh = y_t + (row >= r_t0 ? 1 : 0), x = 0;
        for (size_t col = 0; col < c; col++) {
            size_t sw = x_t + (col >= c_t0 ? 1 : 0);
            xs[n] = (int) x; ys[n] = (int) y; ws[n] = (int) sw; hs[n] = (int) sh; n++;
            x += sw;
        }
        y += sh;
    }
    for (size_t row = 0; row < r - r_t; row++) {
        size_t sh = y_b + (row >= r_b0 ? 1 : 0), x = 0;
        for (size_t col = 0; col < c + 1; col++) {
            size_t sw = x_b + (col >= c_b0 ? 1 : 0);
            xs[n] = (int) x; ys[n] = (int) y; ws[n] = (int) sw; hs[n] = (int) sh; n++;
            x += sw;
        }
        y += sh;
    }
    return n;
}

// ---------------------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------------------
namespace {
struct Packet { int level, subband, lsb, chan; uint64_t prio; int order; };
}

int build_plan(Plan &pl, uint32_t w, uint32_t h, uint32_t chan, uint32_t stages, uint32_t segments, uint32_t planes,
               uint32_t max_stages, uint32_t max_packets, int scratch_bits, bool native_u8) {
    pl.w = w; pl.h = h; pl.chan = chan; pl.stages = stages; pl.segments = segments; pl.planes = planes;
    pl.max_stages = max_stages; pl.scratch_bits = scratch_bits;
    pl.status = ICER_RESULT_OK; pl.enc_status = ICER_RESULT_OK;
    pl.units.clear(); pl.slots.clear(); pl.nslots = 0; pl.prio_order.clear(); pl.rearr_order.clear(); pl.unit_of.clear();
    pl.nstreams = 0; pl.scratch_words = 0;
    // lib_icer keeps the partition geometry in 16-bit fields (icer.h:293-312): larger planes are undefined there, refused here (the
    // bit-slice slot descriptors use 16-bit coordinates as well)
    if (w > 65535u || h > 65535u) { pl.status = ICER_INVALID_INPUT; return pl.status; }
    if (stages < 1 || stages > max_stages) { pl.status = ICER_TOO_MANY_STAGES; return pl.status; }
    if (dim_low(w, stages) < 3 || dim_low(h, stages) < 3) { pl.status = ICER_TOO_MANY_STAGES; return pl.status; }

    // units, big first: level ascending; the LL band rides with the last level
    int xs[40], ys[40], ws[40], hs[40];
    for (uint32_t lv = 1; lv <= stages; lv++)
        for (uint32_t ch = 0; ch < chan; ch++)
            for (int sb = 3; sb >= 0; sb--) {
                if (sb == 0 && lv != stages) continue;
                size_t x0 = (sb == 1 || sb == 3) ? dim_low(w, lv) : 0, y0 = (sb == 2 || sb == 3) ? dim_low(h, lv) : 0;
                size_t sw = (sb == 1 || sb == 3) ? dim_high(w, lv) : dim_low(w, lv);
                size_t sh = (sb == 2 || sb == 3) ? dim_high(h, lv) : dim_low(h, lv);
                int n = partition_rects(sw, sh, (int) segments, xs, ys, ws, hs, nullptr);
                if (n < 0) { pl.status = n; return n; }   // lib_icer: undefined behaviour (icer_compress.c:401)
                for (int sg = 0; sg < n; sg++) {
                    Unit u;
                    u.x = (int) x0 + xs[sg]; u.y = (int) y0 + ys[sg]; u.w = ws[sg]; u.h = hs[sg];
                    u.chan = (uint8_t) ch; u.level = (uint8_t) lv; u.subband = (uint8_t) sb; u.seg = (uint8_t) sg;
                    u.px = (uint32_t) ws[sg] * (uint32_t) hs[sg];
                    pl.unit_of[std::make_tuple((int) ch, (int) lv, sb, sg)] = (uint32_t) pl.units.size();
                    pl.units.push_back(u);
                }
            }
    pl.nstreams = (uint32_t) pl.units.size() * planes;
    // bit-slice slots (coder.cu): per unit row, chunks of ICER_CHUNK columns
    pl.slots.clear();
    for (size_t k = 0; k < pl.units.size(); k++) {
        Unit &q = pl.units[k];
        q.cw = (uint32_t) ((q.w + ICER_CHUNK - 1) / ICER_CHUNK);
        q.slot0 = (uint32_t) pl.slots.size();
        for (int r = 0; r < q.h; r++)
            for (uint32_t j = 0; j < q.cw; j++) {
                SlotDesc d;
                d.x0 = (uint16_t) (q.x + (int) j * ICER_CHUNK); d.y = (uint16_t) (q.y + r); d.unit = (uint16_t) k;
                const int left = q.w - (int) j * ICER_CHUNK;
                d.ncols = (uint8_t) (left < ICER_CHUNK ? left : ICER_CHUNK);
                d.flags = (uint8_t) ((j > 0 ? 1 : 0) | (j + 1 < q.cw ? 2 : 0) | (q.chan << 2));
                pl.slots.push_back(d);
            }
    }
    pl.nslots = (uint32_t) pl.slots.size();
    pl.unit_lut.assign((size_t) chan * (max_stages + 1) * 4 * (ICER_MAX_SEGMENTS + 1), 0xFFFFFFFFu);
    for (size_t k = 0; k < pl.units.size(); k++) {
        const Unit &q = pl.units[k];
        pl.unit_lut[(((size_t) q.chan * (max_stages + 1) + q.level) * 4 + q.subband) * (ICER_MAX_SEGMENTS + 1) + q.seg] = (uint32_t) k;
    }

    // packets in the reference's insertion order, then a stable sort by (priority desc, subband asc)
    std::vector<Packet> pk;
    auto add = [&](int lv, int sb, int lsb, int ch, uint64_t pr) { pk.push_back(Packet{lv, sb, lsb, ch, pr, (int) pk.size()}); };
    if (chan == 1) {
        for (uint32_t st = 1; st <= stages; st++) {
            uint64_t pr = (uint64_t) 1 << st;
            for (uint32_t lsb = 0; lsb < planes; lsb++) {
                add(st, 1, lsb, 0, pr << lsb);
                add(st, 2, lsb, 0, pr << lsb);
                add(st, 3, lsb, 0, ((pr / 2) << lsb) + 1);
            }
        }
        uint64_t pr = (uint64_t) 1 << stages;
        for (uint32_t lsb = 0; lsb < planes; lsb++) add(stages, 0, lsb, 0, (2 * pr) << lsb);
    } else {
        // colour: a uint32_t priority doubled at every Y packet, cumulatively over the lsb loop (icer_color.c:401-404)
        for (uint32_t st = 1; st <= stages; st++) {
            uint32_t pr = 1u << st;
            for (uint32_t lsb = 0; lsb < planes; lsb++)
                for (uint32_t ch = 0; ch < 3; ch++) {
                    if (ch == 0) pr *= 2;
                    add(st, 1, lsb, ch, (uint32_t) (pr << lsb));
                    add(st, 2, lsb, ch, (uint32_t) (pr << lsb));
                    add(st, 3, lsb, ch, (uint32_t) (((pr / 2) << lsb) + 1));
                }
        }
        uint32_t pr = 1u << stages;
        for (uint32_t lsb = 0; lsb < planes; lsb++)
            for (uint32_t ch = 0; ch < 3; ch++) {
                if (ch == 0) pr *= 2;
                add(stages, 0, lsb, ch, (uint32_t) ((2 * pr) << lsb));
            }
    }
    // lib_icer: `ind++; if (ind >= ICER_MAX_PACKETS_16) return ICER_PACKET_COUNT_EXCEEDED` (icer_compress.c:327)
    // -- an ENCODER limit only: the decoder never builds the packet list, so the plan stays usable for decoding
    pl.enc_status = pk.size() >= max_packets ? ICER_PACKET_COUNT_EXCEEDED : ICER_RESULT_OK;
    std::stable_sort(pk.begin(), pk.end(), [](const Packet &a, const Packet &b) {
        if (a.prio != b.prio) return a.prio > b.prio;
        return a.subband < b.subband;
    });
    for (auto &p : pk)
        for (int sg = 0;; sg++) {
            auto it = pl.unit_of.find(std::make_tuple(p.chan, p.level, p.subband, sg));
            if (it == pl.unit_of.end()) break;
            pl.prio_order.push_back(it->second * planes + (uint32_t) p.lsb);
        }
    // rearranged order: segment asc, subband 3..0, level hi..lo, lsb hi..lo, channel asc -- except the native uint8
    // COLOUR encoder, whose loops run subband, level and lsb ASCENDING (icer_color.c:184-203 vs :508-527)
    const bool asc = native_u8 && chan == 3;
    for (int sg = 0; sg <= ICER_MAX_SEGMENTS; sg++)
        for (int a = 0; a <= 3; a++)
            for (int b = 0; b <= (int) max_stages; b++)
                for (int cidx = 0; cidx < (int) planes; cidx++)
                    for (int ch = 0; ch < (int) chan; ch++) {
                        const int sb = asc ? a : 3 - a, lv = asc ? b : (int) max_stages - b, lsb = asc ? cidx : (int) planes - 1 - cidx;
                        auto it = pl.unit_of.find(std::make_tuple(ch, lv, sb, sg));
                        if (it != pl.unit_of.end()) pl.rearr_order.push_back(it->second * planes + (uint32_t) lsb);
                    }
    // payload scratch layout
    pl.scratch_off.resize(pl.nstreams);
    pl.scratch_cap.resize(pl.nstreams);
    uint64_t off = 0;
    for (size_t u = 0; u < pl.units.size(); u++)
        for (uint32_t l = 0; l < planes; l++) {
            uint64_t cap = ((uint64_t) pl.units[u].px * (uint64_t) scratch_bits + 31) / 32 + 2;
            pl.scratch_off[u * planes + l] = (uint32_t) off;
            pl.scratch_cap[u * planes + l] = (uint32_t) cap;
            off += cap;
        }
    pl.scratch_words = off;
    return ICER_RESULT_OK;
}

// Where lib_icer's icer_interleave_uint8 (icer_wavelet.c:570-628) puts element i of a length-n lows|highs array.  For
// even n this is the perfect in-shuffle (i -> 2i / 2(i - n/2) + 1); for odd n the function's `halfleft` is one too large
// (:599) and the result is a different, fixed permutation, which the native uint8 decoder therefore applies
// after every inverse 1-D pass over an odd-length line.  Derived by replaying the function's index operations (odd-tail
// rotation, then per 3^k+1 slice: three reversals + cycle-leader walk, slice sizes from icer_find_k's binary search
// :823-847, which is not always the largest fitting slice) on an identity array.
void u8_interleave_dst(uint32_t len, std::vector<uint32_t> &dst) {
    std::vector<uint32_t> a(len);
    for (uint32_t i = 0; i < len; i++) a[i] = i;
    size_t slice[12];
    { size_t p3 = 1; for (int k = 0; k < 12; k++) { slice[k] = p3 + 1; p3 *= 3; } }
    auto find_k = [&](size_t want) -> int {
        int lo = 0, hi = 11, res = 0;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (want > slice[mid]) { lo = mid + 1; res = mid; }
            else if (want < slice[mid]) hi = mid - 1;
            else break;
        }
        return res;
    };
    auto reverse = [&](size_t s, size_t e) { while (s < e) { std::swap(a[s], a[e]); s++; e--; } };
    const bool odd = len & 1;
    size_t n = odd ? len - 1 : len;
    if (odd) {
        const uint32_t keep = a[n / 2];
        for (size_t i = n / 2; i < n; i++) a[i] = a[i + 1];
        a[len - 1] = keep;
    }
    size_t done = 0;
    while (done < n) {
        const size_t seg = slice[find_k(n - done)], half = seg / 2, left = n - done;
        const size_t halfleft = left / 2 - (odd ? 0 : 1);
        reverse(done + half, done + halfleft + half);
        reverse(done + half, done + seg - 1);
        reverse(done + seg, done + halfleft + half);
        for (size_t i = 1; i < seg; i *= 3) {
            size_t j = i;
            uint32_t carry = a[j + done];
            do {
                j = j < half ? 2 * j : (j - half) * 2 + 1;
                std::swap(a[j + done], carry);
            } while (j != i);
        }
        done += seg;
    }
    dst.assign(len, 0);
    for (uint32_t pos = 0; pos < len; pos++) dst[a[pos]] = pos;
}

void free_plan_device(Plan &pl) {
    cudaFree(pl.d_slots); pl.d_slots = nullptr;
    cudaFree(pl.d_units); cudaFree(pl.d_prio); cudaFree(pl.d_rearr); cudaFree(pl.d_soff); cudaFree(pl.d_scap); cudaFree(pl.d_unit_lut);
    pl.d_unit_lut = nullptr;
    cudaFree(pl.d_list); pl.d_list = nullptr; pl.list_prefix = 0;
    pl.d_units = nullptr; pl.d_prio = pl.d_rearr = pl.d_soff = pl.d_scap = nullptr;
}

int upload_plan(Plan &pl) {
    free_plan_device(pl);
    if (pl.status != ICER_RESULT_OK || pl.units.empty()) return ICER_RESULT_OK;
    CUDA_TRY(cudaMalloc(&pl.d_units, pl.units.size() * sizeof(Unit)));
    CUDA_TRY(cudaMalloc(&pl.d_prio, pl.nstreams * 4));
    CUDA_TRY(cudaMalloc(&pl.d_rearr, pl.nstreams * 4));
    CUDA_TRY(cudaMalloc(&pl.d_soff, pl.nstreams * 4));
    CUDA_TRY(cudaMalloc(&pl.d_scap, pl.nstreams * 4));
    CUDA_TRY(cudaMemcpy(pl.d_units, pl.units.data(), pl.units.size() * sizeof(Unit), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(pl.d_prio, pl.prio_order.data(), pl.nstreams * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(pl.d_rearr, pl.rearr_order.data(), pl.nstreams * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(pl.d_soff, pl.scratch_off.data(), pl.nstreams * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(pl.d_scap, pl.scratch_cap.data(), pl.nstreams * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&pl.d_slots, pl.slots.size() * sizeof(SlotDesc)));
    CUDA_TRY(cudaMemcpy(pl.d_slots, pl.slots.data(), pl.slots.size() * sizeof(SlotDesc), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&pl.d_unit_lut, pl.unit_lut.size() * 4));
    CUDA_TRY(cudaMemcpy(pl.d_unit_lut, pl.unit_lut.data(), pl.unit_lut.size() * 4, cudaMemcpyHostToDevice));
    return ICER_RESULT_OK;
}

// priority-prefix encoding: work lists of the coder.  Units are numbered big-first, so ascending stream ids = big streams first.
int upload_prefix_lists(Plan &pl, uint32_t prefix, cudaStream_t st) {
    if (pl.d_list && pl.list_prefix == prefix) return ICER_RESULT_OK;
    if (!pl.d_list) CUDA_TRY(cudaMalloc(&pl.d_list, (size_t) pl.nstreams * 4));
    std::vector<uint32_t> l(pl.prio_order);
    std::sort(l.begin(), l.begin() + prefix);
    std::sort(l.begin() + prefix, l.end());
    CUDA_TRY(cudaMemcpyAsync(pl.d_list, l.data(), (size_t) pl.nstreams * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));       // `l` is a local
    pl.list_prefix = prefix;
    return ICER_RESULT_OK;
}
