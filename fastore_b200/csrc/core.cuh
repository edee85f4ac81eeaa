// core.cuh -- shared definitions of the device path: parameters, views of the staged batch, and
// the small bit-vector toolkit the per-thread routines are written in.
//
// Everything marked FSB_HD is plain integer code that also compiles as host C++ (no CUDA headers
// needed): tests/emul/ builds those routines for the CPU and checks them against the oracle in
// the CPU-only test tier, so kernel logic is verified before it ever reaches a GPU.  The product
// library never runs them on the host -- there is no CPU fallback.
#pragma once

#include <cstdint>

#include "../../include/fastore_b200.h"

#if defined(__CUDACC__)
#define FSB_HD __host__ __device__ __forceinline__
#else
#define FSB_HD inline
#endif

namespace fsb {

struct DeviceParams
{
    uint32_t k;              // signature_len
    uint32_t s;              // skip_zone_len
    uint32_t cutoff_bits;    // signatureMaskCutoffBits
    uint32_t nbin;           // 4^k
    uint32_t kmer_mask;      // 4^k - 1
    uint32_t paired;
    uint32_t qua_method;
    uint32_t qua_offset;
    uint32_t qua_threshold;
    uint32_t qua_bits;       // 6, 1, 3, 6
    uint32_t has_headers;
    uint32_t key_bits;       // 2k + 1: bits of a signature incl. the N-bin value
};

struct BatchView
{
    const uint8_t* text[2];               // concatenated chunk texts (device)
    const fsb_record* rec[2];             // concatenated record tables (device)
    const uint64_t* chunk_text_base[2];   // [n_chunks] byte offset of each chunk's text inside text[m]
    const uint64_t* chunk_first_rec;      // [n_chunks + 1]
    uint32_t n_chunks;
    uint64_t n_records;
};

inline DeviceParams make_device_params(const fsb_params& p)
{
    DeviceParams d{};
    d.k = p.signature_len; d.s = p.skip_zone_len;
    d.cutoff_bits = p.signature_mask_cutoff_bits;
    d.nbin = 1u << (2 * d.k); d.kmer_mask = d.nbin - 1;
    d.paired = p.paired_end ? 1 : 0;
    d.qua_method = p.quality_method; d.qua_offset = p.quality_offset; d.qua_threshold = p.binary_threshold;
    const uint32_t bpb[4] = {6, 1, 3, 6};                        // QualityCompressionParams::BitsPerBase (Quality.h:58-64)
    d.qua_bits = bpb[p.quality_method & 3];
    d.has_headers = p.reads_have_headers ? 1 : 0;
    d.key_bits = 2 * d.k + 1;
    return d;
}

// chunk index of record i (records are stored chunk-major)
FSB_HD uint32_t find_chunk(const BatchView& b, uint64_t i)
{
    uint32_t lo = 0, hi = b.n_chunks;      // invariant: first[lo] <= i < first[hi]
    while (hi - lo > 1)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (b.chunk_first_rec[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- the per-record card carried through the sort -------------------------------------------------------
// K1 knows everything later stages need about a record except where it lands; instead of leaving it
// in per-record tables that would have to be gathered by sorted index (one 128-byte DRAM line per
// 16-byte entry on B200), it rides through the radix sort as the 64-bit value of the (key, value) pair:
//   bits  0..3   flags   FSB_INFO_REVERSE | SWAPPED | PLAIN_A | PLAIN_B  (>> 16)
//   bits  4..11  minimPos in the stored orientation
//   bits 12..19  length of stored mate A        bits 20..27  length of stored mate B (0 for SE)
//   bits 28..35  headLen (0 when titles are not kept)
//   bits 36..63  record index inside the batch (< 2^28)
constexpr uint32_t kMaxBatchRecords = (1u << 28) - 1u;
FSB_HD uint64_t card_make(uint32_t rec, uint32_t info, uint32_t lenA, uint32_t lenB, uint32_t H)
{
    return ((uint64_t)rec << 36) | ((uint64_t)(H & 0xFFu) << 28) | ((uint64_t)(lenB & 0xFFu) << 20) | ((uint64_t)(lenA & 0xFFu) << 12) |
           ((uint64_t)(info & 0xFFu) << 4) | (uint64_t)((info >> 16) & 0xFu);
}
FSB_HD uint32_t card_rec(uint64_t c) { return (uint32_t)(c >> 36); }
FSB_HD uint32_t card_info(uint64_t c) { return (((uint32_t)c & 0xFu) << 16) | (((uint32_t)c >> 4) & 0xFFu); }      // minimPos | FSB_INFO_*
FSB_HD uint32_t card_lenA(uint64_t c) { return ((uint32_t)c >> 12) & 0xFFu; }
FSB_HD uint32_t card_lenB(uint64_t c) { return ((uint32_t)c >> 20) & 0xFFu; }
FSB_HD uint32_t card_head(uint64_t c) { return (uint32_t)(c >> 28) & 0xFFu; }

// ---- the per-record slot K1 writes and K4 gathers -----------------------------------------------------------
// One slot per record (pair) in input order, 128-byte aligned so that a gather by sorted index only
// touches lines it uses completely.  It holds the record's contribution to the quality, title and
// DNA streams exactly as they will appear in the output (stored orientation, MSB-first):
//   [0, wqa)          quality of stored mate A, from bit 0       lenA * q bits
//   [wqa, 2 wqa)      quality of stored mate B (PE), from bit 0  lenB * q bits
//   [qw, qw + tw)     title (8 bits headLen + 7 bits/char of the mate-1 title) immediately followed by
//                     the DNA of mate A without the signature, then mate B, 2 or 3 bits per symbol
// wqa is a multiple of 4 words, so both quality regions are 16-byte aligned; qw = mates * wqa.
struct SlotGeom
{
    uint32_t wqa;            // words of one mate's quality region (multiple of 4)
    uint32_t qw, tw;         // words of the quality regions together / of the title + DNA region
    uint32_t words;          // slot stride in words, multiple of 32 (128 bytes)
};
inline SlotGeom make_slot_geom(const DeviceParams& P, uint32_t max_len, uint32_t max_head)
{
    SlotGeom g{};
    const uint32_t mates = P.paired ? 2u : 1u;
    auto words = [](uint32_t bits) { return (bits + 31u) >> 5; };
    g.wqa = (words(max_len * P.qua_bits) + 3u) & ~3u;
    g.qw = mates * g.wqa;
    g.tw = words((P.has_headers ? 8u + 7u * (max_head ? max_head - 1u : 0u) : 0u) + mates * max_len * 3u);
    g.words = (g.qw + g.tw + 31u) & ~31u;
    return g;
}

// ---- intrinsics with host equivalents --------------------------------------------------------------
FSB_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s)      // low word of (hi:lo) >> s, s in [0, 31]
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
#endif
}
FSB_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s)      // high word of (hi:lo) << s, s in [0, 31]
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s ? (hi << s) | (lo >> (32u - s)) : hi;
#endif
}
FSB_HD uint32_t popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
FSB_HD uint32_t ctz32(uint32_t x)                                   // x != 0
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__ffs((int)x) - 1);
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}
FSB_HD uint32_t clz32(uint32_t x)                                   // x != 0
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)x);
#else
    return (uint32_t)__builtin_clz(x);
#endif
}
FSB_HD uint32_t bswap32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}
FSB_HD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t sel)     // result byte i = pool[(sel >> 4i) & 7], pool = x (0..3), y (4..7)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, sel);
#else
    const uint64_t pool = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
#endif
}
FSB_HD uint32_t bit_length_u32(uint32_t x) { return x ? 32u - clz32(x) : 0u; }     // Utils.h:235-243 for x < 2^31

// ---- bit vectors of 32*NW positions: position p is bit (p & 31) of word (p >> 5) -----------------
template <int NW> struct BV { uint32_t w[NW]; };

template <int NW> FSB_HD BV<NW> bv_shr(const BV<NW>& x, uint32_t s)       // r[p] = x[p + s], s in [0, 31]
{
    BV<NW> r;
#pragma unroll
    for (int j = 0; j < NW; ++j) r.w[j] = funnel_r(x.w[j], j + 1 < NW ? x.w[j + 1] : 0u, s);
    return r;
}
template <int NW> FSB_HD BV<NW> bv_shl(const BV<NW>& x, uint32_t s)       // r[p] = x[p - s], s in [0, 31]
{
    BV<NW> r;
#pragma unroll
    for (int j = 0; j < NW; ++j) r.w[j] = funnel_l(j ? x.w[j - 1] : 0u, x.w[j], s);
    return r;
}
template <int NW> FSB_HD BV<NW> bv_and(const BV<NW>& a, const BV<NW>& b) { BV<NW> r;
#pragma unroll
    for (int j = 0; j < NW; ++j) r.w[j] = a.w[j] & b.w[j]; return r; }
template <int NW> FSB_HD BV<NW> bv_andn(const BV<NW>& a, const BV<NW>& b) { BV<NW> r;      // a & ~b
#pragma unroll
    for (int j = 0; j < NW; ++j) r.w[j] = a.w[j] & ~b.w[j]; return r; }
template <int NW> FSB_HD BV<NW> bv_or(const BV<NW>& a, const BV<NW>& b) { BV<NW> r;
#pragma unroll
    for (int j = 0; j < NW; ++j) r.w[j] = a.w[j] | b.w[j]; return r; }
template <int NW> FSB_HD bool bv_any(const BV<NW>& a) { uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) o |= a.w[j]; return o != 0; }
template <int NW> FSB_HD BV<NW> bv_select(bool c, const BV<NW>& a, const BV<NW>& b) { BV<NW> r;
#pragma unroll
    for (int j = 0; j < NW; ++j) r.w[j] = c ? a.w[j] : b.w[j]; return r; }
template <int NW> FSB_HD uint32_t bv_popc(const BV<NW>& a) { uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) o += popc32(a.w[j]); return o; }

// bits [0, hi) of one word for any hi (negative: none, 32 and more: all).  The device form leans on
// shl.b32 clamping shift counts above 31 (the result is 0).
FSB_HD uint32_t below_mask(int32_t hi)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(0xFFFFFFFFu), "r"((uint32_t)max(hi, 0)));
    return ~r;
#else
    return hi <= 0 ? 0u : (hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u));
#endif
}
// bits [a, b) set (a, b may be negative or beyond the vector)
template <int NW> FSB_HD BV<NW> bv_range(int32_t a, int32_t b)
{
    BV<NW> r;
#pragma unroll
    for (int j = 0; j < NW; ++j) r.w[j] = below_mask(b - 32 * j) & ~below_mask(a - 32 * j);
    return r;
}
// bits [0, b) set
template <int NW> FSB_HD BV<NW> bv_below(int32_t b)
{
    BV<NW> r;
#pragma unroll
    for (int j = 0; j < NW; ++j) r.w[j] = below_mask(b - 32 * j);
    return r;
}
// r[p] = OR of x[p .. p + width), width in [1, 32]
template <int NW> FSB_HD BV<NW> bv_slide_or(BV<NW> x, uint32_t width)
{
    uint32_t w = 1;
    while (2 * w <= width) { x = bv_or(x, bv_shr(x, w)); w *= 2; }
    if (w < width) x = bv_or(x, bv_shr(x, width - w));
    return x;
}
template <int NW> FSB_HD uint32_t bv_lowest(const BV<NW>& a)          // index of the lowest set bit (a != 0)
{
    uint32_t idx = 0;
#pragma unroll
    for (int j = NW - 1; j >= 0; --j) if (a.w[j]) idx = 32u * j + ctz32(a.w[j]);
    return idx;
}
template <int NW> FSB_HD uint32_t bv_highest(const BV<NW>& a)         // index of the highest set bit (a != 0)
{
    uint32_t idx = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) if (a.w[j]) idx = 32u * j + 31u - clz32(a.w[j]);
    return idx;
}

} // namespace fsb
