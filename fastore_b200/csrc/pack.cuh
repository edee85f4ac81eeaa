// pack.cuh -- K4: bit-pack every record into the four streams at its scanned bit offset (sm_100a).
//
// Replaces FastqRecordsPackerSE/PE::StoreRecords, IFastqPacker::StoreNextRecord / StoreDna /
// StoreQuality / StoreHeader and BitMemoryWriter (FastqPacker.cpp:113-287, 734-759, 815-859;
// BitMemory.h:216-433).  The reference appends records one after the other to four sequential
// MSB-first bit writers; here every record's bit offset in every stream is already known
// (layout.cuh), so records are written independently: a warp owns one record, each lane assembles
// whole 32-bit output words (big-endian bit order, byte-swapped on store), interior words are
// plain stores and only the first/last word of a segment - shared with the neighbouring record -
// is merged with atomicOr into zero-initialised memory.
#pragma once

#include "layout.cuh"

namespace fsb {

struct OutStreams
{
    uint32_t* w[4];          // meta, dna, qua, head as 32-bit words (zero-initialised)
};



// OR `nbits` (<= 32) bits of `value` into the stream at absolute bit offset `off`
__device__ __forceinline__ void or_bits(uint32_t* __restrict__ words, uint64_t off, uint32_t value, uint32_t nbits)
{
    if (nbits == 0) return;
    const uint64_t w = off >> 5;
    const uint32_t rel = (uint32_t)(off & 31);
    const int32_t sh = 32 - (int32_t)rel - (int32_t)nbits;
    if (sh >= 0) atomicOr(&words[w], bswap32(value << sh));
    else
    {
        atomicOr(&words[w], bswap32(value >> (-sh)));
        atomicOr(&words[w + 1], bswap32(value << (32 + sh)));
    }
}

// One stored mate: where its symbols come from.
struct MateSrc
{
    const uint8_t* seq;      // first base of the source mate in the chunk text
    const uint8_t* qua;
    uint32_t len;
    bool rev;                // stored = reverse complement of the source (quality reversed)
};

__device__ __forceinline__ uint32_t dna_code(uint8_t c)        // A,C,G,T,N -> 0..4 (dnaToIdx, FastqPacker.cpp:24-30)
{
    const uint32_t x = (c >> 1) & 3u;
    return (c == 'N') ? 4u : (x ^ (x >> 1));
}
__device__ __forceinline__ uint32_t stored_base(const MateSrc& m, uint32_t t)
{
    if (!m.rev) return dna_code(__ldg(m.seq + t));
    const uint32_t c = dna_code(__ldg(m.seq + (m.len - 1 - t)));
    return c == 4u ? 4u : 3u - c;                                // rcCodes, FastqRecord.h:62-76
}
__device__ __forceinline__ uint32_t stored_qual(const MateSrc& m, uint32_t t, const DeviceParams& P)
{
    const uint32_t c = (uint32_t)__ldg(m.qua + (m.rev ? (m.len - 1 - t) : t)) - P.qua_offset;
    switch (P.qua_method)                                        // StoreQuality, FastqPacker.cpp:205-269
    {
    case FSB_QUA_BINARY: return c >= P.qua_threshold ? 1u : 0u;
    case FSB_QUA_8BIN:
    {   // quaToIdx_8bin (FastqPacker.cpp:41-64): [0,1]->0 [2,9]->1 [10,19]->2 [20,24]->3 [25,29]->4 [30,34]->5 [35,39]->6 >=40->7
        const uint32_t q = c & 63u;
        return (q >= 2u) + (q >= 10u) + (q >= 20u) + (q >= 25u) + (q >= 30u) + (q >= 35u) + (q >= 40u);
    }
    default: return c & 63u;
    }
}

// Emit `count` symbols of `bits` bits each starting at absolute bit offset `off`; sym(i) gives
// symbol i.  Lanes own whole output words.  Whole warp participates.
template <typename SymFn>
__device__ __forceinline__ void emit_segment(uint32_t* __restrict__ words, uint64_t off, uint32_t count, uint32_t bits, SymFn sym)
{
    if (count == 0) return;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t end = off + (uint64_t)count * bits;
    const uint64_t w0 = off >> 5, w1 = (end - 1) >> 5;
    for (uint64_t w = w0 + lane; w <= w1; w += 32)
    {
        const uint64_t lo = max(off, w << 5), hi = min(end, (w << 5) + 32);
        const uint32_t i0 = (uint32_t)((lo - off) / bits), i1 = (uint32_t)((hi - 1 - off) / bits);
        uint32_t acc = 0;
        for (uint32_t i = i0; i <= i1; ++i)
        {
            const uint32_t v = sym(i);
            const int64_t rel = (int64_t)(off + (uint64_t)i * bits) - (int64_t)(w << 5);    // may be negative
            const int32_t sh = 32 - (int32_t)rel - (int32_t)bits;
            acc |= (sh >= 0) ? ((sh < 32) ? (v << sh) : 0u) : (v >> (-sh));
        }
        const bool full = (lo == (w << 5)) && (hi == (w << 5) + 32);
        if (full) words[w] = bswap32(acc);
        else atomicOr(&words[w], bswap32(acc));
    }
}

struct PackArgs
{
    BatchView B;
    DeviceParams P;
    SortedView S;
    BinArrays A;
    StreamScans SC;
    BinOffsets BO;
    OutStreams O;
};

__global__ void __launch_bounds__(256) pack_kernel(PackArgs a)
{
    const DeviceParams& P = a.P;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const bool pe = P.paired != 0;
    for (uint64_t i = warp0; i < a.B.n_records; i += nwarps)
    {
        const uint32_t r = a.S.perm[i];
        const uint32_t key = a.S.skeys[i];
        const uint32_t ch = key >> P.key_bits;
        const bool nbin = (key & ((1u << P.key_bits) - 1)) == P.nbin;
        const uint32_t info = a.S.info[r];
        const uint32_t bin = a.A.bin_of[i];
        const uint64_t start = a.A.bin_start[bin];
        const uint32_t bmin = a.A.bin_min[bin], bmax = a.A.bin_max[bin];
        uint64_t off[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) off[s] = 8ull * a.BO.B[s][bin] + (a.SC.P[s][i] - a.SC.P[s][start]);
        off[0] += 17;

        const fsb_record ra = a.B.rec[0][r];
        const uint8_t* t0 = a.B.text[0] + a.B.chunk_text_base[0][ch];
        MateSrc m1{t0 + ra.seq_off, t0 + ra.qua_off, ra.seq_len, false}, m2{nullptr, nullptr, 0, false};
        if (pe)
        {
            const fsb_record rb = a.B.rec[1][r];
            const uint8_t* t1 = a.B.text[1] + a.B.chunk_text_base[1][ch];
            m2 = MateSrc{t1 + rb.seq_off, t1 + rb.qua_off, rb.seq_len, false};
        }
        const bool rev = (info & FSB_INFO_REVERSE) != 0, swp = (info & FSB_INFO_SWAPPED) != 0;
        // stored pair: forward [m1|m2]; reversed [rc(m2)|rc(m1)]; a swap exchanges the halves
        const bool a_is_m2 = pe && (rev != swp);
        MateSrc A = a_is_m2 ? m2 : m1, Bm = a_is_m2 ? m1 : m2;
        A.rev = rev; Bm.rev = rev;
        const uint32_t pos = info & FSB_INFO_POS_MASK;
        const uint32_t sfx = nbin ? 0u : P.k;

        // ---- meta ---------------------------------------------------------------------------------
        if (lane == 0)
        {
            if (i == start)
            {   // PackToBin header (FastqPacker.cpp:581-583)
                const uint64_t hb = 8ull * a.BO.B[0][bin];
                or_bits(a.O.w[0], hb, bmin & 0xFFu, 8);
                or_bits(a.O.w[0], hb + 8, bmax & 0xFFu, 8);
                // hasReadGroups bit = 0
            }
            uint64_t o = off[0];
            if (bmin != bmax)
            {
                const uint32_t bpl = bit_length_u32(bmax - bmin);
                or_bits(a.O.w[0], o, (A.len - bmin) & ((1u << bpl) - 1), bpl); o += bpl;       // rec->seqLen - minLen
                if (pe) { or_bits(a.O.w[0], o, (Bm.len - bmin) & ((1u << bpl) - 1), bpl); o += bpl; }
            }
            if (!nbin)
            {
                if (pe) { or_bits(a.O.w[0], o, swp ? 1u : 0u, 1); o += 1; }
                or_bits(a.O.w[0], o, rev ? 1u : 0u, 1); o += 1;
                or_bits(a.O.w[0], o, pos & 0xFFu, 8); o += 8;
            }
            or_bits(a.O.w[0], o, (info & FSB_INFO_PLAIN_A) ? 1u : 0u, 1); o += 1;
            if (pe) { or_bits(a.O.w[0], o, (info & FSB_INFO_PLAIN_B) ? 1u : 0u, 1); o += 1; }
        }

        // ---- dna ----------------------------------------------------------------------------------
        {
            const uint32_t bitsA = (info & FSB_INFO_PLAIN_A) ? 2u : 3u;
            const uint32_t cntA = A.len - sfx;
            emit_segment(a.O.w[1], off[1], cntA, bitsA, [&](uint32_t t) { return stored_base(A, t < pos ? t : t + sfx); });
            if (pe)
            {
                const uint32_t bitsB = (info & FSB_INFO_PLAIN_B) ? 2u : 3u;
                emit_segment(a.O.w[1], off[1] + (uint64_t)cntA * bitsA, Bm.len, bitsB, [&](uint32_t t) { return stored_base(Bm, t); });
            }
        }
        // ---- qua ----------------------------------------------------------------------------------
        emit_segment(a.O.w[2], off[2], A.len + (pe ? Bm.len : 0u), P.qua_bits,
                     [&](uint32_t t) { return t < A.len ? stored_qual(A, t, P) : stored_qual(Bm, t - A.len, P); });
        // ---- head ---------------------------------------------------------------------------------
        if (P.has_headers)
        {
            const uint32_t H = ra.head_len;
            const uint8_t* hp = t0 + ra.head_off;
            if (lane == 0) or_bits(a.O.w[3], off[3], H, 8);
            if (H > 1) emit_segment(a.O.w[3], off[3] + 8, H - 1, 7, [&](uint32_t t) { return (uint32_t)__ldg(hp + 1 + t) & 0x7Fu; });
        }
    }
}

} // namespace fsb
