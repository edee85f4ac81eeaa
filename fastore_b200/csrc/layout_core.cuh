// layout_core.cuh -- the framing rules of PackToBins / PackToBin / StoreRecords as one associative scan.
//
// The reference appends bin after bin, record after record, to four sequential bit writers
// (FastqPacker.cpp:417-491, 541-602, 734-759, 815-859): every bin starts on a byte boundary of every
// stream (FlushPartialWordBuffer, :593-596), the meta stream carries 17 header bits per bin (:581-583),
// and the records of a bin follow each other bit by bit.  Where a record lands therefore depends on
// everything in front of it -- but only through a function of a very small family.  Walking over a run
// of sorted records moves a stream position x to
//
//     x + a                        if no bin starts inside the run, or
//     roundup8(x + a) + b          if one does  (a: bits in front of the first bin start of the run,
//                                                b: everything after that byte boundary; later byte
//                                                boundaries are relative to the first one),
//
// and these functions compose into functions of the same form.  So the positions of all records, the
// byte sizes of all bins and the record counts / raw sizes of the bin descriptors come out of ONE
// exclusive scan over the sorted records with LayState as the element type (three streaming kernels,
// layout_fused.cuh) instead of flag / scan / per-bin statistics / bit-length / scan / bin-size / scan /
// placement passes with a dozen launches.
//
// This form needs the bit length of a record to be known from the record alone.  The one framing rule that
// looks at the whole bin -- the length field of bins whose records differ in length (bitsPerLen,
// FastqPacker.cpp:553-566) -- is zero bits wide when all reads of the batch have one length, which the
// staging statistics tell; batches with reads of different lengths take the general kernels of layout.cuh.
//
// FSB_HD code: tests/emul/ runs it on the host against the oracle (CPU tier).
#pragma once

#include "core.cuh"
#include "pack_core.cuh"

namespace fsb {

struct LayState
{
    uint32_t has_start;      // a bin starts inside the run
    uint32_t nstarts;        // bins that start inside the run
    uint32_t cnt;            // records since the last bin start (all records of the run if there is none)
    uint32_t pad;
    uint64_t raw_dna, raw_head;   // raw sizes since the last bin start (the whole run if there is none)
    uint64_t a[4];           // bits in front of the first bin start (all bits of the run if there is none)
    uint64_t b[4];           // has_start: position at the end of the run - roundup8(x + a)
    uint64_t ls[4];          // has_start: position of the last bin start - roundup8(x + a)   (a multiple of 8)
};

FSB_HD uint64_t roundup8(uint64_t x) { return (x + 7u) & ~7ull; }

FSB_HD LayState lay_identity()
{
    LayState s;
    s.has_start = 0; s.nstarts = 0; s.cnt = 0; s.pad = 0; s.raw_dna = 0; s.raw_head = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { s.a[k] = 0; s.b[k] = 0; s.ls[k] = 0; }
    return s;
}

// the run A followed by the run B
FSB_HD LayState lay_combine(const LayState& A, const LayState& B)
{
    LayState r;
    r.pad = 0;
    r.nstarts = A.nstarts + B.nstarts;
    if (!B.has_start)
    {
        r.has_start = A.has_start;
        r.cnt = A.cnt + B.cnt; r.raw_dna = A.raw_dna + B.raw_dna; r.raw_head = A.raw_head + B.raw_head;
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            r.a[k] = A.has_start ? A.a[k] : A.a[k] + B.a[k];
            r.b[k] = A.has_start ? A.b[k] + B.a[k] : 0;
            r.ls[k] = A.ls[k];
        }
    }
    else
    {
        r.has_start = 1;
        r.cnt = B.cnt; r.raw_dna = B.raw_dna; r.raw_head = B.raw_head;
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            if (A.has_start)
            {   // B's first byte boundary, counted from A's: roundup8(R_A + A.b + B.a) = R_A + roundup8(A.b + B.a)
                const uint64_t rb = roundup8(A.b[k] + B.a[k]);
                r.a[k] = A.a[k]; r.b[k] = rb + B.b[k]; r.ls[k] = rb + B.ls[k];
            }
            else { r.a[k] = A.a[k] + B.a[k]; r.b[k] = B.b[k]; r.ls[k] = B.ls[k]; }
        }
    }
    return r;
}

// One sorted record as the scan sees it.
struct LayRec
{
    uint32_t start;          // opens its bin: first record, or its key (chunk : signature) differs from the one in front
    uint32_t chunk_start;    // opens its chunk
    uint32_t nbin;           // lies in the N-bin
    uint32_t bits[4];        // meta, dna, qua, head bits (pack_core.cuh: read_bit_lengths)
    uint32_t raw_dna, raw_head;
};
FSB_HD LayRec lay_record(const DeviceParams& P, bool first, uint32_t key, uint32_t prev_key, uint64_t card, uint32_t uniform_len)
{
    LayRec r;
    r.start = (first || key != prev_key) ? 1u : 0u;
    r.chunk_start = (first || (key >> P.key_bits) != (prev_key >> P.key_bits)) ? 1u : 0u;
    r.nbin = ((key & ((1u << P.key_bits) - 1u)) == P.nbin) ? 1u : 0u;
    const uint32_t lenA = card_lenA(card), lenB = card_lenB(card), H = card_head(card);
    const ReadBits rb = read_bit_lengths(P, r.nbin != 0, card_info(card), lenA, lenB, H, uniform_len, uniform_len);
    r.bits[0] = rb.meta; r.bits[1] = rb.dna; r.bits[2] = rb.qua; r.bits[3] = rb.head;
    r.raw_dna = lenA + lenB;
    r.raw_head = P.has_headers ? H : 0u;
    return r;
}
// append one record to a run
FSB_HD void lay_push(LayState& s, const LayRec& r)
{
    if (r.start)
    {
        s.nstarts++;
        s.cnt = 1; s.raw_dna = r.raw_dna; s.raw_head = r.raw_head;
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const uint64_t hdr = k == 0 ? 17u : 0u;                 // PackToBin: 8 bits minLen, 8 bits maxLen, 1 bit hasReadGroups
            if (s.has_start) { const uint64_t rb = roundup8(s.b[k]); s.ls[k] = rb; s.b[k] = rb + hdr + r.bits[k]; }
            else { s.ls[k] = 0; s.b[k] = hdr + r.bits[k]; }
        }
        s.has_start = 1;
    }
    else
    {
        s.cnt++; s.raw_dna += r.raw_dna; s.raw_head += r.raw_head;
#pragma unroll
        for (int k = 0; k < 4; ++k) { if (s.has_start) s.b[k] += r.bits[k]; else s.a[k] += r.bits[k]; }
    }
}

// The absolute walk: what a thread knows in front of its first record (the exclusive prefix of the scan,
// taken from stream position 0 -- the first record of a batch always opens a bin) and how a record moves it.
struct LayCursor
{
    uint64_t pos[4];         // next free bit of every stream
    uint64_t ls[4];          // first bit of the open bin
    uint32_t nb;             // bins opened so far
    uint32_t cnt;            // records of the open bin so far
    uint64_t raw_dna, raw_head;
};
FSB_HD LayCursor lay_cursor(const LayState& excl)
{
    LayCursor c;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        const uint64_t r0 = roundup8(excl.a[k]);                    // the run in front starts at stream position 0
        c.pos[k] = excl.has_start ? r0 + excl.b[k] : excl.a[k];
        c.ls[k] = excl.has_start ? r0 + excl.ls[k] : 0;
    }
    c.nb = excl.nstarts; c.cnt = excl.cnt; c.raw_dna = excl.raw_dna; c.raw_head = excl.raw_head;
    return c;
}
// Step over one record: `at[k]` receives the record's first bit in stream k.
FSB_HD void lay_step(LayCursor& c, const LayRec& r, uint64_t (&at)[4])
{
    if (r.start)
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) { c.pos[k] = roundup8(c.pos[k]); c.ls[k] = c.pos[k]; }
        c.pos[0] += 17;
        c.nb++; c.cnt = 0; c.raw_dna = 0; c.raw_head = 0;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { at[k] = c.pos[k]; c.pos[k] += r.bits[k]; }
    c.cnt++; c.raw_dna += r.raw_dna; c.raw_head += r.raw_head;
}
// the descriptor of the open bin once its last record has been stepped over
FSB_HD fsb_bin_descriptor lay_descriptor(const LayCursor& c, uint32_t signature)
{
    fsb_bin_descriptor d;
    d.signature = signature;
    d.meta_size = (roundup8(c.pos[0]) - c.ls[0]) >> 3;
    d.dna_size = (roundup8(c.pos[1]) - c.ls[1]) >> 3;
    d.qua_size = (roundup8(c.pos[2]) - c.ls[2]) >> 3;
    d.head_size = (roundup8(c.pos[3]) - c.ls[3]) >> 3;
    d.records_count = c.cnt;
    d.raw_dna_size = c.raw_dna;
    d.raw_head_size = c.raw_head;
    return d;
}

} // namespace fsb
