// pack_core.cuh -- per-thread bit packing of one stored mate (and of a title / meta fields).
//
// Replaces IFastqPacker::StoreNextRecord / StoreDna / StoreQuality / StoreHeader and the
// BitMemoryWriter they append to (FastqPacker.cpp:113-287; BitMemory.h:216-433).  The reference
// pushes symbol after symbol into a sequential MSB-first bit writer.  Here the bit offset of every
// record in every stream is known beforehand (layout.cuh), so one thread packs one stored mate
// independently: it turns four ASCII symbols at a time into 8 / 12 / 24 output bits with word-wide
// arithmetic, assembles 32-bit stream words, shifts them to the record's bit phase and stores them
// into the tile's staging buffer.  Only the first and last word of a segment, which are shared
// with the neighbouring records, are merged with OR; everything in between is a plain store.
//
// FSB_HD code: also compiled for the host by tests/emul/ (CPU-tier check against the oracle).
#pragma once

#include "core.cuh"

namespace fsb {

// ---- word-level helpers with host equivalents ----------------------------------------------------------
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
FSB_HD void or_word(uint32_t* p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}

// ---- output: one thread's bit segment inside a word buffer --------------------------------------------
// The segment is `nbits` bits starting at bit `off` of `words` (bit 0 = most significant bit of
// word 0; words are kept in big-endian *bit* order and byte-swapped when they leave for memory).
// The producer pushes its stream as consecutive 32-bit words; the sink shifts them by the phase.
struct BitSink
{
    uint32_t* words;
    uint32_t idx;        // next output word
    uint32_t last;       // last output word of the segment
    uint32_t phi;        // off & 31
    uint32_t prev;       // previous stream word
    uint32_t nx;         // stream words still to come
    uint32_t lastmask;   // valid bits of the final stream word
    bool head_shared, tail_shared;
};

FSB_HD BitSink sink_open(uint32_t* words, uint32_t off, uint32_t nbits)
{
    BitSink s;
    s.words = words;
    s.idx = off >> 5;
    s.phi = off & 31u;
    s.prev = 0;
    s.nx = (nbits + 31u) >> 5;
    const uint32_t tail = nbits & 31u;
    s.lastmask = tail ? ~(0xFFFFFFFFu >> tail) : 0xFFFFFFFFu;
    const uint32_t end = off + nbits;                       // nbits > 0
    s.last = (end - 1u) >> 5;
    s.head_shared = s.phi != 0;
    s.tail_shared = (end & 31u) != 0;
    return s;
}
FSB_HD void sink_emit(BitSink& s, uint32_t v)
{
    if (s.idx > s.last) return;
    uint32_t* p = s.words + s.idx;
    if ((s.idx == s.last && s.tail_shared) || s.head_shared) or_word(p, v); else *p = v;
    s.head_shared = false;                                   // only the first emitted word starts mid-word
    s.idx++;
}
// next 32 bits of the stream (bits past the end of the segment may hold anything in the last word)
FSB_HD void sink_push(BitSink& s, uint32_t x)
{
    if (s.nx == 0) return;
    if (--s.nx == 0) x &= s.lastmask;
    sink_emit(s, s.phi ? funnel_r(x, s.prev, s.phi) : x);
    s.prev = x;
}
FSB_HD void sink_close(BitSink& s)                            // the bits of the last stream word that spilled over
{
    if (s.phi) sink_emit(s, s.prev << (32u - s.phi));
}

// small right-aligned values (meta fields): OR `nbits` (1..32) bits of v at bit offset off
FSB_HD void or_bits(uint32_t* words, uint32_t off, uint32_t v, uint32_t nbits)
{
    const uint32_t w = off >> 5, rel = off & 31u;
    const int32_t sh = 32 - (int32_t)rel - (int32_t)nbits;
    if (sh >= 0) or_word(words + w, v << sh);
    else
    {
        or_word(words + w, v >> (-sh));
        or_word(words + w + 1, v << (32 + sh));
    }
}

// ---- input: the stored symbols of one mate, four at a time ----------------------------------------------
// `w` is a word pointer, `addr` the byte offset of the mate's first source byte from it.  Stored
// order is the source order, or the reverse of it for a reversed read (FastqRecord::ComputeRC,
// FastqRecord.h:80-111: the quality is reversed alongside).  next() returns stored symbols
// 4j .. 4j+3 with the first one in the most significant byte; one PRMT does the unaligned
// extraction and the byte order for both directions.
struct SymReader
{
    const uint32_t* p;
    int32_t step;
    uint32_t sel;
    uint32_t carry;
};
FSB_HD SymReader reader_open(const uint32_t* w, uint32_t addr, uint32_t len, bool rev)
{
    SymReader r;
    if (!rev)
    {
        const uint32_t o = addr & 3u;
        r.p = w + (addr >> 2);
        r.step = 1;
        r.sel = ((o + 3u) & 7u) | (((o + 2u) & 7u) << 4) | (((o + 1u) & 7u) << 8) | (o << 12);      // result byte 3-t = pool[o + t]
    }
    else
    {
        const uint32_t e = addr + len, o = e & 3u;
        r.p = w + (e >> 2);
        r.step = -1;
        r.sel = ((4u + o) & 7u) | (((5u + o) & 7u) << 4) | (((6u + o) & 7u) << 8) | (((7u + o) & 7u) << 12);   // result byte t = pool[(4 + o + t) & 7]
    }
    r.carry = *r.p;
    return r;
}
FSB_HD uint32_t reader_next(SymReader& r)
{
    r.p += r.step;
    const uint32_t n = *r.p;
    const uint32_t x = byte_perm(r.carry, n, r.sel);
    r.carry = n;
    return x;
}

// ---- symbol coding --------------------------------------------------------------------------------------
// four ASCII bases (first in the top byte) -> their 2-bit codes per byte: A,C,G,T -> 0,1,2,3 (dnaToIdx, FastqPacker.cpp:24-30)
FSB_HD uint32_t base_codes4(uint32_t b)
{
    const uint32_t x = (b >> 1) & 0x03030303u;                   // A 0, C 1, G 3, T 2
    return x ^ ((x >> 1) & 0x01010101u);
}
// one code per byte -> 8 bits, first symbol in the top two bits.  Byte j lands at 24 + 2j; cross
// terms fall below bit 24 or above bit 31.
FSB_HD uint32_t gather4x2(uint32_t c) { return (c * 0x01041040u) >> 24; }
// one 3-bit value per byte -> 12 bits, first symbol on top
FSB_HD uint32_t gather4x3(uint32_t v)
{
    const uint32_t y = (v & 0x00070007u) | ((v & 0x07000700u) >> 5);      // per half: [s_even s_odd] in 6 bits
    return (y & 0x3Fu) | ((y >> 10) & 0xFC0u);
}
// one bit per byte (bit 0) -> 4 bits, first symbol on top
FSB_HD uint32_t gather4x1(uint32_t f) { return (f * 0x10204080u) >> 28; }
// one 6-bit value per byte -> 24 bits, first symbol on top
FSB_HD uint32_t gather4x6(uint32_t x)
{
    const uint32_t c = ((x & 0x3F003F00u) >> 2) | (x & 0x003F003Fu);
    return ((c & 0x0FFF0000u) >> 4) | (c & 0x00000FFFu);
}

// StoreQuality (FastqPacker.cpp:205-269): four quality bytes -> four values of P.qua_bits bits.
template <int Q>
FSB_HD uint32_t quality4(uint32_t b, uint32_t off4 /* offset * 0x01010101 */, uint32_t thr4 /* threshold * 0x01010101 */)
{
    const uint32_t d = ((b | 0x80808080u) - off4) ^ 0x80808080u;        // per byte (q - offset) mod 256, no borrow between bytes
    if (Q == 6) return gather4x6(d);                                      // c & 63
    if (Q == 1)
    {   // c >= binaryThreshold on the unsigned difference: a "negative" difference is large
        const uint32_t ge = ((((d | 0x80808080u) - thr4) | d) >> 7) & 0x01010101u;
        return gather4x1(ge);
    }
    // quaToIdx_8bin (FastqPacker.cpp:41-64) of c & 63: [0,1]->0 [2,9]->1 [10,19]->2 [20,24]->3 [25,29]->4 [30,34]->5 [35,39]->6 >=40->7
    const uint32_t y = (d & 0x3F3F3F3Fu) | 0x80808080u;
    uint32_t v = 0;
    v += ((y - 0x02020202u) >> 7) & 0x01010101u;
    v += ((y - 0x0A0A0A0Au) >> 7) & 0x01010101u;
    v += ((y - 0x14141414u) >> 7) & 0x01010101u;
    v += ((y - 0x19191919u) >> 7) & 0x01010101u;
    v += ((y - 0x1E1E1E1Eu) >> 7) & 0x01010101u;
    v += ((y - 0x23232323u) >> 7) & 0x01010101u;
    v += ((y - 0x28282828u) >> 7) & 0x01010101u;
    return gather4x3(v);
}

// ---- emission of a whole segment held in registers ------------------------------------------------------------
// X[0 .. NX) are the segment's stream words (MSB-first; bits past `nbits` may hold anything).  The
// segment goes to bit offset `off` of `words`: interior words are plain stores, the first and the
// last word -- shared with the neighbouring segments -- are ORed into the zero-initialised buffer.
template <int NX>
FSB_HD void emit_words(const uint32_t (&X)[NX], uint32_t nbits, uint32_t* words, uint32_t off)
{
    if (nbits == 0) return;
    const uint32_t phi = off & 31u;
    uint32_t* w = words + (off >> 5);
    const uint32_t end = phi + nbits;
    const uint32_t last = (end - 1u) >> 5;                       // index of the last output word
    const uint32_t tailbits = end - 32u * last;                  // 1..32 valid bits in it
    const uint32_t lastmask = tailbits >= 32u ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> tailbits);
    uint32_t prev = 0;
#pragma unroll
    for (int j = 0; j <= NX; ++j)
    {
        if ((j & 3) == 0 && (uint32_t)j > last) break;
        const uint32_t x = j < NX ? X[j] : 0u;
        const uint32_t v = funnel_r(x, prev, phi);              // stream bits [32j - phi, 32j - phi + 32)
        prev = x;
        if ((uint32_t)j == last) or_word(w + j, v & lastmask);
        else if (j == 0) or_word(w, v);
        else if ((uint32_t)j < last) w[j] = v;
    }
}

// ---- K4: move one prepacked segment to its final bit position ------------------------------------------------
// `src` (16-byte aligned) holds `nbits` bits starting at bit 0 of src[0], MSB-first; they go to bit
// offset `off` of `words`.  Interior words are plain stores, the first and the last word -- shared
// with the neighbouring segments -- are ORed into the zero-initialised buffer.  Reads whole groups
// of four source words (up to 3 words past the segment; the caller's buffer covers them).
FSB_HD void load4(const uint32_t* p, uint32_t (&x)[4])
{
#if defined(__CUDA_ARCH__)
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
#else
    for (int u = 0; u < 4; ++u) x[u] = p[u];
#endif
}
FSB_HD void shift_copy(const uint32_t* src, uint32_t nbits, uint32_t* words, uint32_t off)
{
    if (nbits == 0) return;
    const uint32_t phi = off & 31u;
    uint32_t* w = words + (off >> 5);
    const uint32_t end = phi + nbits;
    const uint32_t last = (end - 1u) >> 5;                       // index of the last output word
    const uint32_t tailbits = end - 32u * last;                  // 1..32 valid bits in it
    const uint32_t lastmask = tailbits >= 32u ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> tailbits);
    uint32_t prev = 0, j = 0, x[4];
    if (last >= 4u)
    {   // first group: output word 0 is shared with the previous segment
        load4(src, x);
        or_word(w, funnel_r(x[0], 0u, phi));
        w[1] = funnel_r(x[1], x[0], phi); w[2] = funnel_r(x[2], x[1], phi); w[3] = funnel_r(x[3], x[2], phi);
        prev = x[3];
        for (j = 4; j + 4u <= last; j += 4)                       // groups whose four output words all lie strictly inside the segment
        {
            load4(src + j, x);
            w[j] = funnel_r(x[0], prev, phi); w[j + 1] = funnel_r(x[1], x[0], phi);
            w[j + 2] = funnel_r(x[2], x[1], phi); w[j + 3] = funnel_r(x[3], x[2], phi);
            prev = x[3];
        }
    }
    // last group: output words j .. last (1 to 4 of them; bits past the segment in the source are masked off)
    load4(src + j, x);
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u)
    {
        const uint32_t jj = j + u;
        if (jj <= last)
        {
            const uint32_t v = funnel_r(x[u], prev, phi);       // stream bits [32 jj - phi, 32 jj - phi + 32)
            prev = x[u];
            if (jj == last) or_word(w + jj, v & lastmask);
            else if (jj == 0) or_word(w, v);
            else w[jj] = v;
        }
    }
}

// ---- quality stream of one stored mate --------------------------------------------------------------------
// 32 symbols -> Q stream words per round.
template <int NW, int Q>
FSB_HD void pack_quality(SymReader rd, uint32_t len, const DeviceParams& P, uint32_t* words, uint32_t off)
{
    uint32_t X[NW * Q];
    const uint32_t off4 = P.qua_offset * 0x01010101u, thr4 = P.qua_threshold * 0x01010101u;
#pragma unroll
    for (int j = 0; j < NW; ++j)
    {
        if (32u * j < len)
        {
            uint32_t t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = quality4<Q>(reader_next(rd), off4, thr4);
            if (Q == 6)
            {
                X[6 * j + 0] = (t[0] << 8) | (t[1] >> 16);
                X[6 * j + 1] = (t[1] << 16) | (t[2] >> 8);
                X[6 * j + 2] = (t[2] << 24) | t[3];
                X[6 * j + 3] = (t[4] << 8) | (t[5] >> 16);
                X[6 * j + 4] = (t[5] << 16) | (t[6] >> 8);
                X[6 * j + 5] = (t[6] << 24) | t[7];
            }
            else if (Q == 3)
            {
                X[3 * j + 0] = (t[0] << 20) | (t[1] << 8) | (t[2] >> 4);
                X[3 * j + 1] = (t[2] << 28) | (t[3] << 16) | (t[4] << 4) | (t[5] >> 8);
                X[3 * j + 2] = (t[5] << 24) | (t[6] << 12) | t[7];
            }
            else
                X[j] = (t[0] << 28) | (t[1] << 24) | (t[2] << 20) | (t[3] << 16) | (t[4] << 12) | (t[5] << 8) | (t[6] << 4) | t[7];
        }
        else
        {
#pragma unroll
            for (int u = 0; u < Q; ++u) X[Q * j + u] = 0;
        }
    }
    emit_words<NW * Q>(X, len * (uint32_t)Q, words, off);
}

// ---- DNA stream of one stored mate (StoreDna, FastqPacker.cpp:157-202) ----------------------------------
// All `len` symbols are coded into stream words first; then the k signature symbols at
// [cut_pos, cut_pos + cut_len) are cut out (they are implied by the bin) and the rest is emitted.
// SB = bits per symbol: 2 for a mate without 'N', 3 otherwise (A,C,G,T,N -> 0..4).
template <int NW, int SB>
FSB_HD void pack_dna(SymReader rd, uint32_t len, bool rev, uint32_t cut_pos, uint32_t cut_len, uint32_t* words, uint32_t off)
{
    constexpr int NX = SB * NW;                                  // stream words of 32*NW symbols
    uint32_t D[NX + 2];
    const uint32_t comp = rev ? 0x03030303u : 0u;                // rcCodes (FastqRecord.h:62-76): A<->T, C<->G, N stays
#pragma unroll
    for (int j = 0; j < NW; ++j)                                  // 32 symbols per round
    {
        if (32u * j < len)
        {
            uint32_t t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
            {
                const uint32_t b = reader_next(rd);
                const uint32_t c = base_codes4(b) ^ comp;
                if (SB == 2) t[u] = gather4x2(c);
                else
                {
                    const uint32_t f = (b >> 3) & 0x01010101u;   // 'N'
                    t[u] = gather4x3((c & ~(f * 3u)) | (f << 2));
                }
            }
            if (SB == 2)
            {
                D[2 * j] = (t[0] << 24) | (t[1] << 16) | (t[2] << 8) | t[3];
                D[2 * j + 1] = (t[4] << 24) | (t[5] << 16) | (t[6] << 8) | t[7];
            }
            else
            {
                D[3 * j] = (t[0] << 20) | (t[1] << 8) | (t[2] >> 4);
                D[3 * j + 1] = (t[2] << 28) | (t[3] << 16) | (t[4] << 4) | (t[5] >> 8);
                D[3 * j + 2] = (t[5] << 24) | (t[6] << 12) | t[7];
            }
        }
        else
        {
#pragma unroll
            for (int u = 0; u < SB; ++u) D[SB * j + u] = 0;
        }
    }
    D[NX] = 0; D[NX + 1] = 0;
    if (cut_len)
    {
        const uint32_t cb = cut_pos * (uint32_t)SB;              // bit where the cut starts
        const uint32_t cw = cut_len * (uint32_t)SB;              // bits removed (< 64)
        const uint32_t q = cw >> 5, r = cw & 31u;
#pragma unroll
        for (int j = 0; j < NX; ++j)
        {
            const uint32_t shifted = q ? funnel_l(D[j + 2], D[j + 1], r) : funnel_l(D[j + 1], D[j], r);      // stream bits 32j + cw ..
            const int32_t keep = (int32_t)cb - 32 * j;            // leading bits of this word that precede the cut
            const uint32_t m = keep <= 0 ? 0u : (keep >= 32 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> keep));
            D[j] = (D[j] & m) | (shifted & ~m);
        }
    }
    uint32_t E[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) E[j] = D[j];
    emit_words<NX>(E, (len - cut_len) * (uint32_t)SB, words, off);
}

// ---- title (StoreHeader, FastqPacker.cpp:272-287): 8 bits headLen, then 7 bits per char after '@' ---------
FSB_HD void pack_head(const uint32_t* w, uint32_t addr, uint32_t H, uint32_t* words, uint32_t off)
{
    const uint32_t nbits = 8u + 7u * (H ? H - 1u : 0u);
    BitSink s = sink_open(words, off, nbits);
    uint64_t acc = H & 0xFFu;
    uint32_t n = 8;
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(w) + addr;
    for (uint32_t i = 1; i < H; ++i)
    {
        acc = (acc << 7) | (uint64_t)(bytes[i] & 0x7Fu);
        n += 7;
        if (n >= 32) { sink_push(s, (uint32_t)(acc >> (n - 32))); n -= 32; }
    }
    if (n) sink_push(s, (uint32_t)(acc << (32 - n)));
    sink_close(s);
}

// ---- per-record framing (StoreRecords SE :734-759 / PE :815-859, StoreNextRecord :113-153) --------------
struct ReadBits { uint32_t meta, dna, qua, head; };

FSB_HD ReadBits read_bit_lengths(const DeviceParams& P, bool nbin, uint32_t info, uint32_t L1, uint32_t L2, uint32_t H,
                                 uint32_t bmin, uint32_t bmax)
{
    ReadBits b;
    const bool pe = P.paired != 0;
    const uint32_t bpl = (bmin != bmax) ? bit_length_u32(bmax - bmin) : 0;
    b.meta = (pe ? 2 * bpl : bpl) + (nbin ? 0u : (pe ? 10u : 9u)) + 1u + (pe ? 1u : 0u);
    const uint32_t bitsA = (info & FSB_INFO_PLAIN_A) ? 2u : 3u, bitsB = (info & FSB_INFO_PLAIN_B) ? 2u : 3u;
    b.dna = (L1 - (nbin ? 0u : P.k)) * bitsA + (pe ? L2 * bitsB : 0u);
    b.qua = (L1 + (pe ? L2 : 0u)) * P.qua_bits;
    b.head = P.has_headers ? 8u + 7u * (H ? H - 1u : 0u) : 0u;
    return b;
}

// the record's meta fields, most significant first, as one right-aligned value (at most 28 bits)
FSB_HD uint32_t meta_fields(const DeviceParams& P, bool nbin, uint32_t info, uint32_t lenA, uint32_t lenB, uint32_t bmin, uint32_t bmax, uint32_t& nbits)
{
    const bool pe = P.paired != 0;
    uint32_t v = 0, n = 0;
    if (bmin != bmax)
    {
        const uint32_t bpl = bit_length_u32(bmax - bmin), m = (1u << bpl) - 1u;
        v = (lenA - bmin) & m; n = bpl;                                               // rec->seqLen - minLen
        if (pe) { v = (v << bpl) | ((lenB - bmin) & m); n += bpl; }
    }
    if (!nbin)
    {
        if (pe) { v = (v << 1) | ((info & FSB_INFO_SWAPPED) ? 1u : 0u); n += 1; }
        v = (v << 1) | ((info & FSB_INFO_REVERSE) ? 1u : 0u); n += 1;
        v = (v << 8) | (info & 0xFFu); n += 8;
    }
    v = (v << 1) | ((info & FSB_INFO_PLAIN_A) ? 1u : 0u); n += 1;
    if (pe) { v = (v << 1) | ((info & FSB_INFO_PLAIN_B) ? 1u : 0u); n += 1; }
    nbits = n;
    return v;
}

} // namespace fsb
