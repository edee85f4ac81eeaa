// signature.cuh -- K1: per-read minimizer signature on both strands (sm_100a).
//
// Replaces FastqCategorizerBase::FindMinimizer x2 (SE) / x4 (PE), FastqRecord::ComputeRC and the
// per-record selection logic of FastqCategorizerSE/PE::DistributeToBins
// (FastqCategorizer.cpp:79-106, 197-253, 256-363; FastqRecord.h:80-111).
//
// Formulation (no reverse-complement copy, no validity table):
//   * a warp owns one record (SE) or one pair (PE) at a time; lane l holds bases [8l, 8l+8) of a
//     mate as 2-bit codes (16 bits) plus an N mask (8 bits), both MSB-first;
//   * the k-mer at position p = 8l+j is a bit field of the 48-bit window built from lanes l, l+1,
//     l+2; the reverse-strand k-mer at the same forward position is the same field of the window
//     built from the lane-wise reverse-complemented codes, read from the other end;
//   * forward scan window  p in [0, L-k-s)          (FastqCategorizer.cpp:88),
//     reverse scan window  q = L-k-p in [0, L-k-s)  <=> p in (s, L-k];
//   * "first minimum wins" (strict <, :93) becomes a min over keys (kmer << 8) | position, with
//     the reverse strand's position measured in its own orientation;
//   * validBinSignatures[m] (:34-76) is closed-form bit logic on m;
//   * the N >= L/3 filter (:102) uses the popcount of the N masks.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/fastore_b200.h"

namespace fsb {

struct DeviceParams
{
    uint32_t k;              // signature_len
    uint32_t s;              // skip_zone_len
    uint32_t cutoff_mask;    // (1 << signatureMaskCutoffBits) - 1
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
    const uint8_t* text[2];          // concatenated chunk texts (device)
    const fsb_record* rec[2];        // concatenated record tables (device)
    const uint64_t* chunk_text_base[2];   // [n_chunks] byte offset of each chunk's text inside text[m]
    const uint64_t* chunk_first_rec; // [n_chunks + 1]
    uint32_t n_chunks;
    uint64_t n_records;
};

// chunk index of record i (records are stored chunk-major)
__device__ __forceinline__ uint32_t find_chunk(const BatchView& b, uint64_t i)
{
    uint32_t lo = 0, hi = b.n_chunks;      // invariant: first[lo] <= i < first[hi]
    while (hi - lo > 1)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (b.chunk_first_rec[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// validBinSignatures[m] as closed form (InitializeValidBinSignatures, FastqCategorizer.cpp:34-76):
// invalid iff low cutoff bits set, or the top three symbols are AAA / AAC, or an "AA" sits at
// symbol pair (p, p+1) for some p >= 1 (the loop tests (m >> 2j) & 0xF == 0 for j in [0, k-3]).
__device__ __forceinline__ bool signature_valid(uint32_t m, const DeviceParams& P)
{
    const uint32_t isA = ~(m | (m >> 1)) & 0x55555555u;                 // bit 2i set <=> symbol i (from LSB) is A
    const uint32_t pairs = isA & (isA >> 2);                             // bit 2i set <=> symbols i and i+1 are A
    const uint32_t range = (P.k >= 2) ? ((1u << (2 * (P.k - 2))) - 1) : 0;   // j in [0, k-3]
    return ((m & P.cutoff_mask) == 0) & ((m >> (2 * P.k - 6)) >= 2u) & ((pairs & range) == 0);
}

// 4 ASCII bases (little-endian word, first base in the low byte) -> 8 bits of 2-bit codes,
// first base in the top two bits.  A,C,G,T -> 0,1,2,3 ; N -> (garbage, masked by the N flag).
__device__ __forceinline__ uint32_t pack4_codes(uint32_t w)
{
    const uint32_t x = (w >> 1) & 0x03030303u;                 // A=0 C=1 G=3 T=2
    const uint32_t c = x ^ ((x >> 1) & 0x01010101u);           // A=0 C=1 G=2 T=3
    return (c * 0x40100401u) >> 24;                            // b0<<6 | b1<<4 | b2<<2 | b3
}
// N flags of 4 bases -> 4 bits, first base in the top bit ('N' = 0x4E is the only symbol with bit 3)
__device__ __forceinline__ uint32_t pack4_nflags(uint32_t w)
{
    const uint32_t f = (w >> 3) & 0x01010101u;
    return (f * 0x08040201u) >> 24 & 0xFu;                     // b0<<3 | b1<<2 | b2<<1 | b3
}

// reverse the order of the eight 2-bit groups of a 16-bit value and complement them
__device__ __forceinline__ uint32_t revcomp16(uint32_t c)
{
    uint32_t r = __brev(c) >> 16;                                         // bit reversal of 16 bits
    r = ((r >> 1) & 0x5555u) | ((r & 0x5555u) << 1);                       // restore bit order inside each group
    return (~r) & 0xFFFFu;
}

struct StrandMin { uint32_t sig; uint32_t pos; };

// Both FindMinimizer results of one mate: fwd = FM(x), rev = FM(rc(x)).  Whole warp participates.
// `seq` points at the first base (any alignment), L <= 255.  nN receives the N count of the mate.
__device__ __forceinline__ void scan_mate(const uint8_t* __restrict__ seq, uint32_t L, const DeviceParams& P,
                                          StrandMin& fwd, StrandMin& rev, uint32_t& nN)
{
    const unsigned lane = threadIdx.x & 31;
    // ---- load 8 bases per lane with two aligned 8-byte loads + byte shift -------------------------
    const uint64_t addr = (uint64_t)seq + 8ull * lane;
    const uint64_t a8 = addr & ~7ull;
    const uint32_t sh = (uint32_t)(addr & 7u) * 8u;
    uint64_t w = 0;
    if (8u * lane < L)
    {
        const uint64_t lo = __ldg((const unsigned long long*)a8);
        const uint64_t hi = sh ? __ldg((const unsigned long long*)(a8 + 8)) : 0ull;
        w = sh ? ((lo >> sh) | (hi << (64u - sh))) : lo;
    }
    const uint32_t valid_bases = (8u * lane < L) ? min(8u, L - 8u * lane) : 0u;     // bases of this lane inside the read
    const uint32_t w0 = (uint32_t)w, w1 = (uint32_t)(w >> 32);
    const uint32_t inmask = valid_bases ? (0xFFu << (8u - valid_bases)) & 0xFFu : 0u;   // MSB-first mask of real bases
    uint32_t c16 = (pack4_codes(w0) << 8) | pack4_codes(w1);
    uint32_t n8 = ((pack4_nflags(w0) << 4) | pack4_nflags(w1)) & inmask;
    nN = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(n8));
    n8 |= (~inmask) & 0xFFu;                       // bases past the end behave like N (never inside a window anyway)
    const uint32_t r16 = revcomp16(c16);

    // ---- 24-base windows from lanes l, l+1, l+2 ----------------------------------------------------
    const uint32_t c1 = __shfl_down_sync(0xFFFFFFFFu, c16, 1), c2 = __shfl_down_sync(0xFFFFFFFFu, c16, 2);
    const uint32_t n1 = __shfl_down_sync(0xFFFFFFFFu, n8, 1), n2 = __shfl_down_sync(0xFFFFFFFFu, n8, 2);
    const uint32_t r1 = __shfl_down_sync(0xFFFFFFFFu, r16, 1), r2 = __shfl_down_sync(0xFFFFFFFFu, r16, 2);
    const bool has1 = lane < 31, has2 = lane < 30;
    const uint64_t W = ((uint64_t)c16 << 32) | ((uint64_t)(has1 ? c1 : 0u) << 16) | (uint64_t)(has2 ? c2 : 0u);
    const uint64_t WR = ((uint64_t)(has2 ? r2 : 0u) << 32) | ((uint64_t)(has1 ? r1 : 0u) << 16) | (uint64_t)r16;
    const uint32_t NW = (n8 << 16) | ((has1 ? n1 : 0xFFu) << 8) | (has2 ? n2 : 0xFFu);

    const uint32_t k = P.k;
    const int32_t lim = (int32_t)L - (int32_t)k - (int32_t)P.s;     // forward: p < lim ; reverse: q = L-k-p < lim
    const uint32_t nmask = (1u << k) - 1;
    uint64_t bestF = ~0ull, bestR = ~0ull;
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
        const int32_t p = (int32_t)(8 * lane) + j;
        const bool hasN = ((NW >> (24 - k - j)) & nmask) != 0;
        const int32_t q = (int32_t)L - (int32_t)k - p;
        const uint32_t mf = (uint32_t)(W >> (48 - 2 * k - 2 * j)) & P.kmer_mask;
        const uint32_t mr = (uint32_t)(WR >> (2 * j)) & P.kmer_mask;
        const bool okF = !hasN && (p < lim) && signature_valid(mf, P);
        const bool okR = !hasN && (q >= 0) && (q < lim) && signature_valid(mr, P);
        const uint64_t keyF = ((uint64_t)mf << 8) | (uint32_t)p;
        const uint64_t keyR = ((uint64_t)mr << 8) | (uint32_t)(q & 0xFF);
        if (okF && keyF < bestF) bestF = keyF;
        if (okR && keyR < bestR) bestR = keyR;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        const uint64_t oF = __shfl_xor_sync(0xFFFFFFFFu, bestF, d);
        const uint64_t oR = __shfl_xor_sync(0xFFFFFFFFu, bestR, d);
        bestF = oF < bestF ? oF : bestF;
        bestR = oR < bestR ? oR : bestR;
    }
    // filter (FastqCategorizer.cpp:102): no valid k-mer, or too many N
    const bool tooManyN = nN >= L / 3;
    if (bestF == ~0ull || tooManyN) { fwd.sig = P.nbin; fwd.pos = 0; } else { fwd.sig = (uint32_t)(bestF >> 8); fwd.pos = (uint32_t)(bestF & 0xFF); }
    if (bestR == ~0ull || tooManyN) { rev.sig = P.nbin; rev.pos = 0; } else { rev.sig = (uint32_t)(bestR >> 8); rev.pos = (uint32_t)(bestR & 0xFF); }
}

// K1: one warp per record / pair, grid-stride.  Writes the sort key (chunk : signature) and the
// per-read info word (minimPos | FSB_INFO_* flags).
__global__ void __launch_bounds__(256) signature_kernel(BatchView B, DeviceParams P, uint32_t* __restrict__ keys, uint32_t* __restrict__ info,
                                                         uint32_t* __restrict__ sig_out /* nullable: plain signature per read */)
{
    const unsigned lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp0; i < B.n_records; i += nwarps)
    {
        const uint32_t ch = find_chunk(B, i);
        const fsb_record ra = B.rec[0][i];
        const uint8_t* t0 = B.text[0] + B.chunk_text_base[0][ch];
        StrandMin f1, r2;          // FM(m1), FM(rc(m1))
        uint32_t nN1;
        scan_mate(t0 + ra.seq_off, ra.seq_len, P, f1, r2, nN1);
        uint32_t sig, pos, flags = 0;
        if (!P.paired)
        {
            // FastqCategorizerSE::DistributeToBins (:212-245): forward wins ties
            const bool reverse = !(f1.sig <= r2.sig);
            sig = reverse ? r2.sig : f1.sig;
            pos = reverse ? r2.pos : f1.pos;
            if (sig != P.nbin) flags |= reverse ? FSB_INFO_REVERSE : 0u; else pos = 0;
            if (nN1 == 0) flags |= FSB_INFO_PLAIN_A;
        }
        else
        {
            const fsb_record rb = B.rec[1][i];
            const uint8_t* t1 = B.text[1] + B.chunk_text_base[1][ch];
            StrandMin f2, r1;      // FM(m2), FM(rc(m2))
            uint32_t nN2;
            scan_mate(t1 + rb.seq_off, rb.seq_len, P, f2, r1, nN2);
            // FastqCategorizerPE::DistributeToBins (:289-305): strict < everywhere
            const bool isF1 = f1.sig < f2.sig;
            const StrandMin F = isF1 ? f1 : f2;
            const bool isR1 = r1.sig < r2.sig;
            const StrandMin R = isR1 ? r1 : r2;
            bool isRev, first;
            if (F.sig < R.sig) { sig = F.sig; pos = F.pos; isRev = false; first = isF1; }
            else { sig = R.sig; pos = R.pos; isRev = true; first = isR1; }
            bool a_is_m2 = false;
            if (sig != P.nbin)
            {
                if (isRev) flags |= FSB_INFO_REVERSE;
                if (!first) flags |= FSB_INFO_SWAPPED;
                a_is_m2 = isRev != !first;          // reversed pair is [rc(m2) | rc(m1)]; a swap exchanges the halves
            }
            else pos = 0;
            const bool plainA = (a_is_m2 ? nN2 : nN1) == 0, plainB = (a_is_m2 ? nN1 : nN2) == 0;
            if (plainA) flags |= FSB_INFO_PLAIN_A;
            if (plainB) flags |= FSB_INFO_PLAIN_B;
        }
        if (lane == 0)
        {
            keys[i] = (ch << P.key_bits) | sig;
            info[i] = pos | flags;
            if (sig_out) sig_out[i] = sig;
        }
    }
}

} // namespace fsb
