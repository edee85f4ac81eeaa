// sig_core.cuh -- per-thread minimizer search of one mate, bit-parallel over the whole read.
//
// Replaces FastqCategorizerBase::FindMinimizer on both strands of one mate
// (FastqCategorizer.cpp:79-106), ComputeMinimizer (:138-152), the validBinSignatures table
// (:34-76) and FastqRecord::ComputeRC (FastqRecord.h:80-111).
//
// One thread owns one mate.  The read is held as three bit planes of 32*NW positions (hi and lo
// bit of the 2-bit base code A,C,G,T = 0..3, and an 'N' plane), so every test the reference makes
// per k-mer becomes a handful of word operations for all positions at once:
//
//   * validBinSignatures[m]: m is invalid iff its top three symbols are AAA / AAC, or an "AA" sits
//     at symbols (t, t+1) for some t in [1, k-2], or one of its low cutoff bits is set.  With
//     AA[p] = isA[p] & isA[p+1]: the k-mer at p is invalid iff AA hits [p+1, p+k-2] (sliding OR)
//     or AA[p] and base[p+2] is A or C.  The reverse-strand k-mer at the same forward position is
//     the reverse complement, so the same rules read TT instead of AA from the other end.
//   * k-mers containing 'N' (ComputeMinimizer returns 4^k): sliding OR of the N plane over k.
//   * scan windows: forward i in [0, L-k-s)  (:88); reverse strand position q = L-k-p in the same
//     range, i.e. forward p in (s, L-k].
//   * the minimum itself is a radix descent over the candidate set: for each of the 2k key bits,
//     most significant first, keep the candidates whose bit is 0 if there are any.  What survives
//     all share the smallest k-mer; "first minimum wins" (strict <, :93) picks the lowest forward
//     position on the forward strand and the highest on the reverse strand (lowest q).
//   * N filter (:102): popcount of the N plane >= L/3 -> (4^k, 0).
#pragma once

#include "core.cuh"

namespace fsb {

struct StrandMin { uint32_t sig; uint32_t pos; };

// multipliers that gather one bit per byte of two masked words into the top byte of the product:
// bits 8b+beta of w0 and 8b+beta+4 of (w1 << 4) land at 24+b and 28+b; every cross term falls
// below bit 24 or above bit 31 on a position of its own, so nothing carries into the result.
constexpr uint32_t kGatherBit1 = (1u << 23) | (1u << 16) | (1u << 9) | (1u << 2);
constexpr uint32_t kGatherBit2 = (1u << 22) | (1u << 15) | (1u << 8) | (1u << 1);
constexpr uint32_t kGatherBit3 = (1u << 21) | (1u << 14) | (1u << 7) | (1u << 0);

// ASCII bases -> bit planes.  `w` points at the 4-byte-aligned word holding base 0 (8*NW + 1
// words readable), `bshift` = 8 * (address of base 0 & 3).  Planes beyond the read length hold
// whatever follows the read in the text; callers mask with the length.
//   hi plane = ASCII bit 2 (A 0x41, C 0x43 -> 0;  G 0x47, T 0x54 -> 1)
//   lo plane = ASCII bit 1 ^ bit 2 (A 0, C 1, G 1^1 = 0, T 0^1 = 1)
//   N  plane = ASCII bit 3 ('N' = 0x4E is the only symbol of the alphabet with it)
// The loop over the plane words is kept rolled (K1 is bound by instruction fetch, not by issue
// slots): every round produces the next word of each plane, and the planes move through their
// registers like a shift register, so no register array is indexed dynamically.
template <int NW>
FSB_HD void ascii_to_planes(const uint32_t* w, uint32_t bshift, BV<NW>& H, BV<NW>& Lo, BV<NW>& Nm)
{
    uint32_t prev = w[0];
    const uint32_t* p = w + 1;
#pragma unroll
    for (int i = 0; i < NW; ++i) { H.w[i] = 0; Lo.w[i] = 0; Nm.w[i] = 0; }
#pragma unroll 1
    for (int j = 0; j < NW; ++j)
    {
        uint32_t h = 0, b1 = 0, n = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            const uint32_t a = p[2 * q], b = p[2 * q + 1];
            const uint32_t w0 = funnel_r(prev, a, bshift);      // bases 32j + 8q .. +3
            const uint32_t w1 = funnel_r(a, b, bshift);         // bases 32j + 8q + 4 .. +7
            prev = b;
            // ASCII bits 1..3 of eight bases in one word: the first four in the low nibbles, the next four in the high ones
            const uint32_t t = (w0 & 0x0E0E0E0Eu) | ((w1 * 16u) & 0xE0E0E0E0u);
            // the gathered byte enters at the top, the earlier ones move down: after four steps byte q holds bases 8q .. 8q+7
            h = byte_perm(h, (t & 0x44444444u) * kGatherBit2, 0x7321u);
            b1 = byte_perm(b1, (t & 0x22222222u) * kGatherBit1, 0x7321u);
            n = byte_perm(n, (t & 0x88888888u) * kGatherBit3, 0x7321u);
        }
        p += 8;
#pragma unroll
        for (int i = 0; i + 1 < NW; ++i) { H.w[i] = H.w[i + 1]; Lo.w[i] = Lo.w[i + 1]; Nm.w[i] = Nm.w[i + 1]; }
        H.w[NW - 1] = h; Lo.w[NW - 1] = b1 ^ h; Nm.w[NW - 1] = n;
    }
}

// Candidate positions of both strands (forward coordinates): in the scan window, no 'N' in the
// k-mer, signature valid.
template <int NW>
FSB_HD void candidate_masks(const BV<NW>& H, const BV<NW>& Lo, const BV<NW>& Nm, uint32_t L, const DeviceParams& P,
                            BV<NW>& Cf, BV<NW>& Cr)
{
    const uint32_t k = P.k;
    BV<NW> isA, isT;
#pragma unroll
    for (int j = 0; j < NW; ++j) { isA.w[j] = ~(H.w[j] | Lo.w[j]); isT.w[j] = H.w[j] & Lo.w[j]; }
    const BV<NW> AA = bv_and(isA, bv_shr(isA, 1));
    const BV<NW> TT = bv_and(isT, bv_shr(isT, 1));
    const BV<NW> Nw = bv_slide_or(Nm, k);

    // forward strand: AA at any of p+1 .. p+k-2, or AA at p followed by A/C
    BV<NW> badF = bv_or(bv_shr(bv_slide_or(AA, k - 2), 1), bv_andn(AA, bv_shr(H, 2)));
    // reverse strand: rc symbols (t, t+1) are bases (p+k-1-t, p+k-2-t): TT at any of p .. p+k-3, or
    // TT at p+k-2 (rc symbols 0,1 = AA) preceded by T/G at p+k-3 (rc symbol 2 = A/C)
    BV<NW> badR = bv_or(bv_slide_or(TT, k - 2), bv_and(bv_shr(TT, k - 2), bv_shr(H, k - 3)));
    // signatureMaskCutoffBits (FastqCategorizer.cpp:49-53): the low c bits of the k-mer must be 0.
    // Bit b of the k-mer is the lo (even b) / hi (odd b) bit of the symbol b/2 places from its end.
    for (uint32_t b = 0; b < P.cutoff_bits; ++b)
    {
        const BV<NW>& plane = (b & 1u) ? H : Lo;
        badF = bv_or(badF, bv_shr(plane, k - 1 - (b >> 1)));
        const BV<NW> sh = bv_shr(plane, b >> 1);                // rc symbol = 3 - base: the bit is complemented
#pragma unroll
        for (int j = 0; j < NW; ++j) badR.w[j] |= ~sh.w[j];
    }
    const int32_t lim = (int32_t)L - (int32_t)k - (int32_t)P.s;  // FindMinimizer scans i < L - k - s
    Cf = bv_andn(bv_andn(bv_below<NW>(lim), badF), Nw);
    Cr = bv_andn(bv_andn(bv_range<NW>((int32_t)P.s + 1, (int32_t)L - (int32_t)k + 1), badR), Nw);
}

// One bit step of the descent: keep the candidates whose key bit is 0 if there are any.  `plane`
// holds the key bit (INV = false) or its complement (INV = true) at the candidates' positions.
// A step is NW and-nots into an OR tree, one compare and NW three-input logic ops; the mask form is
// spelled out as LOP3s because, left to itself, the compiler turns it into a select per word.
FSB_HD uint32_t mask_nonzero(uint32_t o)                         // o != 0 ? 0xFFFFFFFF : 0
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("set.ne.u32.u32 %0, %1, 0;" : "=r"(r) : "r"(o));
    return r;
#else
    return o ? 0xFFFFFFFFu : 0u;
#endif
}
template <bool INV> FSB_HD uint32_t drop_ones(uint32_t d, uint32_t am, uint32_t plane)     // d & ~(am & key bit)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    if (INV) asm("lop3.b32 %0, %1, %2, %3, 0xB0;" : "=r"(r) : "r"(d), "r"(am), "r"(plane));      // a & ~(b & ~c)
    else asm("lop3.b32 %0, %1, %2, %3, 0x70;" : "=r"(r) : "r"(d), "r"(am), "r"(plane));          // a & ~(b & c)
    return r;
#else
    return d & ~(am & (INV ? ~plane : plane));
#endif
}
template <int NW, bool INV>
FSB_HD uint32_t descend_step_t(BV<NW>& D, const BV<NW>& plane)
{
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) o |= D.w[j] & (INV ? plane.w[j] : ~plane.w[j]);
    const uint32_t am = mask_nonzero(o);                        // any candidate with a 0 bit: drop those with a 1 bit
#pragma unroll
    for (int j = 0; j < NW; ++j) D.w[j] = drop_ones<INV>(D.w[j], am, plane.w[j]);
    return am;                                                  // all ones if a candidate had a 0 bit (the key bit is 0), else 0
}
template <int NW>
FSB_HD uint32_t descend_step(BV<NW>& D, const BV<NW>& plane, bool inv)
{
    return inv ? descend_step_t<NW, true>(D, plane) : descend_step_t<NW, false>(D, plane);
}

// Radix descent, forward strand.  D tracks the candidates shifted onto the symbol under test.
template <int NW>
FSB_HD StrandMin descend_forward(BV<NW> D, const BV<NW>& H, const BV<NW>& Lo, const DeviceParams& P)
{
    uint32_t m = 0;
    for (uint32_t d = 0; d < P.k; ++d)
    {
        if (d) D = bv_shl(D, 1);
        m = 2 * m + descend_step<NW>(D, H, false);                // accumulates -(key bit == 0); the ones are added at the end
        m = 2 * m + descend_step<NW>(D, Lo, false);
    }
    StrandMin r;
    r.sig = m + P.kmer_mask;
    r.pos = bv_lowest(D) - (P.k - 1);
    return r;
}

// Radix descent, reverse strand: symbol d of the rc k-mer at forward position p is the complement
// of base p+k-1-d.  The first minimum in reverse-strand order is the highest forward position.
template <int NW>
FSB_HD StrandMin descend_reverse(const BV<NW>& C, const BV<NW>& H, const BV<NW>& Lo, uint32_t L, const DeviceParams& P)
{
    BV<NW> D = bv_shl(C, P.k - 1);
    uint32_t m = 0;
    for (uint32_t d = 0; d < P.k; ++d)
    {
        if (d) D = bv_shr(D, 1);
        m = 2 * m + descend_step<NW>(D, H, true);
        m = 2 * m + descend_step<NW>(D, Lo, true);
    }
    StrandMin r;
    r.sig = m + P.kmer_mask;
    r.pos = L - P.k - bv_highest(D);
    return r;
}

// Both strands of one mate in one descent: the candidates of the two strands keep their own sets (Df moves up
// through the k-mer, Dr down), but every bit step takes one decision for the union -- keep the candidates whose
// key bit is 0 if either set has any.  What survives shares the smaller of the two strand minima; a strand whose
// set ends up empty had the larger one.  For the selection rules of DistributeToBins (FastqCategorizer.cpp:212-245,
// 289-336) the exact value of a losing strand does not matter -- every comparison it enters is decided by its
// being larger than the winner -- so it is reported as 4^k; ties leave both sets populated and both exact.
FSB_HD uint32_t key_zero_mix(uint32_t df, uint32_t dr, uint32_t plane)      // candidates whose key bit is 0 (one three-input logic op)
{
    return (df & ~plane) | (dr & plane);
}
// a | b | c as one three-input logic op that the compiler will not fold into a longer chain
FSB_HD uint32_t or3(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xFE;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return a | b | c;
#endif
}
// The step's decision needs the OR over all words.  Written as a running `o |= ...` it becomes a chain of 2 NW dependent
// logic ops per bit step -- the longest dependency of the kernel, 16 times per mate; as NW independent ops and a tree of
// three-input ORs the chain is 1 + ceil(log3 NW) deep and two ops shorter.
template <int NW>
FSB_HD uint32_t descend_step_joint(BV<NW>& Df, BV<NW>& Dr, const BV<NW>& plane)
{
#ifdef FSB_K1_DESCENT_CHAIN                                          // (the running OR, for comparison)
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) o |= key_zero_mix(Df.w[j], Dr.w[j], plane.w[j]);
    const uint32_t am0 = mask_nonzero(o);
#pragma unroll
    for (int j = 0; j < NW; ++j) { Df.w[j] = drop_ones<false>(Df.w[j], am0, plane.w[j]); Dr.w[j] = drop_ones<true>(Dr.w[j], am0, plane.w[j]); }
    return am0;
#endif
    uint32_t z[NW + 2];
#pragma unroll
    for (int j = 0; j < NW; ++j) z[j] = key_zero_mix(Df.w[j], Dr.w[j], plane.w[j]);
    z[NW] = 0; z[NW + 1] = 0;
    int n = NW;
#pragma unroll
    for (int round = 0; round < 3; ++round)                      // 3 rounds reduce up to 27 words
    {
        if (n > 1)
        {
            const int m = (n + 2) / 3;
#pragma unroll
            for (int j = 0; j < m; ++j)
            {
                const uint32_t a = z[3 * j], b = (3 * j + 1 < n) ? z[3 * j + 1] : 0u, c = (3 * j + 2 < n) ? z[3 * j + 2] : 0u;
                z[j] = (3 * j + 2 < n) ? or3(a, b, c) : (a | b);
            }
            n = m;
        }
    }
    const uint32_t am = mask_nonzero(z[0]);
#pragma unroll
    for (int j = 0; j < NW; ++j) { Df.w[j] = drop_ones<false>(Df.w[j], am, plane.w[j]); Dr.w[j] = drop_ones<true>(Dr.w[j], am, plane.w[j]); }
    return am;
}
template <int NW>
FSB_HD void descend_joint(const BV<NW>& Cf, const BV<NW>& Cr, const BV<NW>& H, const BV<NW>& Lo, uint32_t L, const DeviceParams& P,
                          StrandMin& fwd, StrandMin& rev)
{
    BV<NW> Df = Cf, Dr = bv_shl(Cr, P.k - 1);
    uint32_t m = descend_step_joint<NW>(Df, Dr, H);               // symbol 0 in place, then k - 1 times: move on, two bit steps
    m = 2 * m + descend_step_joint<NW>(Df, Dr, Lo);
#if defined(__CUDA_ARCH__)
#pragma unroll 2
#endif
    for (uint32_t d = 1; d < P.k; ++d)
    {
        Df = bv_shl(Df, 1); Dr = bv_shr(Dr, 1);
        m = 2 * m + descend_step_joint<NW>(Df, Dr, H);
        m = 2 * m + descend_step_joint<NW>(Df, Dr, Lo);
    }
    m += P.kmer_mask;
    if (bv_any(Df)) { fwd.sig = m; fwd.pos = bv_lowest(Df) - (P.k - 1); }
    if (bv_any(Dr)) { rev.sig = m; rev.pos = L - P.k - bv_highest(Dr); }
}

// FM(x) and FM(rc(x)) of one mate from its bit planes (N plane already cut to the read length), plus
// its N count.  L <= 32 * NW.  (A strand that loses to the other strand of the same mate reports 4^k, see above.)
template <int NW>
FSB_HD void plane_minimizers(const BV<NW>& H, const BV<NW>& Lo, const BV<NW>& Nm, uint32_t L, const DeviceParams& P,
                             StrandMin& fwd, StrandMin& rev, uint32_t& nN)
{
    nN = bv_popc(Nm);
    BV<NW> Cf, Cr;
    candidate_masks<NW>(H, Lo, Nm, L, P, Cf, Cr);
    const bool tooManyN = nN >= L / 3;                           // FastqCategorizer.cpp:102
    fwd.sig = P.nbin; fwd.pos = 0;
    rev.sig = P.nbin; rev.pos = 0;
    if (!tooManyN && (bv_any(Cf) || bv_any(Cr))) descend_joint<NW>(Cf, Cr, H, Lo, L, P, fwd, rev);
}
template <int NW>
FSB_HD void mate_planes(const uint32_t* words, uint32_t bshift, uint32_t L, BV<NW>& H, BV<NW>& Lo, BV<NW>& Nm)
{
    ascii_to_planes<NW>(words, bshift, H, Lo, Nm);
    Nm = bv_and(Nm, bv_below<NW>((int32_t)L));
}
template <int NW>
FSB_HD void mate_minimizers(const uint32_t* words, uint32_t bshift, uint32_t L, const DeviceParams& P,
                            StrandMin& fwd, StrandMin& rev, uint32_t& nN, BV<NW>& H, BV<NW>& Lo, BV<NW>& Nm)
{
    mate_planes<NW>(words, bshift, L, H, Lo, Nm);
    plane_minimizers<NW>(H, Lo, Nm, L, P, fwd, rev, nN);
}

// FastqCategorizerSE::DistributeToBins (FastqCategorizer.cpp:212-245): forward wins ties.
FSB_HD void select_se(const StrandMin& f, const StrandMin& r, uint32_t nN, const DeviceParams& P, uint32_t& sig, uint32_t& info)
{
    const bool reverse = !(f.sig <= r.sig);
    sig = reverse ? r.sig : f.sig;
    uint32_t pos = reverse ? r.pos : f.pos, flags = 0;
    if (sig != P.nbin) flags |= reverse ? FSB_INFO_REVERSE : 0u; else pos = 0;
    if (nN == 0) flags |= FSB_INFO_PLAIN_A;
    info = pos | flags;
}

// FastqCategorizerPE::DistributeToBins (:289-336): strict < everywhere, so ties go to mate 2 over
// mate 1 and to the reverse strand over the forward one.  f1 = FM(m1), f2 = FM(m2),
// r1 = FM(rc(m2)) ("minRev_1": first half of the reversed pair), r2 = FM(rc(m1)).
FSB_HD void select_pe(const StrandMin& f1, const StrandMin& f2, const StrandMin& r1, const StrandMin& r2, uint32_t nN1, uint32_t nN2,
                      const DeviceParams& P, uint32_t& sig, uint32_t& info)
{
    const bool isF1 = f1.sig < f2.sig;
    const StrandMin F = isF1 ? f1 : f2;
    const bool isR1 = r1.sig < r2.sig;
    const StrandMin R = isR1 ? r1 : r2;
    bool isRev, first;
    uint32_t pos, flags = 0;
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
    if ((a_is_m2 ? nN2 : nN1) == 0) flags |= FSB_INFO_PLAIN_A;
    if ((a_is_m2 ? nN1 : nN2) == 0) flags |= FSB_INFO_PLAIN_B;
    info = pos | flags;
}

// fastore_rebin's second scan (DnaRebalancer::FindMinimizerHR, fastore_rebin/DnaRebalancer.cpp:570-601) is FindMinimizer with two more
// conditions per k-mer: m != curSignature and m % divisor == 0 (divisor a power of two: the low log2(divisor) bits are zero -- the
// same test as signatureMaskCutoffBits, taken by candidate_masks).  This removes the positions whose k-mer IS `cur` from both
// candidate sets: symbol d of the forward k-mer at p is base p+d, symbol d of the reverse-strand k-mer at forward position p is the
// complement of base p+k-1-d.
template <int NW>
FSB_HD void exclude_signature(const BV<NW>& H, const BV<NW>& Lo, uint32_t k, uint32_t cur, BV<NW>& Cf, BV<NW>& Cr)
{
    BV<NW> eqF = Cf, eqR = Cr;                                   // positions that still equal `cur` on every symbol looked at so far
    for (uint32_t d = 0; d < k; ++d)
    {
        const uint32_t sym = (cur >> (2u * (k - 1u - d))) & 3u, hi = sym >> 1, lo = sym & 1u;
        const BV<NW> hf = bv_shr(H, d), lf = bv_shr(Lo, d);
        const BV<NW> hr = bv_shr(H, k - 1u - d), lr = bv_shr(Lo, k - 1u - d);
#pragma unroll
        for (int j = 0; j < NW; ++j)
        {
            eqF.w[j] &= (hi ? hf.w[j] : ~hf.w[j]) & (lo ? lf.w[j] : ~lf.w[j]);
            eqR.w[j] &= (hi ? ~hr.w[j] : hr.w[j]) & (lo ? ~lr.w[j] : lr.w[j]);      // complement: 3 - base
        }
    }
    Cf = bv_andn(Cf, eqF);
    Cr = bv_andn(Cr, eqR);
}

// FindNewMinimizer (DnaRebalancer.cpp:604-616) of one read from its bit planes: FindMinimizerHR on the read and on its reverse
// complement, the reverse strand only if its signature is strictly smaller.  P.cutoff_bits must already include log2(divisor).
template <int NW>
FSB_HD void plane_new_minimizer(const BV<NW>& H, const BV<NW>& Lo, const BV<NW>& Nm, uint32_t L, const DeviceParams& P, uint32_t cur,
                                uint32_t& sig, uint32_t& info)
{
    const uint32_t nN = bv_popc(Nm);
    BV<NW> Cf, Cr;
    candidate_masks<NW>(H, Lo, Nm, L, P, Cf, Cr);
    exclude_signature<NW>(H, Lo, P.k, cur, Cf, Cr);
    StrandMin f, r;
    f.sig = r.sig = P.nbin; f.pos = r.pos = 0;
    if (!(nN >= L / 3) && (bv_any(Cf) || bv_any(Cr))) descend_joint<NW>(Cf, Cr, H, Lo, L, P, f, r);
    select_se(f, r, nN, P, sig, info);
    info &= (FSB_INFO_POS_MASK | FSB_INFO_REVERSE);
}

} // namespace fsb
