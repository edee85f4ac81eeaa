// signature.cuh -- K1: per-read minimizer signature on both strands (sm_100a).
//
// Replaces FastqCategorizerBase::FindMinimizer x2 (SE) / x4 (PE), FastqRecord::ComputeRC and the
// per-record selection logic of FastqCategorizerSE/PE::DistributeToBins
// (FastqCategorizer.cpp:79-106, 197-253, 256-363; FastqRecord.h:80-111).
//
// Mapping: one thread per mate (SE: per read; PE: lanes 2i / 2i+1 hold mate 1 / mate 2 of pair i),
// one warp per 32 mates.  The sequences are not 16-byte aligned inside FASTQ text, so the warp first
// copies the aligned 16-byte pieces around every mate's sequence into a private shared-memory slot
// with cp.async -- half a warp per mate, i.e. coalesced 16-byte requests over whole sectors, and
// only the sequence lines are touched (titles, '+' lines and qualities stay in HBM).  Each thread
// then reads its slot as 4-byte words and runs the bit-parallel search of sig_core.cuh.
#pragma once

#include <cuda_runtime.h>

#include "sig_core.cuh"

namespace fsb {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

constexpr int kSigWarps = 4;                                   // warps per block (warps are independent)
template <int NW> constexpr int sig_slot_pieces() { return 2 * NW + 1; }     // 16-byte pieces per mate slot
template <int NW> constexpr size_t sig_smem_bytes() { return (size_t)kSigWarps * 32 * sig_slot_pieces<NW>() * 16; }

// K1.  NW = ceil(longest read of the batch / 32).  keys[i] = chunk : signature, info[i] = minimPos | FSB_INFO_*.
template <int NW>
__global__ void __launch_bounds__(kSigWarps * 32) signature_kernel(BatchView B, DeviceParams P, uint32_t* __restrict__ keys,
                                                                    uint32_t* __restrict__ info, uint32_t* __restrict__ sig_out)
{
    constexpr int PIECES = sig_slot_pieces<NW>();
    extern __shared__ uint4 sig_smem[];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* wslots = sig_smem + (size_t)warp * 32 * PIECES;

    const uint64_t n_mates = P.paired ? 2 * B.n_records : B.n_records;
    const uint64_t g = ((uint64_t)blockIdx.x * kSigWarps + warp) * 32 + lane;      // this thread's mate
    const bool live = g < n_mates;
    const uint64_t i = P.paired ? (g >> 1) : g;
    const unsigned m = P.paired ? (unsigned)(g & 1) : 0u;

    uint32_t L = 0, piece0 = 0, npieces = 0, a = 0, ch = 0;
    if (live)
    {
        ch = find_chunk(B, i);
        const fsb_record r = (m ? B.rec[1] : B.rec[0])[i];           // (no dynamic indexing of kernel parameters)
        const uint64_t off = (m ? B.chunk_text_base[1] : B.chunk_text_base[0])[ch] + r.seq_off;         // byte offset of base 0 inside text[m]
        L = r.seq_len;
        a = (uint32_t)(off & 15u);
        piece0 = (uint32_t)(off >> 4);
        npieces = (a + L + 15u) >> 4;
    }
    // ---- stage: half a warp copies one mate's pieces (whole warp when a slot has more than 16 pieces) ----
    constexpr int LANES_PER_MATE = PIECES <= 16 ? 16 : 32;
    constexpr int MATES_PER_ITER = 32 / LANES_PER_MATE;
#pragma unroll 4
    for (int it = 0; it < 32 / MATES_PER_ITER; ++it)
    {
        const int src = it * MATES_PER_ITER + (int)(lane / LANES_PER_MATE);
        const uint32_t p0 = __shfl_sync(0xFFFFFFFFu, piece0, src);
        const uint32_t np = __shfl_sync(0xFFFFFFFFu, npieces, src);
        const unsigned j = lane % LANES_PER_MATE;
        const uint8_t* base = (P.paired && (src & 1)) ? B.text[1] : B.text[0];
        if (j < np) cp_async16(&wslots[src * PIECES + j], base + ((uint64_t)(p0 + j) << 4));
    }
    cp_async_commit_wait_all();
    __syncwarp();

    StrandMin f, r;
    uint32_t nN = 0;
    f.sig = r.sig = P.nbin; f.pos = r.pos = 0;
    if (live)
    {
        const uint32_t* words = reinterpret_cast<const uint32_t*>(&wslots[lane * PIECES]) + (a >> 2);
        mate_minimizers<NW>(words, 8u * (a & 3u), L, P, f, r, nN);
    }
    uint32_t sig, inf;
    if (!P.paired) select_se(f, r, nN, P, sig, inf);
    else
    {
        // lanes 2i and 2i+1 exchange their results: the even lane holds f1 = FM(m1), r2 = FM(rc(m1));
        // the odd lane f2 = FM(m2), r1 = FM(rc(m2))
        StrandMin of, orv;
        of.sig = __shfl_xor_sync(0xFFFFFFFFu, f.sig, 1); of.pos = __shfl_xor_sync(0xFFFFFFFFu, f.pos, 1);
        orv.sig = __shfl_xor_sync(0xFFFFFFFFu, r.sig, 1); orv.pos = __shfl_xor_sync(0xFFFFFFFFu, r.pos, 1);
        const uint32_t onN = __shfl_xor_sync(0xFFFFFFFFu, nN, 1);
        select_pe(f, of, orv, r, nN, onN, P, sig, inf);          // meaningful on even lanes only
    }
    if (live && m == 0)
    {
        keys[i] = (ch << P.key_bits) | sig;
        info[i] = inf;
        if (sig_out) sig_out[i] = sig;
    }
}

} // namespace fsb
