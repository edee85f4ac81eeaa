// rebin_sig.cuh -- the signature scan of fastore_rebin on the device: DnaRebalancer::FindNewMinimizer
// (fastore_rebin/DnaRebalancer.cpp:570-616) for a table of reads, with the bit-plane search of K1 (sig_core.cuh).
//
// fastore_rebin moves the reads of a bin whose signature is not a multiple of the level's divisor to a bin whose
// signature is: for every such read it looks for the smallest valid k-mer that differs from the bin's signature and
// is a multiple of the divisor, on both strands.  That is FindMinimizer with two extra conditions per k-mer, i.e. the
// same kernel core with two more candidate masks.  (The variant that only admits signatures present in the input bin
// file -- BinBalanceParameters::validBinSignatures filled from the file, RebinModule.cpp:62-68 -- needs a table look-up
// per k-mer and is not covered; the default admits all signatures, :72.)
#pragma once

#include <cuda_runtime.h>

#include "sig_core.cuh"

namespace fsb {

// one thread per read; the sequence is read as aligned words straight from the text (a utility kernel: no staging)
template <int NW>
__global__ void __launch_bounds__(128) find_new_minimizer_kernel(const uint8_t* __restrict__ text, const fsb_record* __restrict__ rec, uint64_t n, DeviceParams P,
                                                                 uint32_t cur, uint32_t* __restrict__ sig_out, uint32_t* __restrict__ info_out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fsb_record r = rec[i];
    const uint32_t L = r.seq_len;
    uint32_t win[8 * NW + 2];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(text + (r.seq_off & ~3u));
    const uint32_t nwords = ((r.seq_off & 3u) + L + 3u) >> 2;
#pragma unroll 1
    for (uint32_t j = 0; j < 8 * NW + 2; ++j) win[j] = j < nwords ? src[j] : 0u;
    BV<NW> H, Lo, Nm;
    mate_planes<NW>(win, 8u * (r.seq_off & 3u), L, H, Lo, Nm);
    uint32_t sig, info;
    plane_new_minimizer<NW>(H, Lo, Nm, L, P, cur, sig, info);
    sig_out[i] = sig; info_out[i] = info;
}

} // namespace fsb
