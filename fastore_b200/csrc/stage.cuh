// stage.cuh -- device-side check of the staged record tables and the batch statistics the host
// needs to size buffers and pick kernel instantiations.
//
// The reference only ASSERTs its input contract (FastqRecord.h:87,192; FastqParser.cpp:130): read
// length 1..255, equal mate lengths in a pair, views inside the chunk.  Violations would make the
// kernels read outside the staged text, so fsb_stage rejects them (FSB_ERR_INPUT).
#pragma once

#include <cuda_runtime.h>

#include "core.cuh"

namespace fsb {

struct StageStats
{
    unsigned long long bases;        // sum of sequence lengths, both mates
    unsigned long long heads;        // sum of head_len (mate 1)
    unsigned long long first_bad;    // lowest record index violating the contract (~0 if none)
    uint32_t min_len, max_len;       // over all mates
    uint32_t max_head;
    uint32_t n_bad;
};

__global__ void __launch_bounds__(256) stage_stats_kernel(BatchView B, DeviceParams P, const uint64_t* __restrict__ text_size0,
                                                          const uint64_t* __restrict__ text_size1, StageStats* __restrict__ st)
{
    __shared__ unsigned long long sh_bases, sh_heads;
    __shared__ uint32_t sh_min, sh_max, sh_hmax;
    if (threadIdx.x == 0) { sh_bases = 0; sh_heads = 0; sh_min = 0xFFFFFFFFu; sh_max = 0; sh_hmax = 0; }
    __syncthreads();
    uint64_t bases = 0, heads = 0;
    uint32_t mn = 0xFFFFFFFFu, mx = 0, hmx = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B.n_records; i += (uint64_t)gridDim.x * blockDim.x)
    {
        const uint32_t ch = find_chunk(B, i);
        const fsb_record a = B.rec[0][i];
        const uint64_t ts0 = text_size0[ch];
        bool ok = a.seq_len >= 1 && a.seq_len <= 255 && (uint64_t)a.seq_off + a.seq_len <= ts0 && (uint64_t)a.qua_off + a.seq_len <= ts0 &&
                  (uint64_t)a.head_off + a.head_len <= ts0;
        bases += a.seq_len; heads += a.head_len;
        mn = min(mn, (uint32_t)a.seq_len); mx = max(mx, (uint32_t)a.seq_len); hmx = max(hmx, (uint32_t)a.head_len);
        if (P.paired)
        {
            const fsb_record b = B.rec[1][i];
            const uint64_t ts1 = text_size1[ch];
            ok = ok && b.seq_len == a.seq_len && (uint64_t)b.seq_off + b.seq_len <= ts1 && (uint64_t)b.qua_off + b.seq_len <= ts1;
            bases += b.seq_len;
        }
        if (!ok) { atomicAdd(&st->n_bad, 1u); atomicMin(&st->first_bad, (unsigned long long)i); }
    }
    atomicAdd(&sh_bases, (unsigned long long)bases); atomicAdd(&sh_heads, (unsigned long long)heads);
    atomicMin(&sh_min, mn); atomicMax(&sh_max, mx); atomicMax(&sh_hmax, hmx);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        atomicAdd(&st->bases, sh_bases); atomicAdd(&st->heads, sh_heads);
        atomicMin(&st->min_len, sh_min); atomicMax(&st->max_len, sh_max); atomicMax(&st->max_head, sh_hmax);
    }
}

} // namespace fsb
