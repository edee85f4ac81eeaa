// layout.cuh -- bin boundaries, per-bin length statistics, per-read bit lengths and the bit-offset
// scans that place every record in the four output streams.
//
// Restates the framing rules of FastqRecordsPackerSE::PackToBins / PackToBin / StoreRecords
// (FastqPacker.cpp:417-491, 541-602, 734-759, 815-859) as arithmetic on lengths:
//   per bin    17 header bits in meta (8 minLen, 8 maxLen, 1 hasReadGroups), then the records in
//              parse order, then each of the four writers is flushed to a byte boundary;
//   per record meta: [len-minLen in bit_length(max-min) bits, twice for PE] if the bin has variable
//              length, [PE & not N-bin: 1 swap bit], [not N-bin: 1 reverse bit + 8 bits minimPos],
//              1 isDnaPlain bit per mate;
//              dna : (L - k) symbols for the mate holding the signature (L in the N-bin), L for the
//              other mate, 2 bits each if the mate has no N else 3;
//              qua : q bits per base; head: 8 + 7 * (headLen - 1) bits (mate-1 header only).
#pragma once

#include "ingest.cuh"
#include "pack_core.cuh"

namespace fsb {

struct SortedView
{
    const uint32_t* skeys;              // [n] sorted keys (chunk : signature)
    const unsigned long long* cards;    // [n] the records' cards in sorted order (core.cuh: card_make)
};

struct BinArrays
{
    uint32_t* bin_of;           // [n]   bin index of each sorted position
    uint32_t* bin_start;        // [nb_max + 1] first sorted position of each bin
    uint32_t* bin_min;          // [nb_max] min seqLen
    uint32_t* bin_max;          // [nb_max] max seqLen
    unsigned long long* bin_raw_dna;   // [nb_max]
    unsigned long long* bin_raw_head;  // [nb_max]
};



// flags[i] = 1 where a new (chunk, signature) bin starts
__global__ void bin_flags_kernel(const uint32_t* __restrict__ skeys, uint64_t n, uint32_t* __restrict__ flags)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || skeys[i] != skeys[i - 1]) ? 1u : 0u;
}

// bin index per sorted position, bin starts, per-bin min/max seqLen and raw sizes.  Records of a
// bin are neighbours in the sorted order, so a warp first combines its lanes per bin
// (__match_any_sync) and only one lane per (warp, bin) touches the per-bin counters.
__global__ void bin_stats_kernel(uint64_t n, DeviceParams P, SortedView S, const uint32_t* __restrict__ flags,
                                 const uint32_t* __restrict__ flags_excl, BinArrays A)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    uint32_t bin = 0xFFFFFFFFu, L = 0, raw = 0, hl = 0;
    if (live)
    {
        const uint32_t f = flags[i];
        bin = flags_excl[i] + f - 1;
        A.bin_of[i] = bin;
        if (f) A.bin_start[bin] = (uint32_t)i;
        const uint64_t c = S.cards[i];
        L = card_lenA(c);                       // PE mates have equal lengths (checked by fsb_stage), so A's length is "seqLen"
        raw = L + card_lenB(c);
        hl = card_head(c);
    }
    const unsigned lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, bin);
    const uint32_t mn = __reduce_min_sync(peers, L), mx = __reduce_max_sync(peers, L);
    const uint32_t sraw = __reduce_add_sync(peers, raw), shl = __reduce_add_sync(peers, hl);
    if (live && (unsigned)(__ffs(peers) - 1) == lane)
    {
        atomicMin(&A.bin_min[bin], mn);
        atomicMax(&A.bin_max[bin], mx);
        atomicAdd(&A.bin_raw_dna[bin], (unsigned long long)sraw);
        if (P.has_headers) atomicAdd(&A.bin_raw_head[bin], (unsigned long long)shl);
    }
}

__global__ void read_bits_kernel(uint64_t n, DeviceParams P, SortedView S, BinArrays A,
                                 uint32_t* __restrict__ bm, uint32_t* __restrict__ bd, uint32_t* __restrict__ bq, uint32_t* __restrict__ bh)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t c = S.cards[i];
    const uint32_t bin = A.bin_of[i];
    const bool nbin = (S.skeys[i] & ((1u << P.key_bits) - 1)) == P.nbin;
    const ReadBits b = read_bit_lengths(P, nbin, card_info(c), card_lenA(c), card_lenB(c), card_head(c), A.bin_min[bin], A.bin_max[bin]);
    bm[i] = b.meta; bd[i] = b.dna; bq[i] = b.qua; bh[i] = b.head;
}

struct StreamScans
{
    const uint64_t* P[4];       // [n + 1] exclusive prefix of per-read bits in sorted order (meta, dna, qua, head)
};

// per bin: byte sizes of the four streams + descriptor.  bins >= nb (up to nb_max) get zero sizes.
__global__ void bin_sizes_kernel(DeviceParams P, uint64_t n, const uint32_t* __restrict__ nb_ptr, uint64_t nb_max, const uint32_t* __restrict__ skeys,
                                 BinArrays A, StreamScans SC, uint64_t* __restrict__ by_m, uint64_t* __restrict__ by_d,
                                 uint64_t* __restrict__ by_q, uint64_t* __restrict__ by_h, fsb_bin_descriptor* __restrict__ desc)
{
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb_max) return;
    const uint64_t nb = *nb_ptr;
    if (b >= nb) { by_m[b] = by_d[b] = by_q[b] = by_h[b] = 0; return; }
    const uint64_t s = A.bin_start[b];
    const uint64_t e = (b + 1 < nb) ? (uint64_t)A.bin_start[b + 1] : n;
    const uint64_t m = (SC.P[0][e] - SC.P[0][s] + 17 + 7) >> 3;          // + bin header, FlushPartialWordBuffer
    const uint64_t d = (SC.P[1][e] - SC.P[1][s] + 7) >> 3;
    const uint64_t q = (SC.P[2][e] - SC.P[2][s] + 7) >> 3;
    const uint64_t h = (SC.P[3][e] - SC.P[3][s] + 7) >> 3;
    by_m[b] = m; by_d[b] = d; by_q[b] = q; by_h[b] = h;
    fsb_bin_descriptor D;
    D.signature = skeys[s] & ((1u << P.key_bits) - 1);
    D.meta_size = m; D.dna_size = d; D.qua_size = q; D.head_size = h;
    D.records_count = e - s;
    D.raw_dna_size = A.bin_raw_dna[b];
    D.raw_head_size = A.bin_raw_head[b];
    desc[b] = D;
}

// Per-chunk summary the host needs to slice the batch result into blocks.
struct ChunkSummary
{
    uint64_t first_bin;          // index of the chunk's first bin in the batch-wide descriptor array
    uint64_t n_bins;
    uint64_t off[4];             // byte offset of the chunk's streams inside the batch-wide streams
    uint64_t size[4];
    uint64_t raw_dna, raw_head;
};

struct BinOffsets { const uint64_t* B[4]; };    // [nb_max + 1] exclusive prefix of per-bin bytes

__global__ void chunk_summary_kernel(BatchView Bv, const uint32_t* __restrict__ nb_ptr, const uint32_t* __restrict__ bin_of, BinOffsets BO,
                                     const fsb_bin_descriptor* __restrict__ desc, ChunkSummary* __restrict__ out)
{
    const uint32_t c = blockIdx.x;                 // one block per chunk
    const uint64_t nb = *nb_ptr;
    const uint64_t r0 = Bv.chunk_first_rec[c], r1 = Bv.chunk_first_rec[c + 1];
    // sorted ranges coincide with the input ranges: the key is chunk-major
    const uint64_t b0 = (r0 < Bv.n_records) ? bin_of[r0] : nb;
    const uint64_t b1 = (r1 < Bv.n_records) ? bin_of[r1] : nb;
    __shared__ unsigned long long raw[2];
    if (threadIdx.x == 0) { raw[0] = 0; raw[1] = 0; }
    __syncthreads();
    unsigned long long a = 0, h = 0;
    for (uint64_t b = b0 + threadIdx.x; b < b1; b += blockDim.x) { a += desc[b].raw_dna_size; h += desc[b].raw_head_size; }
    atomicAdd(&raw[0], a); atomicAdd(&raw[1], h);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        ChunkSummary s;
        s.first_bin = (r0 == r1) ? b0 : b0; s.n_bins = (r0 == r1) ? 0 : b1 - b0;
        for (int k = 0; k < 4; ++k) { s.off[k] = BO.B[k][b0]; s.size[k] = (r0 == r1) ? 0 : BO.B[k][b1] - BO.B[k][b0]; }
        s.raw_dna = raw[0]; s.raw_head = raw[1];
        out[c] = s;
    }
}

} // namespace fsb
