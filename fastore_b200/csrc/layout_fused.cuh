// layout_fused.cuh -- the layout of a batch whose reads all have one length, as one scan (sm_100a).
//
// Replaces, for such batches, the kernels of layout.cuh and the placement kernel of place.cuh: three
// streaming launches turn the sorted (key, card) pairs into everything K4 needs -- every record's bit
// position inside its K4 tile for the four streams, the tiles' first bits, the per-record bin info --
// plus the bin descriptors and the per-chunk summary.  The arithmetic is layout_core.cuh's scan.
//
// Mapping: a thread owns four consecutive sorted records -- one 16-byte vector of every array it reads or writes, so a
// warp's accesses are fully coalesced -- and a block of 512 threads owns 2048 records.
//   lay_reduce   every block's LayState                              -> tile_states[block]
//   lay_scan     one block: exclusive scan over the block states      (in place; [nblocks] = the whole batch)
//   lay_apply    thread prefix = block prefix (+) threads in front; absolute walk that writes loc / binfo / tbase,
//                clears the stream words two K4 tiles share, and emits the descriptor of every bin at its last record
//   chunk_summary_fused   per chunk: first bin, stream offsets and sizes, raw sizes (block per chunk)
#pragma once

#include <cuda_runtime.h>

#include "layout_core.cuh"
#include "layout.cuh"
#include "place.cuh"

namespace fsb {

constexpr uint32_t kLayThreads = 512;
constexpr uint32_t kLayPerThread = 4;                               // a thread walks over four consecutive records: one 16-byte vector of every array
constexpr uint32_t kLayBlock = kLayThreads * kLayPerThread;         // 2048 records per block
constexpr uint32_t kLayTileLanes = kPlaceTile / kLayPerThread;      // threads that share a K4 tile (8 neighbouring lanes)
static_assert(kLayPerThread == 4 && kPlaceTile % kLayPerThread == 0 && 32 % kLayTileLanes == 0, "a K4 tile is a whole group of lanes of one warp");

// ---- moving a LayState between lanes ---------------------------------------------------------------------------
constexpr int kLayWords = (int)(sizeof(LayState) / 4);
static_assert(sizeof(LayState) % 4 == 0, "LayState moves as 32-bit words");
__device__ __forceinline__ LayState lay_shfl_up(const LayState& s, unsigned d)
{
    LayState r;
    const uint32_t* in = reinterpret_cast<const uint32_t*>(&s);
    uint32_t* out = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int w = 0; w < kLayWords; ++w) out[w] = __shfl_up_sync(0xFFFFFFFFu, in[w], d);
    return r;
}
// inclusive scan over the lanes of a warp (lane order = record order)
__device__ __forceinline__ LayState lay_warp_scan(LayState s)
{
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (unsigned d = 1; d < 32; d <<= 1)
    {
        const LayState o = lay_shfl_up(s, d);
        if (lane >= d) s = lay_combine(o, s);
    }
    return s;
}
// Exclusive prefix of `mine` over the threads of the block (thread order = record order); `total` = the block's state.
// THREADS / 32 entries of shared memory.
template <uint32_t THREADS>
__device__ __forceinline__ LayState lay_block_scan(const LayState& mine, LayState& total, LayState* sm)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const LayState inc = lay_warp_scan(mine);
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    LayState before = lay_identity();                               // the warps in front of this one
    total = lay_identity();
    for (unsigned w = 0; w < THREADS / 32; ++w)
    {
        if (w == warp) before = total;
        total = lay_combine(total, sm[w]);
    }
    LayState excl = lay_shfl_up(inc, 1);                            // the lanes in front inside the warp
    if (lane == 0) excl = lay_identity();
    __syncthreads();
    return lay_combine(before, excl);
}

// A thread's four sorted records (fewer at the end of the batch), the key in front of them and the key behind them.
struct LayRun
{
    uint64_t i0;             // first record
    uint32_t cnt;            // records (0 .. 4)
    uint32_t key[4];
    unsigned long long card[4];
    uint32_t prev, next;     // keys of the records i0 - 1 and i0 + cnt (any value where there is none: lay_record / the walk know)
};
__device__ __forceinline__ LayRun lay_run(uint64_t n, const SortedView& S)
{
    LayRun r;
    const unsigned lane = threadIdx.x & 31;
    r.i0 = ((uint64_t)blockIdx.x * kLayThreads + threadIdx.x) * kLayPerThread;
    r.cnt = r.i0 < n ? (uint32_t)min((uint64_t)kLayPerThread, n - r.i0) : 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) { r.key[u] = 0; r.card[u] = 0; }
    if (r.cnt == kLayPerThread)
    {
        const uint4 kk = *reinterpret_cast<const uint4*>(S.skeys + r.i0);
        const ulonglong2 ca = *reinterpret_cast<const ulonglong2*>(S.cards + r.i0), cb = *reinterpret_cast<const ulonglong2*>(S.cards + r.i0 + 2);
        r.key[0] = kk.x; r.key[1] = kk.y; r.key[2] = kk.z; r.key[3] = kk.w;
        r.card[0] = ca.x; r.card[1] = ca.y; r.card[2] = cb.x; r.card[3] = cb.y;
    }
    else
    {
#pragma unroll
        for (int u = 0; u < 4; ++u) if ((uint32_t)u < r.cnt) { r.key[u] = S.skeys[r.i0 + u]; r.card[u] = S.cards[r.i0 + u]; }
    }
    // neighbours: from the lanes next door, from memory at the ends of the warp
    uint32_t last = r.key[0];
#pragma unroll
    for (int u = 1; u < 4; ++u) if ((uint32_t)u < r.cnt) last = r.key[u];
    r.prev = __shfl_up_sync(0xFFFFFFFFu, last, 1);
    r.next = __shfl_down_sync(0xFFFFFFFFu, r.key[0], 1);
    if (lane == 0) r.prev = (r.cnt && r.i0) ? S.skeys[r.i0 - 1] : 0u;
    if (lane == 31) r.next = (r.cnt == kLayPerThread && r.i0 + kLayPerThread < n) ? S.skeys[r.i0 + kLayPerThread] : 0u;
    return r;
}
__device__ __forceinline__ LayState lay_run_state(const DeviceParams& P, const LayRun& run, uint32_t uniform_len)
{
    LayState st = lay_identity();
    uint32_t prev = run.prev;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if ((uint32_t)u < run.cnt)
        {
            lay_push(st, lay_record(P, run.i0 + u == 0, run.key[u], prev, run.card[u], uniform_len));
            prev = run.key[u];
        }
    return st;
}

__global__ void __launch_bounds__(kLayThreads) lay_reduce_kernel(uint64_t n, DeviceParams P, SortedView S, uint32_t uniform_len, LayState* __restrict__ tile_states)
{
    __shared__ LayState sm[kLayThreads / 32];
    const LayState mine = lay_run_state(P, lay_run(n, S), uniform_len);
    LayState total;
    lay_block_scan<kLayThreads>(mine, total, sm);
    if (threadIdx.x == 0) tile_states[blockIdx.x] = total;
}

// one block: states[j] <- combination of states[0 .. j)  (j = 0 .. count; entry [count] is the whole batch).  Every thread takes
// a run of consecutive states, so one block-wide scan covers any count.
constexpr uint32_t kLayScanThreads = 512;
__global__ void __launch_bounds__(kLayScanThreads) lay_scan_kernel(LayState* __restrict__ states, uint32_t count)
{
    __shared__ LayState sm[kLayScanThreads / 32];
    const uint32_t per = (count + kLayScanThreads - 1) / kLayScanThreads;
    const uint32_t j0 = min(count, threadIdx.x * per), j1 = min(count, j0 + per);
    LayState mine = lay_identity();
    for (uint32_t j = j0; j < j1; ++j) mine = lay_combine(mine, states[j]);
    LayState total;
    LayState run = lay_block_scan<kLayScanThreads>(mine, total, sm);
    for (uint32_t j = j0; j < j1; ++j)
    {
        const LayState s = states[j];
        states[j] = run;
        run = lay_combine(run, s);
    }
    if (threadIdx.x == 0) states[count] = total;
}

// first bin and first stream bytes of every chunk that holds records; entry [n_chunks] = bins and stream bytes of the whole batch
struct ChunkStart
{
    unsigned long long first_bin;
    unsigned long long off[4];
};

struct LayOut
{
    Placement pm;            // loc[4], binfo, tbase (place.cuh)
    OutStreams O;            // the streams: words shared by two K4 tiles are cleared here
    fsb_bin_descriptor* desc;
    ChunkStart* chunk_start; // [n_chunks + 1]
    uint32_t* nb_out;        // number of bins
};

__global__ void __launch_bounds__(kLayThreads) lay_apply_kernel(uint64_t n, uint32_t n_chunks, DeviceParams P, SortedView S, uint32_t uniform_len,
                                                                const LayState* __restrict__ tile_prefix, LayOut out)
{
    __shared__ LayState sm[kLayThreads / 32];
    const unsigned lane = threadIdx.x & 31;
    const LayRun run = lay_run(n, S);
    const LayState mine = lay_run_state(P, run, uniform_len);
    LayState total;
    const LayState excl = lay_combine(tile_prefix[blockIdx.x], lay_block_scan<kLayThreads>(mine, total, sm));
    LayCursor cur = lay_cursor(excl);
    const uint32_t key_mask = (1u << P.key_bits) - 1u;
    const bool has_head = P.has_headers != 0;
    // ---- the walk over the thread's records -------------------------------------------------------------------------
    uint64_t at[4][4];                                              // [record][stream]
    uint32_t binfo[4] = {0, 0, 0, 0};
    unsigned long long first_b0[4] = {0, 0, 0, 0};                  // first bit of a K4 tile whose first record is this thread's first record
    uint32_t prev = run.prev;
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) at[u][k] = 0;
        if ((uint32_t)u < run.cnt)
        {
            const uint64_t i = run.i0 + u;
            const uint32_t key = run.key[u];
            const uint32_t next = (uint32_t)u + 1u < run.cnt ? run.key[(u + 1) & 3] : run.next;
            const bool last_of_batch = i + 1 == n;
            const LayRec r = lay_record(P, i == 0, key, prev, run.card[u], uniform_len);
            lay_step(cur, r, at[u]);
            if (r.chunk_start)
            {
                ChunkStart cs;
                cs.first_bin = cur.nb - 1u;
#pragma unroll
                for (int k = 0; k < 4; ++k) cs.off[k] = cur.ls[k] >> 3;
                out.chunk_start[key >> P.key_bits] = cs;
            }
            if (u == 0)
            {   // a tile starts at its first record, or at the first byte of the bin that record opens (bin header and all)
#pragma unroll
                for (int k = 0; k < 4; ++k) first_b0[k] = r.start ? cur.ls[k] : at[0][k];
            }
            binfo[u] = (uniform_len & 0xFFu) | ((uniform_len & 0xFFu) << 8) | (r.start ? 0x10000u : 0u) | (r.nbin ? 0x20000u : 0u);
            if (last_of_batch || next != key) out.desc[cur.nb - 1u] = lay_descriptor(cur, key & key_mask);   // last record of its bin
            prev = key;
        }
    }
    // ---- K4 tiles: the eight lanes of a tile take its first bit from the first of them ------------------------------
    const uint64_t tile = run.i0 / kPlaceTile;
    const bool leader = (lane & (kLayTileLanes - 1u)) == 0 && run.cnt != 0;
    uint32_t loc[4][4];                                             // [stream][record]
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        const bool has = !(k == 3 && !has_head);
        const unsigned long long b0 = __shfl_sync(0xFFFFFFFFu, has ? first_b0[k] : 0ull, lane & ~(kLayTileLanes - 1u));
        if (leader)
        {
            out.pm.tbase[4 * tile + k] = b0;
            if (b0 & 31u) out.O.w[k][b0 >> 5] = 0;                  // write_out ORs into the word it shares with the tile in front
        }
        const uint32_t lo = (uint32_t)(b0 & 127u);                  // staging_bit: bits from the 16-byte group the tile starts in
#pragma unroll
        for (int u = 0; u < 4; ++u) loc[k][u] = has ? (uint32_t)(at[u][k] - b0) + lo : 0u;
    }
    if (run.cnt == kLayPerThread)
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(out.pm.loc[k] + run.i0) = make_uint4(loc[k][0], loc[k][1], loc[k][2], loc[k][3]);
        *reinterpret_cast<uint4*>(out.pm.binfo + run.i0) = make_uint4(binfo[0], binfo[1], binfo[2], binfo[3]);
    }
    else
    {
#pragma unroll
        for (int u = 0; u < 3; ++u)
            if ((uint32_t)u < run.cnt)
            {
#pragma unroll
                for (int k = 0; k < 4; ++k) out.pm.loc[k][run.i0 + u] = loc[k][u];
                out.pm.binfo[run.i0 + u] = binfo[u];
            }
    }
    if (run.cnt && run.i0 + run.cnt == n)
    {   // end of the batch: the end of the last tile, the totals
        const uint64_t tiles = (n + kPlaceTile - 1) / kPlaceTile;
        ChunkStart cs;
        cs.first_bin = cur.nb;
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const bool has = !(k == 3 && !has_head);
            const unsigned long long e = has ? roundup8(cur.pos[k]) : 0ull;
            out.pm.tbase[4 * tiles + k] = e;
            if (e & 31u) out.O.w[k][e >> 5] = 0;
            cs.off[k] = roundup8(cur.pos[k]) >> 3;
        }
        out.chunk_start[n_chunks] = cs;
        *out.nb_out = cur.nb;
    }
}

// ChunkSummary (layout.cuh) of every chunk from the chunk starts: a chunk without records takes the start of the next chunk
// that has some (or the end of the batch), so its sizes come out as zero.
__global__ void __launch_bounds__(128) chunk_summary_fused_kernel(BatchView Bv, const ChunkStart* __restrict__ cs, const fsb_bin_descriptor* __restrict__ desc,
                                                                  ChunkSummary* __restrict__ out)
{
    const uint32_t c = blockIdx.x;                 // one block per chunk
    auto start_of = [&](uint32_t ch) { while (ch < Bv.n_chunks && Bv.chunk_first_rec[ch] == Bv.chunk_first_rec[ch + 1]) ++ch; return cs[ch]; };
    const ChunkStart s0 = start_of(c), s1 = start_of(c + 1);
    __shared__ unsigned long long raw[2];
    if (threadIdx.x == 0) { raw[0] = 0; raw[1] = 0; }
    __syncthreads();
    unsigned long long a = 0, h = 0;
    for (uint64_t b = s0.first_bin + threadIdx.x; b < s1.first_bin; b += blockDim.x) { a += desc[b].raw_dna_size; h += desc[b].raw_head_size; }
    atomicAdd(&raw[0], a); atomicAdd(&raw[1], h);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        ChunkSummary s;
        s.first_bin = s0.first_bin; s.n_bins = s1.first_bin - s0.first_bin;
        for (int k = 0; k < 4; ++k) { s.off[k] = s0.off[k]; s.size[k] = s1.off[k] - s0.off[k]; }
        s.raw_dna = raw[0]; s.raw_head = raw[1];
        out[c] = s;
    }
}

} // namespace fsb
