// place.cuh -- K4: scatter the prepacked records into the four output streams (sm_100a).
//
// Replaces the sequential appends of FastqRecordsPackerSE/PE::StoreRecords / StoreNextRecord into the
// four BitMemoryWriters and the per-bin framing of PackToBin (FastqPacker.cpp:113-153, 541-602,
// 734-759, 815-859; BitMemory.h:216-433).
//
// K1 (ingest.cuh) has left every record's quality, title and DNA bits in a 128-byte aligned slot in
// input order.  Bins are laid out back to back in the sorted order, so the T consecutive sorted
// records a block owns cover one contiguous bit range of every stream.  A block
//   1. looks up its records (sorted key + card, bin, bit offsets from the layout scans);
//   2. gathers their slots -- whole 128-byte lines -- into shared memory with cp.async;
//   3. funnel-shifts every segment to its bit phase into per-stream staging buffers (one thread per
//      (record, segment); the record's meta fields and the 17-bit bin headers are produced here);
//   4. writes the staging buffers out with coalesced 16-byte stores; only the first / last word of a
//      tile, shared with the neighbouring tiles, is merged with atomicOr into pre-zeroed words.
#pragma once

#include "layout.cuh"
#include "pack_core.cuh"

namespace fsb {

struct OutStreams
{
    uint32_t* w[4];          // meta, dna, qua, head as 32-bit words (tile-boundary words zero-initialised)
};

struct PlaceArgs
{
    BatchView B;
    DeviceParams P;
    SlotGeom G;
    SortedView S;
    BinArrays A;
    StreamScans SC;
    BinOffsets BO;
    OutStreams O;
    const uint32_t* slots;   // [n_records][G.words]
    const uint32_t* nb_ptr;  // number of bins (device)
};

// bit offset of the record at sorted position i in stream s
__device__ __forceinline__ uint64_t stream_offset(const PlaceArgs& a, int s, uint64_t i, uint32_t bin, uint64_t start)
{
    return 8ull * a.BO.B[s][bin] + (a.SC.P[s][i] - a.SC.P[s][start]) + (s == 0 ? 17ull : 0ull);
}

// The tile's bit range in stream s: from its first record (or the start of that record's bin,
// header and all) to the same point of the next tile.
__device__ __forceinline__ void tile_range(const PlaceArgs& a, int s, uint64_t i0, uint32_t T, unsigned long long& b0, unsigned long long& b1)
{
    const uint64_t n = a.B.n_records;
    {
        const uint32_t bin = a.A.bin_of[i0];
        const uint64_t start = a.A.bin_start[bin];
        b0 = (i0 == start) ? 8ull * a.BO.B[s][bin] : stream_offset(a, s, i0, bin, start);
    }
    const uint64_t in = i0 + T;
    if (in < n)
    {
        const uint32_t bin = a.A.bin_of[in];
        const uint64_t start = a.A.bin_start[bin];
        b1 = (in == start) ? 8ull * a.BO.B[s][bin] : stream_offset(a, s, in, bin, start);
    }
    else b1 = 8ull * a.BO.B[s][*a.nb_ptr];
}

// ---- TMA bulk copies with mbarrier completion (sm_90+): one instruction moves a whole slot -------------------------------
#ifndef FSB_K4_BULK
#define FSB_K4_BULK 1
#endif
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one arrival that also announces `bytes` of asynchronous copies to come
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
                 ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is counted on the mbarrier
__device__ __forceinline__ void bulk_copy_g2s(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

#ifndef FSB_K4_TILE
#define FSB_K4_TILE 32
#endif
constexpr uint32_t kPlaceTile = FSB_K4_TILE;  // records per tile: one lane per record, one warp per role
constexpr uint32_t kPlaceRoles = 4;          // quality of mate A, quality of mate B, DNA, title + meta
constexpr uint32_t kPlaceThreads = 128;     // kPlaceRoles * kPlaceTile
static_assert(kPlaceTile == 32, "place_kernel maps the records of a tile onto the lanes of a warp");


// Write a tile's staging buffer to its stream.  Staging word j is stream word base + j with base a
// multiple of 4 (16-byte aligned), so whole groups of four go out as vector stores; the few words
// at both ends are handled one by one, and the first / last word are merged with atomicOr when
// they are shared with the neighbouring tile.  The ranges are worked out once per tile and stream
// (one thread) and read from shared memory by everybody.
struct WritePlan
{
    unsigned long long base;     // stream word of staging word 0
    uint32_t ws, we;             // first / last staging word with bits of this tile
    uint32_t vs, ve;             // 16-byte vectors [vs, ve) belong to this tile alone
    uint32_t nlo, hi_begin, nhi; // leftover words: [ws, ws + nlo) and [hi_begin, hi_begin + nhi)
    uint32_t shared;             // bit 0: word ws is shared with the previous tile, bit 1: word we with the next; bit 2: anything to write
};
__device__ __forceinline__ WritePlan make_write_plan(uint64_t b0, uint64_t b1)
{
    WritePlan w{};
    if (b1 <= b0) return w;
    w.base = (b0 >> 5) & ~3ull;
    w.ws = (uint32_t)((b0 >> 5) - w.base); w.we = (uint32_t)(((b1 - 1) >> 5) - w.base);
    const bool head_shared = (b0 & 31u) != 0, tail_shared = (b1 & 31u) != 0;
    const uint32_t fs = w.ws + (head_shared ? 1u : 0u), fe1 = w.we + 1u - (tail_shared ? 1u : 0u);  // words [fs, fe1) belong to this tile alone
    w.vs = (fs + 3u) >> 2; w.ve = fe1 >> 2;
    // leftovers: [ws, min(4 vs, we + 1)) and [max(4 ve, 4 vs), we]; at most 3 + 3 + 2 words
    const uint32_t lo_end = w.ve > w.vs ? 4u * w.vs : w.we + 1u;
    w.hi_begin = w.ve > w.vs ? 4u * w.ve : w.we + 1u;
    w.nlo = lo_end > w.ws ? lo_end - w.ws : 0u; w.nhi = w.we + 1u > w.hi_begin ? w.we + 1u - w.hi_begin : 0u;
    w.shared = (head_shared ? 1u : 0u) | (tail_shared ? 2u : 0u) | 4u;
    return w;
}
// Every word that is read is cleared behind the reader, so the staging is all zero again when the next tile's
// segments are ORed into it.
//
// The four streams of a tile differ a lot in size (quality : DNA : titles : meta is about 450 : 150 : 30 : 3 vectors at
// 2 x 150 bp), so the vectors of all four are dealt out together: they form one line of `total` vectors, warp w takes
// the w-th quarter of it and works through the (one or two) streams its quarter covers.  The few words at the ends of
// a stream's range that do not fill a vector -- and may be shared with the neighbouring tile -- go to warp s for stream s.
__device__ __forceinline__ void write_out_all(uint32_t* const (&stg)[4], const OutStreams& O, const WritePlan* __restrict__ wplan)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t nvec[4], total = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) { nvec[s] = (wplan[s].shared & 4u) ? wplan[s].ve - min(wplan[s].vs, wplan[s].ve) : 0u; total += nvec[s]; }
    const uint32_t lo = total * warp / kPlaceRoles, hi = total * (warp + 1u) / kPlaceRoles;      // this warp's part of the line
    uint32_t first = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s)
    {
        const uint32_t a = max(lo, first), b = min(hi, first + nvec[s]);                           // the part of it that lies in stream s
        if (a < b)
        {
            uint4* sv = reinterpret_cast<uint4*>(stg[s]);
            uint4* gv = reinterpret_cast<uint4*>(O.w[s] + wplan[s].base);
            const uint32_t shift = wplan[s].vs - first;
#pragma unroll 1
            for (uint32_t j = a + lane + shift; j < b + shift; j += 32)
            {
                uint4 v = sv[j];
                sv[j] = make_uint4(0, 0, 0, 0);
                v.x = bswap32(v.x); v.y = bswap32(v.y); v.z = bswap32(v.z); v.w = bswap32(v.w);
                gv[j] = v;
            }
        }
        first += nvec[s];
    }
    {
        const WritePlan w = wplan[warp];                              // stream `warp`: its leftover words
        if ((w.shared & 4u) && lane < w.nlo + w.nhi)
        {
            uint32_t* st = warp == 0 ? stg[0] : (warp == 1 ? stg[1] : (warp == 2 ? stg[2] : stg[3]));
            uint32_t* g = (warp == 0 ? O.w[0] : (warp == 1 ? O.w[1] : (warp == 2 ? O.w[2] : O.w[3]))) + w.base;
            const uint32_t j = lane < w.nlo ? w.ws + lane : w.hi_begin + (lane - w.nlo);
            const uint32_t v = bswap32(st[j]);
            st[j] = 0;
            if ((j == w.ws && (w.shared & 1u)) || (j == w.we && (w.shared & 2u))) atomicOr(g + j, v);
            else g[j] = v;
        }
    }
}
// bit offset of a global stream position inside the tile's staging buffer
__device__ __forceinline__ uint32_t staging_bit(uint64_t off, uint64_t tile_b0) { return (uint32_t)(off - (((tile_b0 >> 5) & ~3ull) << 5)); }

// ---- shared-memory plan, computed on the host from the batch statistics --------------------------------------
struct PlacePlan
{
    uint32_t T;              // records per tile
    uint32_t threads;        // kPlaceRoles * T
    uint32_t slot_stride;    // words per slot in shared memory (odd number of 16-byte units: conflict-free 16-byte reads)
    uint32_t slot_bytes;     // one buffer of T slots (there are two: the next tile's slots arrive while this one is placed)
    uint32_t off_plan, off_slots, off_staging[4], staging_bytes, total_bytes;
    uint32_t tiles_per_block;  // 0: persistent blocks striding over all tiles; R > 0: block b owns the tiles [b * R, (b + 1) * R)
};
inline uint32_t staging_words(uint32_t T, uint32_t bits_per_record) { return ((T * bits_per_record + 31u) / 32u + 6u + 3u) & ~3u; }

struct RecPlan
{
    uint32_t loc[4];         // bit offset of the record inside the tile's staging buffers (meta, dna, qua, head)
    uint32_t nbits[4];
    uint32_t meta_val;
    uint32_t bin_header;     // 0, or 0x80000000 | the bin's 17 header bits when the record opens its bin
    uint32_t qa_bits;        // quality bits of stored mate A (mate B's follow in the stream)
};

inline PlacePlan make_place_plan(const DeviceParams& P, const SlotGeom& G, uint32_t max_len, uint32_t max_head)
{
    PlacePlan pl{};
    const uint32_t mates = P.paired ? 2u : 1u;
    pl.slot_stride = (G.qw + G.tw + 3u) & ~3u;
    if ((pl.slot_stride & 7u) == 0) pl.slot_stride += 4u;
    pl.T = kPlaceTile;
    pl.threads = kPlaceRoles * pl.T;
    uint32_t o = 0;
    pl.off_plan = o; o += 2u * pl.T * (uint32_t)sizeof(RecPlan);            // two plans: warps still writing a tile out must not see the next one
    o = (o + 15u) & ~15u;
    pl.slot_bytes = pl.T * pl.slot_stride * 4u + 16u;                                   // + slack: shift_copy may read a few words past a segment
    pl.off_slots = o; o += 2u * pl.slot_bytes;
    pl.off_staging[0] = o; o += staging_words(pl.T, 28u + 17u) * 4u;
    pl.off_staging[1] = o; o += staging_words(pl.T, mates * max_len * 3u + 7u) * 4u;
    pl.off_staging[2] = o; o += staging_words(pl.T, mates * max_len * P.qua_bits + 7u) * 4u;
    pl.off_staging[3] = o; o += staging_words(pl.T, P.has_headers ? 8u + 7u * (max_head ? max_head - 1u : 0u) + 7u : 0u) * 4u;
    pl.staging_bytes = o - pl.off_staging[0];
    pl.total_bytes = o;
    return pl;
}

// ---- placement tables -------------------------------------------------------------------------------------------
// Everything K4 has to look up per record is worked out beforehand by a streaming kernel with
// plenty of threads to hide the dependent loads (bin of the record -> start of the bin -> scans), so
// that K4 itself only reads arrays indexed by the sorted position:
//   tbase[tile][s]   first bit of the tile in stream s (the bin's first byte when its first record opens a bin);
//                    entry [tiles] is the end of the stream
//   loc[s][i]        bit offset of record i inside its tile's staging buffer of stream s
//   binfo[i]         bin minLen | bin maxLen << 8 | record opens its bin << 16 | N-bin << 17
struct Placement
{
    uint32_t* loc[4];
    uint32_t* binfo;
    unsigned long long* tbase;
};

// One kernel, one warp per tile of T = 32 sorted records: every lane works out its record's stream positions,
// lane 0's record gives the tile's first bit (the bin's first byte when the record opens a bin), which is also where
// the tile-boundary word may have to be cleared: write_out merges the first / last word of a tile with atomicOr when
// it is shared with the neighbouring tile; those words -- and only those -- must be zero beforehand (this replaces
// a memset of the whole output).  The last warp also writes the end-of-stream entry tbase[tiles].
__global__ void __launch_bounds__(256) placement_kernel(PlaceArgs a, Placement pm, uint64_t tiles)
{
    const uint64_t n = a.B.n_records;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;       // blockDim is a multiple of the tile size
    const uint64_t tile = i / kPlaceTile;
    if (tile >= tiles) return;                                                  // whole warps
    const unsigned lane = threadIdx.x & 31;
    const bool live = i < n;
    uint32_t bin = 0;
    uint64_t start = 0;
    if (live) { bin = a.A.bin_of[i]; start = a.A.bin_start[bin]; }
#pragma unroll
    for (int s = 0; s < 4; ++s)
    {
        const bool has = !(s == 3 && !a.P.has_headers);
        unsigned long long off = 0, b0 = 0;
        if (live && has)
        {
            off = stream_offset(a, s, i, bin, start);
            b0 = (i == start) ? 8ull * a.BO.B[s][bin] : off;                    // meaningful on lane 0: the tile's first bit
        }
        b0 = __shfl_sync(0xFFFFFFFFu, b0, 0);
        if (live) pm.loc[s][i] = has ? staging_bit(off, b0) : 0u;
        if (lane == 0)
        {
            pm.tbase[4 * tile + s] = b0;
            if (b0 & 31u) a.O.w[s][b0 >> 5] = 0;
            if (tile + 1 == tiles)
            {   // end of the stream
                const unsigned long long e = has ? 8ull * a.BO.B[s][*a.nb_ptr] : 0ull;
                pm.tbase[4 * tiles + s] = e;
                if (e & 31u) a.O.w[s][e >> 5] = 0;
            }
        }
    }
    if (live)
    {
        const bool nbin = (a.S.skeys[i] & ((1u << a.P.key_bits) - 1u)) == a.P.nbin;
        pm.binfo[i] = (a.A.bin_min[bin] & 0xFFu) | ((a.A.bin_max[bin] & 0xFFu) << 8) | (i == start ? 0x10000u : 0u) | (nbin ? 0x20000u : 0u);
    }
}

// ---- K4 --------------------------------------------------------------------------------------------------------------
// Persistent blocks of four warps; warp = role (quality A, quality B, DNA, title + meta), lane = record of
// the tile.  Per tile a thread needs three coalesced values (card, loc, binfo); they are loaded two tiles
// ahead, the slot gather of the next tile is issued before the current tile is placed (two slot buffers),
// so the long latencies overlap with the shifting and the write-out.
struct TileRegs
{
    unsigned long long card;
    uint32_t loc, binfo;
    unsigned long long tb0, tb1;     // the tile's bit range in the thread's stream
};

__global__ void __launch_bounds__(kPlaceRoles * kPlaceTile) place_kernel(PlaceArgs a, PlacePlan pl, Placement pm, uint64_t tiles)
{
    extern __shared__ uint4 place_smem[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(place_smem);
    __shared__ WritePlan wplans[2][4];
#if FSB_K4_BULK
    __shared__ uint64_t slots_ready[2];          // one mbarrier per slot buffer: completes when all slots of a tile have landed
    if (threadIdx.x == 0) { mbar_init(&slots_ready[0], 1); mbar_init(&slots_ready[1], 1); mbar_init_fence(); }
    __syncthreads();
    uint32_t ready_parity = 0;                   // bit b: parity of the phase buffer b completes next
#endif
    const DeviceParams& P = a.P;
    const SlotGeom& G = a.G;
    constexpr uint32_t T = kPlaceTile;
    const uint32_t tid = threadIdx.x, lane = tid & 31, role = tid >> 5;
    const uint64_t n = a.B.n_records;
    RecPlan* plans = reinterpret_cast<RecPlan*>(smem + pl.off_plan);
    uint32_t* stg[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) stg[s] = reinterpret_cast<uint32_t*>(smem + pl.off_staging[s]);
    const uint32_t npieces = (G.qw + G.tw + 3u) >> 2;
    const uint32_t* my_loc = role == 0 ? pm.loc[0] : (role == 1 ? pm.loc[1] : (role == 2 ? pm.loc[2] : pm.loc[3]));

    // persistent: tiles blockIdx.x, + gridDim.x, ..; otherwise a run of consecutive tiles per block (see ingest_kernel)
    const uint64_t stride = pl.tiles_per_block ? 1u : gridDim.x;
    const uint64_t tile_end = pl.tiles_per_block ? min(tiles, ((uint64_t)blockIdx.x + 1u) * pl.tiles_per_block) : tiles;
    auto load_tile = [&](uint64_t tile) -> TileRegs
    {
        TileRegs r{};
        if (tile < tile_end)
        {
            const uint64_t i = tile * T + lane;
            if (i < n) { r.card = a.S.cards[i]; r.loc = my_loc[i]; r.binfo = pm.binfo[i]; }
            r.tb0 = pm.tbase[4 * tile + role]; r.tb1 = pm.tbase[4 * tile + 4 + role];
        }
        return r;
    };
#if FSB_K4_BULK
    // One bulk copy (TMA, cp.async.bulk) per slot: lane q of warp (q mod 4) asks for the slot of the tile's record q; the copies
    // report to the buffer's mbarrier, on which one thread has announced the tile's bytes.
    auto gather = [&](uint64_t tile, const TileRegs& r, uint32_t buf)
    {
        if (tile < tile_end)
        {
            const uint32_t ntile = (uint32_t)min((uint64_t)T, n - tile * T);
            const uint32_t bytes = npieces * 16u;
            if (tid == 0) mbar_arrive_expect_tx(&slots_ready[buf], ntile * bytes);
            if (lane < ntile && (lane & (kPlaceRoles - 1u)) == role)
            {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem + pl.off_slots + (size_t)buf * pl.slot_bytes) + lane * pl.slot_stride * 4u;
                bulk_copy_g2s(dst, a.slots + (uint64_t)card_rec(r.card) * G.words, bytes, &slots_ready[buf]);
            }
        }
    };
#else
    // warp w fetches the slots of records w, w + 4, .. of the tile: lane p copies 16-byte piece p, p + 32, ..
    auto gather = [&](uint64_t tile, const TileRegs& r, uint32_t buf)
    {
        if (tile < tile_end)
        {
            const uint32_t ntile = (uint32_t)min((uint64_t)T, n - tile * T);
            const uint32_t my_rec = card_rec(r.card);
            uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem + pl.off_slots + (size_t)buf * pl.slot_bytes) + 16u * lane + role * pl.slot_stride * 4u;
            const uint4* const slots4 = reinterpret_cast<const uint4*>(a.slots) + lane;
            const uint32_t pieces_per_slot = G.words >> 2;
#pragma unroll 1
            for (uint32_t q = role; q < ntile; q += kPlaceRoles)
            {
                const uint32_t rec = __shfl_sync(0xFFFFFFFFu, my_rec, q);
                const uint4* src = slots4 + (uint64_t)rec * pieces_per_slot;
                cp_async16_if(lane < npieces, sa, src);
                if (npieces > 32u)
                {
#pragma unroll 1
                    for (uint32_t pc = lane + 32u; pc < npieces; pc += 32) cp_async16_s(sa + 16u * (pc - lane), src + (pc - lane));
                }
                sa += kPlaceRoles * pl.slot_stride * 4u;
            }
        }
        cp_async_commit();
    };

#endif
    {   // the staging starts out zero; afterwards write_out clears what it reads
        uint4* z = reinterpret_cast<uint4*>(smem + pl.off_staging[0]);
        for (uint32_t j = tid; j < (pl.staging_bytes >> 4); j += blockDim.x) z[j] = make_uint4(0, 0, 0, 0);
    }
    uint64_t tile = pl.tiles_per_block ? (uint64_t)blockIdx.x * pl.tiles_per_block : (uint64_t)blockIdx.x;
    TileRegs cur = load_tile(tile), nxt = load_tile(tile + stride);
    uint32_t buf = 0;
    gather(tile, cur, buf);
    for (; tile < tile_end; tile += stride)
    {
        gather(tile + stride, nxt, buf ^ 1u);                         // in flight while this tile is placed
        const TileRegs nn = load_tile(tile + 2 * stride);             // used in the next round
        const uint64_t i0 = tile * T;
        const uint32_t ntile = (uint32_t)min((uint64_t)T, n - i0);
        const bool live = lane < ntile;
        const uint32_t* slot_buf = reinterpret_cast<const uint32_t*>(smem + pl.off_slots + (size_t)buf * pl.slot_bytes);

        // ---- 1. the tile's plan (in the plan buffer the previous tile did not use) -----------------------------------------
        RecPlan* plan = plans + buf * T;
        WritePlan* wplan = wplans[buf];
        if (lane == 0) wplan[role] = make_write_plan(cur.tb0, cur.tb1);
        if (live)
        {
            RecPlan& q = plan[lane];
            q.loc[role] = cur.loc;
            if (role == 0)
            {
                const uint32_t bmin = cur.binfo & 0xFFu, bmax = (cur.binfo >> 8) & 0xFFu;
                const bool nbin = (cur.binfo & 0x20000u) != 0;
                const uint32_t info = card_info(cur.card), lenA = card_lenA(cur.card), lenB = card_lenB(cur.card), H = card_head(cur.card);
                const ReadBits rb = read_bit_lengths(P, nbin, info, lenA, lenB, H, bmin, bmax);
                q.nbits[0] = rb.meta; q.nbits[1] = rb.dna; q.nbits[2] = rb.qua; q.nbits[3] = rb.head;
                uint32_t mbits;
                q.meta_val = meta_fields(P, nbin, info, lenA, lenB, bmin, bmax, mbits);
                q.bin_header = (cur.binfo & 0x10000u) ? (0x80000000u | (bmin << 9) | (bmax << 1)) : 0u;   // PackToBin (FastqPacker.cpp:581-583)
                q.qa_bits = lenA * P.qua_bits;
            }
        }
#if FSB_K4_BULK
        mbar_wait(&slots_ready[buf], (ready_parity >> buf) & 1u);      // this tile's slots have arrived (the next tile's may still travel)
        ready_parity ^= 1u << buf;
#else
        cp_async_wait_group1();                                        // this tile's slots have arrived (the next tile's may still travel)
#endif
        __syncthreads();

        // ---- 2. every (record, segment) to its bit phase ---------------------------------------------------------------------
        if (live)
        {
            const RecPlan& q = plan[lane];
            const uint32_t* slot = slot_buf + (size_t)lane * pl.slot_stride;
            if (role == 0) shift_copy_aligned(slot, q.qa_bits, stg[2], q.loc[2]);
            else if (role == 1) shift_copy_aligned(slot + G.wqa, q.nbits[2] - q.qa_bits, stg[2], q.loc[2] + q.qa_bits);
            else if (role == 2) shift_copy(slot + G.qw, q.nbits[3], q.nbits[1], stg[1], q.loc[1]);      // the DNA follows the title
            else
            {
                if (q.bin_header) or_bits(stg[0], q.loc[0] - 17u, q.bin_header & 0x1FFFFu, 17);
                or_bits(stg[0], q.loc[0], q.meta_val, q.nbits[0]);
                if (P.has_headers) shift_copy_aligned(slot + G.qw, q.nbits[3], stg[3], q.loc[3]);
            }
        }
        __syncthreads();

        // ---- 3. out ------------------------------------------------------------------------------------------------------------
        write_out_all(stg, a.O, wplan);
        // no barrier here: the next round only touches the other plan buffer before its own barrier, and nothing is
        // ORed into the staging before every warp has passed that barrier, i.e. has finished writing this tile out
        cur = nxt; nxt = nn; buf ^= 1u;
    }
#if !FSB_K4_BULK
    cp_async_wait_all();
#endif
}

} // namespace fsb
