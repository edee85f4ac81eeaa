// ingest.cuh -- K1: the one pass over the FASTQ text (sm_100a).
//
// Replaces, per record: FastqCategorizerBase::FindMinimizer x2 (SE) / x4 (PE), FastqRecord::ComputeRC
// and the selection logic of FastqCategorizerSE/PE::DistributeToBins (FastqCategorizer.cpp:79-106,
// 197-253, 256-363; FastqRecord.h:80-111), and the symbol coding of IFastqPacker::StoreDna /
// StoreQuality / StoreHeader (FastqPacker.cpp:157-287).
//
// B200 fills whole 128-byte lines from HBM on every read miss (profiles/r01b_dram_granularity_
// microbench.txt), so a record's sequence, quality and title cannot be gathered again later without
// paying for the whole text a second time.  K1 therefore does everything that needs the text while
// the lines are there: it finds the signature, decides orientation and mate order, and immediately
// codes the record's DNA, quality and title bits in the *stored* orientation into a 128-byte
// aligned per-record slot (core.cuh: SlotGeom).  What remains for K4 is to gather whole slots by
// sorted index and shift them to their final bit position.
//
// Mapping: one thread per mate (SE: per read; PE: lanes 2i / 2i+1 hold mate 1 / mate 2 of pair i),
// one warp per 32 mates, warps independent.  Sequence, title and quality are copied into per-thread
// shared-memory windows with cp.async 16-byte pieces (the quality reuses the sequence window once
// the sequence has been consumed); the slots of a warp are assembled in a shared-memory staging
// area and leave as one contiguous, fully coalesced block.
#pragma once

#include <cuda_runtime.h>

#include "sig_core.cuh"
#include "pack_core.cuh"

namespace fsb {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t smem_addr, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gmem_src) : "memory");
}
// the same, predicated inside the instruction (no branch, no reconvergence bookkeeping around it)
__device__ __forceinline__ void cp_async16_if(bool pred, uint32_t smem_addr, const void* gmem_src)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p cp.async.cg.shared.global [%0], [%1], 16;\n}\n"
                 ::"r"(smem_addr), "l"(gmem_src), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t shared_addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(shared_addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t shared_addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(shared_addr) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <int NW> constexpr uint32_t win_pieces() { return 2 * NW + 1; }      // 16-byte pieces of the aligned window around L <= 32 NW bytes
// An odd number of pieces per window keeps the lanes' windows on different banks.  A lane's pieces
// start 16 bytes into its window slot, i.e. they end 16 bytes into the next lane's slot: what a
// reversed reader sees when it looks 4 bytes below its first piece is the previous lane's last
// piece (or the front pad), which is harmless.
template <int NW> constexpr uint32_t win_slot_bytes() { return win_pieces<NW>() * 16; }

#ifndef FSB_K1_WARPS
#define FSB_K1_WARPS 8
#endif
#ifndef FSB_K1_MINBLOCKS
#define FSB_K1_MINBLOCKS 2
#endif
constexpr uint32_t kIngestMaxWarps = FSB_K1_WARPS;      // warps per block (launch bound)

// The lanes of a warp read windows that other lanes' cp.async requests have filled: every lane waits for its own requests
// (cp.async.wait_group 0), then the warp synchronises.  compute-sanitizer's racecheck does not take __syncwarp() as ordering
// for asynchronous copies and reports the reads that follow; the diagnostic build -DFSB_K1_NAMED_BARRIER replaces exactly
// these two synchronisations by a 32-thread named barrier (bar.sync, one id per warp) to show that nothing else is flagged.
__device__ __forceinline__ void sync_after_window_copies(unsigned warp)
{
#ifdef FSB_K1_NAMED_BARRIER
    asm volatile("bar.sync %0, 32;" ::"r"(warp + 1u) : "memory");
#else
    (void)warp;
    __syncwarp();
#endif
}

// The first radix pass of the sort needs the digit counts of every sort tile (scan_sort.cuh).  K1 knows every key the moment
// it writes it, so it counts them itself (one L2 reduction per record) and the pass needs no histogram kernel: counts (zeroed by
// the caller, null = off) in the layout [chunk][digit][tile of the chunk]; chunk_tiles[c] = sort tiles of the chunks in front of c.
struct SortSeed
{
    uint32_t* counts;
    const uint32_t* chunk_tiles;     // [n_chunks + 1]
    uint32_t mask, radix, tile;      // digit mask and radix of the first pass, keys per sort tile
};

struct IngestPlan
{
    uint32_t warps;          // warps per block
    uint32_t stg_stride;     // words per record in the slot staging: the slot's used words rounded up to 4, an odd number of 16-byte units
    uint32_t head_pieces;    // 16-byte pieces per title window incl. the guard piece (0: no titles)
    uint32_t off_head, off_staging, off_lut, off_chunks, total_bytes;
    uint32_t blocks_per_sm;
    uint32_t batches_per_warp;   // 0: persistent warps striding over the whole batch; R > 0: block b owns the warp batches [b * warps * R, (b + 1) * warps * R)
};

// the bit-spreading tables of pack_core.cuh; every block copies them into shared memory
__device__ const SpreadLut g_spread_lut = SpreadLut();

// Shared memory of a block of `warps` warps:
//   sequence / quality windows   32 per warp; once the qualities are packed they also stage the slots' quality regions
//   title windows                one per record
//   slot staging                 one slot per record (quality regions, title + DNA region): the slots leave as whole lines
//   spreading tables, chunk tables
template <int NW>
inline IngestPlan make_ingest_plan(const DeviceParams& P, const SlotGeom& G, uint32_t max_head, uint32_t smem_per_sm = 227u * 1024u)
{
    IngestPlan best{};
    uint32_t best_warps_per_sm = 0;
    const uint32_t recs = P.paired ? 16u : 32u;                      // records per warp
    for (uint32_t warps = kIngestMaxWarps; warps >= 1; warps >>= 1)
    {
        IngestPlan pl{};
        pl.warps = warps;
        pl.head_pieces = P.has_headers ? ((15u + max_head + 15u) >> 4) + 1u : 0u;
        pl.stg_stride = (G.qw + G.tw + 3u) & ~3u;
        if ((pl.stg_stride & 7u) == 0) pl.stg_stride += 4u;
        uint32_t o = 16;                                             // front pad: reversed readers may look 4 bytes below a window
        o += warps * 32u * win_slot_bytes<NW>() + 16u;               // the last lane's pieces end 16 bytes past its slot
        pl.off_head = o;
        o += warps * recs * pl.head_pieces * 16u;
        o += 32;                                                     // back pad: forward readers run up to 19 bytes past a window
        pl.off_staging = o;
        o += warps * recs * pl.stg_stride * 4u;
        o = (o + 15u) & ~15u;
        pl.off_lut = o;
        o += (uint32_t)sizeof(SpreadLut);
        pl.off_chunks = o;
        pl.total_bytes = o;
        pl.blocks_per_sm = std::min(std::min(smem_per_sm / (o + 1024u), 2048u / (warps * 32u)), 32u);
        const uint32_t wps = pl.blocks_per_sm * warps;
        if (wps > best_warps_per_sm) { best = pl; best_warps_per_sm = wps; }
    }
    return best;
}

// Every lane copies the aligned window around its own span, one 16-byte piece per step.  (Spreading a
// span over several lanes would coalesce the requests, but costs two shuffles and an index
// computation per piece; and every line is still fetched from HBM exactly once.)
// The windows of a warp are copied by the warp together: GL lanes per window, lane k of a group copies
// pieces k, k + GL, ..., so that one instruction requests runs of 16 GL bytes (32 / GL windows at a time) instead of
// 32 separate half-used sectors -- the memory side of K1 is bound by L2 requests, not by bytes.  Fewer lanes per
// window mean fewer trips through the loop (its shuffles and address arithmetic are paid per trip) but shorter
// runs per request; FSB_K1_GATHER_LANES picks the balance (measured: profiles/).  Must be called by all lanes of the warp.
#ifndef FSB_K1_GATHER_LANES
#define FSB_K1_GATHER_LANES 4
#endif
template <int NW>
__device__ __forceinline__ void gather_window(uint8_t* my_window, uint64_t piece0, uint32_t npieces, const uint8_t* text)
{
    constexpr unsigned GL = FSB_K1_GATHER_LANES;                           // lanes per window
    constexpr unsigned PPL = (win_pieces<NW>() + GL - 1) / GL;             // pieces per lane
    static_assert(GL == 2 || GL == 4 || GL == 8 || GL == 16, "lanes per window: a power of two");
    const unsigned lane = threadIdx.x & 31, lg = lane & (GL - 1u), grp = lane / GL;
    const unsigned long long my_src = (unsigned long long)(text + (piece0 << 4));
    // shared address of the first piece (below 2^24) and the number of pieces in one word
    const uint32_t my_dst = ((uint32_t)__cvta_generic_to_shared(my_window) + 16u) | (min(npieces, win_pieces<NW>()) << 24);
#pragma unroll 1
    for (uint32_t w0 = 0; w0 < 32u; w0 += 32u / GL)
    {
        const unsigned from = w0 + grp;
        const unsigned long long src = __shfl_sync(0xFFFFFFFFu, my_src, from);
        const uint32_t dn = __shfl_sync(0xFFFFFFFFu, my_dst, from);
        const uint32_t np = dn >> 24, dst = (dn & 0xFFFFFFu) + 16u * lg;
        const uint8_t* sp = reinterpret_cast<const uint8_t*>(src) + 16u * lg;
#ifndef FSB_EXP_NOGATHER
#pragma unroll
        for (unsigned q = 0; q < PPL; ++q) cp_async16_if(lg + q * GL < np, dst + 16u * q * GL, sp + 16u * q * GL);
#else
        (void)dst; (void)sp; (void)np;
#endif
    }
}

// What a lane knows about its mate of one warp batch before the text arrives: the record table
// entry as loaded (nothing is derived from it before it is needed, so the load stays in flight).
struct MateMeta
{
    uint4 rec;               // fsb_record: head_off, seq_off, qua_off, seq_len | head_len << 16
    uint64_t text_base;      // of the mate's chunk inside text[m]
    uint32_t ch;
    bool live;
};

// The chunk tables: first record of every chunk and the chunks' text offsets (shared-memory copies for small batches).
struct ChunkTables
{
    const uint64_t* first;                // [n_chunks + 1]
    const uint64_t* text_base[2];         // [n_chunks]
    uint32_t n_chunks;
};
__device__ __forceinline__ uint32_t chunk_of(const ChunkTables& T, uint64_t i)
{
    uint32_t lo = 0, hi = T.n_chunks;      // invariant: first[lo] <= i < first[hi]
    while (hi - lo > 1)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (T.first[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// K1.  NW = ceil(longest read of the batch / 32).
//   keys[i]  = chunk : signature            cards[i] = card_make(...)  (core.cuh)
//   slots    = [n_records][G.words] words   sig_out / info_out: optional per-read output (parity tests)
// Q = quality bits per symbol (6, 3 or 1).
//
// Persistent warps: a warp walks over warp batches of 32 mates (stride = warps of the grid); the record
// table entries of the next batch are loaded while the signature search of the current one runs, and
// its sequence / title windows are requested as soon as the current quality regions have left.
// With pl.batches_per_warp = R > 0 the grid is not persistent: a block owns R * warps consecutive warp
// batches and the hardware hands blocks out as SM resources free up -- the mode used when another
// sub-batch's kernels share the GPU (a persistent grid that could not start all its blocks at once would
// end with a long tail).
template <int NW, int Q>
__global__ void __launch_bounds__(kIngestMaxWarps * 32, FSB_K1_MINBLOCKS) ingest_kernel(BatchView B, DeviceParams P, SlotGeom G, IngestPlan pl, uint32_t* __restrict__ keys,
                                                                      unsigned long long* __restrict__ cards, uint32_t* __restrict__ slots,
                                                                      uint32_t* __restrict__ sig_out, uint32_t* __restrict__ info_out, SortSeed seed)
{
    constexpr uint32_t SB = win_slot_bytes<NW>();
    extern __shared__ uint4 ingest_smem[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(ingest_smem);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t recs_per_warp = P.paired ? 16u : 32u;
    uint8_t* warp_windows = smem + (size_t)warp * 32 * SB;
    uint8_t* my_window = warp_windows + (size_t)lane * SB;                            // pieces at + 16
    uint8_t* hslots = smem + pl.off_head + (size_t)warp * recs_per_warp * pl.head_pieces * 16u;
    uint32_t* stg = reinterpret_cast<uint32_t*>(smem + pl.off_staging) + (size_t)warp * recs_per_warp * pl.stg_stride;

    // ---- once per block: the spreading tables ------------------------------------------------------------------------
    {
        const uint4* src = reinterpret_cast<const uint4*>(g_spread_lut.v);
        uint4* dst = reinterpret_cast<uint4*>(smem + pl.off_lut);
        for (uint32_t j = threadIdx.x; j < sizeof(SpreadLut) / 16u; j += blockDim.x) dst[j] = src[j];
    }
    __syncthreads();
    // ---- the chunk tables: with at most 32 chunks lane c keeps chunk c's first record and text offsets in registers
    //      for the whole kernel, and a record's chunk is a ballot away; larger batches search the tables in global memory
    ChunkTables T;
    T.n_chunks = B.n_chunks; T.first = B.chunk_first_rec; T.text_base[0] = B.chunk_text_base[0]; T.text_base[1] = B.chunk_text_base[1];
    const bool chunks_in_lanes = B.n_chunks <= 32u;
    uint64_t c_first = ~0ull, c_tb0 = 0, c_tb1 = 0;
    uint32_t c_tiles0 = 0, c_tiles1 = 0;                                              // sort tiles in front of the lane's chunk / up to its end
    const bool seed_sort = seed.counts != nullptr && chunks_in_lanes;
    if (seed_sort && lane < B.n_chunks) { c_tiles0 = seed.chunk_tiles[lane]; c_tiles1 = seed.chunk_tiles[lane + 1]; }
    if (chunks_in_lanes && lane < B.n_chunks)
    {
        c_first = B.chunk_first_rec[lane];
        c_tb0 = B.chunk_text_base[0][lane];
        c_tb1 = P.paired ? B.chunk_text_base[1][lane] : 0ull;
    }
    const LutShared lut{(uint32_t)__cvta_generic_to_shared(smem + pl.off_lut)};

    const uint64_t n_mates = P.paired ? 2 * B.n_records : B.n_records;
    const uint64_t n_wb = (n_mates + 31u) >> 5;                                       // warp batches
    const uint32_t R = pl.batches_per_warp;
    const uint64_t wb_stride = R ? (uint64_t)pl.warps : (uint64_t)gridDim.x * pl.warps;
    const uint64_t wb_end = R ? min(n_wb, ((uint64_t)blockIdx.x + 1u) * pl.warps * R) : n_wb;
    uint64_t wb = (R ? (uint64_t)blockIdx.x * pl.warps * R : (uint64_t)blockIdx.x * pl.warps) + warp;
    if (wb >= wb_end) return;                                                         // (no block-wide barriers below)
    const unsigned m = P.paired ? (lane & 1u) : 0u;
    const uint32_t lrec = P.paired ? (lane >> 1) : lane;                              // record index inside the warp
    const uint8_t* my_text = m ? B.text[1] : B.text[0];
    const uint32_t* win_words = reinterpret_cast<const uint32_t*>(my_window);
    uint32_t* my_stage = stg + (size_t)lrec * pl.stg_stride;

    auto load_meta = [&](uint64_t batch) -> MateMeta
    {
        MateMeta mm{};
        const uint64_t g = batch * 32u + lane;
        mm.live = g < n_mates;
        if (mm.live)
        {
            const uint64_t i = P.paired ? (g >> 1) : g;
            mm.rec = *reinterpret_cast<const uint4*>((m ? B.rec[1] : B.rec[0]) + i);
        }
        if (chunks_in_lanes)
        {   // chunk of the batch's first record by ballot; the lanes' records lie in it or (rarely) in later chunks
            const uint64_t i0 = P.paired ? (batch * 16u) : (batch * 32u);
            const uint32_t ch0 = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, c_first <= i0)) - 1u;
            const uint64_t i = P.paired ? (g >> 1) : g;
            uint32_t ch = ch0;
            for (uint32_t c = ch0 + 1u; c < B.n_chunks; ++c)       // warp uniform trip count: until no lane moves on
            {
                const uint64_t f = __shfl_sync(0xFFFFFFFFu, c_first, c);
                if (!__any_sync(0xFFFFFFFFu, mm.live && f <= i)) break;
                if (mm.live && f <= i) ch = c;
            }
            const uint64_t tb0 = __shfl_sync(0xFFFFFFFFu, c_tb0, ch), tb1 = __shfl_sync(0xFFFFFFFFu, c_tb1, ch);
            mm.ch = ch; mm.text_base = m ? tb1 : tb0;
        }
        else if (mm.live)
        {
            const uint64_t i = P.paired ? (g >> 1) : g;
            mm.ch = chunk_of(T, i);
            mm.text_base = (m ? T.text_base[1] : T.text_base[0])[mm.ch];
        }
        return mm;
    };
    // the sequence windows and the title windows of one batch
    auto gather_seq_and_titles = [&](const MateMeta& mm)
    {
        const uint64_t seq_at = mm.text_base + mm.rec.y;
        const uint32_t L = mm.rec.w & 0xFFFFu, H = (m == 0 && P.has_headers) ? ((mm.rec.w >> 16) & 0xFFu) : 0u;
        gather_window<NW>(my_window, seq_at >> 4, mm.live ? ((uint32_t)(seq_at & 15u) + L + 15u) >> 4 : 0u, my_text);
        if (mm.live && H)
        {   // the lane of mate 1 copies the pieces of its record's title
            const uint64_t head_at = mm.text_base + mm.rec.x;
            const uint32_t np = min(((uint32_t)(head_at & 15u) + H + 15u) >> 4, pl.head_pieces - 1u);
            const uint8_t* src = B.text[0] + ((head_at >> 4) << 4);
            uint8_t* dst = hslots + (size_t)lrec * pl.head_pieces * 16u + 16u;
#pragma unroll 1
            for (uint32_t j = 0; j < np; ++j) cp_async16(dst + 16u * j, src + 16u * j);
        }
        cp_async_commit();
    };
    MateMeta cur = load_meta(wb);
    gather_seq_and_titles(cur);
    for (;;)
    {
        const uint64_t wb_next = wb + wb_stride;
        const bool more = wb_next < wb_end;                                           // warp uniform
        const bool live = cur.live;
        const uint32_t L = cur.rec.w & 0xFFFFu, H = (m == 0 && P.has_headers) ? ((cur.rec.w >> 16) & 0xFFu) : 0u;
        const uint64_t i = P.paired ? ((wb * 32u + lane) >> 1) : (wb * 32u + lane);   // record (pair) index
        const uint64_t rec0 = P.paired ? (wb * 16u) : (wb * 32u);
        const uint32_t nrec = (uint32_t)min((uint64_t)recs_per_warp, B.n_records - rec0);
        const uint64_t seq_at = cur.text_base + cur.rec.y, qua_at = cur.text_base + cur.rec.z;
        const uint32_t a_seq = (uint32_t)(seq_at & 15u), a_qua = (uint32_t)(qua_at & 15u), a_head = (uint32_t)((cur.text_base + cur.rec.x) & 15u);
        cp_async_wait_all();
        sync_after_window_copies(warp);

        // ---- bit planes of the sequence; from here on the windows belong to the qualities ----------------------------
        BV<NW> Hp, Lp, Np;
#pragma unroll
        for (int j = 0; j < NW; ++j) { Hp.w[j] = 0; Lp.w[j] = 0; Np.w[j] = 0; }
        if (live) mate_planes<NW>(win_words + 4 + (a_seq >> 2), 8u * (a_seq & 3u), L, Hp, Lp, Np);
        __syncwarp();                                                 // every lane is done with its sequence window
        gather_window<NW>(my_window, qua_at >> 4, live ? (a_qua + L + 15u) >> 4 : 0u, my_text);
        cp_async_commit();
        MateMeta nxt{};
        if (more) nxt = load_meta(wb_next);                           // in flight during the signature search

        // ---- signature ------------------------------------------------------------------------------------------------
        StrandMin f, r;
        uint32_t nN = 0;
        f.sig = r.sig = P.nbin; f.pos = r.pos = 0;
#ifndef FSB_EXP_NOCOMPUTE
        if (live) plane_minimizers<NW>(Hp, Lp, Np, L, P, f, r, nN);
#endif
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
            sig = __shfl_sync(0xFFFFFFFFu, sig, lane & ~1u);
            inf = __shfl_sync(0xFFFFFFFFu, inf, lane & ~1u);
        }
        const uint32_t Lother = P.paired ? __shfl_xor_sync(0xFFFFFFFFu, L, 1) : 0u;
        const uint32_t Hrec = P.paired ? __shfl_sync(0xFFFFFFFFu, H, lane & ~1u) : H;          // the record's title length (mate 1)
        const bool nbin = sig == P.nbin;
        const bool rev = (inf & FSB_INFO_REVERSE) != 0, swp = (inf & FSB_INFO_SWAPPED) != 0;
        // stored pair: forward [m1|m2]; reversed [rc(m2)|rc(m1)]; a swap exchanges the halves
        const bool a_is_m2 = P.paired && (rev != swp);
        const bool roleB = P.paired && ((m == 1) != a_is_m2);
        const uint32_t lenA = roleB ? Lother : L, lenB = P.paired ? (roleB ? L : Lother) : 0u;
        const bool plainA = (inf & FSB_INFO_PLAIN_A) != 0;
        const uint32_t sfx = nbin ? 0u : P.k;
        const uint32_t head_bits = P.has_headers ? 8u + 7u * (Hrec ? Hrec - 1u : 0u) : 0u;

        // ---- DNA of this mate in the stored orientation (StoreDna), straight from the planes: it follows the title ----------
        SegEmit ed = seg_open(my_stage + G.qw, 0, 0);
        if (live)
        {
            const uint32_t cut_len = roleB ? 0u : sfx, cut_pos = (roleB || nbin) ? 0u : (inf & FSB_INFO_POS_MASK);
            const uint32_t off = head_bits + (roleB ? (lenA - sfx) * (plainA ? 2u : 3u) : 0u);
            ed = seg_open(my_stage + G.qw, off, (L - cut_len) * (nN == 0 ? 2u : 3u));
#ifndef FSB_EXP_NOCOMPUTE
            pack_dna_planes<NW>(Hp, Lp, Np, L, rev, nN == 0, cut_pos, cut_len, lut, ed);
#endif
        }
        // ---- title, key and card --------------------------------------------------------------------------------------
        if (P.has_headers)
        {   // the two lanes of a pair take one part of the title each (a single-end lane takes both)
            const uint32_t* hw = reinterpret_cast<const uint32_t*>(hslots + (size_t)lrec * pl.head_pieces * 16u);
            const uint32_t a_head_rec = P.paired ? __shfl_sync(0xFFFFFFFFu, a_head, lane & ~1u) : a_head;
            if (live)
            {
                pack_head_part(hw, 16u + a_head_rec, Hrec, m, my_stage + G.qw);
                if (!P.paired) pack_head_part(hw, 16u + a_head_rec, Hrec, 1u, my_stage + G.qw);
            }
        }
        if (live && m == 0)
        {
            keys[i] = (cur.ch << P.key_bits) | sig;
            cards[i] = card_make((uint32_t)i, inf, lenA, lenB, H);
            if (sig_out) { sig_out[i] = sig; info_out[i] = inf; }
        }
        if (seed_sort)
        {   // the record's sort tile: its place inside its chunk / keys per tile (the chunk's tables sit in lane cur.ch)
            const uint32_t ch = live ? cur.ch : 0u;
            const uint64_t first = __shfl_sync(0xFFFFFFFFu, c_first, ch);
            const uint32_t t0 = __shfl_sync(0xFFFFFFFFu, c_tiles0, ch), t1 = __shfl_sync(0xFFFFFFFFu, c_tiles1, ch);
            if (live && m == 0)
                atomicAdd(&seed.counts[(uint64_t)seed.radix * t0 + (uint64_t)(sig & seed.mask) * (t1 - t0) + (uint32_t)((i - first) / seed.tile)], 1u);
        }
        cp_async_wait_all();
        sync_after_window_copies(warp);

        // ---- quality of this mate in the stored orientation (StoreQuality): every lane packs its mate's stream as
        //      16-byte vectors into the mate's quality region of the staged slot ----------------------------------------
        if (live) pack_quality_to<Q>(reader_open(win_words, 16u + a_qua, L, rev), L, P, my_stage + (roleB ? G.wqa : 0u));
        __syncwarp();                                                 // windows and title windows are free again
        if (more) gather_seq_and_titles(nxt);
        // word 0 of the DNA segments: mate A merges into the title's last word, then mate B into A's
        if (live && !roleB) seg_finish(ed, true);
        __syncwarp();
        if (live && roleB) seg_finish(ed, true);
        __syncwarp();
        {   // the warp's slots leave as whole lines: 8 lanes per record, 16-byte vectors; a lane's (up to four) vectors of a record
            // are loaded together and then stored, so that the stores do not each wait for a shared-memory load of their own
            const uint32_t nvec = (G.qw + G.tw + 3u) >> 2, sub = lane >> 3, l8 = lane & 7u;
            uint32_t sa = (uint32_t)__cvta_generic_to_shared(stg) + 16u * l8 + sub * pl.stg_stride * 4u;
            uint32_t* const batch_slots = slots + rec0 * G.words;
            uint32_t di = sub * G.words + 4u * l8;
#pragma unroll 1
            for (uint32_t r = sub; r < nrec; r += 4)
            {
#ifndef FSB_EXP_NOCOPYOUT
#pragma unroll 1
                for (uint32_t v = l8; v < nvec; v += 32)
                {
                    uint4 t[4];
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u) if (v + 8u * u < nvec) t[u] = lds128(sa + 16u * (v + 8u * u - l8));
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u) if (v + 8u * u < nvec) *reinterpret_cast<uint4*>(batch_slots + di + 4u * (v + 8u * u - l8)) = t[u];
                }
#endif
                sa += 16u * pl.stg_stride; di += 4u * G.words;
            }
        }
        if (!more) break;
        __syncwarp();                                                 // the staging is rewritten by the next batch
        cur = nxt; wb = wb_next;
    }
}

} // namespace fsb
