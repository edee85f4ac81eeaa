// pack.cuh -- K4: bit-pack and scatter the records of one tile of the bin-sorted order (sm_100a).
//
// Replaces FastqRecordsPackerSE/PE::StoreRecords, IFastqPacker::StoreNextRecord / StoreDna /
// StoreQuality / StoreHeader and BitMemoryWriter (FastqPacker.cpp:113-287, 734-759, 815-859;
// BitMemory.h:216-433).
//
// Three kernels, one per kind of stream (quality; DNA; meta + titles), so that each keeps its
// shared-memory footprint -- and with it the number of resident warps -- where a gather-heavy
// kernel needs it.  In each, a block owns T consecutive records of the sorted order.  Because bins
// are laid out back to back in that order, the tile's output is one contiguous bit range of the
// stream:
//   1. every thread looks up its record (sorted position -> record index, flags, bit offset);
//   2. the warps gather the aligned 16-byte pieces around each source quality / sequence / title
//      into shared-memory slots with cp.async (coalesced requests over whole sectors);
//   3. one thread per stored mate (per record for meta + titles) packs its segment into the tile's
//      staging buffer (pack_core.cuh);
//   4. the block writes the staging buffer to the stream with coalesced 16-byte stores; only the
//      first and last word of the tile, shared with the neighbouring tiles, are merged with
//      atomicOr into zero-initialised memory.
#pragma once

#include "layout.cuh"
#include "pack_core.cuh"

namespace fsb {

struct OutStreams
{
    uint32_t* w[4];          // meta, dna, qua, head as 32-bit words (boundary words zero-initialised)
};

struct PackArgs
{
    BatchView B;
    DeviceParams P;
    SortedView S;
    BinArrays A;
    StreamScans SC;
    BinOffsets BO;
    OutStreams O;
    const uint32_t* nb_ptr;  // number of bins (device)
};

// ---- what every pack kernel needs to know about one sorted position -----------------------------------
struct TileRec
{
    uint32_t info, bin, bmin, bmax, lenA, lenB, H, ch;
    uint64_t start;                  // sorted position of the first record of the bin
    fsb_record r1, rA, rB;           // mate-1 record (title), records of the stored mates A and B
    bool nbin, first_of_bin, rev, a_is_m2;
};

__device__ __forceinline__ TileRec tile_lookup(const PackArgs& a, uint64_t i)
{
    const DeviceParams& P = a.P;
    TileRec t;
    const uint32_t r = a.S.perm[i], key = a.S.skeys[i];
    t.ch = key >> P.key_bits;
    t.nbin = (key & ((1u << P.key_bits) - 1u)) == P.nbin;
    t.info = a.S.info[r];
    t.bin = a.A.bin_of[i];
    t.start = a.A.bin_start[t.bin];
    t.first_of_bin = i == t.start;
    t.bmin = a.A.bin_min[t.bin]; t.bmax = a.A.bin_max[t.bin];
    t.rev = (t.info & FSB_INFO_REVERSE) != 0;
    const bool swp = (t.info & FSB_INFO_SWAPPED) != 0;
    // stored pair: forward [m1|m2]; reversed [rc(m2)|rc(m1)]; a swap exchanges the halves
    t.a_is_m2 = P.paired && (t.rev != swp);
    t.r1 = a.B.rec[0][r];
    fsb_record r2 = t.r1;
    if (P.paired) r2 = a.B.rec[1][r];
    t.rA = t.a_is_m2 ? r2 : t.r1;
    t.rB = t.a_is_m2 ? t.r1 : r2;
    t.lenA = t.rA.seq_len; t.lenB = P.paired ? t.rB.seq_len : 0u;
    t.H = P.has_headers ? t.r1.head_len : 0u;
    return t;
}

// bit offset of the record at sorted position i in stream s
__device__ __forceinline__ uint64_t stream_offset(const PackArgs& a, int s, uint64_t i, uint32_t bin, uint64_t start)
{
    return 8ull * a.BO.B[s][bin] + (a.SC.P[s][i] - a.SC.P[s][start]) + (s == 0 ? 17ull : 0ull);
}

// The tile's bit range in stream s: from its first record (or the start of that record's bin,
// header and all) to the same point of the next tile.
__device__ __forceinline__ void tile_range(const PackArgs& a, int s, uint64_t i0, uint32_t T, unsigned long long& b0, unsigned long long& b1)
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

__device__ __forceinline__ void cp_async16_pack(void* smem_dst, const void* gmem_src)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all_pack() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

// A warp copies the windows of its 32 lanes' spans into their slots: lane -> (span, piece) pairs,
// consecutive lanes on consecutive 16-byte pieces of the same span (coalesced requests).  A slot is
// one guard piece followed by the window.
template <int PW>
__device__ __forceinline__ void gather_spans(uint8_t* warp_slots, uint32_t slot_bytes, uint32_t piece0, uint32_t npieces_m /* npieces | m << 8 */,
                                             const uint8_t* text0, const uint8_t* text1)
{
    const unsigned lane = threadIdx.x & 31;
    uint32_t span = lane / PW, j = lane % PW;
#pragma unroll
    for (int it = 0; it < PW; ++it)                             // 32 * PW (span, piece) pairs, 32 per round
    {
        const uint32_t p0 = __shfl_sync(0xFFFFFFFFu, piece0, span);
        const uint32_t nm = __shfl_sync(0xFFFFFFFFu, npieces_m, span);
        if (j < (nm & 0xFFu))
            cp_async16_pack(warp_slots + span * slot_bytes + 16u + 16u * j, ((nm >> 8) ? text1 : text0) + ((uint64_t)(p0 + j) << 4));
        j += 32 % PW; span += 32 / PW;
        if (j >= PW) { j -= PW; span += 1; }
    }
}
// same with a run-time window size (titles)
__device__ __forceinline__ void gather_spans_rt(uint8_t* warp_slots, uint32_t slot_bytes, uint32_t pw, uint32_t piece0, uint32_t npieces, const uint8_t* text)
{
    const unsigned lane = threadIdx.x & 31;
    for (uint32_t idx = lane; idx < 32u * pw; idx += 32)
    {
        const uint32_t span = idx / pw, j = idx - span * pw;
        const uint32_t p0 = __shfl_sync(0xFFFFFFFFu, piece0, span);
        const uint32_t np = __shfl_sync(0xFFFFFFFFu, npieces, span);
        if (j < np) cp_async16_pack(warp_slots + span * slot_bytes + 16u + 16u * j, text + ((uint64_t)(p0 + j) << 4));
    }
}

__device__ __forceinline__ void zero_staging(uint8_t* begin, uint32_t bytes)
{
    uint4* st = reinterpret_cast<uint4*>(begin);
    for (uint32_t j = threadIdx.x; j < (bytes >> 4); j += blockDim.x) st[j] = make_uint4(0, 0, 0, 0);
}

// Write a tile's staging buffer to its stream.  Staging word j is stream word base + j with base a
// multiple of 4 (16-byte aligned), so whole groups of four go out as vector stores; the few words
// at both ends are handled one by one, and the first / last word are merged with atomicOr when
// they are shared with the neighbouring tile.
__device__ __forceinline__ void write_out(const uint32_t* stg, uint32_t* stream_words, uint64_t b0, uint64_t b1)
{
    if (b1 <= b0) return;
    const uint32_t tid = threadIdx.x;
    const uint64_t base = (b0 >> 5) & ~3ull;
    const uint32_t ws = (uint32_t)((b0 >> 5) - base), we = (uint32_t)(((b1 - 1) >> 5) - base);       // first / last word with bits of this tile
    const bool head_shared = (b0 & 31u) != 0, tail_shared = (b1 & 31u) != 0;
    const uint32_t fs = ws + (head_shared ? 1u : 0u), fe1 = we + 1u - (tail_shared ? 1u : 0u);      // words [fs, fe1) belong to this tile alone
    const uint32_t vs = (fs + 3u) >> 2, ve = fe1 >> 2;                                              // vectors [vs, ve)
    uint32_t* g = stream_words + base;
    const uint4* sv = reinterpret_cast<const uint4*>(stg);
    uint4* gv = reinterpret_cast<uint4*>(g);
    for (uint32_t j = vs + tid; j < ve; j += blockDim.x)
    {
        uint4 v = sv[j];
        v.x = bswap32(v.x); v.y = bswap32(v.y); v.z = bswap32(v.z); v.w = bswap32(v.w);
        gv[j] = v;
    }
    // leftovers: [ws, min(4 vs, we + 1)) and [max(4 ve, 4 vs), we]; at most 3 + 3 + 2 words
    const uint32_t lo_end = ve > vs ? 4u * vs : we + 1u, hi_begin = ve > vs ? 4u * ve : we + 1u;
    const uint32_t nlo = lo_end > ws ? lo_end - ws : 0u, nhi = we + 1u > hi_begin ? we + 1u - hi_begin : 0u;
    if (tid < nlo + nhi)
    {
        const uint32_t j = tid < nlo ? ws + tid : hi_begin + (tid - nlo);
        const uint32_t v = bswap32(stg[j]);
        if ((j == ws && head_shared) || (j == we && tail_shared)) atomicOr(g + j, v);
        else g[j] = v;
    }
}
// bit offset of a global stream position inside the tile's staging buffer
__device__ __forceinline__ uint32_t staging_bit(uint64_t off, uint64_t tile_b0) { return (uint32_t)(off - (((tile_b0 >> 5) & ~3ull) << 5)); }

// ---- shared-memory plans, computed on the host from the batch statistics ----------------------------------
constexpr uint32_t kPackThreads = 128;
template <int NW> constexpr uint32_t pack_slot_pieces() { return 2 * NW + 2; }     // guard piece + aligned window

struct PackPlan
{
    uint32_t T;              // records per tile
    uint32_t threads;
    uint32_t head_pieces;    // 16-byte pieces per title slot incl. the guard piece (aux kernel)
    uint32_t off_staging[2], staging_bytes, total_bytes;
};
inline uint32_t staging_words(uint32_t T, uint32_t bits_per_record) { return ((T * bits_per_record + 31u) / 32u + 6u + 3u) & ~3u; }

// quality / DNA kernels: one thread per stored mate
template <int NW>
inline PackPlan make_mate_plan(const DeviceParams& P, uint32_t max_len, uint32_t bits_per_symbol)
{
    PackPlan pl{};
    const uint32_t roles = P.paired ? 2u : 1u;
    pl.threads = kPackThreads; pl.T = kPackThreads / roles;
    for (;;)
    {
        uint32_t o = 16;                                                       // front pad: reversed readers may look 4 bytes below a slot
        o += pl.threads * pack_slot_pieces<NW>() * 16u;
        o += 32;                                                               // back pad: forward readers run up to 19 bytes past a window
        pl.off_staging[0] = o;
        pl.staging_bytes = staging_words(pl.T, bits_per_symbol * max_len * roles + 7u) * 4u;
        pl.total_bytes = o + pl.staging_bytes;
        if (pl.total_bytes <= 200u * 1024u || pl.threads <= 32u) break;
        pl.threads >>= 1; pl.T >>= 1;
    }
    return pl;
}
// aux kernel (meta fields + titles): one thread per record
inline PackPlan make_aux_plan(const DeviceParams& P, uint32_t max_head)
{
    PackPlan pl{};
    pl.threads = kPackThreads; pl.T = kPackThreads;
    pl.head_pieces = P.has_headers ? ((15u + max_head + 15u) >> 4) + 1u : 0u;
    uint32_t o = 16;
    o += pl.T * pl.head_pieces * 16u;
    o += 32;
    pl.off_staging[0] = o;
    o += staging_words(pl.T, 52u) * 4u;
    pl.off_staging[1] = o;
    o += staging_words(pl.T, P.has_headers ? 8u + 7u * (max_head ? max_head - 1u : 0u) + 7u : 0u) * 4u;
    pl.staging_bytes = o - pl.off_staging[0];
    pl.total_bytes = o;
    return pl;
}

// the stored mate a thread of the quality / DNA kernels owns: threads [0, T) hold mate A of record
// tid, threads [T, 2T) mate B of record tid - T
struct MyMate { uint32_t len, m; uint64_t seq_at, qua_at; bool plain, roleB; };
__device__ __forceinline__ MyMate my_mate(const PackArgs& a, const TileRec& t, bool roleB)
{
    MyMate mm;
    mm.roleB = roleB;
    mm.m = roleB ? (t.a_is_m2 ? 0u : 1u) : (t.a_is_m2 ? 1u : 0u);
    const fsb_record mine = roleB ? t.rB : t.rA;
    mm.len = mine.seq_len;
    mm.plain = (t.info & (roleB ? FSB_INFO_PLAIN_B : FSB_INFO_PLAIN_A)) != 0;
    const uint64_t tb = (mm.m ? a.B.chunk_text_base[1] : a.B.chunk_text_base[0])[t.ch];
    mm.seq_at = tb + mine.seq_off; mm.qua_at = tb + mine.qua_off;
    return mm;
}

// ---- boundary words ---------------------------------------------------------------------------------------------
// The pack kernels merge the first / last word of a tile with atomicOr when it is shared with the
// neighbouring tile; those words -- and only those -- must be zero beforehand (this replaces a
// memset of the whole output).  blockIdx.y = stream; tile sizes per stream in T[].
struct TileSizes { uint32_t T[4]; };
__global__ void __launch_bounds__(256) zero_boundary_words_kernel(PackArgs a, TileSizes ts)
{
    const int s = blockIdx.y;
    const uint32_t T = ts.T[s];
    const uint64_t n = a.B.n_records;
    const uint64_t tile = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t i0 = tile * T;
    if (i0 >= n) return;
    unsigned long long b0, b1;
    tile_range(a, s, i0, T, b0, b1);
    if (b0 & 31u) a.O.w[s][b0 >> 5] = 0;
    if (i0 + T >= n && (b1 & 31u)) a.O.w[s][b1 >> 5] = 0;
}

// ---- K4q: quality stream -------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(kPackThreads) pack_quality_kernel(PackArgs a, PackPlan pl)
{
    constexpr uint32_t SP = pack_slot_pieces<NW>(), PW = 2 * NW + 1;
    extern __shared__ uint4 pack_smem[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(pack_smem);
    __shared__ unsigned long long tb0, tb1;
    const DeviceParams& P = a.P;
    const uint32_t T = pl.T, tid = threadIdx.x, lane = tid & 31;
    const bool roleB = tid >= T;
    const uint64_t i0 = (uint64_t)blockIdx.x * T, i = i0 + (roleB ? tid - T : tid);
    const bool live = i < a.B.n_records;

    TileRec t{};
    MyMate mm{};
    uint64_t off = 0;
    if (live)
    {
        t = tile_lookup(a, i);
        mm = my_mate(a, t, roleB);
        off = stream_offset(a, 2, i, t.bin, t.start) + (roleB ? (uint64_t)t.lenA * P.qua_bits : 0ull);
    }
    if (tid == 0) tile_range(a, 2, i0, T, tb0, tb1);
    uint8_t* slots = smem + 16;
    uint32_t* stg = reinterpret_cast<uint32_t*>(smem + pl.off_staging[0]);
    zero_staging(smem + pl.off_staging[0], pl.staging_bytes);
    const uint32_t np = live ? (((uint32_t)(mm.qua_at & 15u) + mm.len + 15u) >> 4) : 0u;
    gather_spans<PW>(slots + (size_t)(tid - lane) * SP * 16u, SP * 16u, (uint32_t)(mm.qua_at >> 4), np | (mm.m << 8), a.B.text[0], a.B.text[1]);
    cp_async_wait_all_pack();
    __syncthreads();
    if (live)
    {
        const uint32_t* qw = reinterpret_cast<const uint32_t*>(slots + (size_t)tid * SP * 16u);
        const SymReader rq = reader_open(qw, 16u + (uint32_t)(mm.qua_at & 15u), mm.len, t.rev);
        const uint32_t loc = staging_bit(off, tb0);
        if (P.qua_bits == 6) pack_quality<NW, 6>(rq, mm.len, P, stg, loc);
        else if (P.qua_bits == 3) pack_quality<NW, 3>(rq, mm.len, P, stg, loc);
        else pack_quality<NW, 1>(rq, mm.len, P, stg, loc);
    }
    __syncthreads();
    write_out(stg, a.O.w[2], tb0, tb1);
}

// ---- K4d: DNA stream ---------------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(kPackThreads) pack_dna_kernel(PackArgs a, PackPlan pl)
{
    constexpr uint32_t SP = pack_slot_pieces<NW>(), PW = 2 * NW + 1;
    extern __shared__ uint4 pack_smem[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(pack_smem);
    __shared__ unsigned long long tb0, tb1;
    __shared__ uint4 mate_desc[kPackThreads];                  // DNA work items, indexed by the natural owner
    __shared__ uint16_t dna_order[kPackThreads];               // thread -> item: mates without 'N' first
    __shared__ uint32_t class_count[2][kPackThreads / 32];
    const DeviceParams& P = a.P;
    const uint32_t T = pl.T, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool roleB = tid >= T;
    const uint64_t i0 = (uint64_t)blockIdx.x * T, i = i0 + (roleB ? tid - T : tid);
    const bool live = i < a.B.n_records;

    TileRec t{};
    MyMate mm{};
    mm.plain = true;
    uint64_t off = 0;
    uint32_t cut_len = 0, cut_pos = 0;
    if (live)
    {
        t = tile_lookup(a, i);
        mm = my_mate(a, t, roleB);
        const uint32_t sfx = t.nbin ? 0u : P.k;
        const uint32_t bitsA = (t.info & FSB_INFO_PLAIN_A) ? 2u : 3u;
        off = stream_offset(a, 1, i, t.bin, t.start) + (roleB ? (uint64_t)(t.lenA - sfx) * bitsA : 0ull);
        cut_len = roleB ? 0u : sfx;
        cut_pos = roleB ? 0u : (t.nbin ? 0u : (t.info & FSB_INFO_POS_MASK));
    }
    if (tid == 0) tile_range(a, 1, i0, T, tb0, tb1);
    // DNA work is handed out by class so that warps run one code path: mates without 'N' (2 bits
    // per symbol) first, then the others (3 bits)
    const unsigned b2 = __ballot_sync(0xFFFFFFFFu, live && mm.plain), b3 = __ballot_sync(0xFFFFFFFFu, live && !mm.plain);
    if (lane == 0) { class_count[0][warp] = __popc(b2); class_count[1][warp] = __popc(b3); }
    uint8_t* slots = smem + 16;
    uint32_t* stg = reinterpret_cast<uint32_t*>(smem + pl.off_staging[0]);
    zero_staging(smem + pl.off_staging[0], pl.staging_bytes);
    const uint32_t np = live ? (((uint32_t)(mm.seq_at & 15u) + mm.len + 15u) >> 4) : 0u;
    gather_spans<PW>(slots + (size_t)(tid - lane) * SP * 16u, SP * 16u, (uint32_t)(mm.seq_at >> 4), np | (mm.m << 8), a.B.text[0], a.B.text[1]);
    cp_async_wait_all_pack();
    __syncthreads();
    uint32_t total2 = 0, n_items = 0;
    {
        mate_desc[tid] = make_uint4(staging_bit(off, tb0), mm.len | (cut_pos << 16), cut_len | (t.rev ? 0x100u : 0u) | ((uint32_t)(mm.seq_at & 15u) << 16), 0u);
        uint32_t before2 = 0, before3 = 0;
        const uint32_t nwarps = blockDim.x >> 5;
        for (uint32_t w = 0; w < nwarps; ++w)
        {
            const uint32_t c2 = class_count[0][w], c3 = class_count[1][w];
            if (w < warp) { before2 += c2; before3 += c3; }
            total2 += c2; n_items += c2 + c3;
        }
        const uint32_t lt = (1u << lane) - 1u;
        if (live) dna_order[mm.plain ? before2 + __popc(b2 & lt) : total2 + before3 + __popc(b3 & lt)] = (uint16_t)tid;
    }
    __syncthreads();
    if (tid < n_items)
    {
        const uint32_t item = dna_order[tid];
        const uint4 d = mate_desc[item];
        const uint32_t len = d.y & 0xFFFFu, cpos = d.y >> 16, clen = d.z & 0xFFu;
        const bool rev = (d.z & 0x100u) != 0;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(slots + (size_t)item * SP * 16u);
        const SymReader rs = reader_open(sw, 16u + (d.z >> 16), len, rev);
        if (tid < total2) pack_dna<NW, 2>(rs, len, rev, cpos, clen, stg, d.x);
        else pack_dna<NW, 3>(rs, len, rev, cpos, clen, stg, d.x);
    }
    __syncthreads();
    write_out(stg, a.O.w[1], tb0, tb1);
}

// ---- K4a: meta stream (bin headers + record fields) and titles ------------------------------------------------
__global__ void __launch_bounds__(kPackThreads) pack_aux_kernel(PackArgs a, PackPlan pl)
{
    extern __shared__ uint4 pack_smem[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(pack_smem);
    __shared__ unsigned long long tb0[2], tb1[2];
    const DeviceParams& P = a.P;
    const uint32_t T = pl.T, tid = threadIdx.x, lane = tid & 31;
    const uint64_t i0 = (uint64_t)blockIdx.x * T, i = i0 + tid;
    const bool live = i < a.B.n_records;

    TileRec t{};
    uint64_t off_m = 0, off_h = 0, head_at = 0;
    if (live)
    {
        t = tile_lookup(a, i);
        off_m = stream_offset(a, 0, i, t.bin, t.start);
        if (P.has_headers)
        {
            off_h = stream_offset(a, 3, i, t.bin, t.start);
            head_at = a.B.chunk_text_base[0][t.ch] + t.r1.head_off;
        }
    }
    if (tid == 0) tile_range(a, 0, i0, T, tb0[0], tb1[0]);
    if (tid == 32) { if (P.has_headers) tile_range(a, 3, i0, T, tb0[1], tb1[1]); else { tb0[1] = 0; tb1[1] = 0; } }
    uint8_t* slots = smem + 16;
    uint32_t* stg_m = reinterpret_cast<uint32_t*>(smem + pl.off_staging[0]);
    uint32_t* stg_h = reinterpret_cast<uint32_t*>(smem + pl.off_staging[1]);
    zero_staging(smem + pl.off_staging[0], pl.staging_bytes);
    if (pl.head_pieces)
    {
        const uint32_t nh = live ? (((uint32_t)(head_at & 15u) + t.H + 15u) >> 4) : 0u;
        gather_spans_rt(slots + (size_t)(tid - lane) * pl.head_pieces * 16u, pl.head_pieces * 16u, pl.head_pieces - 1u, (uint32_t)(head_at >> 4), nh, a.B.text[0]);
    }
    cp_async_wait_all_pack();
    __syncthreads();
    if (live)
    {
        if (t.first_of_bin)      // PackToBin header (FastqPacker.cpp:581-583): minLen, maxLen, hasReadGroups = 0
            or_bits(stg_m, staging_bit(off_m - 17, tb0[0]), ((t.bmin & 0xFFu) << 9) | ((t.bmax & 0xFFu) << 1), 17);
        uint32_t mbits;
        const uint32_t mv = meta_fields(P, t.nbin, t.info, t.lenA, t.lenB, t.bmin, t.bmax, mbits);
        or_bits(stg_m, staging_bit(off_m, tb0[0]), mv, mbits);
        if (P.has_headers)
            pack_head(reinterpret_cast<const uint32_t*>(slots + (size_t)tid * pl.head_pieces * 16u), 16u + (uint32_t)(head_at & 15u), t.H, stg_h, staging_bit(off_h, tb0[1]));
    }
    __syncthreads();
    write_out(stg_m, a.O.w[0], tb0[0], tb1[0]);
    write_out(stg_h, a.O.w[3], tb0[1], tb1[1]);
}

} // namespace fsb
