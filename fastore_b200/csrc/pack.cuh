// pack.cuh -- K4: bit-pack and scatter the records of one tile of the bin-sorted order (sm_100a).
//
// Replaces FastqRecordsPackerSE/PE::StoreRecords, IFastqPacker::StoreNextRecord / StoreDna /
// StoreQuality / StoreHeader and BitMemoryWriter (FastqPacker.cpp:113-287, 734-759, 815-859;
// BitMemory.h:216-433).
//
// A block owns T consecutive records of the sorted order.  Because bins are laid out back to back
// in that order, the tile's output is one contiguous bit range in each of the four streams:
//   1. every thread looks up its record (sorted position -> record index, flags, bit offsets);
//   2. the warps gather the aligned 16-byte pieces around each source sequence, quality and title
//      into shared-memory slots with cp.async (coalesced requests over whole sectors; the text is
//      read from HBM exactly once here);
//   3. one thread per stored mate packs its DNA and quality segment into the tile's staging
//      buffers (pack_core.cuh); the mate-B thread (SE: the only thread) also writes the record's
//      meta fields, the title and, for the first record of a bin, the 17-bit bin header;
//   4. the block writes the staging buffers to the streams with coalesced 32-bit stores; only the
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

// shared-memory plan of one tile, computed on the host from the batch statistics
struct PackTilePlan
{
    uint32_t T;              // records per tile
    uint32_t threads;        // T (SE) or 2T (PE)
    uint32_t head_pieces;    // 16-byte pieces per title slot incl. the guard piece (0: no titles)
    uint32_t cap_words[4];   // staging capacity per stream
    uint32_t off_qua_slots, off_head_slots, off_staging[4], total_bytes;
};

template <int NW> constexpr uint32_t pack_slot_pieces() { return 2 * NW + 2; }     // guard piece + aligned window

template <int NW>
inline PackTilePlan make_pack_plan(const DeviceParams& P, uint32_t T, uint32_t max_len, uint32_t max_head)
{
    PackTilePlan pl{};
    const uint32_t roles = P.paired ? 2u : 1u;
    pl.T = T; pl.threads = T * roles;
    pl.head_pieces = P.has_headers ? ((15u + max_head + 15u) >> 4) + 1u : 0u;
    const uint32_t Lsum = max_len * roles;
    const uint32_t bits[4] = {52u, 3u * Lsum + 7u, P.qua_bits * Lsum + 7u, P.has_headers ? 8u + 7u * (max_head ? max_head - 1u : 0u) + 7u : 0u};
    uint32_t o = 16;                                                           // front pad: reversed readers may look 4 bytes below a slot
    o += pl.threads * pack_slot_pieces<NW>() * 16u;                            // sequence slots
    pl.off_qua_slots = o;
    o += pl.threads * pack_slot_pieces<NW>() * 16u;
    pl.off_head_slots = o;
    o += T * pl.head_pieces * 16u;
    o += 32;                                                                   // back pad: forward readers run up to 19 bytes past a window
    for (int s = 0; s < 4; ++s)
    {
        pl.cap_words[s] = ((T * bits[s] + 31u) / 32u + 2u + 3u) & ~3u;
        pl.off_staging[s] = o;
        o += pl.cap_words[s] * 4u;
    }
    pl.total_bytes = o;
    return pl;
}

__device__ __forceinline__ void cp_async16_pack(void* smem_dst, const void* gmem_src)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}

// a warp copies the windows of its 32 lanes' spans: lane -> (span, piece) pairs, consecutive lanes
// on consecutive pieces of the same span
template <int PW>
__device__ __forceinline__ void gather_spans(uint8_t* slots, uint32_t slot_bytes, uint32_t my_slot, uint32_t piece0, uint32_t npieces_m /* npieces | m << 8 */,
                                             const uint8_t* text0, const uint8_t* text1)
{
    const unsigned lane = threadIdx.x & 31;
#pragma unroll 2
    for (uint32_t idx = lane; idx < 32u * PW; idx += 32)
    {
        const uint32_t span = idx / PW, j = idx - span * PW;
        const uint32_t p0 = __shfl_sync(0xFFFFFFFFu, piece0, span);
        const uint32_t nm = __shfl_sync(0xFFFFFFFFu, npieces_m, span);
        const uint32_t slot = __shfl_sync(0xFFFFFFFFu, my_slot, span);
        if (j < (nm & 0xFFu))
            cp_async16_pack(slots + (size_t)slot * slot_bytes + 16u + 16u * j, ((nm >> 8) ? text1 : text0) + ((uint64_t)(p0 + j) << 4));
    }
}

template <int NW>
__global__ void __launch_bounds__(128) pack_kernel(PackArgs a, PackTilePlan pl)
{
    constexpr uint32_t SP = pack_slot_pieces<NW>();            // pieces per sequence / quality slot
    constexpr uint32_t PW = 2 * NW + 1;                        // window pieces
    extern __shared__ uint4 pack_smem[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(pack_smem);
    __shared__ unsigned long long tile_start[4], tile_end[4];

    const DeviceParams& P = a.P;
    const bool pe = P.paired != 0;
    const uint32_t T = pl.T, tid = threadIdx.x;
    const uint32_t p = tid < T ? tid : tid - T;                // record of the tile
    const bool roleB = tid >= T;                                // PE: second stored mate
    const bool aux = pe ? roleB : true;                         // writes meta, title, bin header
    const uint64_t n = a.B.n_records;
    const uint64_t i0 = (uint64_t)blockIdx.x * T, i = i0 + p;
    const bool live = i < n;

    // ---- 1. record lookup -----------------------------------------------------------------------------
    uint32_t info = 0, bin = 0, bmin = 0, bmax = 0, lenA = 0, lenB = 0, H = 0, myLen = 0;
    bool nbin = false, first_of_bin = false, myRev = false, myPlain = true;
    uint64_t off[4] = {0, 0, 0, 0}, bin_bit0 = 0;
    uint64_t seq_at = 0, qua_at = 0, head_at = 0;               // byte offsets inside text[m]
    uint32_t my_m = 0;
    if (live)
    {
        const uint32_t r = a.S.perm[i], key = a.S.skeys[i];
        const uint32_t ch = key >> P.key_bits;
        nbin = (key & ((1u << P.key_bits) - 1u)) == P.nbin;
        info = a.S.info[r];
        bin = a.A.bin_of[i];
        const uint64_t start = a.A.bin_start[bin];
        first_of_bin = i == start;
        bmin = a.A.bin_min[bin]; bmax = a.A.bin_max[bin];
#pragma unroll
        for (int s = 0; s < 4; ++s) off[s] = 8ull * a.BO.B[s][bin] + (a.SC.P[s][i] - a.SC.P[s][start]);
        bin_bit0 = 8ull * a.BO.B[0][bin];
        off[0] += 17;
        const bool rev = (info & FSB_INFO_REVERSE) != 0, swp = (info & FSB_INFO_SWAPPED) != 0;
        // stored pair: forward [m1|m2]; reversed [rc(m2)|rc(m1)]; a swap exchanges the halves
        const bool a_is_m2 = pe && (rev != swp);
        const fsb_record r1 = a.B.rec[0][r];
        fsb_record r2 = r1;
        if (pe) r2 = a.B.rec[1][r];
        const fsb_record rA = a_is_m2 ? r2 : r1;
        const fsb_record rB = a_is_m2 ? r1 : r2;
        lenA = rA.seq_len; lenB = pe ? rB.seq_len : 0u;
        H = P.has_headers ? r1.head_len : 0u;
        my_m = roleB ? (a_is_m2 ? 0u : 1u) : (a_is_m2 ? 1u : 0u);
        const fsb_record mine = roleB ? rB : rA;
        myLen = mine.seq_len;
        myRev = rev;
        myPlain = (info & (roleB ? FSB_INFO_PLAIN_B : FSB_INFO_PLAIN_A)) != 0;
        const uint64_t tb = (my_m ? a.B.chunk_text_base[1] : a.B.chunk_text_base[0])[ch];
        seq_at = tb + mine.seq_off; qua_at = tb + mine.qua_off;
        head_at = a.B.chunk_text_base[0][ch] + r1.head_off;
    }
    if (tid == 0)
    {
        // the tile's bit range in each stream: from its first record (or the start of that record's
        // bin, header and all) to the same point of the next tile
#pragma unroll
        for (int s = 0; s < 4; ++s) tile_start[s] = first_of_bin ? 8ull * a.BO.B[s][bin] : off[s];
        const uint64_t in = i0 + T;
        if (in < n)
        {
            const uint32_t nbin_i = a.A.bin_of[in];
            const uint64_t nstart = a.A.bin_start[nbin_i];
#pragma unroll
            for (int s = 0; s < 4; ++s)
                tile_end[s] = (in == nstart) ? 8ull * a.BO.B[s][nbin_i] : 8ull * a.BO.B[s][nbin_i] + (a.SC.P[s][in] - a.SC.P[s][nstart]) + (s == 0 ? 17ull : 0ull);
        }
        else
        {
            const uint32_t nb = *a.nb_ptr;
#pragma unroll
            for (int s = 0; s < 4; ++s) tile_end[s] = 8ull * a.BO.B[s][nb];
        }
    }
    // ---- 2. zero the staging buffers, gather the source windows ---------------------------------------------
    {
        uint4* st = reinterpret_cast<uint4*>(smem + pl.off_staging[0]);
        const uint32_t nvec = (pl.total_bytes - pl.off_staging[0]) >> 4;
        for (uint32_t j = tid; j < nvec; j += blockDim.x) st[j] = make_uint4(0, 0, 0, 0);
    }
    uint8_t* seq_slots = smem + 16;
    uint8_t* qua_slots = smem + pl.off_qua_slots;
    uint8_t* head_slots = smem + pl.off_head_slots;
    {
        const uint32_t sa = (uint32_t)(seq_at & 15u), qa = (uint32_t)(qua_at & 15u);
        const uint32_t ns = live ? ((sa + myLen + 15u) >> 4) : 0u, nq = live ? ((qa + myLen + 15u) >> 4) : 0u;
        gather_spans<PW>(seq_slots, SP * 16u, tid, (uint32_t)(seq_at >> 4), ns | (my_m << 8), a.B.text[0], a.B.text[1]);
        gather_spans<PW>(qua_slots, SP * 16u, tid, (uint32_t)(qua_at >> 4), nq | (my_m << 8), a.B.text[0], a.B.text[1]);
        if (pl.head_pieces)
        {
            // titles belong to the aux threads; every warp runs the loop over its own lanes' spans
            const uint32_t ha = (uint32_t)(head_at & 15u);
            const uint32_t nh = (live && aux) ? ((ha + H + 15u) >> 4) : 0u;
            const unsigned lane = tid & 31;
            const uint32_t hp = pl.head_pieces - 1u;            // window pieces
            for (uint32_t idx = lane; idx < 32u * hp; idx += 32)
            {
                const uint32_t span = idx / hp, j = idx - span * hp;
                const uint32_t p0 = __shfl_sync(0xFFFFFFFFu, (uint32_t)(head_at >> 4), span);
                const uint32_t np = __shfl_sync(0xFFFFFFFFu, nh, span);
                const uint32_t slot = __shfl_sync(0xFFFFFFFFu, p, span);
                if (j < np) cp_async16_pack(head_slots + (size_t)slot * pl.head_pieces * 16u + 16u + 16u * j, a.B.text[0] + ((uint64_t)(p0 + j) << 4));
            }
        }
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();

    // ---- 3. pack ------------------------------------------------------------------------------------------
    uint32_t* stg[4];
    uint32_t loc[4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
    {
        stg[s] = reinterpret_cast<uint32_t*>(smem + pl.off_staging[s]);
        loc[s] = (uint32_t)(off[s] - ((tile_start[s] >> 5) << 5));          // bit offset inside the staging buffer
    }
    if (live)
    {
        const uint32_t sfx = nbin ? 0u : P.k, mpos = info & FSB_INFO_POS_MASK;
        const uint32_t bitsA = (info & FSB_INFO_PLAIN_A) ? 2u : 3u;
        const uint32_t dna_off = roleB ? loc[1] + (lenA - sfx) * bitsA : loc[1];
        const uint32_t qua_off = roleB ? loc[2] + lenA * P.qua_bits : loc[2];
        const uint32_t cut_len = roleB ? 0u : sfx, cut_pos = roleB ? 0u : (nbin ? 0u : mpos);
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(seq_slots + (size_t)tid * SP * 16u);
        const uint32_t* qw = reinterpret_cast<const uint32_t*>(qua_slots + (size_t)tid * SP * 16u);
        const uint32_t saddr = 16u + (uint32_t)(seq_at & 15u), qaddr = 16u + (uint32_t)(qua_at & 15u);
        if (myPlain) pack_dna<NW, 2>(reader_open(sw, saddr, myLen, myRev), myLen, myRev, cut_pos, cut_len, stg[1], dna_off);
        else pack_dna<NW, 3>(reader_open(sw, saddr, myLen, myRev), myLen, myRev, cut_pos, cut_len, stg[1], dna_off);
        const SymReader rq = reader_open(qw, qaddr, myLen, myRev);
        if (P.qua_bits == 6) pack_quality<6>(rq, myLen, P, stg[2], qua_off);
        else if (P.qua_bits == 3) pack_quality<3>(rq, myLen, P, stg[2], qua_off);
        else pack_quality<1>(rq, myLen, P, stg[2], qua_off);
        if (aux)
        {
            if (first_of_bin)      // PackToBin header (FastqPacker.cpp:581-583): minLen, maxLen, hasReadGroups = 0
                or_bits(stg[0], (uint32_t)(bin_bit0 - ((tile_start[0] >> 5) << 5)), ((bmin & 0xFFu) << 9) | ((bmax & 0xFFu) << 1), 17);
            uint32_t mbits;
            const uint32_t mv = meta_fields(P, nbin, info, lenA, lenB, bmin, bmax, mbits);
            or_bits(stg[0], loc[0], mv, mbits);
            if (P.has_headers)
                pack_head(reinterpret_cast<const uint32_t*>(head_slots + (size_t)p * pl.head_pieces * 16u), 16u + (uint32_t)(head_at & 15u), H, stg[3], loc[3]);
        }
    }
    __syncthreads();

    // ---- 4. write the tile out ----------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < 4; ++s)
    {
        const uint64_t b0 = tile_start[s], b1 = tile_end[s];
        if (b1 <= b0) continue;
        const uint64_t w0 = b0 >> 5;
        const uint32_t nw = (uint32_t)(((b1 - 1) >> 5) - w0) + 1u;
        const bool head_shared = (b0 & 31u) != 0, tail_shared = (b1 & 31u) != 0;
        uint32_t* g = a.O.w[s] + w0;
        for (uint32_t j = tid; j < nw; j += blockDim.x)
        {
            const uint32_t v = bswap32(stg[s][j]);
            if ((j == 0 && head_shared) || (j == nw - 1 && tail_shared)) atomicOr(g + j, v);
            else g[j] = v;
        }
    }
}

} // namespace fsb
