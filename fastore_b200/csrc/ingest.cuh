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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <int NW> constexpr uint32_t win_pieces() { return 2 * NW + 1; }      // 16-byte pieces of the aligned window around L <= 32 NW bytes
// + one guard piece in front; an odd number of pieces per window keeps the lanes' windows on different banks
template <int NW> constexpr uint32_t win_slot_bytes() { return ((win_pieces<NW>() + 1) | 1u) * 16; }

struct IngestPlan
{
    uint32_t warps;          // warps per block
    uint32_t stg_stride;     // words per record in the slot staging: G.words + 4, so that the lanes' stores spread over the banks
    uint32_t head_pieces;    // 16-byte pieces per title window incl. the guard piece (0: no titles)
    uint32_t off_head, off_staging, total_bytes;
};

template <int NW>
inline IngestPlan make_ingest_plan(const DeviceParams& P, const SlotGeom& G, uint32_t max_head)
{
    IngestPlan pl{};
    pl.head_pieces = P.has_headers ? ((15u + max_head + 15u) >> 4) + 1u : 0u;
    pl.stg_stride = G.words + 4u;
    const uint32_t recs = P.paired ? 16u : 32u;                      // records per warp
    for (pl.warps = 4; ; pl.warps >>= 1)
    {
        uint32_t o = 16;                                             // front pad: reversed readers may look 4 bytes below a window
        o += pl.warps * 32u * win_slot_bytes<NW>();
        pl.off_head = o;
        o += pl.warps * recs * pl.head_pieces * 16u;
        o += 32;                                                     // back pad: forward readers run up to 19 bytes past a window
        pl.off_staging = o;
        o += pl.warps * recs * pl.stg_stride * 4u;
        pl.total_bytes = o;
        if (o <= 56u * 1024u || pl.warps == 1) break;
    }
    return pl;
}

// Every lane copies the aligned window around its own span, one 16-byte piece per step.  (Spreading a
// span over several lanes would coalesce the requests, but costs two shuffles and an index
// computation per piece; K1 is bound by the integer pipe, not by the load path, and every line is
// still fetched from HBM exactly once.)
template <int NW>
__device__ __forceinline__ void gather_window(uint8_t* my_window, uint64_t piece0, uint32_t npieces, const uint8_t* text)
{
    constexpr uint32_t PW = win_pieces<NW>();
    const uint8_t* src = text + (piece0 << 4);
#pragma unroll
    for (uint32_t j = 0; j < PW; ++j)
        if (j < npieces) cp_async16(my_window + 16u + 16u * j, src + 16u * j);
}

// K1.  NW = ceil(longest read of the batch / 32).
//   keys[i]  = chunk : signature            cards[i] = card_make(...)  (core.cuh)
//   slots    = [n_records][G.words] words   sig_out / info_out: optional per-read output (parity tests)
// Q = quality bits per symbol (6, 3 or 1).
template <int NW, int Q>
__global__ void __launch_bounds__(128) ingest_kernel(BatchView B, DeviceParams P, SlotGeom G, IngestPlan pl, uint32_t* __restrict__ keys,
                                                      unsigned long long* __restrict__ cards, uint32_t* __restrict__ slots,
                                                      uint32_t* __restrict__ sig_out, uint32_t* __restrict__ info_out)
{
    constexpr uint32_t SB = win_slot_bytes<NW>();
    extern __shared__ uint4 ingest_smem[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(ingest_smem);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t recs_per_warp = P.paired ? 16u : 32u;
    uint8_t* wslots = smem + 16 + (size_t)warp * 32 * SB;
    uint8_t* hslots = smem + pl.off_head + (size_t)warp * recs_per_warp * pl.head_pieces * 16u;
    uint32_t* stg = reinterpret_cast<uint32_t*>(smem + pl.off_staging) + (size_t)warp * recs_per_warp * pl.stg_stride;

    const uint64_t n_mates = P.paired ? 2 * B.n_records : B.n_records;
    const uint64_t g0 = ((uint64_t)blockIdx.x * pl.warps + warp) * 32;             // first mate of this warp
    if (g0 >= n_mates) return;                                                     // whole warp out of range (no block-wide barriers below)
    const uint64_t g = g0 + lane;
    const bool live = g < n_mates;
    const uint64_t i = P.paired ? (g >> 1) : g;                                     // record (pair) index
    const unsigned m = P.paired ? (unsigned)(g & 1) : 0u;
    const uint32_t lrec = P.paired ? (lane >> 1) : lane;                            // record index inside the warp

    uint32_t L = 0, a_seq = 0, a_qua = 0, a_head = 0, H = 0, ch = 0;
    uint64_t seq_p0 = 0, qua_p0 = 0;
    uint32_t seq_np = 0, qua_np = 0, head_p0 = 0, head_np = 0;
    if (live)
    {
        ch = find_chunk(B, i);
        const fsb_record r = (m ? B.rec[1] : B.rec[0])[i];
        const uint64_t tb = (m ? B.chunk_text_base[1] : B.chunk_text_base[0])[ch];
        L = r.seq_len;
        const uint64_t so = tb + r.seq_off, qo = tb + r.qua_off;
        a_seq = (uint32_t)(so & 15u); seq_p0 = so >> 4; seq_np = (a_seq + L + 15u) >> 4;
        a_qua = (uint32_t)(qo & 15u); qua_p0 = qo >> 4; qua_np = (a_qua + L + 15u) >> 4;
        if (m == 0 && P.has_headers)
        {
            const uint64_t ho = tb + r.head_off;
            H = r.head_len;
            a_head = (uint32_t)(ho & 15u); head_p0 = (uint32_t)(ho >> 4); head_np = (a_head + H + 15u) >> 4;
        }
    }
    // ---- stage the sequences and the titles; clear the slot staging meanwhile ------------------------------
    uint8_t* my_window = wslots + (size_t)lane * SB;
    const uint8_t* my_text = m ? B.text[1] : B.text[0];
    gather_window<NW>(my_window, seq_p0, seq_np, my_text);
    if (pl.head_pieces)
    {
        const uint32_t hp = pl.head_pieces - 1u, step = P.paired ? 2u : 1u;
        for (uint32_t base = 0; base < recs_per_warp * hp; base += 32)           // uniform trip count: the shuffles need every lane
        {
            const uint32_t idx = base + lane;
            const uint32_t rec = min(idx / hp, recs_per_warp - 1u), j = idx - rec * hp;
            const uint32_t p0 = __shfl_sync(0xFFFFFFFFu, head_p0, rec * step);
            const uint32_t np = __shfl_sync(0xFFFFFFFFu, head_np, rec * step);
            if (j < np && j < hp) cp_async16(hslots + (size_t)rec * pl.head_pieces * 16u + 16u + 16u * j, B.text[0] + ((uint64_t)(p0 + j) << 4));
        }
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncwarp();

    // ---- bit planes of the sequence; from here on the windows belong to the qualities -------------------------
    const uint32_t* win_words = reinterpret_cast<const uint32_t*>(wslots + (size_t)lane * SB);
    BV<NW> Hp, Lp, Np;
#pragma unroll
    for (int j = 0; j < NW; ++j) { Hp.w[j] = 0; Lp.w[j] = 0; Np.w[j] = 0; }
    if (live) mate_planes<NW>(win_words + 4 + (a_seq >> 2), 8u * (a_seq & 3u), L, Hp, Lp, Np);
    __syncwarp();                                                 // every lane is done with its sequence window
    gather_window<NW>(my_window, qua_p0, qua_np, my_text);
    cp_async_commit();

    // ---- signature -------------------------------------------------------------------------------------------
    StrandMin f, r;
    uint32_t nN = 0;
    f.sig = r.sig = P.nbin; f.pos = r.pos = 0;
    if (live) plane_minimizers<NW>(Hp, Lp, Np, L, P, f, r, nN);
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
    const bool nbin = sig == P.nbin;
    const bool rev = (inf & FSB_INFO_REVERSE) != 0, swp = (inf & FSB_INFO_SWAPPED) != 0;
    // stored pair: forward [m1|m2]; reversed [rc(m2)|rc(m1)]; a swap exchanges the halves
    const bool a_is_m2 = P.paired && (rev != swp);
    const bool roleB = P.paired && ((m == 1) != a_is_m2);
    const uint32_t lenA = roleB ? Lother : L, lenB = P.paired ? (roleB ? L : Lother) : 0u;
    const bool plainA = (inf & FSB_INFO_PLAIN_A) != 0;
    const uint32_t sfx = nbin ? 0u : P.k;
    uint32_t* my_slot = stg + (size_t)lrec * pl.stg_stride;

    // ---- DNA of this mate in the stored orientation (StoreDna), straight from the planes ----------------------------
    SegEmit ed = seg_open(my_slot, 0, 0), eh = ed, eq = ed;
    if (live)
    {
        const uint32_t cut_len = roleB ? 0u : sfx, cut_pos = (roleB || nbin) ? 0u : (inf & FSB_INFO_POS_MASK);
        const uint32_t off = 32u * (G.qw + G.hw) + (roleB ? (lenA - sfx) * (plainA ? 2u : 3u) : 0u);
        ed = seg_open(my_slot, off, (L - cut_len) * (nN == 0 ? 2u : 3u));
        pack_dna_planes<NW>(Hp, Lp, Np, L, rev, nN == 0, cut_pos, cut_len, ed);
    }
    // ---- title, key and card --------------------------------------------------------------------------------------
    if (live && m == 0)
    {
        if (P.has_headers)
        {
            eh = seg_open(my_slot, 32u * G.qw, 8u + 7u * (H ? H - 1u : 0u));
            pack_head(reinterpret_cast<const uint32_t*>(hslots + (size_t)lrec * pl.head_pieces * 16u), 16u + a_head, H, eh);
            seg_finish(eh, false);
        }
        keys[i] = (ch << P.key_bits) | sig;
        cards[i] = card_make((uint32_t)i, inf, lenA, lenB, H);
        if (sig_out) { sig_out[i] = sig; info_out[i] = inf; }
    }
    cp_async_wait_all();
    __syncwarp();

    // ---- quality of this mate in the stored orientation (StoreQuality) --------------------------------------------
    if (live)
    {
        eq = seg_open(my_slot, roleB ? lenA * Q : 0u, L * Q);
        pack_quality<Q>(reader_open(win_words, 16u + a_qua, L, rev), L, P, eq);
    }
    // word 0 of every segment: the mates A first, then the mates B merge into what A has stored
    if (live && !roleB) { seg_finish(ed, false); seg_finish(eq, false); }
    __syncwarp();
    if (live && roleB) { seg_finish(ed, true); seg_finish(eq, true); }
    __syncwarp();

    // ---- the warp's slots leave as one contiguous block -------------------------------------------------------------
    {
        const uint64_t rec0 = P.paired ? (g0 >> 1) : g0;
        const uint64_t nrec = min((uint64_t)recs_per_warp, B.n_records - rec0);
        uint4* gv = reinterpret_cast<uint4*>(slots + rec0 * G.words);
        const uint32_t vpr = G.words >> 2, nvec = (uint32_t)nrec * vpr;            // 16-byte vectors per record / in all
        uint32_t rec = lane / vpr, piece = lane - rec * vpr;                        // staging keeps 4 pad words per record
        const uint32_t drec = 32u / vpr, dpiece = 32u - drec * vpr;
        for (uint32_t j = lane; j < nvec; j += 32)
        {
            gv[j] = *reinterpret_cast<const uint4*>(stg + (size_t)rec * pl.stg_stride + 4u * piece);
            rec += drec; piece += dpiece;
            if (piece >= vpr) { piece -= vpr; rec += 1; }
        }
    }
}

} // namespace fsb
