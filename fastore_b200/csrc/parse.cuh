// parse.cuh -- FASTQ chunk text -> record table on the device (sm_100a).
//
// Replaces SingleFastqRecordParser::ReadNextRecord / SkipLine over a whole chunk
// (FastqParser.cpp:46-68, 118-165; driven by FastqRecordsParserSE/PE::ParseFrom, :315-343, 501-586) for callers
// that hand fsb_stage / fsb_bin_chunks the chunk text alone (fsb_chunk.records == NULL): the 16-byte-per-mate
// record table then never crosses PCIe and no host thread has to walk over the text.
//
// The reference parser is a sequential scan: four lines per record; a line ends at LF, at CR LF or at a lone CR;
// the chunk ends silently at the first record whose title is empty or does not start with '@', whose '+' line
// is empty or whose quality length differs from its sequence length (the latter two compared as 16-bit values,
// as the reference's uint16 locals do).  Nothing in those rules looks further back than the start of the
// record, so the scan turns into three data-parallel passes over the text of every (chunk, mate) segment:
//
//   parse_count    a 16-bit line-end mask per 16 bytes of text (a byte ends a line if it is LF, or CR not followed by LF;
//                  found four bytes at a time with word arithmetic) and the line ends per 16 KB tile
//                  (exclusive scan over the tile counts: scan_sort.cuh)
//   parse_lines    from the masks: every line end writes the start of the next line at its rank: line_start[k + 1] = p + 1
//   parse_records  thread r takes lines 4r .. 4r+3: lengths (CR of a CR LF taken off, missing lines count as
//                  empty, exactly what SkipLine returns at the end of the memory), the acceptance rules, the
//                  fsb_record (-C: title cut at the first space); the lowest rejected r of a segment is where the
//                  reference's loop stops, found with atomicMin together with the reason.
//
// Host logic (fastore_b200.cu) sizes the tables between the passes and, for PE, keeps min(n1, n2) pairs.
#pragma once

#include <cuda_runtime.h>

#include "core.cuh"
#include "parse_core.cuh"

namespace fsb {

constexpr uint32_t kParseThreads = 256;
constexpr uint32_t kParseVecs = 4;                                   // 16-byte vectors per thread
constexpr uint32_t kParseTile = kParseThreads * kParseVecs * 16;     // bytes per block
constexpr uint32_t kParseTileVecs = kParseTile / 16;                 // line-end masks (16 bits each) per tile

// one chunk text of one mate
struct ParseSeg
{
    unsigned long long text_base;   // offset of the text in the batch's text buffer of that mate
    unsigned long long size;
    unsigned long long tile0;       // first tile (tiles of all segments are numbered through)
    unsigned long long line0;       // first entry of the segment in the line-start table (set after the count pass)
    unsigned long long rec0;        // first record of the segment's chunk in the record tables (dense numbering by capacity)
    uint32_t mate, chunk;
    uint32_t n_ends, n_lines, cap;  // line ends, lines (one more if the text does not end with a line end), record candidates = ceil(n_lines / 4)
    uint32_t pad;                   // (all three set after the count pass)
};

// result of a segment: the first rejected record candidate and why (FSH_STOP_* of host_api.h), the first record outside
// the device contract (length 1..255, title <= 255)
struct ParseResult
{
    unsigned long long first_bad;   // (candidate << 8) | reason; ~0: every candidate is a record
    unsigned long long first_invalid;   // candidate; ~0: none
};

__device__ __forceinline__ uint32_t parse_seg_of_tile(const ParseSeg* __restrict__ segs, uint32_t n_segs, unsigned long long tile)
{
    uint32_t lo = 0, hi = n_segs;                  // invariant: segs[lo].tile0 <= tile < segs[hi].tile0
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (segs[mid].tile0 <= tile) lo = mid; else hi = mid; }
    return lo;
}

// vector `vec` (0 .. kParseTileVecs) of a tile, the byte behind it, and how many of its bytes belong to the segment; must be
// called by whole warps with consecutive `vec` in consecutive lanes
struct ParseVec { uint4 v; uint32_t next; uint32_t valid; unsigned long long off; };
__device__ __forceinline__ ParseVec parse_load(const uint8_t* __restrict__ text, const ParseSeg& s, unsigned long long tile, uint32_t vec)
{
    ParseVec r;
    r.off = (tile - s.tile0) * kParseTile + 16ull * vec;                           // offset inside the segment
    r.valid = r.off < s.size ? (uint32_t)min(16ull, s.size - r.off) : 0u;
    r.v = make_uint4(0, 0, 0, 0);
    if (r.valid) r.v = *reinterpret_cast<const uint4*>(text + s.text_base + r.off);   // segments start 256-byte aligned and are padded behind
    // the byte behind the vector: the next lane's first byte; the last lane of the warp reads it (inside the segment only)
    uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, r.v.x, 1) & 0xFFu;
    if ((threadIdx.x & 31u) == 31u) nxt = (r.off + 16 < s.size) ? text[s.text_base + r.off + 16] : 0u;
    r.next = (r.off + 16 < s.size) ? nxt : 0u;                                     // a CR at the very end of the text ends its line
    return r;
}

// Pass 1: the line-end mask of every 16-byte vector (16 bits, end_mask[tile * kParseTileVecs + vector]) and the line ends per tile.
// Thread t takes the vectors t, t + 256, ... of its tile (coalesced loads).
__global__ void __launch_bounds__(kParseThreads) parse_count_kernel(const uint8_t* __restrict__ text0, const uint8_t* __restrict__ text1,
                                                                    const ParseSeg* __restrict__ segs, uint32_t n_segs, uint32_t* __restrict__ tile_count,
                                                                    uint16_t* __restrict__ end_mask)
{
    __shared__ uint32_t warp_sum[kParseThreads / 32];
    const unsigned long long tile = blockIdx.x;
    const ParseSeg s = segs[parse_seg_of_tile(segs, n_segs, tile)];
    const uint8_t* text = s.mate ? text1 : text0;
    uint32_t cnt = 0;
#pragma unroll
    for (uint32_t k = 0; k < kParseVecs; ++k)
    {
        const uint32_t vec = k * kParseThreads + threadIdx.x;
        const ParseVec pv = parse_load(text, s, tile, vec);
        const uint32_t vw[4] = {pv.v.x, pv.v.y, pv.v.z, pv.v.w};
        const uint32_t m = line_end_mask(vw, pv.next, pv.valid);
        end_mask[tile * kParseTileVecs + vec] = (uint16_t)m;
        cnt += __popc(m);
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if ((threadIdx.x & 31u) == 0) warp_sum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t t = 0;
#pragma unroll
        for (uint32_t w = 0; w < kParseThreads / 32; ++w) t += warp_sum[w];
        tile_count[tile] = t;
    }
}

// line ends of every segment: the difference of the scanned tile counts at its ends (tile_prefix[all tiles] = the total)
__global__ void parse_seg_ends_kernel(const ParseSeg* __restrict__ segs, uint32_t n_segs, const uint32_t* __restrict__ tile_prefix, uint32_t* __restrict__ seg_ends)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_segs) seg_ends[k] = tile_prefix[segs[k + 1].tile0] - tile_prefix[segs[k].tile0];
}

// Pass 2, from the masks alone (the text is not read again): thread t takes the four consecutive vectors 4t .. 4t+3 of its tile,
// i.e. 64 mask bits = 64 bytes of text.  tile_prefix: exclusive scan of tile_count over all tiles.  line_start holds, per
// segment, n_line_ends + 1 entries from line0.
__global__ void __launch_bounds__(kParseThreads) parse_lines_kernel(const uint16_t* __restrict__ end_mask, const ParseSeg* __restrict__ segs, uint32_t n_segs,
                                                                    const uint32_t* __restrict__ tile_prefix, uint32_t* __restrict__ line_start)
{
    static_assert(kParseVecs == 4, "a thread's masks are one 64-bit load");
    __shared__ uint32_t warp_sum[kParseThreads / 32];
    const unsigned long long tile = blockIdx.x;
    const ParseSeg s = segs[parse_seg_of_tile(segs, n_segs, tile)];
    unsigned long long m = reinterpret_cast<const unsigned long long*>(end_mask + tile * kParseTileVecs)[threadIdx.x];
    const uint32_t cnt = (uint32_t)__popcll(m);
    // exclusive rank of the thread's first line end inside the block
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (unsigned)d) inc += o; }
    if (lane == 31u) warp_sum[warp] = inc;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; ++w) before += warp_sum[w];
    // rank inside the segment: line ends of the segment's earlier tiles + those in front inside the block
    uint32_t k = tile_prefix[tile] - tile_prefix[s.tile0] + before + inc - cnt;
    uint32_t* ls = line_start + s.line0;
    if (tile == s.tile0 && threadIdx.x == 0) ls[0] = 0;
    const unsigned long long off = (tile - s.tile0) * kParseTile + 64ull * threadIdx.x;      // offset of the thread's 64 bytes inside the segment
    while (m)
    {
        const uint32_t b = (uint32_t)__ffsll((long long)m) - 1u;
        m &= m - 1ull;
        ls[++k] = (uint32_t)(off + b + 1u);
    }
}

// records: [mate] the record tables of the batch.  One thread per record candidate of a segment (blockIdx.y = segment).
__global__ void __launch_bounds__(256) parse_records_kernel(const uint8_t* __restrict__ text0, const uint8_t* __restrict__ text1, const ParseSeg* __restrict__ segs,
                                                            const uint32_t* __restrict__ line_start, uint32_t keep_headers, uint32_t keep_comments,
                                                            fsb_record* __restrict__ rec0, fsb_record* __restrict__ rec1, ParseResult* __restrict__ results)
{
    const ParseSeg s = segs[blockIdx.y];
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= s.cap) return;
    const uint8_t* text = (s.mate ? text1 : text0) + s.text_base;
    fsb_record o;
    bool invalid = false;
    const uint32_t reason = parse_candidate(text, (uint32_t)s.size, line_start + s.line0, s.n_ends, s.n_lines, r, keep_headers != 0, keep_comments != 0, o, invalid);
    if (reason != kStopNone) { atomicMin(&results[blockIdx.y].first_bad, ((unsigned long long)r << 8) | reason); return; }
    if (invalid) atomicMin(&results[blockIdx.y].first_invalid, (unsigned long long)r);
    (s.mate ? rec1 : rec0)[s.rec0 + r] = o;
}

} // namespace fsb
