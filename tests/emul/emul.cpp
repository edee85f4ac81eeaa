// tests/emul/emul.cpp -- TEST INFRASTRUCTURE: runs the per-thread routines of the CUDA kernels
// (fastore_b200/csrc/*_core.cuh, plain integer code marked FSB_HD) on the CPU, one "thread" after
// the other, over a whole chunk.  The CPU-only test tier compares the result with the oracle so
// that kernel logic is checked before GPU time is spent.  Built only by tests/ (g++, no CUDA);
// nothing in the product loads it.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../fastore_b200/csrc/sig_core.cuh"

using namespace fsb;

namespace {

// What the kernel's staging does for one mate: 16-byte pieces of the aligned window around the
// sequence land in a shared-memory slot; the thread then reads 4-byte words from it.
template <int NW>
void stage_slot(const uint8_t* text, uint64_t text_size, uint32_t seq_off, uint32_t L, uint32_t* slot /* (2NW+1)*4 words */)
{
    const int64_t a0 = (int64_t)(seq_off & ~15u);
    const uint32_t npieces = ((seq_off & 15u) + L + 15u) >> 4;
    uint8_t* dst = reinterpret_cast<uint8_t*>(slot);
    std::memset(dst, 0xAB, (2 * NW + 1) * 16);                    // stale shared memory
    for (uint32_t j = 0; j < npieces && j < 2 * NW + 1; ++j)
        for (int b = 0; b < 16; ++b)
        {
            const int64_t src = a0 + 16 * j + b;
            dst[16 * j + b] = (src >= 0 && (uint64_t)src < text_size) ? text[src] : 0x5A;   // device pad bytes
        }
}

template <int NW>
void mate_scan(const uint8_t* text, uint64_t text_size, const fsb_record& r, const DeviceParams& P, StrandMin& f, StrandMin& rv, uint32_t& nN)
{
    uint32_t slot[(2 * NW + 1) * 4 + 4];
    stage_slot<NW>(text, text_size, r.seq_off, r.seq_len, slot);
    const uint32_t a = r.seq_off & 15u;
    mate_minimizers<NW>(slot + (a >> 2), 8 * (a & 3u), r.seq_len, P, f, rv, nN);
}

template <int NW>
void run_signatures(const DeviceParams& P, const fsb_chunk* ch, uint32_t* sig_out, uint32_t* info_out)
{
    for (uint64_t i = 0; i < ch->n_records; ++i)
    {
        StrandMin f1, r2;
        uint32_t n1;
        mate_scan<NW>(ch->text[0], ch->text_size[0], ch->records[0][i], P, f1, r2, n1);
        if (!P.paired) select_se(f1, r2, n1, P, sig_out[i], info_out[i]);
        else
        {
            StrandMin f2, r1;
            uint32_t n2;
            mate_scan<NW>(ch->text[1], ch->text_size[1], ch->records[1][i], P, f2, r1, n2);
            select_pe(f1, f2, r1, r2, n1, n2, P, sig_out[i], info_out[i]);
        }
    }
}

} // namespace

extern "C" int emul_signatures(const fsb_params* p, const fsb_chunk* ch, uint32_t* sig_out, uint32_t* info_out)
{
    const DeviceParams P = make_device_params(*p);
    uint32_t maxL = 1;
    for (int m = 0; m < (P.paired ? 2 : 1); ++m)
        for (uint64_t i = 0; i < ch->n_records; ++i) if (ch->records[m][i].seq_len > maxL) maxL = ch->records[m][i].seq_len;
    switch ((maxL + 31) / 32)
    {
    case 1: run_signatures<1>(P, ch, sig_out, info_out); break;
    case 2: run_signatures<2>(P, ch, sig_out, info_out); break;
    case 3: run_signatures<3>(P, ch, sig_out, info_out); break;
    case 4: run_signatures<4>(P, ch, sig_out, info_out); break;
    case 5: run_signatures<5>(P, ch, sig_out, info_out); break;
    case 6: run_signatures<6>(P, ch, sig_out, info_out); break;
    case 7: run_signatures<7>(P, ch, sig_out, info_out); break;
    case 8: run_signatures<8>(P, ch, sig_out, info_out); break;
    default: return FSB_ERR_INPUT;
    }
    return FSB_OK;
}
