/*
 * ref_rebin_harness.cpp -- the *compiled reference* behind orc_find_new_minimizer (TEST INFRASTRUCTURE ONLY, see oracle_api.h).
 *
 * No algorithm here: a subclass of the reference's DnaRebalancer (fastore_rebin/DnaRebalancer.h:28-70) makes the protected member
 * FindNewMinimizer callable, builds the two FastqRecord views it takes (the read and its reverse complement, the way
 * DnaRebalancer::Rebalance prepares them with FastqRecord::ComputeRC, DnaRebalancer.cpp:246-258) and returns what it returns.
 * Compiled by oracle/Makefile against the sources where they lie under /root/reference; output oracle/_ref/libfastore_ref_rebin.so.
 */
#include "Globals.h"
#include "FastqRecord.h"
#include "Params.h"
#include "../fastore_rebin/Params.h"
#include "../fastore_rebin/DnaRebalancer.h"

#include <cstring>
#include <map>
#include <tuple>

#include "oracle_api.h"

namespace {

struct OpenRebalancer : public DnaRebalancer
{
    OpenRebalancer(const MinimizerParameters& m, const BinBalanceParameters& b) : DnaRebalancer(m, b, false) {}
    std::tuple<uint32, uint16, bool> Call(const FastqRecord& f, const FastqRecord& r, uint32 cur) { return FindNewMinimizer(f, r, cur); }
};

} // namespace

extern "C" void refrebin_find_new_minimizer(const fsb_params* p, const uint8_t* seq, uint32_t len, uint32_t cur, uint32_t divisor,
                                            uint32_t* sig, uint32_t* pos, uint32_t* is_rev)
{
    // one rebalancer per parameter set (its base class builds the 4^k-entry validity table)
    static std::map<uint64, OpenRebalancer*> cache;
    const uint64 key = (uint64)p->signature_len | ((uint64)p->skip_zone_len << 8) | ((uint64)p->signature_mask_cutoff_bits << 16) | ((uint64)divisor << 24);
    auto it = cache.find(key);
    if (it == cache.end())
    {
        MinimizerParameters* m = new MinimizerParameters();
        m->signatureLen = p->signature_len; m->skipZoneLen = p->skip_zone_len; m->signatureMaskCutoffBits = p->signature_mask_cutoff_bits;
        std::memcpy(m->dnaSymbolOrder, p->dna_symbol_order, 5);
        BinBalanceParameters* b = new BinBalanceParameters();
        b->signatureParity = divisor;
        b->validBinSignatures.resize(m->TotalMinimizersCount(), true);          // RebinModule.cpp:72
        it = cache.insert(std::make_pair(key, new OpenRebalancer(*m, *b))).first;
    }
    char fwd[256], rev[256], qf[256], qr[256];
    std::memcpy(fwd, seq, len);
    std::memset(qf, 'I', len);
    FastqRecord rf, rr;
    rf.seq = fwd; rf.qua = qf; rf.seqLen = (uint16)len;
    rr.seq = rev; rr.qua = qr;
    rf.ComputeRC(rr);                                                           // FastqRecord.h:80-111
    auto res = it->second->Call(rf, rr, cur);
    *sig = std::get<0>(res); *pos = std::get<1>(res); *is_rev = std::get<2>(res) ? 1u : 0u;
}
