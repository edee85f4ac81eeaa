/*
 * ref_harness.cpp -- the *compiled reference* behind the oracle interface.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h).  This file contains no algorithm: it builds the
 * reference's own objects (FastqCategorizerSE/PE, FastqRecordsPackerSE/PE, BinaryBinBlock) from a
 * record table, calls Categorize + PackToBins exactly as BinModule.cpp:130-133 / :379-382 do, and
 * copies the result out.  It is compiled by oracle/Makefile against the headers and sources where
 * they lie under /root/reference (nothing is copied into the repo); the output goes to
 * oracle/_ref/libfastore_ref.so, which is git-ignored but travels to the GPU box.
 */
#include "Globals.h"
#include "FastqRecord.h"
#include "FastqCategorizer.h"
#include "FastqPacker.h"
#include "BinBlockData.h"
#include "Params.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#include "oracle_api.h"

namespace {

BinModuleConfig make_config(const fsb_params* p)
{
    BinModuleConfig c;
    c.minimizer.signatureLen = p->signature_len;
    c.minimizer.skipZoneLen = p->skip_zone_len;
    c.minimizer.signatureMaskCutoffBits = p->signature_mask_cutoff_bits;
    std::memcpy(c.minimizer.dnaSymbolOrder, p->dna_symbol_order, 5);
    c.archiveType.readType = p->paired_end ? ArchiveType::READ_PE : ArchiveType::READ_SE;
    c.archiveType.qualityOffset = p->quality_offset;
    c.archiveType.readsHaveHeaders = p->reads_have_headers != 0;
    c.quaParams.method = p->quality_method;
    c.quaParams.binaryThreshold = p->binary_threshold;
    return c;
}

// Build FastqRecord views over a private, mutable copy (Categorize rewrites seq/qua in place):
// per record seq (m1|m2 for PE) then qua (q1|q2), the layout FastqRecordsParserPE::ParseFrom
// produces (FastqParser.cpp:527-553); SE records are plain views in the reference
// (FastqParser.cpp:315-343), which the categoriser and packer cannot tell apart from this.
struct Records
{
    std::vector<FastqRecord> recs;
    std::vector<char> buf;         // private mutable copy of the sequences and qualities

    Records(const fsb_params* p, const fsb_chunk* ch, uint64 first, uint64 n)
    {
        recs.resize(n);
        const bool pe = p->paired_end != 0;
        uint64 total = 0;
        for (uint64 i = 0; i < n; ++i)
        {
            total += 2ull * ch->records[0][first + i].seq_len;
            if (pe) total += 2ull * ch->records[1][first + i].seq_len;
        }
        buf.resize(total + 16);
        char* o = buf.data();
        const char* t1 = (const char*)ch->text[0];
        const char* t2 = (const char*)ch->text[1];
        for (uint64 i = 0; i < n; ++i)
        {
            const fsb_record& a = ch->records[0][first + i];
            FastqRecord& r = recs[i];
            r.seq = o;
            std::memcpy(o, t1 + a.seq_off, a.seq_len); o += a.seq_len;
            if (pe)
            {
                const fsb_record& b = ch->records[1][first + i];
                std::memcpy(o, t2 + b.seq_off, b.seq_len); o += b.seq_len;
                r.auxLen = b.seq_len;
            }
            r.qua = o;
            std::memcpy(o, t1 + a.qua_off, a.seq_len); o += a.seq_len;
            if (pe)
            {
                const fsb_record& b = ch->records[1][first + i];
                std::memcpy(o, t2 + b.qua_off, b.seq_len); o += b.seq_len;
            }
            r.seqLen = a.seq_len;
            // the header is never modified on this path; view it in the caller's chunk
            if (p->reads_have_headers) { r.head = const_cast<char*>(t1) + a.head_off; r.headLen = a.head_len; }
        }
    }
};

void run_reference(const BinModuleConfig& cfg, bool pe, std::vector<FastqRecord>& recs, BinaryBinBlock& block)
{
    std::map<uint32, FastqRecordsPtrBin> bins;
    if (pe)
    {
        FastqCategorizerPE cat(cfg.minimizer, cfg.minFilter, cfg.catParams);
        FastqRecordsPackerPE packer(cfg);
        cat.Categorize(recs, bins);
        block.Clear();
        packer.PackToBins(bins, block);
    }
    else
    {
        FastqCategorizerSE cat(cfg.minimizer, cfg.minFilter, cfg.catParams);
        FastqRecordsPackerSE packer(cfg);
        cat.Categorize(recs, bins);
        block.Clear();
        packer.PackToBins(bins, block);
    }
}

uint8_t* dup_bytes(const byte* p, uint64 n)
{
    uint8_t* o = (uint8_t*)std::malloc(n ? n : 1);
    if (n) std::memcpy(o, p, n);
    return o;
}

} // namespace

extern "C" int ref_bin_chunk(const fsb_params* p, const fsb_chunk* ch, orc_block* out)
{
    std::memset(out, 0, sizeof(*out));
    const uint64 n = ch->n_records;
    out->n_records = n;
    if (n == 0) return FSB_OK;
    const BinModuleConfig cfg = make_config(p);
    const bool pe = p->paired_end != 0;
    Records R(p, ch, 0, n);

    // per-read tuples: categorise once more on a private copy so the packed run below sees
    // pristine input (Categorize mutates the reads)
    {
        Records R2(p, ch, 0, n);
        std::map<uint32, FastqRecordsPtrBin> bins;
        if (pe) { FastqCategorizerPE cat(cfg.minimizer, cfg.minFilter, cfg.catParams); cat.Categorize(R2.recs, bins); }
        else    { FastqCategorizerSE cat(cfg.minimizer, cfg.minFilter, cfg.catParams); cat.Categorize(R2.recs, bins); }
        out->read_signature = (uint32_t*)std::malloc(n * sizeof(uint32_t));
        out->read_info = (uint32_t*)std::malloc(n * sizeof(uint32_t));
        for (const auto& kv : bins)
        {
            for (const FastqRecord* r : kv.second.records)
            {
                const uint64 i = (uint64)(r - R2.recs.data());
                uint32_t info = r->minimPos;
                if (r->IsReadReverse()) info |= FSB_INFO_REVERSE;
                if (r->IsPairSwapped()) info |= FSB_INFO_SWAPPED;
                bool plainA = std::find(r->seq, r->seq + r->seqLen, 'N') == r->seq + r->seqLen;
                if (plainA) info |= FSB_INFO_PLAIN_A;
                if (pe)
                {
                    const char* b = r->seq + r->seqLen;
                    if (std::find(b, b + r->auxLen, 'N') == b + r->auxLen) info |= FSB_INFO_PLAIN_B;
                }
                out->read_signature[i] = kv.first;
                out->read_info[i] = info;
            }
        }
    }

    // one block per calling thread for the life of the process: its FastqRawBlockStats tables are never
    // freed by the reference (Stats.cpp:33-36).  Thread-safe: the full-size GPU test checks all chunks of a
    // batch on a pool of threads.
    static thread_local BinaryBinBlock* blockPtr = new BinaryBinBlock();
    BinaryBinBlock& block = *blockPtr;
    run_reference(cfg, pe, R.recs, block);

    out->meta = dup_bytes(block.metaData.Pointer(), block.metaSize); out->meta_size = block.metaSize;
    out->dna = dup_bytes(block.dnaData.Pointer(), block.dnaSize);    out->dna_size = block.dnaSize;
    out->qua = dup_bytes(block.quaData.Pointer(), block.quaSize);    out->qua_size = block.quaSize;
    out->head = dup_bytes(block.headData.Pointer(), block.headSize); out->head_size = block.headSize;
    out->raw_dna_size = block.rawDnaSize;
    out->raw_head_size = block.rawHeadSize;
    out->n_bins = block.descriptors.size();
    out->bins = (fsb_bin_descriptor*)std::calloc(out->n_bins ? out->n_bins : 1, sizeof(fsb_bin_descriptor));
    uint64 b = 0;
    for (const auto& kv : block.descriptors)      // std::map: ascending signature, N-bin (4^k) last
    {
        fsb_bin_descriptor& d = out->bins[b++];
        d.signature = kv.first;
        d.meta_size = kv.second.metaSize;
        d.dna_size = kv.second.dnaSize;
        d.qua_size = kv.second.quaSize;
        d.head_size = kv.second.headSize;
        d.records_count = kv.second.recordsCount;
        d.raw_dna_size = kv.second.rawDnaSize;
        d.raw_head_size = kv.second.rawHeadSize;
    }
    return FSB_OK;
}

extern "C" void ref_block_free(orc_block* b)
{
    if (!b) return;
    std::free(b->meta); std::free(b->dna); std::free(b->qua); std::free(b->head); std::free(b->bins);
    std::free(b->read_signature); std::free(b->read_info);
    std::memset(b, 0, sizeof(*b));
}

extern "C" void ref_find_minimizer(const fsb_params* p, const uint8_t* seq, uint32_t len, uint32_t* sig, uint32_t* pos)
{
    // the categoriser's ctor builds the 4^k-entry validity table (FastqCategorizer.cpp:34-76):
    // keep one instance per parameter set
    static std::map<uint32, std::pair<BinModuleConfig*, FastqCategorizerBase*>> cache;
    const uint32 key = p->signature_len | (p->skip_zone_len << 8) | (p->signature_mask_cutoff_bits << 16);
    auto it = cache.find(key);
    if (it == cache.end())
    {
        BinModuleConfig* cfg = new BinModuleConfig(make_config(p));
        FastqCategorizerBase* cat = new FastqCategorizerBase(cfg->minimizer, cfg->minFilter, cfg->catParams);
        it = cache.insert(std::make_pair(key, std::make_pair(cfg, cat))).first;
    }
    FastqRecord r;
    r.seq = (char*)seq;
    r.seqLen = (uint16)len;
    auto m = it->second.second->FindMinimizer(r);
    *sig = m.first;
    *pos = m.second;
}

// Wall seconds for `reps` passes of Categorize + PackToBins with `threads` workers, each over its
// own contiguous slice of the records (its own "chunk", as the reference's -t N workers have).
// Record materialisation (the parser's job) is outside the timed region.
extern "C" double ref_time_bin_chunk(const fsb_params* p, const fsb_chunk* ch, int threads, int reps)
{
    if (threads < 1) threads = 1;
    const BinModuleConfig cfg = make_config(p);
    const bool pe = p->paired_end != 0;
    const uint64 n = ch->n_records;
    double total = 0.0;
    // BinaryBinBlock owns a FastqRawBlockStats whose QVZ tables are never freed by the reference
    // (Stats.cpp:33-36): allocate the blocks once per call, not per repetition.
    static std::vector<BinaryBinBlock*> blocks;
    while ((int)blocks.size() < threads) blocks.push_back(new BinaryBinBlock());
    for (int rep = 0; rep < reps; ++rep)
    {
        std::vector<Records*> parts(threads);
        for (int t = 0; t < threads; ++t)
        {
            const uint64 first = n * (uint64)t / (uint64)threads;
            const uint64 cnt = n * (uint64)(t + 1) / (uint64)threads - first;
            parts[t] = new Records(p, ch, first, cnt);
        }
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t)
            th.emplace_back([&, t]() { if (!parts[t]->recs.empty()) run_reference(cfg, pe, parts[t]->recs, *blocks[t]); });
        for (auto& x : th) x.join();
        auto t1 = std::chrono::steady_clock::now();
        total += std::chrono::duration<double>(t1 - t0).count();
        for (int t = 0; t < threads; ++t) delete parts[t];
    }
    return total;
}
