// GpuBinEncoder.h -- the binding a FaStore maintainer adds to run Categorize + PackToBins on a B200.
//
// Reference-side code: it is compiled against FaStore's own headers (FastqRecord.h, BinBlockData.h, Params.h) and
// the C ABI of include/fastore_b200.h, and replaces the pair of calls every chunk goes through
//
//     categorizer.Categorize(reads, dnaBins);  packer.PackToBins(dnaBins, binBins);
//     (BinModule.cpp:130-133 / :379-382)
//
// by one call, GpuBinEncoder::CategorizeAndPack(reads, binBins).  oracle/Makefile (target ref_gpu) builds a second
// fastore_bin from the UNMODIFIED reference sources with this binding in place (integration/gpu_bin_shim.h), and
// tests/test_reference_binding.py byte-compares its bin files with the stock binary's.
//
// Record tables: FastqRecordsParserSE::ParseFrom leaves pointers into the chunk buffer (FastqParser.cpp:315-343), so
// the table is those pointers minus the address of the first title -- no copy.  FastqRecordsParserPE::ParseFrom
// (:527-553) copies every pair as m1|m2|q1|q2 into a third buffer while the titles stay in the first input chunk:
// mate 2's table points into that third buffer as it is, mate 1's text (title + m1 + q1) is gathered into a private
// buffer.  (A maintainer would rather tell the PE parser to skip its copy; this binding does not touch the parser.)
#ifndef H_GPU_BIN_ENCODER
#define H_GPU_BIN_ENCODER

#include <algorithm>
#include <cstring>
#include <vector>

#include "fastore_b200.h"

#include "BinBlockData.h"
#include "Exception.h"
#include "FastqRecord.h"
#include "Params.h"

class GpuBinEncoder
{
public:
    GpuBinEncoder(const BinModuleConfig& cfg, int device = 0) : ctx(NULL), paired(cfg.archiveType.readType == ArchiveType::READ_PE), headers(cfg.archiveType.readsHaveHeaders)
    {
        fsb_params p;
        std::memset(&p, 0, sizeof(p));
        p.signature_len = cfg.minimizer.signatureLen;
        p.skip_zone_len = cfg.minimizer.skipZoneLen;
        p.signature_mask_cutoff_bits = cfg.minimizer.signatureMaskCutoffBits;
        p.paired_end = paired ? 1 : 0;
        p.quality_method = cfg.quaParams.method;
        p.quality_offset = cfg.archiveType.qualityOffset;
        p.binary_threshold = cfg.quaParams.binaryThreshold;
        p.reads_have_headers = headers ? 1 : 0;
        std::copy(cfg.minimizer.dnaSymbolOrder, cfg.minimizer.dnaSymbolOrder + 5, p.dna_symbol_order);
        if (fsb_create(&p, device, NULL, &ctx) != FSB_OK)
            throw Exception(std::string("GpuBinEncoder: ") + fsb_last_error(NULL));
    }
    ~GpuBinEncoder() { fsb_destroy(ctx); }

    // reads: as filled by FastqRecordsParserSE/PE::ParseFrom for one chunk
    void CategorizeAndPack(const std::vector<FastqRecord>& reads, BinaryBinBlock& out)
    {
        const size_t n = reads.size();
        fsb_chunk ch;
        std::memset(&ch, 0, sizeof(ch));
        ch.n_records = n;
        table[0].resize(n);
        if (!paired)
        {
            // every view lies in the chunk buffer, in parse order: offsets from the lowest address
            const char* base = reads[0].seq;
            const char* end = reads[0].qua + reads[0].seqLen;
            for (size_t i = 0; i < n; ++i)
            {
                const FastqRecord& r = reads[i];
                if (headers && r.head < base) base = r.head;
                base = std::min(base, std::min((const char*)r.seq, (const char*)r.qua));
                end = std::max(end, std::max((const char*)r.seq, (const char*)r.qua) + r.seqLen);
                if (headers) end = std::max(end, (const char*)r.head + r.headLen);
            }
            for (size_t i = 0; i < n; ++i)
            {
                const FastqRecord& r = reads[i];
                fsb_record& t = table[0][i];
                t.head_off = headers ? (uint32_t)(r.head - base) : 0;
                t.seq_off = (uint32_t)(r.seq - base);
                t.qua_off = (uint32_t)(r.qua - base);
                t.seq_len = r.seqLen;
                t.head_len = headers ? r.headLen : 0;
                t.reserved = 0;
            }
            ch.text[0] = (const uint8_t*)base; ch.text_size[0] = (uint64_t)(end - base); ch.records[0] = table[0].data();
        }
        else
        {
            // mate 2: views into the parser's m1|m2|q1|q2 buffer; mate 1: title + m1 + q1 gathered into `text1`
            table[1].resize(n);
            size_t bytes = 0;
            for (size_t i = 0; i < n; ++i) bytes += (headers ? reads[i].headLen : 0) + 2u * reads[i].seqLen;
            text1.resize(bytes + 16);
            const char* base2 = reads[0].seq;
            const char* end2 = reads[0].qua + reads[0].seqLen + reads[0].auxLen;
            for (size_t i = 0; i < n; ++i)
            {
                base2 = std::min(base2, (const char*)reads[i].seq);
                end2 = std::max(end2, (const char*)reads[i].qua + reads[i].seqLen + reads[i].auxLen);
            }
            size_t o = 0;
            for (size_t i = 0; i < n; ++i)
            {
                const FastqRecord& r = reads[i];
                fsb_record& a = table[0][i];
                fsb_record& b = table[1][i];
                a.head_off = (uint32_t)o; a.head_len = headers ? r.headLen : 0;
                if (headers) { std::memcpy(&text1[o], r.head, r.headLen); o += r.headLen; }
                a.seq_off = (uint32_t)o; std::memcpy(&text1[o], r.seq, r.seqLen); o += r.seqLen;
                a.qua_off = (uint32_t)o; std::memcpy(&text1[o], r.qua, r.seqLen); o += r.seqLen;
                a.seq_len = r.seqLen; a.reserved = 0;
                b.head_off = 0; b.head_len = 0; b.reserved = 0;
                b.seq_off = (uint32_t)(r.seq + r.seqLen - base2);
                b.qua_off = (uint32_t)(r.qua + r.seqLen - base2);
                b.seq_len = r.auxLen;
            }
            ch.text[0] = (const uint8_t*)text1.data(); ch.text_size[0] = o; ch.records[0] = table[0].data();
            ch.text[1] = (const uint8_t*)base2; ch.text_size[1] = (uint64_t)(end2 - base2); ch.records[1] = table[1].data();
        }
        fsb_block b;
        if (fsb_bin_chunks(ctx, &ch, 1, &b) != FSB_OK)
            throw Exception(std::string("GpuBinEncoder: ") + fsb_last_error(ctx));

        // library-owned pinned memory -> the block's own Buffers (never hand CUDA memory to Buffer: its dtor delete[]s)
        out.Clear();
        out.blockType = BinaryBinBlock::MultiSignatureType;
        Put(out.metaData, b.meta, b.meta_size); out.metaSize = b.meta_size;
        Put(out.dnaData, b.dna, b.dna_size);    out.dnaSize = b.dna_size;
        Put(out.quaData, b.qua, b.qua_size);    out.quaSize = b.qua_size;
        Put(out.headData, b.head, b.head_size); out.headSize = b.head_size;
        out.rawDnaSize = b.raw_dna_size; out.rawHeadSize = b.raw_head_size;
        for (uint64_t i = 0; i < b.n_bins; ++i)
        {
            const fsb_bin_descriptor& d = b.bins[i];
            BinaryBinDescriptor& o = out.descriptors[(uint32)d.signature];
            o.metaSize = d.meta_size; o.dnaSize = d.dna_size; o.quaSize = d.qua_size; o.headSize = d.head_size;
            o.recordsCount = d.records_count; o.rawDnaSize = d.raw_dna_size; o.rawHeadSize = d.raw_head_size;
        }
    }

private:
    static void Put(Buffer& dst, const uint8_t* src, uint64_t n)
    {
        if (dst.Size() < n) dst.Extend(n);
        if (n) std::memcpy(dst.Pointer(), src, n);
    }
    fsb_ctx* ctx;
    bool paired, headers;
    std::vector<fsb_record> table[2];
    std::vector<char> text1;
};

#endif // H_GPU_BIN_ENCODER
