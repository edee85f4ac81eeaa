// fastore_bin_b200 -- FASTQ reads binning tool with the categorise + pack stage on B200 GPUs.
//
// Same command line as the reference's `fastore_bin e` (main.cpp:44-98, 164-336) and byte-identical
// output to `fastore_bin e -t1` with the same -b: the chunk reader, record parser and bin-file
// writer (libfastore_host.so) stay on the host, Categorize + PackToBins run behind the C ABI
// (libfastore_b200.so).  What changes is the chunk dispatch layer (reference: TFastqChunkReader /
// BinEncoderSE,PE / BinChunkWriter operators, BinOperator.cpp:25-600):
//
//     reader thread  ->  parser threads  ->  one worker thread per GPU  ->  ordered writer turn
//
// Chunk i goes to GPU i mod G; blocks are committed in chunk order (the reference's -t N writes
// them in arrival order, which is why only its -t1 output is reproducible).  Extra options:
//     -G<n>  GPUs to use (default: all sm_100 devices)      -P<n>  parser threads (default 4)
//     -W<n>  workers (contexts) per GPU (default 2: one worker's copies overlap the other's kernels)
//     -D     parse the chunks on the GPU as well (the record tables come back for the title statistics; no parser threads)
//     -K<n>  chunks a worker hands to one fsb_bin_chunks call when that many are already parsed
//            (default 256 MiB / block size, 1..16: small -b chunks are batched into full-size launches)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host_api.h"

namespace {

struct Args
{
    std::vector<std::string> in, out;
    fsh_bin_config cfg{};
    int gpus = 0, parsers = 4, threads = 1, workers = 2, per_call = 0;   // (the memchr parser does ~2 GB/s per thread: four keep up with the reader)
    bool verbose = false, gz = false, device_parse = false;
};

void split_list(const char* s, std::vector<std::string>& out)                  // main.cpp:198-232
{
    std::string cur;
    for (const char* p = s; ; ++p)
    {
        if (*p == ' ' || *p == '\n' || *p == 0)
        {
            if (!cur.empty()) out.push_back(cur);
            cur.clear();
            if (*p == 0) break;
        }
        else cur.push_back(*p);
    }
}

bool parse_args(int argc, const char** argv, Args& a)
{
    fsb_params& p = a.cfg.params;
    p.signature_len = 8; p.skip_zone_len = 8;            // Params.h:36-40: the skip zone defaults to SignatureLength; scripts always pass -s
    p.signature_mask_cutoff_bits = 0; p.paired_end = 0; p.quality_method = FSB_QUA_NONE; p.quality_offset = 33;
    p.binary_threshold = 20; p.reads_have_headers = 0;
    std::memcpy(p.dna_symbol_order, "ACGTN", 5);
    a.cfg.min_block_bin_size = 8; a.cfg.keep_comments = 1; a.cfg.fastq_block_size = 1ull << 28;
    if (argc < 2 || argv[1][0] != 'e') { std::fprintf(stderr, "Error: only the 'e' (binning) mode is provided; decode bin files with the reference's fastore_bin d\n"); return false; }
    for (int i = 2; i < argc; ++i)
    {
        const char* s = argv[i];
        if (s[0] != '-') continue;
        const size_t len = std::strlen(s);
        const long v = (len > 2 && len < 10) ? std::strtol(s + 2, nullptr, 10) : -1;
        switch (s[1])
        {
        case 'i': split_list(s + 2, a.in); break;
        case 'o': split_list(s + 2, a.out); break;
        case 'g': a.gz = true; break;
        case 'b': a.cfg.fastq_block_size = (uint64_t)v << 20; break;
        case 't': a.threads = (int)v; break;
        case 'v': a.verbose = true; a.cfg.verbose = 1; break;
        case 'z': p.paired_end = 1; break;
        case 'p': p.signature_len = (uint8_t)v; break;
        case 's': p.skip_zone_len = (uint8_t)v; break;
        case 'm': a.cfg.min_block_bin_size = (uint32_t)v; break;
        case 'H': p.reads_have_headers = 1; break;
        case 'C': a.cfg.keep_comments = 0; break;
        case 'q': p.quality_method = (uint8_t)v; break;
        case 'w': p.binary_threshold = (uint8_t)v; break;
        case 'I': p.quality_offset = 64; break;
        case 'G': a.gpus = (int)v; break;
        case 'P': a.parsers = (int)std::max(1L, v); break;
        case 'D': a.device_parse = true; break;
        case 'W': a.workers = (int)std::min(8L, std::max(1L, v)); break;
        case 'K': a.per_call = (int)std::min(64L, std::max(1L, v)); break;
        }
    }
    if (a.in.empty()) { std::fprintf(stderr, "Error: no input file(s) specified\n"); return false; }
    if (a.out.empty()) { std::fprintf(stderr, "Error: no output file specified\n"); return false; }
    if (p.paired_end && a.in.size() % 2) { std::fprintf(stderr, "Error: invalid number of input files specified in PE mode\n"); return false; }
    if (a.gz) { std::fprintf(stderr, "Error: .gz input is not supported by this tool\n"); return false; }
    if (a.threads <= 0 || a.threads > 64) { std::fprintf(stderr, "Error: invalid number of threads specified\n"); return false; }
    return true;
}

struct Chunk
{
    uint64_t idx = 0;
    uint8_t* text[2] = {nullptr, nullptr};
    uint64_t size[2] = {0, 0};
    std::vector<fsb_record> rec[2];
    fsh_titles* titles = nullptr;                 // header-field statistics of the chunk's titles (both mates), gathered by the parser thread
    bool bad = false;
    std::string err;
};

struct Pipeline
{
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Chunk*> pool;                     // free chunk buffers
    std::vector<Chunk*> raw;                      // read, not yet parsed
    std::map<uint64_t, Chunk*> parsed;            // by chunk index
    uint64_t n_read = 0, next_write = 0;
    bool read_done = false, failed = false;
    std::string error;

    void fail(const std::string& e) { std::lock_guard<std::mutex> l(mu); if (!failed) { failed = true; error = e; } cv.notify_all(); }
};

} // namespace

int main(int argc, const char** argv)
{
    Args a;
    if (argc < 2 || !parse_args(argc, argv, a)) { std::fprintf(stderr, "usage: fastore_bin_b200 e -i<files> -o<out> [-z] [-H] [-C] [-q<0-2>] [-w<n>] [-I] [-p<n>] [-s<n>] [-m<n>] [-b<MB>] [-G<gpus>] [-W<workers per GPU>] [-K<chunks per call>] [-P<parser threads>] [-D] [-v]\n"); return -1; }
    const bool pe = a.cfg.params.paired_end != 0;
    auto wall = [] { return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count(); };
    if (a.verbose) std::fprintf(stderr, "[wall %.3f] main entered\n", wall());
    const int ndev = fsb_device_count();
    if (a.verbose) std::fprintf(stderr, "[wall %.3f] CUDA driver initialised, %d device(s)\n", wall(), ndev);
    if (ndev == 0) { std::fprintf(stderr, "Error: no sm_100 GPU available (this tool has no CPU fallback)\n"); return -1; }
    int G = a.gpus > 0 ? std::min(a.gpus, ndev) : ndev;
    if (a.gpus <= 0)
    {   // a CUDA context costs about half a second per device: do not open more devices than the input has pairs of chunks
        uint64_t bytes = 0;
        const size_t half_n = a.cfg.params.paired_end ? a.in.size() / 2 : a.in.size();
        for (size_t i = 0; i < half_n; ++i) { FILE* f = std::fopen(a.in[i].c_str(), "rb"); if (f) { std::fseek(f, 0, SEEK_END); const long long sz = ftello(f); if (sz > 0) bytes += (uint64_t)sz; std::fclose(f); } }
        const uint64_t n_chunks = bytes / std::max<uint64_t>(a.cfg.fastq_block_size, 1) + 1;
        G = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)G, (n_chunks + 1) / 2));
    }

    std::vector<const char*> f1, f2;
    const size_t half = pe ? a.in.size() / 2 : a.in.size();
    for (size_t i = 0; i < half; ++i) f1.push_back(a.in[i].c_str());
    for (size_t i = half; i < a.in.size(); ++i) f2.push_back(a.in[i].c_str());
    fsh_reader* reader = fsh_reader_open(f1.data(), (uint32_t)f1.size(), f2.data(), (uint32_t)f2.size(), a.cfg.fastq_block_size);
    if (!reader) { std::fprintf(stderr, "Error: %s\n", fsh_last_error()); return -1; }
    fsh_writer* writer = fsh_writer_open(a.out[0].c_str(), &a.cfg);
    if (!writer) { std::fprintf(stderr, "Error: %s\n", fsh_last_error()); return -1; }

    const auto t0 = std::chrono::steady_clock::now();
    const int NWK = G * a.workers;                                  // workers: worker w drives GPU w mod G
    const int K = a.per_call > 0 ? a.per_call : (int)std::min<uint64_t>(16, std::max<uint64_t>(1, (256ull << 20) / std::max<uint64_t>(a.cfg.fastq_block_size, 1)));
    Pipeline P;
    const int nbuf = NWK * K + a.parsers + 1;
    std::vector<Chunk> chunks((size_t)nbuf);
    // the chunk buffers (pinned host memory) are allocated when the reader first uses them: a short input pins a few, a long one all
    // of them.  (A thread of its own for the pinning was measured slower: it fights the context creation and the workers' first
    // allocations for the driver lock, r02j / r02k.)
    for (Chunk& c : chunks) P.pool.insert(P.pool.begin(), &c);
    std::atomic<int> contexts_ready{0};

    // ---- reader: one chunk after the other, exactly the reference's cuts ------------------------------------
    std::thread t_read([&] {
        for (;;)
        {
            Chunk* c = nullptr;
            {
                std::unique_lock<std::mutex> l(P.mu);
                P.cv.wait(l, [&] { return !P.pool.empty() || P.failed; });
                if (P.failed) break;
                c = P.pool.back(); P.pool.pop_back();               // last in, first out: buffers that exist are reused before new ones are pinned
            }
            bool have = true;
            for (int m = 0; m < (pe ? 2 : 1) && have; ++m)
                if (!c->text[m]) { c->text[m] = (uint8_t*)fsb_host_alloc(a.cfg.fastq_block_size + 64); have = c->text[m] != nullptr; }
            if (!have) { P.fail("cannot allocate pinned chunk buffers"); break; }
            const int rc = fsh_reader_next(reader, c->text[0], &c->size[0], c->text[1], &c->size[1]);
            std::unique_lock<std::mutex> l(P.mu);
            if (rc <= 0) { P.pool.push_back(c); P.read_done = true; if (rc < 0 && !P.failed) { P.failed = true; P.error = "read error"; } P.cv.notify_all(); break; }
            c->idx = P.n_read++;
            P.raw.push_back(c);
            P.cv.notify_all();
        }
    });
    // ---- parsers: chunk text -> record tables ----------------------------------------------------------------
    std::vector<std::thread> t_parse;
    for (int t = 0; t < a.parsers; ++t)
        t_parse.emplace_back([&] {
            for (;;)
            {
                Chunk* c = nullptr;
                {
                    std::unique_lock<std::mutex> l(P.mu);
                    P.cv.wait(l, [&] { return !P.raw.empty() || P.read_done || P.failed; });
                    if (P.failed || (P.raw.empty() && P.read_done)) break;
                    if (P.raw.empty()) continue;
                    c = P.raw.front(); P.raw.erase(P.raw.begin());
                }
                c->bad = false;
                if (a.device_parse)
                {   // the library parses the text on the GPU: hand the chunk on as it is
                    for (int m = 0; m < 2; ++m) c->rec[m].clear();
                    if (c->titles) { fsh_titles_free(c->titles); c->titles = nullptr; }
                    std::lock_guard<std::mutex> l(P.mu);
                    P.parsed[c->idx] = c;
                    P.cv.notify_all();
                    continue;
                }
                for (int m = 0; m < (pe ? 2 : 1); ++m)
                {
                    c->rec[m].resize(fsh_max_records(c->text[m], c->size[m]));
                    fsh_parse_stats st;
                    // symbols, quality range and title characters are checked on the device (FSB_OPT_VALIDATE): the parser only finds the lines
                    const int rc = fsh_parse_chunk_ex(c->text[m], c->size[m], a.cfg.params.reads_have_headers, a.cfg.keep_comments, a.cfg.params.quality_offset,
                                                      a.cfg.params.quality_method, 0, c->rec[m].data(), c->rec[m].size(), &st);
                    c->rec[m].resize(st.n_records);
                    if (st.stop_reason == FSH_STOP_CAPACITY) { c->bad = true; c->err = "chunk " + std::to_string(c->idx) + ": record table too small (internal error)"; }
                    else if (rc != FSB_OK) { c->bad = true; c->err = "chunk " + std::to_string(c->idx) + ": " + std::to_string(st.invalid_records) + " record(s) outside the input contract (symbols ACGTN, length <= 255, quality range)"; }
                }
                if (pe && c->rec[0].size() != c->rec[1].size())
                {   // the reference stops at the shorter of the two (FastqParser.cpp:527)
                    const size_t n = std::min(c->rec[0].size(), c->rec[1].size());
                    c->rec[0].resize(n); c->rec[1].resize(n);
                }
                if (a.cfg.params.reads_have_headers && !c->bad)
                {   // FastqRawBlockStats of the chunk (Stats.cpp:90-169), merged into the file's statistics in chunk order by the writer turn
                    if (c->titles) fsh_titles_free(c->titles);
                    c->titles = fsh_titles_new();
                    for (int m = 0; m < (pe ? 2 : 1); ++m) fsh_titles_add(c->titles, c->text[m], c->rec[m].data(), c->rec[m].size());
                }
                std::lock_guard<std::mutex> l(P.mu);
                P.parsed[c->idx] = c;
                P.cv.notify_all();
            }
        });
    // ---- workers: chunk i -> worker i mod (G * W) on GPU (i mod G * W) mod G, blocks committed in chunk order -----------
    std::atomic<uint64_t> total_records{0};
    std::vector<std::thread> t_gpu;
    for (int w = 0; w < NWK; ++w)
        t_gpu.emplace_back([&, w] {
            fsb_ctx* ctx = nullptr;
            if (fsb_create(&a.cfg.params, w % G, nullptr, &ctx) != FSB_OK) { P.fail(std::string("GPU ") + std::to_string(w % G) + ": " + fsb_last_error(nullptr)); return; }
            if (a.device_parse) { fsb_set_option(ctx, FSB_OPT_KEEP_RECORDS, 1); fsb_set_option(ctx, FSB_OPT_KEEP_COMMENTS, a.cfg.keep_comments ? 1 : 0); }
            auto since = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
            if (a.verbose) std::fprintf(stderr, "[%.2f s] worker %d: context on GPU %d ready\n", since(), w, w % G);
            { std::lock_guard<std::mutex> l(P.mu); contexts_ready++; P.cv.notify_all(); }
            std::vector<Chunk*> mine;
            std::vector<fsb_chunk> in;
            bool stop = false;
            for (uint64_t idx = (uint64_t)w; !stop; )
            {
                // the worker's next chunk (blocking), plus the ones after it that are parsed already
                mine.clear();
                {
                    std::unique_lock<std::mutex> l(P.mu);
                    P.cv.wait(l, [&] { return P.parsed.count(idx) || (P.read_done && idx >= P.n_read) || P.failed; });
                    if (P.failed || !P.parsed.count(idx)) break;
                    for (uint64_t j = idx; (int)mine.size() < K && P.parsed.count(j); j += (uint64_t)NWK) { mine.push_back(P.parsed[j]); P.parsed.erase(j); }
                }
                in.clear();
                std::vector<size_t> slot(mine.size(), (size_t)-1);       // chunk -> position in the call (empty chunks are not sent)
                for (size_t q = 0; q < mine.size() && !stop; ++q)
                {
                    Chunk* c = mine[q];
                    if (c->bad) { P.fail(c->err); stop = true; break; }
                    if (a.device_parse ? c->size[0] == 0 : c->rec[0].empty()) continue;
                    fsb_chunk ch;
                    std::memset(&ch, 0, sizeof(ch));
                    for (int m = 0; m < (pe ? 2 : 1); ++m) { ch.text[m] = c->text[m]; ch.text_size[m] = c->size[m]; ch.records[m] = a.device_parse ? nullptr : c->rec[m].data(); }
                    ch.n_records = a.device_parse ? 0 : c->rec[0].size();
                    slot[q] = in.size();
                    in.push_back(ch);
                }
                if (stop) break;
                std::vector<fsb_block> out(in.size());
                const double t_call = since();
                if (!in.empty() && fsb_bin_chunks(ctx, in.data(), (uint32_t)in.size(), out.data()) != FSB_OK)
                { P.fail(std::string("chunk ") + std::to_string(idx) + ": " + fsb_last_error(ctx)); break; }
                if (a.verbose) std::fprintf(stderr, "[%.2f s] worker %d: chunks %llu.. (%zu in one call) binned in %.3f s\n", since(), w, (unsigned long long)idx, in.size(), since() - t_call);
                if (a.device_parse)
                {   // the tables the device built, for the title statistics (FastqRawBlockStats of the chunk) -- here, outside the writer turn
                    for (size_t q = 0; q < mine.size() && !stop; ++q)
                    {
                        Chunk* c = mine[q];
                        if (slot[q] == (size_t)-1) continue;
                        const uint64_t n = out[slot[q]].n_records;
                        for (int m = 0; m < (pe ? 2 : 1) && !stop; ++m)
                        {
                            c->rec[m].resize(n);
                            uint64_t got = 0;
                            if (fsb_get_records(ctx, (uint32_t)slot[q], m, c->rec[m].data(), n, &got) != FSB_OK || got != n)
                            { P.fail(std::string("chunk ") + std::to_string(c->idx) + ": " + fsb_last_error(ctx)); stop = true; }
                        }
                        if (!stop && a.cfg.params.reads_have_headers)
                        {
                            c->titles = fsh_titles_new();
                            for (int m = 0; m < (pe ? 2 : 1); ++m) fsh_titles_add(c->titles, c->text[m], c->rec[m].data(), c->rec[m].size());
                        }
                    }
                    if (stop) break;
                }
                for (size_t q = 0; q < mine.size() && !stop; ++q, idx += (uint64_t)NWK)
                {
                    Chunk* c = mine[q];
                    const uint64_t n = c->rec[0].size();
                    {   // the writer turn: BinFileWriter is single-threaded state (BinFile.cpp:103-146), and order defines the bytes
                        std::unique_lock<std::mutex> l(P.mu);
                        P.cv.wait(l, [&] { return P.next_write == idx || P.failed; });
                        if (P.failed) { stop = true; break; }
                    }
                    int rc = FSB_OK;
                    if (n)
                    {
                        if (c->titles && fsh_titles_consistent(c->titles)) rc = fsh_writer_merge_titles(writer, c->titles);
                        else
                        {   // statistics that depend on the record order: record by record, as the reference's single-thread loop does
                            rc = fsh_writer_add_titles(writer, c->text[0], c->rec[0].data(), n);
                            if (rc == FSB_OK && pe) rc = fsh_writer_add_titles(writer, c->text[1], c->rec[1].data(), n);
                        }
                        if (rc == FSB_OK) rc = fsh_writer_add_block(writer, &out[slot[q]]);
                    }
                    if (rc != FSB_OK) { P.fail(std::string("writing chunk ") + std::to_string(idx) + ": " + fsh_last_error()); stop = true; break; }
                    total_records += n;
                    if (a.verbose) std::fprintf(stderr, "\rchunk %llu: %llu records, %llu bins   ", (unsigned long long)idx, (unsigned long long)n, (unsigned long long)(n ? out[slot[q]].n_bins : 0));
                    std::lock_guard<std::mutex> l(P.mu);
                    P.next_write = idx + 1;
                    P.pool.push_back(c);
                    P.cv.notify_all();
                }
            }
            fsb_destroy(ctx);
        });

    t_read.join();
    const double t_read_done = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& t : t_parse) t.join();
    for (auto& t : t_gpu) t.join();
    fsh_reader_close(reader);
    const int wrc = fsh_writer_close(writer);
    if (P.failed) { std::fprintf(stderr, "Error: %s\n", P.error.c_str()); return -1; }
    if (wrc != FSB_OK) { std::fprintf(stderr, "Error: %s\n", fsh_last_error()); return -1; }
    if (a.verbose)
    {
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::fprintf(stderr, "\nreader finished after %.2f s", t_read_done);
        std::fprintf(stderr, "\n%llu records in %llu chunks on %d GPU(s) x %d worker(s): %.2f s, %.0f records/s\n", (unsigned long long)total_records.load(),
                     (unsigned long long)P.next_write, G, a.workers, s, total_records.load() / std::max(s, 1e-9));
    }
    // The bin files are closed.  Un-pinning gigabytes of chunk buffers and tearing the CUDA context down takes seconds and gives
    // nothing back that the exit of the process does not: leave at once.
    if (a.verbose) std::fprintf(stderr, "[wall %.3f] leaving\n", wall());
    std::fflush(nullptr);
    std::_Exit(0);
}
