// synth.cpp -- seeded synthetic FASTQ generator for tests and bench.py.
//
// Every record is a pure function of (seed, record index): the output does not depend on the
// thread count or on how the index range is split, so the CPU baseline, the parity tests and the
// GPU path always see identical bytes, and 100 M-pair workloads can be generated shard by shard.
//
// Model: reads drawn from a synthetic uniform-ACGT genome (itself a hash of the position, nothing
// is stored), uniform start, 50 % reverse-complemented, substitution errors, isolated N, PE insert
// 350 +- 30, qualities round(N(mean, sd)) clipped to [2, 40] + 33, titles "@SYN.<i>" (SE) or
// "@SYN.<i>/1", "@SYN.<i>/2" (PE; the numeric token is what the reference's PE chunk cutter
// re-synchronises on, FastqStream.cpp:231-256).  Optional families stress the categoriser's filter
// rules: N-rich reads (10-60 % N, straddling the N >= L/3 rule, FastqCategorizer.cpp:102),
// low-complexity reads (homopolymers, dinucleotide repeats, poly-A tails: the AA / AAA / AAC
// signature filter, :56-63), all-N reads (the N bin) and directed tie cases (reverse-palindromic
// reads, identical mates, mates that are each other's reverse complement: the <= / < tie rules of
// :217 and :289-304).
#include "host_api.h"

#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct Rng
{
    uint64_t s;
    explicit Rng(uint64_t seed) : s(mix64(seed) | 1) {}
    inline uint64_t next()
    {
        s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
        return s * 0x2545F4914F6CDD1Dull;
    }
    inline uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

const char kBases[4] = {'A', 'C', 'G', 'T'};

inline char genome_base(uint64_t seed, uint64_t pos)
{
    const uint64_t h = mix64(seed * 0x100000001B3ull + (pos >> 5));
    return kBases[(h >> (2 * (pos & 31))) & 3];
}

inline char complement(char c)
{
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; }
    return 'N';
}

struct Generator
{
    fsh_synth_config cfg;
    uint32_t genome_len;
    uint16_t qual_lut[65536 / 16];     // 4096-entry inverse CDF: 12-bit uniform -> quality char
    uint32_t sub_thr, n_thr;           // 20-bit thresholds

    explicit Generator(const fsh_synth_config& c) : cfg(c)
    {
        genome_len = cfg.genome_len ? cfg.genome_len : 100000000u;
        const double mean = (cfg.qual_mean_x10 ? cfg.qual_mean_x10 : 360) / 10.0;
        const double sd = (cfg.qual_sd_x10 ? cfg.qual_sd_x10 : 40) / 10.0;
        // P(round(X) <= q) for X ~ N(mean, sd), clipped to [2, 40]
        double cdf[64];
        for (int q = 0; q < 64; ++q) cdf[q] = 0.5 * std::erfc(-((q + 0.5) - mean) / (sd * std::sqrt(2.0)));
        for (int i = 0; i < 4096; ++i)
        {
            const double u = (i + 0.5) / 4096.0;
            int q = 2;
            while (q < 40 && cdf[q] < u) ++q;
            qual_lut[i] = (uint16_t)(q + 33);
        }
        sub_thr = (uint32_t)(((uint64_t)cfg.sub_rate_ppm << 20) / 1000000ull);
        n_thr = (uint32_t)(((uint64_t)cfg.n_rate_ppm << 20) / 1000000ull);
    }

    uint32_t record_len(uint64_t idx) const
    {
        if (cfg.min_len == 0 || cfg.min_len >= cfg.read_len) return cfg.read_len;
        Rng r(cfg.seed * 0x9E3779B1ull + idx * 2 + 0x51ED);
        return cfg.min_len + r.below(cfg.read_len - cfg.min_len + 1);
    }

    static uint32_t digits(uint64_t v) { uint32_t d = 1; while (v >= 10) { v /= 10; ++d; } return d; }

    uint32_t title_len(uint64_t idx, uint32_t len) const
    {
        uint32_t t = 5 + digits(idx) + (cfg.paired ? 2 : 0);      // "@SYN." + idx [+ "/1"]
        if (cfg.header_comments) t += 5 + digits(len) + 10;        // " len=" + L + " synthetic"
        return t;
    }

    uint64_t record_bytes(uint64_t idx) const
    {
        const uint32_t len = record_len(idx);
        const uint32_t eol = cfg.crlf ? 2 : 1;
        return (uint64_t)title_len(idx, len) + len + 1 + len + 4 * eol;
    }

    // sequence of one mate in read orientation
    void make_mate(Rng& r, uint64_t start, bool revcomp, uint32_t len, char* seq) const
    {
        for (uint32_t i = 0; i < len; ++i)
        {
            char c;
            if (!revcomp) c = genome_base(cfg.seed ^ 0x67E55ull, (start + i) % genome_len);
            else c = complement(genome_base(cfg.seed ^ 0x67E55ull, (start + len - 1 - i) % genome_len));
            const uint64_t x = r.next();
            if ((uint32_t)(x & 0xFFFFF) < sub_thr) c = kBases[((uint32_t)(x >> 20) & 3)];     // may re-draw the same base
            if ((uint32_t)((x >> 24) & 0xFFFFF) < n_thr) c = 'N';
            seq[i] = c;
        }
    }

    void make_quals(Rng& r, uint32_t len, char* q) const
    {
        uint32_t i = 0;
        while (i < len)
        {
            uint64_t x = r.next();
            for (int j = 0; j < 5 && i < len; ++j, x >>= 12) q[i++] = (char)qual_lut[x & 4095];
        }
    }

    void special_family(Rng& r, uint32_t family, uint32_t len, char* s1, char* s2) const
    {
        // family: 1 N-rich, 2 low-complexity, 3 all-N, 4 tie
        const bool pe = cfg.paired != 0;
        if (family == 1)
        {
            for (int m = 0; m < (pe ? 2 : 1); ++m)
            {
                char* s = m ? s2 : s1;
                const uint32_t frac = 100 + r.below(501);          // 10.0 % .. 60.0 % of positions
                for (uint32_t i = 0; i < len; ++i) if (r.below(1000) < frac) s[i] = 'N';
            }
        }
        else if (family == 2)
        {
            for (int m = 0; m < (pe ? 2 : 1); ++m)
            {
                char* s = m ? s2 : s1;
                const uint32_t kind = r.below(4);
                if (kind == 0) { const char c = kBases[r.below(4)]; for (uint32_t i = 0; i < len; ++i) s[i] = c; }
                else if (kind == 1) { const char a = kBases[r.below(4)], b = kBases[r.below(4)]; for (uint32_t i = 0; i < len; ++i) s[i] = (i & 1) ? b : a; }
                else if (kind == 2) { const uint32_t from = len / 4 + r.below(len / 2 + 1); for (uint32_t i = from; i < len; ++i) s[i] = 'A'; }
                else { const uint32_t to = len / 4 + r.below(len / 2 + 1); for (uint32_t i = 0; i < to; ++i) s[i] = 'T'; }
            }
        }
        else if (family == 3)
        {
            std::memset(s1, 'N', len);
            if (pe && r.below(2)) std::memset(s2, 'N', len);
        }
        else if (family == 4)
        {
            const uint32_t kind = r.below(4);
            if (kind == 0)
            {   // reverse-palindromic read: rc(s) == s  -> forward and reverse minima tie
                for (uint32_t i = 0; i < len / 2; ++i) s1[len - 1 - i] = complement(s1[i]);
                if (len & 1) s1[len / 2] = 'N';
                if (pe) std::memcpy(s2, s1, len);
            }
            else if (kind == 1 && pe) std::memcpy(s2, s1, len);                     // identical mates: f1 == f2, r1 == r2
            else if (kind == 2 && pe) { for (uint32_t i = 0; i < len; ++i) s2[i] = complement(s1[len - 1 - i]); }   // m2 = rc(m1): f1 == r1
            else
            {   // tandem repeat of a short unit: the minimal k-mer occurs many times (first-position rule)
                const uint32_t unit = 3 + r.below(10);
                for (uint32_t i = unit; i < len; ++i) s1[i] = s1[i - unit];
                if (pe) for (uint32_t i = 0; i < len; ++i) s2[i] = s1[(i + 1) % len];
            }
        }
    }

    // writes record idx of both files; returns bytes written to file 1 (file 2 gets the same count)
    uint64_t write_record(uint64_t idx, uint8_t* o1, uint8_t* o2, uint64_t off, fsb_record* r1, fsb_record* r2) const
    {
        const uint32_t len = record_len(idx);
        const bool pe = cfg.paired != 0;
        Rng r(cfg.seed * 0xD1B54A32D192ED03ull + idx);
        char s1[256], s2[256], q1[256], q2[256];

        const uint32_t insert = 320 + r.below(61);                     // 350 +- 30
        const uint32_t span = pe ? (insert > len ? insert : len) : len;
        const uint64_t start = r.next() % (genome_len - span);
        const bool flip = r.below(2) != 0;
        if (!pe) make_mate(r, start, flip, len, s1);
        else if (!flip) { make_mate(r, start, false, len, s1); make_mate(r, start + span - len, true, len, s2); }
        else { make_mate(r, start + span - len, true, len, s1); make_mate(r, start, false, len, s2); }
        make_quals(r, len, q1);
        if (pe) make_quals(r, len, q2);

        const uint32_t f = r.below(1000000);
        uint32_t family = 0, acc = cfg.nrich_ppm;
        if (f < acc) family = 1;
        else if (f < (acc += cfg.lowcomplex_ppm)) family = 2;
        else if (f < (acc += cfg.alln_ppm)) family = 3;
        else if (f < (acc += cfg.tie_ppm)) family = 4;
        if (family) special_family(r, family, len, s1, s2);

        const uint32_t eol = cfg.crlf ? 2 : 1;
        uint64_t written = 0;
        for (int m = 0; m < (pe ? 2 : 1); ++m)
        {
            uint8_t* o = (m ? o2 : o1) + off;
            uint8_t* p = o;
            std::memcpy(p, "@SYN.", 5); p += 5;
            char num[24]; uint32_t nd = 0; uint64_t v = idx;
            do { num[nd++] = (char)('0' + v % 10); v /= 10; } while (v);
            while (nd) *p++ = (uint8_t)num[--nd];
            if (pe) { *p++ = '/'; *p++ = (uint8_t)('1' + m); }
            if (cfg.header_comments)
            {
                std::memcpy(p, " len=", 5); p += 5;
                v = len; nd = 0;
                do { num[nd++] = (char)('0' + v % 10); v /= 10; } while (v);
                while (nd) *p++ = (uint8_t)num[--nd];
                std::memcpy(p, " synthetic", 10); p += 10;
            }
            const uint32_t tlen = (uint32_t)(p - o);
            if (cfg.crlf) *p++ = '\r';
            *p++ = '\n';
            const uint64_t seq_off = off + (uint64_t)(p - o);
            std::memcpy(p, m ? s2 : s1, len); p += len;
            if (cfg.crlf) *p++ = '\r';
            *p++ = '\n';
            *p++ = '+';
            if (cfg.crlf) *p++ = '\r';
            *p++ = '\n';
            const uint64_t qua_off = off + (uint64_t)(p - o);
            std::memcpy(p, m ? q2 : q1, len); p += len;
            if (cfg.crlf) *p++ = '\r';
            *p++ = '\n';
            written = (uint64_t)(p - o);
            fsb_record* rec = m ? r2 : r1;
            if (rec)
            {
                rec->head_off = (uint32_t)off;
                rec->seq_off = (uint32_t)seq_off;
                rec->qua_off = (uint32_t)qua_off;
                rec->seq_len = (uint16_t)len;
                rec->head_len = (uint8_t)tlen;
                rec->reserved = 0;
            }
        }
        (void)eol;
        return written;
    }
};

} // namespace

extern "C" int fsh_synth_size(const fsh_synth_config* cfg, uint64_t* bytes1, uint64_t* bytes2)
{
    if (!cfg || cfg->read_len < 1 || cfg->read_len > 255) return FSB_ERR_PARAM;
    Generator g(*cfg);
    uint64_t total = 0;
    for (uint64_t i = 0; i < cfg->n_records; ++i) total += g.record_bytes(cfg->first_index + i);
    if (bytes1) *bytes1 = total;
    if (bytes2) *bytes2 = cfg->paired ? total : 0;
    return FSB_OK;
}

extern "C" int fsh_synth_fill(const fsh_synth_config* cfg, uint8_t* text1, uint8_t* text2,
                              fsb_record* records1, fsb_record* records2, int threads)
{
    if (!cfg || cfg->read_len < 1 || cfg->read_len > 255 || !text1 || (cfg->paired && !text2)) return FSB_ERR_PARAM;
    Generator g(*cfg);
    const uint64_t n = cfg->n_records;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;

    // pass 1: byte offset of each slice
    std::vector<uint64_t> slice_bytes((size_t)threads, 0);
    auto slice_begin = [&](int t) { return n * (uint64_t)t / (uint64_t)threads; };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t)
            th.emplace_back([&, t]() {
                uint64_t b = 0;
                for (uint64_t i = slice_begin(t); i < slice_begin(t + 1); ++i) b += g.record_bytes(cfg->first_index + i);
                slice_bytes[(size_t)t] = b;
            });
        for (auto& x : th) x.join();
    }
    std::vector<uint64_t> slice_off((size_t)threads + 1, 0);
    for (int t = 0; t < threads; ++t) slice_off[(size_t)t + 1] = slice_off[(size_t)t] + slice_bytes[(size_t)t];
    if (slice_off[(size_t)threads] >= 0xFFFFFFFFull && (records1 || records2)) return FSB_ERR_PARAM;   // u32 offsets

    // pass 2: fill
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([&, t]() {
            uint64_t off = slice_off[(size_t)t];
            for (uint64_t i = slice_begin(t); i < slice_begin(t + 1); ++i)
                off += g.write_record(cfg->first_index + i, text1, text2, off,
                                      records1 ? records1 + i : nullptr, (records2 && cfg->paired) ? records2 + i : nullptr);
        });
    for (auto& x : th) x.join();
    return FSB_OK;
}
