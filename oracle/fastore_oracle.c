/*
 * fastore_oracle.c -- CPU restatement ("port") of the reference fastore_bin categorise + pack path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h).  Plain C99, scalar, one record at a time: written
 * to be obviously equal to the reference, not to be fast.  Every function cites the reference
 * lines it follows (paths relative to /root/reference/fastore/fastore_bin/).
 *
 * Parity pinning: the reference ships no golden vectors or unit tests, so this
 * port is pinned against the *compiled reference itself*: oracle/ref_harness.cpp links the
 * reference's FastqCategorizer.o / FastqPacker.o and tests/test_oracle_vs_reference.py requires
 * byte-identical streams, descriptors and per-read tuples on every synthetic family, and
 * tests/golden/ holds vectors produced by that compiled reference (tools/make_golden.py).
 */
#include "oracle_api.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------ */
/* bit writer: BitMemoryWriter (BitMemory.h:216-433).  The reference keeps a 32-bit accumulator,
 * appends MSB-first (PutBit/Put2Bits/PutBits :250-313), emits full words big-endian (Put4Bytes
 * :359-365) and on FlushPartialWordBuffer (:375-390) left-aligns the tail to a byte and emits the
 * used bytes.  Net effect, restated: one MSB-first bit stream, zero-padded to a byte on flush. */
typedef struct {
    uint8_t* buf;
    uint64_t cap;
    uint64_t pos;      /* bytes emitted */
    uint32_t acc;      /* pending bits, right-aligned */
    uint32_t nacc;     /* number of pending bits, < 8 */
} bitw;

static void bw_init(bitw* w) { w->cap = 1 << 16; w->buf = (uint8_t*)malloc(w->cap); w->pos = 0; w->acc = 0; w->nacc = 0; }

static void bw_byte(bitw* w, uint8_t b)
{
    if (w->pos >= w->cap) { w->cap += w->cap >> 1; w->buf = (uint8_t*)realloc(w->buf, w->cap); }
    w->buf[w->pos++] = b;
}

/* PutBits(word, n): the n low bits of word, most significant first (BitMemory.h:294-313) */
static void bw_put(bitw* w, uint32_t value, uint32_t n)
{
    for (int32_t i = (int32_t)n - 1; i >= 0; --i) {
        w->acc = (w->acc << 1) | ((value >> i) & 1u);
        if (++w->nacc == 8) { bw_byte(w, (uint8_t)w->acc); w->acc = 0; w->nacc = 0; }
    }
}

/* FlushPartialWordBuffer (BitMemory.h:375-390) */
static void bw_flush(bitw* w)
{
    if (w->nacc) { bw_byte(w, (uint8_t)(w->acc << (8 - w->nacc))); w->acc = 0; w->nacc = 0; }
}

/* ------------------------------------------------------------------------------------------ */
static uint32_t nbin_value(const fsb_params* p) { return 1u << (2 * p->signature_len); } /* FastqCategorizer.cpp:24-25 */

/* symbolIdxTable (FastqCategorizer.cpp:27-29) / dnaToIdx (FastqPacker.cpp:24-30) */
static int sym_idx(const fsb_params* p, uint8_t c)
{
    for (int i = 0; i < 5; ++i)
        if ((uint8_t)p->dna_symbol_order[i] == c) return i;
    return -1;
}

/* validBinSignatures[m] (InitializeValidBinSignatures, FastqCategorizer.cpp:34-76), evaluated on
 * demand with the same tests in the same order. */
int orc_signature_valid(const fsb_params* p, uint32_t i)
{
    const uint32_t k = p->signature_len;
    const uint32_t loMask = (1u << p->signature_mask_cutoff_bits) - 1;      /* :49 */
    int isInvalid = (i & loMask) != 0;                                       /* :53 */
    uint32_t m = i >> (2 * k - 6);                                           /* :56 */
    isInvalid |= (m == 0u) || (m == 1u);                                     /* :57  AAA / AAC prefix */
    m = i;
    for (uint32_t j = 0; !isInvalid && j < k - 2; ++j) {                     /* :60 */
        isInvalid |= ((m & 0xFu) == 0);                                      /* :63  ..AA.. */
        m >>= 2;
    }
    return !isInvalid;
}

/* ComputeMinimizer (FastqCategorizer.cpp:138-152) */
static uint32_t compute_minimizer(const fsb_params* p, const uint8_t* dna, uint32_t k)
{
    uint32_t r = 0;
    for (uint32_t i = 0; i < k; ++i) {
        if (dna[i] == 'N') return nbin_value(p);
        r = (r << 2) + (uint32_t)sym_idx(p, dna[i]);
    }
    return r;
}

/* FindMinimizer (FastqCategorizer.cpp:79-106) */
void orc_find_minimizer(const fsb_params* p, const uint8_t* seq, uint32_t len, uint32_t* sig, uint32_t* pos_out)
{
    const uint32_t nbin = nbin_value(p);
    uint32_t minimizer = nbin;
    uint32_t pos = 0;
    const int32_t end = (int32_t)len - (int32_t)p->signature_len - (int32_t)p->skip_zone_len;  /* :88 */
    for (int32_t i = 0; i < end; ++i) {
        uint32_t m = compute_minimizer(p, seq + i, p->signature_len);
        if (m < minimizer && orc_signature_valid(p, m)) { minimizer = m; pos = (uint32_t)i; }  /* :93 */
    }
    uint32_t ncount = 0;
    for (uint32_t i = 0; i < len; ++i) ncount += (seq[i] == 'N');
    if (minimizer >= nbin || ncount >= len / 3) { *sig = nbin; *pos_out = 0; return; }          /* :102 */
    *sig = minimizer; *pos_out = pos;
}

/* DnaRebalancer::FindMinimizerHR (fastore_rebin/DnaRebalancer.cpp:570-601): FindMinimizer with m != curSignature and
 * m % curDivisor == 0 added (binParams.validBinSignatures is all true by default, RebinModule.cpp:72) */
static void find_minimizer_hr(const fsb_params* p, const uint8_t* seq, uint32_t len, uint32_t cur, uint32_t divisor, uint32_t* sig, uint32_t* pos_out)
{
    const uint32_t nbin = nbin_value(p);
    uint32_t minimizer = nbin;                                              /* maxLongMinimValue = 4^k */
    uint32_t pos = 0;
    const int32_t end = (int32_t)len - (int32_t)p->signature_len - (int32_t)p->skip_zone_len;  /* :580 */
    for (int32_t i = 0; i < end; ++i) {
        uint32_t m = compute_minimizer(p, seq + i, p->signature_len);
        if (m < minimizer && m != cur && m % divisor == 0 && orc_signature_valid(p, m)) { minimizer = m; pos = (uint32_t)i; }  /* :584-592 */
    }
    uint32_t ncount = 0;
    for (uint32_t i = 0; i < len; ++i) ncount += (seq[i] == 'N');
    if (minimizer >= nbin || ncount >= len / 3) { *sig = nbin; *pos_out = 0; return; }          /* :597-599 */
    *sig = minimizer; *pos_out = pos;
}

static uint8_t rc_code(uint8_t c);
/* DnaRebalancer::FindNewMinimizer (:604-616): the scan on the read and on its reverse complement; the reverse strand only
 * if its signature is strictly smaller */
void orc_find_new_minimizer(const fsb_params* p, const uint8_t* seq, uint32_t len, uint32_t cur, uint32_t divisor, uint32_t* sig, uint32_t* pos, uint32_t* is_rev)
{
    uint8_t rc[256];
    for (uint32_t i = 0; i < len; ++i) rc[len - 1 - i] = rc_code(seq[i]);
    uint32_t fs, fp, rs, rp;
    find_minimizer_hr(p, seq, len, cur, divisor, &fs, &fp);
    find_minimizer_hr(p, rc, len, cur, divisor, &rs, &rp);
    if (fs > rs) { *sig = rs; *pos = rp; *is_rev = 1; } else { *sig = fs; *pos = fp; *is_rev = 0; }
}

/* FastqRecord::ComputeRC (FastqRecord.h:80-111): reverse-complement of the whole span, quality
 * reversed alongside. */
static uint8_t rc_code(uint8_t c)
{
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'N': return 'N'; }
    return 0xFF;
}
static void compute_rc(const uint8_t* seq, const uint8_t* qua, uint32_t len, uint8_t* rseq, uint8_t* rqua)
{
    for (uint32_t i = 0; i < len; ++i) { rseq[len - 1 - i] = rc_code(seq[i]); rqua[len - 1 - i] = qua[i]; }
}

/* ------------------------------------------------------------------------------------------ */
/* one categorised record: the FastqRecord after DistributeToBins mutated it in place */
typedef struct {
    uint8_t* seq;       /* seqLen + auxLen symbols, stored orientation */
    uint8_t* qua;
    const uint8_t* head;
    uint32_t seqLen, auxLen, headLen;
    uint32_t minimPos, flags;    /* flags: bit0 reverse, bit1 swapped */
    uint32_t sig;
} orec;

/* FastqCategorizerSE::DistributeToBins (FastqCategorizer.cpp:197-253), one record */
static void categorize_se(const fsb_params* p, orec* r, uint8_t* rcs, uint8_t* rcq)
{
    const uint32_t nbin = nbin_value(p);
    uint32_t fs, fp, rs, rp;
    compute_rc(r->seq, r->qua, r->seqLen, rcs, rcq);                     /* :208 */
    orc_find_minimizer(p, r->seq, r->seqLen, &fs, &fp);                  /* :212 */
    orc_find_minimizer(p, rcs, r->seqLen, &rs, &rp);                     /* :213 */
    uint32_t sig, pos; int reverse = 0;
    if (fs <= rs) { sig = fs; pos = fp; } else { sig = rs; pos = rp; reverse = 1; }   /* :217-225 */
    r->flags = 0; r->minimPos = 0;
    if (sig != nbin) {                                                   /* :230-239 */
        if (reverse) { r->flags |= 1; memcpy(r->seq, rcs, r->seqLen); memcpy(r->qua, rcq, r->seqLen); }
        r->minimPos = pos;
    }
    r->sig = sig;
}

/* FastqCategorizerPE::DistributeToBins (FastqCategorizer.cpp:256-363), one pair */
static void categorize_pe(const fsb_params* p, orec* r, uint8_t* rcs, uint8_t* rcq)
{
    const uint32_t nbin = nbin_value(p);
    const uint32_t L1 = r->seqLen, L2 = r->auxLen, len = L1 + L2;
    uint32_t f1s, f1p, f2s, f2p, r1s, r1p, r2s, r2p;
    compute_rc(r->seq, r->qua, len, rcs, rcq);        /* :272 -> [rc(m2) | rc(m1)], recRev.seqLen = L2 */
    orc_find_minimizer(p, r->seq, L1, &f1s, &f1p);            /* :281 minFwd_1 */
    orc_find_minimizer(p, rcs, L2, &r1s, &r1p);               /* :282 minRev_1 = rc(m2) */
    orc_find_minimizer(p, r->seq + L1, L2, &f2s, &f2p);       /* :286 minFwd_2 */
    orc_find_minimizer(p, rcs + L2, L1, &r2s, &r2p);          /* :287 minRev_2 = rc(m1) */
    int isF1 = f1s < f2s;                                     /* :289 */
    uint32_t Fs = isF1 ? f1s : f2s, Fp = isF1 ? f1p : f2p;
    int isR1 = r1s < r2s;                                     /* :292 */
    uint32_t Rs = isR1 ? r1s : r2s, Rp = isR1 ? r1p : r2p;
    uint32_t sig, pos; int isRev = 0, isFwdMinim;
    if (Fs < Rs) { sig = Fs; pos = Fp; isFwdMinim = isF1; }   /* :295-299 */
    else { sig = Rs; pos = Rp; isRev = 1; isFwdMinim = isR1; }/* :300-305 */
    r->flags = 0; r->minimPos = 0;                            /* rec.Reset() :263 */
    if (sig != nbin) {                                        /* :321-336 */
        if (isRev) { memcpy(r->seq, rcs, len); memcpy(r->qua, rcq, len); r->flags |= 1; }
        if (!isFwdMinim) {                                    /* SwapReads (FastqRecord.h:190-199) */
            for (uint32_t i = 0; i < L1; ++i) {
                uint8_t t = r->seq[i]; r->seq[i] = r->seq[L1 + i]; r->seq[L1 + i] = t;
                t = r->qua[i]; r->qua[i] = r->qua[L1 + i]; r->qua[L1 + i] = t;
            }
            r->flags ^= 2;
        }
        r->minimPos = pos;
    }
    r->sig = sig;
}

/* ------------------------------------------------------------------------------------------ */
/* quaToIdx_8bin (IFastqPacker ctor, FastqPacker.cpp:41-64) */
static void build_qua8(uint8_t out[64])
{
    static const uint8_t t[64] = { 0, 0, 6, 6, 6, 6, 6, 6, 6, 6, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 22, 22, 22, 22, 22,
                                   27, 27, 27, 27, 27, 33, 33, 33, 33, 33, 37, 37, 37, 37, 37, 40, 40, 40, 40, 40, 40, 40, 40,
                                   40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40 };
    uint8_t prev = 0; uint32_t sym = 0;
    for (uint32_t i = 0; i < 64; ++i) { if (t[i] != prev) { prev = t[i]; sym++; } out[i] = (uint8_t)sym; }
}

static uint32_t bit_length(uint64_t x)     /* Utils.h:235-243 */
{
    for (uint32_t i = 0; i < 32; ++i) if (x < (1ull << i)) return i;
    return 64;
}

static uint32_t quality_bits(const fsb_params* p)   /* QualityCompressionParams::BitsPerBase, Quality.h:58-64 */
{
    static const uint32_t bpb[4] = { 6, 1, 3, 6 };
    return bpb[p->quality_method & 3];
}

typedef struct { bitw meta, dna, qua, head; const fsb_params* p; uint8_t qua8[64]; } packer;

/* StoreNextRecord + StoreDna + StoreQuality + StoreHeader (FastqPacker.cpp:113-287) for one mate */
static void store_next_record(packer* pk, const uint8_t* seq, const uint8_t* qua, uint32_t seqLen, uint32_t minimPos,
                              int isReverse, const uint8_t* head, uint32_t headLen, uint32_t suffixLen, int usesHeaders)
{
    const fsb_params* p = pk->p;
    if (suffixLen != 0) {                                    /* :124-129 */
        bw_put(&pk->meta, (uint32_t)isReverse, 1);
        bw_put(&pk->meta, minimPos, 8);
    }
    int isDnaPlain = 1;                                      /* :164 */
    for (uint32_t i = 0; i < seqLen; ++i) if (seq[i] == 'N') { isDnaPlain = 0; break; }
    bw_put(&pk->meta, (uint32_t)isDnaPlain, 1);              /* :165 */
    const uint32_t bits = isDnaPlain ? 2 : 3;                /* :170-201 */
    for (uint32_t i = 0; i < minimPos; ++i) bw_put(&pk->dna, (uint32_t)sym_idx(p, seq[i]), bits);
    for (uint32_t i = minimPos + suffixLen; i < seqLen; ++i) bw_put(&pk->dna, (uint32_t)sym_idx(p, seq[i]), bits);

    const uint32_t qbits = quality_bits(p);                  /* :214-268 */
    for (uint32_t i = 0; i < seqLen; ++i) {
        uint32_t c = (uint32_t)qua[i] - p->quality_offset;
        switch (p->quality_method) {
        case FSB_QUA_BINARY: bw_put(&pk->qua, c >= p->binary_threshold, 1); break;
        case FSB_QUA_8BIN:   bw_put(&pk->qua, pk->qua8[c & 63], qbits); break;
        default:             bw_put(&pk->qua, c, qbits); break;          /* MET_NONE, MET_QVZ */
        }
    }
    if (usesHeaders) {                                       /* :272-287 */
        bw_put(&pk->head, headLen, 8);
        for (uint32_t i = 1; i < headLen; ++i) bw_put(&pk->head, head[i], 7);
    }
}

static int cmp_u64(const void* a, const void* b)
{
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return (x > y) - (x < y);
}

/* Categorize (FastqCategorizer.cpp:169-192) + PackToBins (FastqPacker.cpp:417-491) over one chunk */
static int bin_records(const fsb_params* p, const fsb_chunk* ch, uint64_t first, uint64_t n, orc_block* out, int keep_per_read)
{
    const int pe = p->paired_end != 0;
    const uint32_t nbin = nbin_value(p);
    memset(out, 0, sizeof(*out));
    out->n_records = n;
    if (n == 0) return FSB_OK;

    /* materialise the records the way the parser does: SE views the chunk (FastqParser.cpp:315-343);
     * PE copies m1|m2 then q1|q2 per pair (FastqParser.cpp:527-553) */
    orec* recs = (orec*)calloc(n, sizeof(orec));
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const fsb_record* a = &ch->records[0][first + i];
        total += a->seq_len;
        if (pe) total += ch->records[1][first + i].seq_len;
    }
    uint8_t* seqbuf = (uint8_t*)malloc(total + 1);
    uint8_t* quabuf = (uint8_t*)malloc(total + 1);
    uint64_t off = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const fsb_record* a = &ch->records[0][first + i];
        orec* r = &recs[i];
        r->seq = seqbuf + off; r->qua = quabuf + off;
        r->seqLen = a->seq_len; r->headLen = a->head_len;
        r->head = ch->text[0] + a->head_off;
        memcpy(r->seq, ch->text[0] + a->seq_off, a->seq_len);
        memcpy(r->qua, ch->text[0] + a->qua_off, a->seq_len);
        off += a->seq_len;
        if (pe) {
            const fsb_record* b = &ch->records[1][first + i];
            if (b->seq_len != a->seq_len) { free(recs); free(seqbuf); free(quabuf); return FSB_ERR_INPUT; }
            r->auxLen = b->seq_len;
            memcpy(r->seq + a->seq_len, ch->text[1] + b->seq_off, b->seq_len);
            memcpy(r->qua + a->seq_len, ch->text[1] + b->qua_off, b->seq_len);
            off += b->seq_len;
        }
    }

    /* categorise in parse order */
    uint8_t rcs[1024], rcq[1024];
    for (uint64_t i = 0; i < n; ++i) {
        if (pe) categorize_pe(p, &recs[i], rcs, rcq); else categorize_se(p, &recs[i], rcs, rcq);
    }
    if (keep_per_read) {
        out->read_signature = (uint32_t*)malloc(n * sizeof(uint32_t));
        out->read_info = (uint32_t*)malloc(n * sizeof(uint32_t));
        for (uint64_t i = 0; i < n; ++i) {
            const orec* r = &recs[i];
            uint32_t info = r->minimPos | ((r->flags & 1) ? FSB_INFO_REVERSE : 0) | ((r->flags & 2) ? FSB_INFO_SWAPPED : 0);
            int plainA = 1, plainB = 1;
            for (uint32_t j = 0; j < r->seqLen; ++j) if (r->seq[j] == 'N') plainA = 0;
            for (uint32_t j = 0; j < r->auxLen; ++j) if (r->seq[r->seqLen + j] == 'N') plainB = 0;
            info |= plainA ? FSB_INFO_PLAIN_A : 0;
            if (pe) info |= plainB ? FSB_INFO_PLAIN_B : 0;
            out->read_signature[i] = r->sig;
            out->read_info[i] = info;
        }
    }

    /* bins_: std::map<uint32, FastqRecordsPtrBin>, records pushed back in parse order
     * (FastqCategorizer.cpp:247,357) == stable grouping by signature, ascending; the N-bin key 4^k
     * is the largest so it comes last (FastqPacker.cpp:430-485). */
    uint64_t* order = (uint64_t*)malloc(n * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) order[i] = ((uint64_t)recs[i].sig << 32) | i;   /* n < 2^32 per chunk */
    qsort(order, n, sizeof(uint64_t), cmp_u64);

    uint64_t nb = 0;
    for (uint64_t i = 0; i < n; ++i) if (i == 0 || (order[i] >> 32) != (order[i - 1] >> 32)) nb++;
    out->bins = (fsb_bin_descriptor*)calloc(nb, sizeof(fsb_bin_descriptor));
    out->n_bins = nb;

    packer pk; pk.p = p; build_qua8(pk.qua8);
    bw_init(&pk.meta); bw_init(&pk.dna); bw_init(&pk.qua); bw_init(&pk.head);
    const int usesHeaders = p->reads_have_headers != 0;

    uint64_t b = 0;
    for (uint64_t s = 0; s < n; ++b) {
        uint64_t e = s;
        const uint32_t sig = (uint32_t)(order[s] >> 32);
        while (e < n && (uint32_t)(order[e] >> 32) == sig) e++;
        /* PackToBin (FastqPacker.cpp:541-602) */
        const int nBin = (sig == nbin);
        uint32_t minLen = 0xFFFFFFFFu, maxLen = 0;       /* FastqRecordBinStats over seqLen (:358-359 / FastqRecord.h:312-319) */
        for (uint64_t j = s; j < e; ++j) {
            const orec* r = &recs[(uint32_t)order[j]];
            if (r->seqLen < minLen) minLen = r->seqLen;
            if (r->seqLen > maxLen) maxLen = r->seqLen;
        }
        const int hasConstLen = (minLen == maxLen);                         /* :566 */
        const uint32_t suffixLen = nBin ? 0 : p->signature_len;             /* :570-573 */
        const uint32_t bitsPerLen = hasConstLen ? 0 : bit_length(maxLen - minLen);   /* :576-577 */
        const uint64_t m0 = pk.meta.pos, d0 = pk.dna.pos, q0 = pk.qua.pos, h0 = pk.head.pos;
        bw_put(&pk.meta, minLen, 8);                                        /* :581-583 */
        bw_put(&pk.meta, maxLen, 8);
        bw_put(&pk.meta, 0, 1);
        fsb_bin_descriptor* d = &out->bins[b];
        d->signature = sig;
        for (uint64_t j = s; j < e; ++j) {
            const orec* r = &recs[(uint32_t)order[j]];
            if (!pe) {                                                       /* StoreRecords SE :734-759 */
                if (!hasConstLen) bw_put(&pk.meta, r->seqLen - minLen, bitsPerLen);
                store_next_record(&pk, r->seq, r->qua, r->seqLen, r->minimPos, r->flags & 1, r->head, r->headLen, suffixLen, usesHeaders);
                d->raw_dna_size += r->seqLen;
            } else {                                                         /* StoreRecords PE :815-859 */
                if (!hasConstLen) {
                    bw_put(&pk.meta, r->seqLen - minLen, bitsPerLen);
                    bw_put(&pk.meta, r->auxLen - minLen, bitsPerLen);
                }
                if (suffixLen != 0) bw_put(&pk.meta, (r->flags >> 1) & 1, 1);
                store_next_record(&pk, r->seq, r->qua, r->seqLen, r->minimPos, r->flags & 1, r->head, r->headLen, suffixLen, usesHeaders);
                /* mate 2: GetPair() view, minimPos 0, suffixLen 0, no header (:823-825,849-851) */
                store_next_record(&pk, r->seq + r->seqLen, r->qua + r->seqLen, r->auxLen, 0, 0, NULL, 0, 0, 0);
                d->raw_dna_size += r->seqLen + r->auxLen;
            }
            d->raw_head_size += r->headLen;
            d->records_count++;
        }
        bw_flush(&pk.meta); bw_flush(&pk.dna); bw_flush(&pk.qua); bw_flush(&pk.head);   /* :593-596 */
        d->meta_size = pk.meta.pos - m0; d->dna_size = pk.dna.pos - d0;
        d->qua_size = pk.qua.pos - q0; d->head_size = pk.head.pos - h0;
        out->raw_dna_size += d->raw_dna_size; out->raw_head_size += d->raw_head_size;
        s = e;
    }
    out->meta = pk.meta.buf; out->meta_size = pk.meta.pos;
    out->dna = pk.dna.buf;   out->dna_size = pk.dna.pos;
    out->qua = pk.qua.buf;   out->qua_size = pk.qua.pos;
    out->head = pk.head.buf; out->head_size = pk.head.pos;
    free(order); free(recs); free(seqbuf); free(quabuf);
    return FSB_OK;
}

int orc_bin_chunk(const fsb_params* p, const fsb_chunk* chunk, orc_block* out)
{
    if (!p || !chunk || !out) return FSB_ERR_PARAM;
    if (p->signature_len < 3 || p->signature_len > 15) return FSB_ERR_PARAM;
    if (memcmp(p->dna_symbol_order, "ACGTN", 5) != 0) return FSB_ERR_PARAM;
    return bin_records(p, chunk, 0, chunk->n_records, out, 1);
}

void orc_block_free(orc_block* b)
{
    if (!b) return;
    free(b->meta); free(b->dna); free(b->qua); free(b->head); free(b->bins);
    free(b->read_signature); free(b->read_info);
    memset(b, 0, sizeof(*b));
}

/* ------------------------------------------------------------------------------------------ */
/* timing: `threads` workers, each categorises + packs a contiguous slice of the records as its
 * own chunk (how the reference's -t N workers split the input, BinOperator.cpp:71-342), `reps`
 * passes; returns wall seconds of the whole thing.  Parsing and I/O are not included. */
typedef struct { const fsb_params* p; const fsb_chunk* ch; uint64_t first, n; int reps; } tjob;

static void* tmain(void* arg)
{
    tjob* j = (tjob*)arg;
    for (int r = 0; r < j->reps; ++r) {
        orc_block b;
        bin_records(j->p, j->ch, j->first, j->n, &b, 0);
        orc_block_free(&b);
    }
    return NULL;
}

double orc_time_bin_chunk(const fsb_params* p, const fsb_chunk* chunk, int threads, int reps)
{
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    tjob* jobs = (tjob*)malloc(sizeof(tjob) * (size_t)threads);
    const uint64_t n = chunk->n_records;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; ++t) {
        jobs[t].p = p; jobs[t].ch = chunk; jobs[t].reps = reps;
        jobs[t].first = n * (uint64_t)t / (uint64_t)threads;
        jobs[t].n = n * (uint64_t)(t + 1) / (uint64_t)threads - jobs[t].first;
        pthread_create(&th[t], NULL, tmain, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
