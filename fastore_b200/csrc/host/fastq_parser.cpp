// fastq_parser.cpp -- host FASTQ record parser: chunk text -> record table for the C ABI.
//
// Same acceptance rules as the reference's SingleFastqRecordParser::ReadNextRecord
// (FastqParser.cpp:118-165) and SkipLine (:46-68): four lines per record, LF / CR / CRLF line ends,
// the title must start with '@', the '+' line must be non-empty, len(quality) == len(sequence);
// parsing of the chunk stops silently at the first record that breaks one of these.  With
// keep_comments == 0 (-C) the header is cut at the first space (:148-155).
//
// The reference leaves a few inputs undefined (out-of-table symbol look-ups FastqRecord.h:95-96,
// 8-bit lengths FastqRecord.h:45-48, FastqPacker.h:59, quality table FastqPacker.cpp:250).  This
// parser flags those records instead (FSB_ERR_INPUT) so that nothing undefined reaches the device.
#include "host_api.h"

#include <cstring>

namespace {

struct Cursor
{
    const uint8_t* mem;
    uint64_t pos;
    uint64_t size;

    // The same result found with memchr: the line ends at the next LF; a CR right in front of it belongs to the line end.
    // A CR anywhere else in the line would end the line for SkipLine: such lines take the byte-wise path.
    uint32_t skip_line_fast()
    {
        if (pos >= size) return 0;
        const uint8_t* p = mem + pos;
        const uint64_t left = size - pos;
        const uint8_t* lf = (const uint8_t*)std::memchr(p, '\n', (size_t)left);
        const uint64_t span = lf ? (uint64_t)(lf - p) : left;          // bytes in front of the LF (or to the end of the chunk)
        const uint8_t* cr = span ? (const uint8_t*)std::memchr(p, '\r', (size_t)span) : nullptr;
        if (cr && !(lf && cr + 1 == lf)) return skip_line();           // a lone CR inside: SkipLine's rule
        pos += span + (lf ? 1 : 0);
        return (uint32_t)(cr ? span - 1 : span);
    }
    // SkipLine (FastqParser.cpp:46-68): returns the line length without the line end
    uint32_t skip_line()
    {
        uint32_t len = 0;
        while (pos < size)
        {
            const uint8_t c = mem[pos++];
            if (c != '\n' && c != '\r') { len++; continue; }
            if (c == '\r' && pos < size && mem[pos] == '\n') pos++;
            break;
        }
        return len;
    }
};

inline bool is_dna(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N'; }

} // namespace

extern "C" uint64_t fsh_max_records(const uint8_t* text, uint64_t size)
{
    // a record has at least 4 line ends.  SkipLine (FastqParser.cpp:46-68) accepts LF, CRLF and a lone CR,
    // so line ends = LFs + CRs that are not followed by an LF (two memchr sweeps; the second finds
    // nothing in an LF-only file)
    auto count = [&](int ch, bool lone_only) {
        uint64_t n = 0;
        const uint8_t* p = text;
        const uint8_t* e = text + size;
        while (p < e)
        {
            const uint8_t* q = (const uint8_t*)std::memchr(p, ch, (size_t)(e - p));
            if (!q) break;
            if (!lone_only || q + 1 >= e || q[1] != '\n') n++;
            p = q + 1;
        }
        return n;
    };
    return (count('\n', false) + count('\r', true)) / 4 + 2;
}

extern "C" int fsh_parse_chunk(const uint8_t* text, uint64_t size, int keep_headers, int keep_comments,
                               int quality_offset, int quality_method,
                               fsb_record* records, uint64_t capacity, fsh_parse_stats* stats)
{
    return fsh_parse_chunk_ex(text, size, keep_headers, keep_comments, quality_offset, quality_method, 1, records, capacity, stats);
}

extern "C" int fsh_parse_chunk_ex(const uint8_t* text, uint64_t size, int keep_headers, int keep_comments,
                                  int quality_offset, int quality_method, int validate_bytes,
                                  fsb_record* records, uint64_t capacity, fsh_parse_stats* stats)
{
    fsh_parse_stats st;
    std::memset(&st, 0, sizeof(st));
    st.min_seq_len = 0xFFFFFFFFu;
    if (size >= 0xFFFFFFFFull) { if (stats) *stats = st; return FSB_ERR_INPUT; }   // u32 offsets

    Cursor c{text, 0, size};
    const bool range_checked = (quality_method == FSB_QUA_NONE || quality_method == FSB_QUA_8BIN || quality_method == FSB_QUA_QVZ);
    uint64_t n = 0;
    while (c.pos < size)                                           // FastqParser.cpp:120
    {
        const uint64_t title = c.pos;
        const uint32_t titleLen = c.skip_line_fast();
        if (titleLen == 0 || text[title] != '@') { st.stop_reason = FSH_STOP_BAD_TITLE; break; }    // :125
        const uint64_t seq = c.pos;
        const uint32_t seqLen = c.skip_line_fast();
        const uint32_t plen = c.skip_line_fast();
        if ((uint16_t)plen == 0) { st.stop_reason = FSH_STOP_EMPTY_PLUS; break; }                   // :132-134 (uint16 plen)
        const uint64_t qua = c.pos;
        const uint32_t qlen = c.skip_line_fast();
        if ((uint16_t)qlen != seqLen) { st.stop_reason = FSH_STOP_LEN_MISMATCH; break; }            // :137-139 (uint16 qlen)
        if (n >= capacity) { st.stop_reason = FSH_STOP_CAPACITY; break; }

        uint32_t headLen = 0;
        if (keep_headers)
        {
            headLen = titleLen;
            if (!keep_comments)                                    // :148-155
            {
                const void* sp = std::memchr(text + title, ' ', titleLen);
                if (sp) headLen = (uint32_t)((const uint8_t*)sp - (text + title));
            }
        }

        bool ok = seqLen >= 1 && seqLen <= 255 && headLen <= 255;
        if (ok && validate_bytes)
        {
            for (uint32_t i = 0; i < seqLen; ++i) ok &= is_dna(text[seq + i]);
            if (range_checked)
                for (uint32_t i = 0; i < seqLen; ++i)
                {
                    const int q = (int)text[qua + i] - quality_offset;
                    ok &= (q >= 0 && q < 64);
                }
            else
                for (uint32_t i = 0; i < seqLen; ++i) ok &= ((int)text[qua + i] >= quality_offset);
            for (uint32_t i = 1; i < headLen; ++i) ok &= (text[title + i] < 128);
        }
        if (!ok) st.invalid_records++;

        fsb_record& r = records[n++];
        r.head_off = (uint32_t)title;
        r.seq_off = (uint32_t)seq;
        r.qua_off = (uint32_t)qua;
        r.seq_len = (uint16_t)seqLen;
        r.head_len = (uint8_t)headLen;
        r.reserved = 0;
        if (seqLen < st.min_seq_len) st.min_seq_len = seqLen;
        if (seqLen > st.max_seq_len) st.max_seq_len = seqLen;
        st.consumed_bytes = c.pos;
    }
    st.n_records = n;
    if (n == 0) st.min_seq_len = 0;
    if (stats) *stats = st;
    return st.invalid_records ? FSB_ERR_INPUT : FSB_OK;
}

// IFastqStreamReaderBase::GetNextRecordPos (FastqStream.cpp:15-40)
extern "C" uint64_t fsh_cut_position(const uint8_t* buf, uint64_t size, uint64_t window)
{
    auto skip_to_eol = [&](uint64_t& p) { while (p < size && buf[p] != '\n' && buf[p] != '\r') ++p; if (p + 1 < size && buf[p] == '\r' && buf[p + 1] == '\n') ++p; };
    uint64_t pos = size - window;
    skip_to_eol(pos); ++pos;
    while (pos < size && buf[pos] != '@') { skip_to_eol(pos); ++pos; }
    const uint64_t pos0 = pos;
    skip_to_eol(pos); ++pos;
    if (pos < size && buf[pos] == '@') return pos;      // pos0 was a quality line
    return pos0;
}
