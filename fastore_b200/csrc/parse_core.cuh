// parse_core.cuh -- the word arithmetic of the device-side FASTQ parse (parse.cuh): which bytes of a 16-byte vector end a line.
//
// A line ends at LF, at CR LF (the CR is not an end of its own) or at a lone CR -- SkipLine, FastqParser.cpp:46-68.
// FSB_HD code: tests/emul/ runs it on the host against a byte-by-byte restatement (CPU tier).
#pragma once

#include "core.cuh"

namespace fsb {

// 0x80 in every byte of w that equals the byte repeated in pattern4 (exact: no borrow crosses a byte)
FSB_HD uint32_t bytes_equal(uint32_t w, uint32_t pattern4)
{
    const uint32_t x = w ^ pattern4;
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}
// The flags of two words (0x80 per byte) as eight mask bits, byte b of word j at bit 4 j + b: with c = f0 >> 7 | f1 >> 3 byte b
// holds its two flags at bits 0 and 4, and one multiplication moves byte b down by 7 b places next to its neighbours (every
// stray product term lands on a position of its own below bit 21 or above bit 28, so nothing carries into the result).
FSB_HD uint32_t gather_flags8(uint32_t f0, uint32_t f1)
{
    const uint32_t c = (f0 >> 7) | (f1 >> 3);
    return ((c * ((1u << 21) | (1u << 14) | (1u << 7) | 1u)) >> 21) & 0xFFu;
}
// bit b of the result: byte b of the vector ends a line -- it is LF, or CR not followed by LF (FastqParser.cpp:46-68).
// `next` = the byte behind the vector (0 if none); bytes from `valid_bytes` on do not belong to the text.
FSB_HD uint32_t line_end_mask(const uint32_t (&w)[4], uint32_t next, uint32_t valid_bytes)
{
    uint32_t lf[5], e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) lf[j] = bytes_equal(w[j], 0x0A0A0A0Au);
    lf[4] = next == '\n' ? 0x80u : 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        const uint32_t lf_behind = (lf[j] >> 8) | (lf[j + 1] << 24);           // the LF flag of the byte behind every byte
        e[j] = lf[j] | (bytes_equal(w[j], 0x0D0D0D0Du) & ~lf_behind);
    }
    const uint32_t m = gather_flags8(e[0], e[1]) | (gather_flags8(e[2], e[3]) << 8);
    return valid_bytes >= 16u ? m : (m & ((1u << valid_bytes) - 1u));
}

enum { kStopNone = 0, kStopBadTitle = 1, kStopEmptyPlus = 2, kStopLenMismatch = 3 };     // FSH_STOP_* (host_api.h)

// Record candidate r of a text = its lines 4r .. 4r+3 (SingleFastqRecordParser::ReadNextRecord, FastqParser.cpp:118-165).
//   ls        line starts: ls[j] = first byte of line j (n_ends + 1 entries; ls[0] = 0)
//   n_ends    line ends in the text; n_lines = n_ends, or one more if the text does not end with a line end
// Returns why the reference's loop would stop at this candidate (kStopNone: it is a record, written to `o`); `invalid` = the
// record is outside the device contract (read length 1..255, title at most 255 bytes).
FSB_HD uint32_t parse_candidate(const uint8_t* text, uint32_t size, const uint32_t* ls, uint32_t n_ends, uint32_t n_lines, uint32_t r,
                                bool keep_headers, bool keep_comments, fsb_record& o, bool& invalid)
{
    // what SkipLine returns for line j and where the line starts; lines past the last one are empty (the scan has hit the end of the memory)
    uint32_t start[4], len[4];
#pragma unroll
    for (uint32_t q = 0; q < 4; ++q)
    {
        const uint32_t j = 4u * r + q;
        if (j >= n_lines) { start[q] = size; len[q] = 0; }
        else
        {
            start[q] = ls[j];
            if (j < n_ends)
            {   // the line has a line end at ls[j + 1] - 1
                const uint32_t term = ls[j + 1] - 1u;
                const bool crlf = text[term] == '\n' && term > start[q] && text[term - 1] == '\r';
                len[q] = term - start[q] - (crlf ? 1u : 0u);
            }
            else len[q] = size - start[q];                          // the last line of a text without a final line end
        }
    }
    const uint32_t titleLen = len[0], seqLen = len[1], plusLen = len[2], quaLen = len[3];
    invalid = false;
    if (titleLen == 0 || text[start[0]] != '@') return kStopBadTitle;                      // FastqParser.cpp:125
    if ((plusLen & 0xFFFFu) == 0) return kStopEmptyPlus;                                   // :132-134 (uint16 plen)
    if ((quaLen & 0xFFFFu) != seqLen) return kStopLenMismatch;                             // :137-139 (uint16 qlen)
    uint32_t headLen = 0;
    if (keep_headers)
    {
        headLen = titleLen;
        if (!keep_comments)                                        // :148-155: the title ends at the first space
        {
            const uint32_t lim = titleLen < 257u ? titleLen : 257u; // beyond 255 the record is outside the contract anyway
            for (uint32_t i = 0; i < lim; ++i) if (text[start[0] + i] == ' ') { headLen = i; break; }
        }
    }
    invalid = seqLen < 1 || seqLen > 255 || headLen > 255;
    o.head_off = start[0]; o.seq_off = start[1]; o.qua_off = start[3];
    o.seq_len = (uint16_t)seqLen; o.head_len = (uint8_t)headLen; o.reserved = 0;
    return kStopNone;
}

} // namespace fsb
