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

} // namespace fsb
