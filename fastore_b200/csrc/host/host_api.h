/*
 * host_api.h -- C interface of the host-side library (libfastore_host.so): the C++ code that stays
 * on the CPU either side of the device path.  No CUDA here.
 *
 *   fsh_parse_*   FASTQ record parser  -> record table for the C ABI   (mirrors FastqParser.cpp:118-165)
 *   fsh_cut_*     FASTQ chunk cutter                                   (mirrors FastqStream.cpp:15-256)
 *   fsh_reader_*  FASTQ chunk reader (SE / PE with read-id re-synchronisation) (FastqStream.cpp:44-256)
 *   fsh_writer_*  bin-file writer: .bmeta / .bdna / .bqua / .bhead          (BinFile.cpp:47-462, Stats.cpp:63-170)
 *   fsh_synth_*   seeded synthetic FASTQ generator used by tests and bench.py
 */
#ifndef FASTORE_HOST_API_H
#define FASTORE_HOST_API_H

#include "../../../include/fastore_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- parser ---------------------------------------------------------------------------------- */

/* Per-chunk parse statistics: the FastqRecordBinStats part of FastqRawBlockStats (Stats.h:26-90). */
typedef struct fsh_parse_stats {
    uint64_t n_records;
    uint32_t min_seq_len;
    uint32_t max_seq_len;
    uint64_t consumed_bytes;     /* offset just past the last accepted record */
    uint32_t stop_reason;        /* 0 = end of chunk, else why parsing stopped early (FSH_STOP_*) */
    uint32_t invalid_records;    /* records violating the device contract (symbols, length, quality) */
} fsh_parse_stats;

enum {
    FSH_STOP_END = 0,
    FSH_STOP_BAD_TITLE = 1,      /* title empty or not starting with '@' (FastqParser.cpp:125-126) */
    FSH_STOP_EMPTY_PLUS = 2,     /* '+' line empty (:133-134) */
    FSH_STOP_LEN_MISMATCH = 3,   /* quality length != sequence length (:138-139) */
    FSH_STOP_CAPACITY = 4        /* record table full */
};

/*
 * Parse one chunk of FASTQ text into a record table.  Same acceptance rules as
 * SingleFastqRecordParser::ReadNextRecord: LF or CRLF line ends, parsing stops silently at the
 * first malformed record.  keep_headers = -H, keep_comments = !-C (header cut at the first space).
 * Returns FSB_OK, or FSB_ERR_INPUT if a record breaks the contract the reference leaves undefined
 * (sequence symbol outside ACGTN, read longer than 255, header longer than 255, quality outside
 * [offset, offset+64) for the 6/3-bit modes); such records are counted in invalid_records.
 */
int fsh_parse_chunk(const uint8_t* text, uint64_t size, int keep_headers, int keep_comments,
                    int quality_offset, int quality_method,
                    fsb_record* records, uint64_t capacity, fsh_parse_stats* stats);

/* The same with a switch for the byte-level contract check (symbols, quality range, 7-bit titles): a caller that hands the
 * table to the device path may leave that check to the library (FSB_OPT_VALIDATE, on by default) and pass validate_bytes = 0;
 * lengths are always checked.  Lines are found with memchr (LF files and CRLF files at memory speed; a line holding a lone
 * CR falls back to the byte-wise rule of SkipLine). */
int fsh_parse_chunk_ex(const uint8_t* text, uint64_t size, int keep_headers, int keep_comments,
                       int quality_offset, int quality_method, int validate_bytes,
                       fsb_record* records, uint64_t capacity, fsh_parse_stats* stats);

/* Upper bound of the number of records in `size` bytes of FASTQ text (for sizing the table). */
uint64_t fsh_max_records(const uint8_t* text, uint64_t size);

/* ---- chunk cutter ---------------------------------------------------------------------------- */

/*
 * Where to cut a full buffer of `size` bytes so that the chunk ends at a record boundary:
 * IFastqStreamReader::GetNextRecordPos / SkipToEol logic (FastqStream.cpp:15-40,73-85): start at
 * size - window, go to the next line start, advance to the next line starting with '@'; if the
 * following line also starts with '@' the record starts there.  Returns the offset of the first
 * byte of the record that begins the *next* chunk.
 */
uint64_t fsh_cut_position(const uint8_t* buf, uint64_t size, uint64_t window);

/* ---- chunk reader ---------------------------------------------------------------------------- */

/*
 * Reads FASTQ files chunk by chunk exactly like IFastqStreamReaderSE/PE::ReadNextChunk
 * (FastqStream.cpp:44-101, 104-228): a buffer of `block_size` bytes per file is filled (carried
 * tail first, then the files one after the other), cut at the first record start after
 * block_size - window (window = 8 KiB SE, 1 MiB PE; FastqStream.h:99,142), in PE mode both cuts are
 * re-synchronised on the numeric read id of the next title (FastqStream.cpp:231-256); the chunk
 * loses its final line end, the rest is carried over.  Chunk boundaries define block boundaries,
 * so they must be the reference's.
 */
typedef struct fsh_reader fsh_reader;
fsh_reader* fsh_reader_open(const char* const* files1, uint32_t n1, const char* const* files2, uint32_t n2, uint64_t block_size);
/* buf1 / buf2: block_size bytes each (buf2 unused for SE).  Returns 1 (chunk delivered), 0 (end of input), -1 (I/O error). */
int  fsh_reader_next(fsh_reader* r, uint8_t* buf1, uint64_t* size1, uint8_t* buf2, uint64_t* size2);
void fsh_reader_close(fsh_reader* r);

/* ---- bin-file writer -------------------------------------------------------------------------- */

/* The fields of BinModuleConfig (Params.h:167-193) that are not binning parameters of the device path. */
typedef struct fsh_bin_config {
    fsb_params params;
    uint32_t min_block_bin_size;   /* -m, CategorizerParameters::minBlockBinSize (8) */
    uint8_t  keep_comments;        /* !-C, HeadersCompressionParams::preserveComments */
    uint8_t  verbose;              /* -v (sets qvzOpts.verbose / stats in the dumped parameters, main.cpp:236) */
    uint8_t  reserved[2];
    uint64_t fastq_block_size;     /* -b in bytes */
} fsh_bin_config;

/*
 * Writes <prefix>.bmeta / .bdna / .bqua [/ .bhead] byte for byte like BinFileWriter
 * (StartCompress / WriteNextBlock / FinishCompress / WriteFileFooter, BinFile.cpp:47-462), except
 * that the never-initialised padding bytes of the 88-byte parameter dump are written as zeros.
 * QVZ (-q3) footers carry a codebook computed by fastore_pack code and are not supported here.
 */
typedef struct fsh_writer fsh_writer;
fsh_writer* fsh_writer_open(const char* prefix, const fsh_bin_config* cfg);
/* FastqRawBlockStats::Update for the titles of one parsed chunk (mate 1, and mate 2 in PE mode): Stats.cpp:90-169 */
int  fsh_writer_add_titles(fsh_writer* w, const uint8_t* text, const fsb_record* records, uint64_t n_records);
/* The same title statistics gathered per chunk away from the writer (e.g. in parser threads) and merged in chunk order, like
 * the per-chunk FastqRawBlockStats the reference merges in BinFileWriter::WriteNextBlock (Stats.cpp:205-236).  A part whose
 * records disagree on whether a field is numeric (fsh_titles_consistent == 0) must be added with fsh_writer_add_titles. */
typedef struct fsh_titles fsh_titles;
fsh_titles* fsh_titles_new(void);
void fsh_titles_free(fsh_titles* t);
int  fsh_titles_add(fsh_titles* t, const uint8_t* text, const fsb_record* records, uint64_t n_records);
int  fsh_titles_consistent(const fsh_titles* t);
int  fsh_writer_merge_titles(fsh_writer* w, const fsh_titles* t);
int  fsh_writer_add_block(fsh_writer* w, const fsb_block* block);          /* WriteNextBlock */
int  fsh_writer_close(fsh_writer* w);                                      /* FinishCompress; frees w */
const char* fsh_last_error(void);

/* ---- synthetic FASTQ --------------------------------------------------------------------------*/

typedef struct fsh_synth_config {
    uint64_t seed;
    uint64_t first_index;        /* index of the first record (records are a pure function of seed+index) */
    uint64_t n_records;          /* reads (SE) or pairs (PE) */
    uint32_t read_len;           /* fixed length, or the maximum when min_len != 0 */
    uint32_t min_len;            /* 0 = fixed length; else lengths uniform in [min_len, read_len] */
    uint32_t paired;             /* 0 SE, 1 PE */
    uint32_t genome_len;         /* synthetic genome size in bases (0 -> 100,000,000) */
    uint32_t sub_rate_ppm;       /* substitution errors per million bases   (10000 = 1 %)  */
    uint32_t n_rate_ppm;         /* isolated N per million bases            (5000 = 0.5 %) */
    uint32_t nrich_ppm;          /* fraction of reads with 10-60 % N                        */
    uint32_t lowcomplex_ppm;     /* fraction of low-complexity reads (homopolymer, dinucleotide, poly-A tail) */
    uint32_t alln_ppm;           /* fraction of all-N reads                                 */
    uint32_t tie_ppm;            /* fraction of directed tie cases (palindromes, identical mates, ...) */
    uint32_t header_comments;    /* 1: append " len=<L> synthetic" comments to the titles    */
    uint32_t crlf;               /* 1: CRLF line ends */
    uint32_t qual_mean_x10;      /* mean quality *10 (0 -> 360) */
    uint32_t qual_sd_x10;        /* sd *10 (0 -> 40) */
    uint32_t reserved;
} fsh_synth_config;

/* Bytes of FASTQ text the configuration produces for file 1 and (PE) file 2. */
int fsh_synth_size(const fsh_synth_config* cfg, uint64_t* bytes1, uint64_t* bytes2);

/* Fill the text buffers (sized by fsh_synth_size) using `threads` host threads; optionally also
 * write the record tables the parser would produce with keep_headers=1, keep_comments=1. */
int fsh_synth_fill(const fsh_synth_config* cfg, uint8_t* text1, uint8_t* text2,
                   fsb_record* records1, fsb_record* records2, int threads);

#ifdef __cplusplus
}
#endif
#endif
