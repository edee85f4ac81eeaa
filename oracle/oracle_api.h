/*
 * oracle_api.h -- interface shared by the two CPU checkers under oracle/.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 *   orc_*  plain-C restatement of the reference algorithm     (oracle/fastore_oracle.c,  "port")
 *   ref_*  the reference's own objects behind the same call    (oracle/ref_harness.cpp -> oracle/_ref/)
 *
 * Both take exactly the input of the product's C ABI (include/fastore_b200.h: fsb_params,
 * fsb_chunk, fsb_record) and return one block with malloc'd buffers.
 */
#ifndef FASTORE_ORACLE_API_H
#define FASTORE_ORACLE_API_H

#include "../include/fastore_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_block {
    uint8_t* meta;
    uint8_t* dna;
    uint8_t* qua;
    uint8_t* head;
    uint64_t meta_size, dna_size, qua_size, head_size;
    uint64_t raw_dna_size, raw_head_size;
    fsb_bin_descriptor* bins;
    uint64_t n_bins;
    uint64_t n_records;
    uint32_t* read_signature;   /* per record, parse order */
    uint32_t* read_info;        /* minimPos | FSB_INFO_* flags */
} orc_block;

/* plain-C port */
int  orc_bin_chunk(const fsb_params* p, const fsb_chunk* chunk, orc_block* out);
void orc_block_free(orc_block* b);
/* per-strand scan alone: FindMinimizer (FastqCategorizer.cpp:79-106) */
void orc_find_minimizer(const fsb_params* p, const uint8_t* seq, uint32_t len, uint32_t* sig, uint32_t* pos);
int  orc_signature_valid(const fsb_params* p, uint32_t m);
/* wall seconds for `reps` passes of categorise+pack over the chunk split across `threads` */
double orc_time_bin_chunk(const fsb_params* p, const fsb_chunk* chunk, int threads, int reps);

/* the compiled reference behind the same interface (only in oracle/_ref/libfastore_ref.so) */
int  ref_bin_chunk(const fsb_params* p, const fsb_chunk* chunk, orc_block* out);
void ref_block_free(orc_block* b);
void ref_find_minimizer(const fsb_params* p, const uint8_t* seq, uint32_t len, uint32_t* sig, uint32_t* pos);

/* DnaRebalancer::FindNewMinimizer (fastore_rebin/DnaRebalancer.cpp:570-616) of one read: the port, and the reference's own
 * member function called through oracle/ref_rebin_harness.cpp (oracle/_ref/libfastore_ref_rebin.so) */
void orc_find_new_minimizer(const fsb_params* p, const uint8_t* seq, uint32_t len, uint32_t cur, uint32_t divisor, uint32_t* sig, uint32_t* pos, uint32_t* is_rev);
void refrebin_find_new_minimizer(const fsb_params* p, const uint8_t* seq, uint32_t len, uint32_t cur, uint32_t divisor, uint32_t* sig, uint32_t* pos, uint32_t* is_rev);
double ref_time_bin_chunk(const fsb_params* p, const fsb_chunk* chunk, int threads, int reps);

#ifdef __cplusplus
}
#endif
#endif
