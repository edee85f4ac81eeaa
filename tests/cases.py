"""Shared parity cases: (name, synthetic config kwargs, binning params kwargs).

They cover BASELINE.json's configs at sizes the CPU checkers finish in seconds:
  C1 SE 100 bp lossless, C2/C3 PE 150 bp lossless, C4 250 bp with -p12/-p10 -s10 and the
  reduced / max profiles, C5 N-rich + low-complexity + all-N + directed ties; plus variable
  lengths, CRLF, commented headers, Phred+64-style offsets and tiny / ragged inputs.
"""
from fastore_b200 import _native as N

STRESS = dict(nrich=0.10, lowcomplex=0.10, alln=0.01, tie=0.05)

CASES = [
    ("c1_se100_lossless", dict(n_records=20000, read_len=100, seed=101), dict(signature_len=8, skip_zone_len=0)),
    ("c2_pe150_lossless", dict(n_records=20000, read_len=150, paired=True, seed=102), dict(signature_len=8, skip_zone_len=0, paired_end=True)),
    ("c2_pe150_fast_s10", dict(n_records=8000, read_len=150, paired=True, seed=103), dict(signature_len=8, skip_zone_len=10, paired_end=True)),
    ("c4_se250_p12_s10_reduced", dict(n_records=6000, read_len=250, seed=104, header_comments=True),
     dict(signature_len=12, skip_zone_len=10, quality_method=N.FSB_QUA_8BIN)),
    ("c4_pe250_p12_s10_max", dict(n_records=4000, read_len=250, paired=True, seed=104),
     dict(signature_len=12, skip_zone_len=10, paired_end=True, quality_method=N.FSB_QUA_BINARY, reads_have_headers=False)),
    ("c4_pe250_p10_s10_lossy", dict(n_records=4000, read_len=250, paired=True, seed=1040),
     dict(signature_len=10, skip_zone_len=10, paired_end=True, quality_method=N.FSB_QUA_QVZ)),
    ("c5_pe150_stress", dict(n_records=12000, read_len=150, paired=True, seed=105, **STRESS), dict(signature_len=8, skip_zone_len=0, paired_end=True)),
    ("c5_se150_stress", dict(n_records=12000, read_len=150, seed=1050, **STRESS), dict(signature_len=8, skip_zone_len=0)),
    ("c5_se100_stress_s8", dict(n_records=8000, read_len=100, seed=1051, **STRESS), dict(signature_len=8, skip_zone_len=8)),
    ("se_varlen", dict(n_records=8000, read_len=151, min_len=30, seed=7, **STRESS), dict(signature_len=8, skip_zone_len=0)),
    ("pe_varlen", dict(n_records=6000, read_len=120, min_len=20, paired=True, seed=8, **STRESS), dict(signature_len=8, skip_zone_len=4, paired_end=True)),
    ("se_short_reads", dict(n_records=3000, read_len=24, min_len=1, seed=9, **STRESS), dict(signature_len=8, skip_zone_len=0)),
    ("pe_short_reads_p6", dict(n_records=3000, read_len=20, min_len=1, paired=True, seed=10, **STRESS), dict(signature_len=6, skip_zone_len=2, paired_end=True)),
    ("se_crlf_comments", dict(n_records=3000, read_len=100, seed=11, crlf=True, header_comments=True), dict(signature_len=8, skip_zone_len=0)),
    ("se_max_len_255", dict(n_records=2000, read_len=255, seed=12, **STRESS), dict(signature_len=13, skip_zone_len=0)),
    ("pe_max_len_255_p4", dict(n_records=2000, read_len=255, paired=True, seed=13, **STRESS), dict(signature_len=4, skip_zone_len=0, paired_end=True)),
    ("se_cutoff_bits", dict(n_records=4000, read_len=100, seed=14), dict(signature_len=8, skip_zone_len=0, cutoff_bits=3)),
    ("se_thr_w30_noheads", dict(n_records=4000, read_len=100, seed=15), dict(signature_len=8, skip_zone_len=0, quality_method=N.FSB_QUA_BINARY, binary_threshold=30, reads_have_headers=False)),
    ("se_single_record", dict(n_records=1, read_len=100, seed=16), dict(signature_len=8, skip_zone_len=0)),
    ("pe_three_records", dict(n_records=3, read_len=100, paired=True, seed=17), dict(signature_len=8, skip_zone_len=0, paired_end=True)),
    ("se_small_genome_dups", dict(n_records=8000, read_len=100, seed=18, genome_len=3000, sub_rate=0.0, n_rate=0.0), dict(signature_len=8, skip_zone_len=0)),
]


def make_case(name):
    from fastore_b200 import synth
    for n, skw, pkw in CASES:
        if n == name:
            cfg = synth.synth_config(**skw)
            params = N.make_params(**pkw)
            t1, t2, r1, r2 = synth.generate(cfg, threads=4)
            if not params.reads_have_headers:
                r1 = r1.copy(); r1["head_len"] = 0
                if r2 is not None:
                    r2 = r2.copy(); r2["head_len"] = 0
            chunk = N.make_chunk(t1, r1, t2, r2)
            return params, chunk, (t1, t2, r1, r2)
    raise KeyError(name)
