#!/usr/bin/env python
"""Randomised CPU stress of raw-template blocks (not collected by pytest; needs oracle/_ref): random depth / read length / fragment size,
heavy indel and soft-clip rates, lone mates, N bases, random -L / -R trimming through the compiled process_template_vector +
call_genotypes_ML and the restatement: normalised bytes, pileup[] and every gt_vcf field bit-identical, or both refuse.
   usage: python tests/fuzz_cpu_blocks.py [first_seed] [n_seeds]"""
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.bindings import Oracle, Reference
from tests import blockgen
import tests.test_oracle_vs_reference as T
first, count = (int(sys.argv[1]) if len(sys.argv) > 1 else 100), (int(sys.argv[2]) if len(sys.argv) > 2 else 50)
ok = bad = refused = 0
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    case = dict(depth=int(rng.integers(3, 80)), read_len=int(rng.integers(25, 151)), paired=bool(rng.random() < 0.7),
                frag_mean=int(rng.integers(40, 400)), frag_sd=int(rng.integers(5, 80)), indel_frac=float(rng.choice([0.0, 0.2, 0.6])),
                clip_frac=float(rng.choice([0.0, 0.2, 0.6])), single_mate_frac=float(rng.choice([0.0, 0.3])), nonconv_frac=float(rng.choice([0.0, 0.3])),
                n_frac=float(rng.choice([0.0, 0.05])))
    lt, rt = (int(rng.integers(0, 12)), int(rng.integers(0, 12))), (int(rng.integers(0, 12)), int(rng.integers(0, 12)))
    ref = blockgen.random_reference(rng, 6000, n_runs=2)
    try:
        Tm, B, M, y = blockgen.make_block(rng, ref, 200, 3000, **case)
    except Exception as e:
        continue
    o, r = Oracle(left_trim=lt, right_trim=rt), Reference(left_trim=lt, right_trim=rt)
    try:
        x, pile_r, vcf_r, ref_r, nt_r, nb_r = r.process_block(Tm, B, M, ref, y)
    except Exception as e:
        # the reference refused (illegal CIGAR after trimming ...): the oracle must refuse too
        try:
            o.process_block(Tm, B, M, blockgen.window_codes(ref, max(int(Tm[0]["forward_position"] or Tm[0]["reverse_position"]) - 2, 1), y + 2), y)
            print("seed", seed, "reference refuses, oracle accepts:", str(e)[:80]); bad += 1
        except Exception:
            refused += 1
        continue
    nt_o, nb_o = o.normalise_block(Tm, B, M)
    xo, pile_o, vcf_o = o.process_block(Tm, B, M, ref_r, y)
    same = nb_o.tobytes() == nb_r.tobytes() and xo == x and pile_o.tobytes() == pile_r.tobytes()
    for f in ("counts", "qual", "gt_prob", "fisher_strand", "mq", "aq", "max_gt"):
        same = same and vcf_o["gtm"][f].tobytes() == vcf_r["gtm"][f].tobytes()
    same = same and vcf_o["skip"].tobytes() == vcf_r["skip"].tobytes()
    if same: ok += 1
    else: bad += 1; print("seed", seed, "MISMATCH", case, lt, rt)
Reference()
print("block fuzz seeds %d..%d: %d identical, %d refused by both, %d problems" % (first, first + count - 1, ok, refused, bad))
