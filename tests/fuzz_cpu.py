#!/usr/bin/env python
"""Randomised CPU stress (not collected by pytest; needs oracle/_ref): seeded record streams with random reader options through
(1) the compiled reference chain and the oracle restatement -- blocks, templates and every gt_vcf field bit-identical, or both
refuse the stream --, (2) the product's host block builder over the oracle's descriptors, and (3) every 4th seed with one contig the whole
reference program from files (random -q -l -k -e -d -A, -D with a synthetic index) against the harness chain.   usage: python tests/fuzz_cpu.py [first_seed] [n_seeds]"""
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from bs_call_b200 import hostio, lib as bslib  # noqa: E402
from oracle.bindings import Oracle, Reference, bcf_diff  # noqa: E402
from tests import bamgen, util  # noqa: E402
from tests.test_full_binary import BIN, chain_records, run_binary, synthetic_dbsnp, write_case  # noqa: E402
from tests.test_cpu_reader import descriptors_from_oracle  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 60
sites = refused = programs = 0
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    bam, n, tl, refs = bamgen.make_stream(seed, dup=float(rng.choice([0.0, 0.1, 0.3])), junk=float(rng.choice([0.0, 0.1, 0.3])))
    o = dict(mapq_thresh=int(rng.integers(0, 40)), max_template_len=int(rng.integers(200, 1500)), keep_unmatched=bool(rng.random() < 0.3),
             ignore_duplicates=bool(rng.random() < 0.3), keep_duplicates=bool(rng.random() < 0.3))
    # model / normalisation parameters: the defaults on even seeds (what the whole-program runs and the child process use), random ones on odd seeds
    mp = dict() if seed % 2 == 0 else dict(under_conv=float(rng.choice([0.0, 0.01, 0.05])), over_conv=float(rng.choice([0.0, 0.05, 0.1])),
                                           ref_bias=float(rng.choice([1.0, 2.0, 5.0])), min_qual=int(rng.integers(5, 35)),
                                           left_trim=(int(rng.integers(0, 8)), int(rng.integers(0, 8))), right_trim=(int(rng.integers(0, 8)), int(rng.integers(0, 8))))
    oracle, ref = Oracle(**mp), Reference(calc_threads=2, **mp)
    # --report-file's counters live on both sides (read_input's tallies, the conversion profile, the normalisation tallies)
    ref.stats_enable(True); ref.stats_reset()
    oracle.profile_enable(True); oracle.profile_reset()
    try:
        wbk, wt, wb, wm, wv = oracle.read_input(bam, tl, refs, run_chain=True, **o)
    except RuntimeError:
        # the restatement refuses the stream: the compiled reference must die on it (fatal error or failed assertion) -- in a child
        code = ("import sys, numpy as np; sys.path.insert(0, %r); from oracle.bindings import Reference; from tests import bamgen; "
                "rng = np.random.default_rng(%d); "
                "bam, n, tl, refs = bamgen.make_stream(%d, dup=float(rng.choice([0.0, 0.1, 0.3])), junk=float(rng.choice([0.0, 0.1, 0.3]))); "
                "Reference(calc_threads=2).read_input(bam, tl, refs, run_chain=True, **%r); print('survived')"
                % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), seed, seed, o))
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        assert r.returncode != 0 and "survived" not in r.stdout, "seed %d: the oracle refuses the stream, the reference does not" % seed
        refused += 1
        ref.stats_enable(False); oracle.profile_enable(False)
        continue
    rbk, rt, rb, rm, rv = ref.read_input(bam, tl, refs, run_chain=True, **o)
    util.same_profile(oracle.profile_read(), ref.stats_read(), "seed %d" % seed, recycled_vectors=True)
    ref.stats_enable(False); oracle.profile_enable(False)
    assert len(rbk) == len(wbk) and all((rbk[f] == wbk[f]).all() for f in ("tid", "x", "y", "first_template", "n_templates", "vcf_off")), seed
    assert bamgen.template_keys(rt, rb, rm) == bamgen.template_keys(wt, wb, wm), seed
    util.assert_gt_meth_close(wv["gtm"], wv["skip"], rv["gtm"], rv["skip"], exact_doubles=True)
    sites += int((wv["skip"] == 0).sum())
    orec, ob, om = oracle.decode_records(bam, o["mapq_thresh"], o["max_template_len"], o["keep_unmatched"], o["ignore_duplicates"])
    hb, ht = bslib.build_blocks(bam, descriptors_from_oracle(orec, ob, bam), bslib.reader_params(keep_unmatched=o["keep_unmatched"], keep_duplicates=o["keep_duplicates"]))
    assert len(hb) == len(wbk) and all((hb[f] == wbk[f]).all() for f in ("tid", "x", "y", "n_templates", "first_template")), seed
    assert bamgen.template_keys(ht, ob, om) == bamgen.template_keys(wt, wb, wm), seed
    if seed % 4 == 0 and len(tl) == 1 and os.path.exists(BIN):
        with tempfile.TemporaryDirectory() as tmp:
            names, fa, bf = write_case(tmp, bam, tl, refs)
            extra = ["-q", str(o["mapq_thresh"]), "-l", str(o["max_template_len"])] + (["-k"] if o["keep_unmatched"] else []) + \
                    (["-e"] if o["ignore_duplicates"] else []) + (["-d"] if o["keep_duplicates"] else [])
            allp, db = bool(rng.random() < 0.3), None
            if allp:
                extra.append("-A")
            if rng.random() < 0.5:                       # -D: an index file in the reference's format through its own reader
                files, db = synthetic_dbsnp(rng, tl, frac=0.03)
                hostio.write_dbsnp_index(os.path.join(tmp, "db.idx"), files, prefixes=("rs", "ss"), bins_per_block=int(rng.integers(1, 40)))
                extra += ["-D", os.path.join(tmp, "db.idx")]
            out = os.path.join(tmp, "o.bcf")
            run_binary(BIN, fa, bf, out, extra=tuple(extra))
            d = bcf_diff(hostio.read_bcf(out)[1], chain_records(ref, bam, tl, refs, all_positions=allp, dbsnp=db, **o))
            assert d["records_a"] == d["records_b"] == d["identical"], (seed, d)
            programs += 1
print("cpu fuzz ok: seeds %d..%d, %d called sites bit-identical between the reference chain and the restatement, host builder = oracle blocks, "
      "%d whole-program runs = harness chain, %d streams refused by both sides" % (first, first + count - 1, sites, programs, refused))
