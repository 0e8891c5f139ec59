#!/usr/bin/env python
"""Randomised stress of the reader path on a GPU box (not collected by pytest): many seeded record streams with random
reader options and random host-pipeline settings (chunked upload, builder pieces, framer threads) through bsgpu_call_bam,
block by block against the oracle chain.   usage: python tests/fuzz_reader.py [first_seed] [n_seeds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from bs_call_b200 import lib as bslib  # noqa: E402
from oracle.bindings import Oracle  # noqa: E402
from tests import bamgen, blockgen, util  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
gpu, oracle = bslib.BsGpu(), Oracle()
sites = refused = records = 0
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    bam, n, tl, refs = bamgen.make_stream(seed, dup=float(rng.choice([0.0, 0.1, 0.3])), junk=float(rng.choice([0.0, 0.1, 0.3])))
    o = dict(mapq_thresh=int(rng.integers(0, 40)), max_template_len=int(rng.integers(200, 1500)), keep_unmatched=bool(rng.random() < 0.3),
             ignore_duplicates=bool(rng.random() < 0.3), keep_duplicates=bool(rng.random() < 0.3))
    for k, v in (("BSGPU_READER_CHUNK_MIN_BYTES", rng.choice(["1", "1000000000"])), ("BSGPU_BUILDER_THREADS", str(int(rng.integers(1, 9)))),
                 ("BSGPU_BUILDER_MIN_RECORDS", rng.choice(["1", "1000000000"])), ("BSGPU_FRAMER_THREADS", str(int(rng.integers(1, 9)))),
                 ("BSGPU_FRAMER_MIN_BYTES", rng.choice(["1", "1000000000"]))):
        os.environ[k] = str(v)
    try:
        wbk, wt, wb, wm, wv = oracle.read_input(bam, tl, refs, run_chain=True, **o)
    except RuntimeError:
        # a stream on which the reference aborts (duplicate waiting read name, mates that disagree, a mate before its
        # block window): the product must refuse it as well
        try:
            gpu.call_bam(bam, tl, refs, bslib.reader_params(**o))
        except bslib.BsGpuError:
            refused += 1
            continue
        raise AssertionError("seed %d: the oracle refuses the stream, the product does not" % seed)
    blocks, vcf = gpu.call_bam(bam, tl, refs, bslib.reader_params(**o))
    assert len(blocks) == len(wbk), (seed, len(blocks), len(wbk))
    for b, w in zip(blocks, wbk):
        assert (b["tid"], b["x"], b["y"], b["n_templates"], b["first_template"]) == (w["tid"], w["x"], w["y"], w["n_templates"], w["first_template"]), seed
        sz = int(w["y"]) - int(w["x"]) + 1
        sites += util.assert_vcf_close(vcf[int(b["vcf_off"]):int(b["vcf_off"]) + sz], wv[int(w["vcf_off"]):int(w["vcf_off"]) + sz])
    # the same stream on to BCF records, against the oracle's writer over the device's own gt_vcf[] block by block, with the
    # --report-file side channels switched on every other seed (the records must not change)
    allp = bool(rng.random() < 0.3)
    parts, total = [], 0
    for b in blocks:
        x, y, tid = int(b["x"]), int(b["y"]), int(b["tid"])
        rb, k = oracle.print_block(vcf[int(b["vcf_off"]):int(b["vcf_off"]) + y - x + 1], blockgen.window_codes(refs[tid], x, y + 2), x, rid=tid,
                                   ctg_end=int(tl[tid]), all_positions=allp)
        parts.append(rb)
        total += k
    want = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
    gpu.profile_enable(seed % 2 == 0)
    _, out, nrec = gpu.call_bam_bcf(bam, tl, refs, bslib.reader_params(**o), bslib.bcf_params(all_positions=allp))
    gpu.profile_enable(False)
    assert nrec == total and np.asarray(out).tobytes() == want.tobytes(), "seed %d: BCF records differ" % seed
    records += nrec
print("fuzz ok: seeds %d..%d, %d called sites compared, %d BCF records byte-identical, %d streams refused by both sides"
      % (first, first + count - 1, sites, records, refused))
