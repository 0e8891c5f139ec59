"""Generates the golden fixtures under tests/golden/ by RUNNING THE REFERENCE'S OWN CODE
(oracle/_ref/libbsref.so = unmodified /root/reference sources, see oracle/Makefile and oracle/ref_harness.c).

Run in the authoring container (where /root/reference exists):  python tests/golden/make_golden.py
The GPU boxes have no reference tree; they check against these committed files.

  sites_v1.npz   per-site model goldens: pileup[] + ref -> gt_meth[] + skip[]   (call_thread body through
                 call_genotypes_ML on blocks built from synthetic reads, plus direct calc_gt_prob/fisher KATs)
  block_*.npz    block goldens: raw templates -> normalised templates, pileup[], gt_vcf[]
                 (process_template_vector -> call_genotypes_ML)
  profile_v1.npz --report-file side channels of the block and reader goldens (meth_profile, base / read tallies)
  writer_v1.npz  the BCF records of the reference's writer for the gt_vcf[] of the block goldens
  reader_*.npz   reader goldens: raw BAM records -> per-record descriptors (get_next_align_details), blocks and
                 templates (read_input), gt_vcf[] of every block (the whole chain)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.bindings import Reference  # noqa: E402
from tests import blockgen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

BLOCK_CASES = {
    "block_pe_plain": dict(seed=11, reflen=5000, start=150, end=4200, trims=((0, 0), (0, 0)),
                           kw=dict(depth=30, read_len=100, paired=True)),
    "block_pe_indel_clip_trim": dict(seed=12, reflen=5000, start=150, end=4200, trims=((5, 0), (3, 2)),
                                     kw=dict(depth=25, read_len=100, paired=True, indel_frac=0.3, clip_frac=0.3, frag_mean=150)),
    "block_se_deep": dict(seed=13, reflen=1200, start=100, end=700, trims=((0, 0), (0, 0)),
                          kw=dict(depth=500, read_len=150, paired=False, snp_rate=0.02)),
    "block_mixed": dict(seed=14, reflen=4000, start=50, end=3300, trims=((2, 2), (2, 2)),
                        kw=dict(depth=15, read_len=75, paired=True, single_mate_frac=0.3, indel_frac=0.2,
                                clip_frac=0.1, nonconv_frac=0.2, n_frac=0.03, frag_mean=120, frag_sd=40)),
}


def make_blocks():
    for name, c in BLOCK_CASES.items():
        rng = np.random.default_rng(c["seed"])
        ref = blockgen.random_reference(rng, c["reflen"], n_runs=2)
        T, B, M, y = blockgen.make_block(rng, ref, c["start"], c["end"], **c["kw"])
        lt, rt = c["trims"]
        r = Reference(left_trim=lt, right_trim=rt)
        x, pile, vcf, refw, nt, nb = r.process_block(T, B, M, ref, y)
        Reference()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), templates=T, bases=B, misms=M, y=np.uint32(y), x=np.uint32(x),
                            left_trim=np.array(lt, dtype=np.uint32), right_trim=np.array(rt, dtype=np.uint32),
                            ref=refw, pileup=pile, vcf=vcf, norm_templates=nt, norm_bases=nb)
        print(name, "sites", len(vcf), "called", int((vcf["skip"] == 0).sum()), "templates", len(T))


def make_sites():
    r = Reference()
    piles, refs, outs = [], [], []
    # (1) site records harvested from reference-built pileups of random blocks: realistic count structure
    for seed, kw in ((21, dict(depth=30, read_len=100, paired=True)),
                     (22, dict(depth=8, read_len=100, paired=True, snp_rate=0.01)),
                     (23, dict(depth=400, read_len=150, paired=False, snp_rate=0.03)),
                     (24, dict(depth=60, read_len=100, paired=True, nonconv_frac=0.5, snp_rate=0.02))):
        rng = np.random.default_rng(seed)
        ref = blockgen.random_reference(rng, 3000, n_runs=1)
        T, B, M, y = blockgen.make_block(rng, ref, 100, 2300 if kw["depth"] < 100 else 500, **kw)
        x, pile, vcf, refw, nt, nb = r.process_block(T, B, M, ref, y)
        piles.append(pile)
        refs.append(refw)
        outs.append(vcf)
    pile = np.concatenate(piles)
    ref = np.concatenate(refs)
    vcf = np.concatenate(outs)
    # (2) direct KATs for calc_gt_prob / fisher on adversarial count vectors
    rng = np.random.default_rng(99)
    n = 3000
    counts = np.zeros((n, 8), dtype=np.uint64)
    for i in range(n):
        k = int(rng.integers(0, 7))
        for c in (rng.choice(8, size=k, replace=False) if k else []):
            counts[i, c] = int(rng.integers(1, 80)) if rng.random() < 0.9 else int(rng.integers(80, 20000))
    qual = np.where(counts > 0, rng.integers(1, 44, size=(n, 8)), 0).astype(np.int32)
    rf = rng.integers(0, 5, size=n).astype(np.uint8)
    kat = r.calc_gt_prob_batch(counts, qual, rf)
    tabs = np.concatenate([rng.integers(0, 40, size=(1500, 4)), rng.integers(0, 600, size=(500, 4)),
                           np.array([[0, 0, 0, 0], [5, 0, 0, 5], [0, 7, 7, 0], [1, 0, 0, 0], [300, 2, 1, 280], [12, 3, 4, 11]])]).astype(np.int32)
    fis = np.array([r.fisher(t) for t in tabs])
    np.savez_compressed(os.path.join(HERE, "sites_v1.npz"), pileup=pile, ref=ref, gt_meth=vcf["gtm"], skip=vcf["skip"],
                        kat_counts=counts, kat_qual=qual, kat_rf=rf, kat_out=kat, fisher_tabs=tabs, fisher_p=fis)
    print("sites", len(pile), "called", int((vcf["skip"] == 0).sum()), "het",
          int(np.isin(vcf["gtm"]["max_gt"], [1, 2, 3, 5, 6, 8])[vcf["skip"] == 0].sum()))


READER_CASES = {
    # paired-end, default options; mixed: singles + pairs, keep_unmatched, its own thresholds
    "reader_pe": dict(seed=31, stream=dict(n_contigs=2, paired=True, depth=14, read_len=70, contig_len=4000),
                      opts=dict(mapq_thresh=20, max_template_len=1000, keep_unmatched=False, ignore_duplicates=False, keep_duplicates=False)),
    "reader_mixed": dict(seed=32, stream=dict(n_contigs=3, dup=0.25, contig_len=3500),
                         opts=dict(mapq_thresh=10, max_template_len=600, keep_unmatched=True, ignore_duplicates=True, keep_duplicates=False)),
}


def make_reader():
    """reader goldens: raw BAM records -> get_next_align_details() per record, read_input() blocks and templates, and the
    gt_vcf[] of every block from the reference's chain read_input -> process_template_vector -> call_genotypes_ML"""
    from tests import bamgen
    r = Reference()
    for name, c in READER_CASES.items():
        bam, n, tl, refs = bamgen.make_stream(c["seed"], **c["stream"])
        o = c["opts"]
        rec, rbases, rmisms = r.decode_records(bam, o["mapq_thresh"], o["max_template_len"], o["keep_unmatched"], o["ignore_duplicates"])
        blocks, tm, bases, misms, vcf = r.read_input(bam, tl, refs, run_chain=True, **o)
        extra = {"ref%d" % i: refs[i] for i in range(len(refs))}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), bam=bam, target_len=tl, rec=rec, rec_bases=rbases, rec_misms=rmisms,
                            blocks=blocks, templates=tm, bases=bases, misms=misms, vcf=vcf,
                            **{k: np.array(v) for k, v in o.items()}, **extra)
        print(name, "records", n, "kept", int((rec["ret"] == 0).sum()), "blocks", len(blocks), "templates", len(tm),
              "sites", len(vcf), "called", int((vcf["skip"] == 0).sum()))


def make_profile():
    """--report-file side channels (non-CpG conversion profile, base / read tallies) of the block and reader goldens,
    from the reference with a live bs_stats; each case starts from a fresh profile"""
    from tests import bamgen
    out = {}

    def keep(name, p):
        for k in ("used", "conv", "base_filter", "filter_cts", "filter_bases"):
            out[name + "__" + k] = np.asarray(p[k])

    for name, c in BLOCK_CASES.items():
        rng = np.random.default_rng(c["seed"])
        ref = blockgen.random_reference(rng, c["reflen"], n_runs=2)
        T, B, M, y = blockgen.make_block(rng, ref, c["start"], c["end"], **c["kw"])
        lt, rt = c["trims"]
        r = Reference(left_trim=lt, right_trim=rt)
        r.stats_enable(True); r.stats_reset()
        x = r.process_block(T, B, M, ref, y)[0]
        keep(name, r.stats_read())
        out[name + "__ref"] = blockgen.window_codes(ref, x, y + 1)          # the profile looks one code past the block
        r.stats_enable(False)
        Reference()
        print(name, "profile used", int(out[name + "__used"]), "counts", int(out[name + "__conv"].sum()))
    r = Reference()
    for name, c in READER_CASES.items():
        bam, n, tl, refs = bamgen.make_stream(c["seed"], **c["stream"])
        r.stats_enable(True); r.stats_reset()
        r.read_input(bam, tl, refs, run_chain=True, **c["opts"])
        keep(name, r.stats_read())
        r.stats_enable(False)
        print(name, "profile used", int(out[name + "__used"]), "counts", int(out[name + "__conv"].sum()), "filter_cts", out[name + "__filter_cts"])
    np.savez_compressed(os.path.join(HERE, "profile_v1.npz"), **out)


def make_writer():
    """writer goldens: the BCF records the reference's own print_vcf_entry / flush_vcf_entries / _print_vcf_entry
    (src/print_vcf.c, compiled unmodified; bcf_write captured by oracle/ref_harness.c) emit for the gt_vcf[] of the block
    goldens, without and with -A (all positions)"""
    out = {}
    r = Reference()
    for name, c in BLOCK_CASES.items():
        g = np.load(os.path.join(HERE, name + ".npz"))
        rng = np.random.default_rng(c["seed"])
        ref = blockgen.random_reference(rng, c["reflen"], n_runs=2)
        x, sz = int(g["x"]), len(g["vcf"])
        refw = blockgen.window_codes(ref, x, x + sz + 1)
        assert (refw[:sz] == g["ref"]).all()
        out[name + "__ref"] = refw
        for allp in (0, 1):
            b, n = r.print_block(g["vcf"], refw, x, rid=2, all_positions=bool(allp))
            out["%s__all%d" % (name, allp)] = b
            out["%s__n%d" % (name, allp)] = np.uint32(n)
            print(name, "all_positions", allp, "records", n, "bytes", len(b))
    # the reader goldens: every block of the reference's chain through the reference's writer, concatenated in stream order
    from tests import bamgen
    for name, c in READER_CASES.items():
        g = np.load(os.path.join(HERE, name + ".npz"))
        _, _, tl, refs = bamgen.make_stream(c["seed"], **c["stream"])
        parts, total = [], 0
        for b in g["blocks"]:
            x, y, tid = int(b["x"]), int(b["y"]), int(b["tid"])
            sz = y - x + 1
            v = g["vcf"][int(b["vcf_off"]):int(b["vcf_off"]) + sz]
            rb, n = r.print_block(v, blockgen.window_codes(refs[tid], x, y + 2), x, rid=tid, ctg_end=int(tl[tid]))
            parts.append(rb); total += n
        out[name + "__bcf"] = np.concatenate(parts)
        out[name + "__nrec"] = np.uint32(total)
        print(name, "records", total, "bytes", len(out[name + "__bcf"]))
    np.savez_compressed(os.path.join(HERE, "writer_v1.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["blocks", "sites", "reader", "profile", "writer"]
    if "writer" in which:
        make_writer()
    if "profile" in which:
        make_profile()
    if "blocks" in which:
        make_blocks()
    if "sites" in which:
        make_sites()
    if "reader" in which:
        make_reader()
