"""The whole program on the GPU (-m gpu): bs_call built from the reference's own main(), option parsing, header / FASTA /
BAM input and VCF / BCF output (over oracle/minihts), with the product's seam files in place of the reference's hot-path
files, run from the command line on the same files as the all-CPU binary (tests/test_full_binary.py):

  oracle/_ref/bs_call_gpu         bsgpu_seam_reader.c for get_template_vector.c + process_template.c + call_genotypes.c
                                  (seam C: gt_vcf[] to the reference's print thread; BSGPU_SEAM_RECORDS=1, seam D: BCF records
                                  built on the device and handed to bcf_write)
  oracle/_ref/bs_call_gpu_narrow  bsgpu_dropin.c for call_genotypes.c only (seam A)

The BCF files must hold the records of the CPU binary's file.  Records may differ only where the device's transcendental
functions put a likelihood within rounding of a tie or an integer boundary (the guard bands, tests/test_gpu_guard.py):
counted, bounded, printed.
"""
import os
import subprocess

import numpy as np
import pytest

from bs_call_b200 import hostio
from tests import bamgen
from tests.test_full_binary import BIN, write_case, run_binary, chain_records, contig_tail_race, _keyed

pytestmark = pytest.mark.gpu

REF_DIR = os.path.dirname(BIN)
VARIANTS = {
    "seamC": ("bs_call_gpu", {}),
    "seamD": ("bs_call_gpu", {"BSGPU_SEAM_RECORDS": "1"}),
    "narrow": ("bs_call_gpu_narrow", {}),
    # the record loop (sam_read1 per record) instead of the bulk input (bgzf_read straight into the session's stage), which is
    # what seams C / D use when every contig of the header is wanted
    "seamC_loop": ("bs_call_gpu", {"BSGPU_SEAM_BULK": "0"}),
    "seamD_loop": ("bs_call_gpu", {"BSGPU_SEAM_RECORDS": "1", "BSGPU_SEAM_BULK": "0"}),
}


def gpu_binary(variant):
    name, env = VARIANTS[variant]
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path) or not os.path.exists(BIN):
        pytest.skip("oracle/_ref/%s not built (reference tree absent at build time)" % name)
    e = dict(os.environ)
    e.update(env)
    return path, e


def compare(got, want, what, tol=0.002):
    g, w = _keyed(got), _keyed(want)
    only = set(g) ^ set(w)
    differ = [k for k in g if k in w and g[k] != w[k]]
    n = max(len(w), 1)
    print("%s: %d records, %d only on one side, %d differ" % (what, len(w), len(only), len(differ)))
    assert len(w) > 100
    assert len(only) + len(differ) <= max(3, int(tol * n)), (what, sorted(only)[:5], differ[:5])
    return len(only), len(differ)


@pytest.mark.parametrize("variant", ["seamC", "seamD", "narrow", "seamD_loop"])
@pytest.mark.parametrize("seed,extra", [(3, ()), (11, ("-k",)), (12, ("-d",)), (13, ("-q", "5", "-l", "400", "-L", "3", "-R", "2,4"))])
def test_gpu_binary_writes_the_cpu_binarys_records(tmp_path, variant, seed, extra):
    path, env = gpu_binary(variant)
    bam, n, tl, refs = bamgen.make_stream(seed, n_contigs=1, contig_len=30000)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    cpu, gpu = os.path.join(str(tmp_path), "cpu.bcf"), os.path.join(str(tmp_path), "gpu.bcf")
    run_binary(BIN, fa, bf, cpu, extra=extra)
    run_binary(path, fa, bf, gpu, extra=extra, env=env)
    tc, rc = hostio.read_bcf(cpu)
    tg, rg = hostio.read_bcf(gpu)
    assert tc == tg                                  # the header is the reference's own code in both
    compare(rg, rc, "%s seed %d %s" % (variant, seed, " ".join(extra)))


@pytest.mark.parametrize("variant", ["seamC", "seamD", "seamC_loop", "seamD_loop"])
def test_gpu_binary_on_several_contigs(reference, tmp_path, variant):
    """three contigs: the wide seams keep every contig's end (the reader does not free a contig under the print thread), so the
    file holds what the harness-driven chain produces -- the CPU binary itself loses records of contig ends to its race
    (tests/test_full_binary.py::contig_tail_race)"""
    path, env = gpu_binary(variant)
    bam, n, tl, refs = bamgen.make_stream(3, n_contigs=3)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    gpu = os.path.join(str(tmp_path), "gpu.bcf")
    run_binary(path, fa, bf, gpu, env=env)
    _, rg = hostio.read_bcf(gpu)
    want = chain_records(reference, bam, tl, refs)
    compare(rg, want, variant + " three contigs")


def test_gpu_binary_all_positions_vcf_text(tmp_path):
    """-A -O v: the text lines of the GPU binary are the CPU binary's (the same formatter over the same records)"""
    path, env = gpu_binary("seamD")
    bam, n, tl, refs = bamgen.make_stream(5, n_contigs=1, contig_len=20000)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    cpu, gpu = os.path.join(str(tmp_path), "cpu.vcf"), os.path.join(str(tmp_path), "gpu.vcf")
    run_binary(BIN, fa, bf, cpu, otype="v", extra=("-A",))
    run_binary(path, fa, bf, gpu, otype="v", extra=("-A",), env=env)
    lc = open(cpu).read().split("\n")
    lg = open(gpu).read().split("\n")
    assert len(lc) == len(lg) and len(lc) > 1000
    bad = [i for i, (a, b) in enumerate(zip(lc, lg)) if a != b]
    print("VCF text: %d lines, %d differ" % (len(lc), len(bad)))
    assert len(bad) <= max(3, len(lc) // 500), [(lc[i], lg[i]) for i in bad[:3]]


@pytest.mark.parametrize("variant", ["seamC", "seamD", "seamD_loop"])
def test_gpu_binary_report_file(tmp_path, variant):
    """--report-file: the JSON statistics of the GPU binary against the CPU binary's.  Seam C: the reference's writer over the
    device's gt_vcf[], the read-level tallies and the conversion profile from the device.  Seam D: the site statistics of
    src/print_vcf.c:382-526 from the device too (k_bcf_stats), folded into bs_stats at join_calc_threads"""
    import json
    path, env = gpu_binary(variant)
    bam, n, tl, refs = bamgen.make_stream(21, n_contigs=1, contig_len=30000, dup=0.1, junk=0.05)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    out = {}
    for tag, b, e in (("cpu", BIN, None), ("gpu", path, env)):
        js = os.path.join(str(tmp_path), tag + ".json")
        run_binary(b, fa, bf, os.path.join(str(tmp_path), tag + ".bcf"), extra=("--report-file", js), env=e)
        out[tag] = json.load(open(js))
    c, g = out["cpu"], out["gpu"]

    def leaves(o, pre=""):
        if isinstance(o, dict):
            for k, v in o.items():
                yield from leaves(v, pre + "/" + str(k))
        elif isinstance(o, list):
            for i, v in enumerate(o):
                yield from leaves(v, pre + "/%d" % i)
        else:
            yield pre, o

    lc, lg = dict(leaves(c)), dict(leaves(g))
    assert set(lc) == set(lg)

    def close(a, b):
        if isinstance(a, (int, float)) and isinstance(b, (int, float)):
            return abs(a - b) <= 1e-6 * max(1.0, abs(a), abs(b))
        return a == b

    bad = [k for k in lc if not close(lc[k], lg[k])]
    print("report: %d values, %d differ: %s" % (len(lc), len(bad), [(k, lc[k], lg[k]) for k in bad[:8]]))
    # the base-level tallies come from the device's normalisation pass, the site statistics from the reference's writer over
    # the device's records; read-level "Passed" is not a function of the input in the reference
    for k in lc:
        if k.startswith("/filterStats/BaseLevel"):
            assert lc[k] == lg[k], (k, lc[k], lg[k])
    assert abs(c["totalStats"]["SNPS"]["All"] - g["totalStats"]["SNPS"]["All"]) <= 3
    hard = [k for k in bad if "ReadLevel/Passed" not in k and not k.startswith("/date")]
    assert len(hard) <= max(10, len(lc) // 50), hard[:10]


@pytest.mark.parametrize("variant", ["seamC", "seamD", "seamD_loop"])
def test_gpu_binary_with_dbsnp_index(tmp_path, variant):
    """-D: the reference's index reader over a synthetic index file (tests/test_full_binary.py::test_binary_with_dbsnp_index).
    Seam C keeps the reference's writer, which looks every site up; on seam D the reader loads the contig's entries and hands
    them to the device writer (bsgpu_set_contig_annotation): ids and always-written sites as in the CPU binary's file"""
    from tests.test_full_binary import synthetic_dbsnp
    path, env = gpu_binary(variant)
    rng = np.random.default_rng(77)
    bam, n, tl, refs = bamgen.make_stream(31, n_contigs=1, contig_len=30000)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    files, arrays = synthetic_dbsnp(rng, tl)
    idx = os.path.join(str(tmp_path), "db.idx")
    hostio.write_dbsnp_index(idx, files, prefixes=("rs", "ss"), bins_per_block=50)
    cpu, gpu = os.path.join(str(tmp_path), "cpu.bcf"), os.path.join(str(tmp_path), "gpu.bcf")
    run_binary(BIN, fa, bf, cpu, extra=("-D", idx))
    run_binary(path, fa, bf, gpu, extra=("-D", idx), env=env)
    plain = os.path.join(str(tmp_path), "plain.bcf")
    run_binary(BIN, fa, bf, plain)
    _, rc = hostio.read_bcf(cpu)
    _, rg = hostio.read_bcf(gpu)
    _, rp = hostio.read_bcf(plain)
    assert len(_keyed(rc)) > len(_keyed(rp))         # the index added always-written sites
    compare(rg, rc, variant + " with -D")


@pytest.mark.parametrize("variant", ["seamD", "seamD_loop"])
def test_gpu_binary_several_contigs_with_dbsnp(reference, tmp_path, variant):
    """three contigs with -D on seam D: every contig's entries reach the device writer when the session gets to the contig (bulk
    input: from the session's worker through bsgpu_bam_on_contig), against the harness chain with the same table"""
    from tests.test_full_binary import synthetic_dbsnp
    path, env = gpu_binary(variant)
    rng = np.random.default_rng(5)
    bam, n, tl, refs = bamgen.make_stream(3, n_contigs=3)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    files, arrays = synthetic_dbsnp(rng, tl, frac=0.03)
    idx = os.path.join(str(tmp_path), "db.idx")
    hostio.write_dbsnp_index(idx, files, prefixes=("rs", "ss"), bins_per_block=20)
    gpu = os.path.join(str(tmp_path), "gpu.bcf")
    run_binary(path, fa, bf, gpu, extra=("-D", idx), env=env)
    _, rg = hostio.read_bcf(gpu)
    want = chain_records(reference, bam, tl, refs, dbsnp=arrays)
    compare(rg, want, variant + " three contigs with -D")
