"""The whole program: the reference's unmodified bs_call binary
(oracle/_ref/bs_call: every file of src/Makefile:15-17 over oracle/minihts, this repository's stand-in for htslib) run from
the command line on a BAM file and a FASTA file, against the checkers everything else in tests/ is pinned to.

CPU part (this file, `reference` marker): the records of the BCF file the binary writes equal the compiled chain
read_input -> process_template_vector -> call_genotypes_ML -> print_vcf_entry of oracle/_ref/libbsref.so driven block by
block through the harness -- i.e. the file-level run and the in-memory harness are the same function, so a result that
holds against the harness holds against the program.  The GPU part is tests/test_gpu_full_binary.py.
"""
import os
import subprocess

import numpy as np
import pytest

from bs_call_b200 import hostio
from tests import bamgen, blockgen

pytestmark = pytest.mark.reference

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "bs_call")

# dictionary ids of PASS fail mac1 CX GT FT GL GQ DP MQ QD MC8 AMQ CS CG FS in the header print_vcf_header() writes
# (src/print_vcf.c:712-733: PASS first, then ids in order of first appearance: CX(INFO) fail q20 qd2 fs60 mq40 mac1 GT FT ...)
HEADER_IDS = (0, 2, 7, 1, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19)


def write_case(tmp, bam, tl, refs):
    names = ["ctg%d" % i for i in range(len(tl))]
    fa, bf = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "in.bam")
    hostio.write_fasta(fa, names, refs)
    hostio.write_bam(bf, names, tl, bam)
    return names, fa, bf


def run_binary(binary, fa, bf, out, otype="u", extra=(), threads="3", env=None):
    cmd = [binary, "-r", fa, "-n", "S", "--benchmark-mode", "-O", otype, "-o", out, "-t", threads, *extra, bf]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, "%s\n%s" % (" ".join(cmd), r.stderr[-3000:])
    return r.stderr


def chain_records(impl, bam, tl, refs, all_positions=False, dbsnp=None, **opts):
    """the BCF records of the harness-driven chain, block by block; dbsnp: per contig dbsnp_arrays(...) or None"""
    blocks, _, _, _, vcf = impl.read_input(bam, tl, refs, run_chain=True, **opts)
    want = []
    for w in blocks:
        c = int(w["tid"])
        n = int(w["y"]) - int(w["x"]) + 1
        rec, _ = impl.print_block(vcf[int(w["vcf_off"]):int(w["vcf_off"]) + n], blockgen.window_codes(refs[c], int(w["x"]), int(w["y"]) + 2),
                                  int(w["x"]), rid=c, ctg_end=int(tl[c]), vcf_ids=HEADER_IDS, all_positions=all_positions,
                                  dbsnp=None if dbsnp is None else dbsnp[c])
        want.append(rec)
    return np.concatenate(want) if want else np.zeros(0, dtype=np.uint8)


@pytest.fixture(scope="module")
def binary():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/bs_call not built (reference tree absent at build time)")
    return BIN


@pytest.mark.parametrize("seed,extra,opts", [
    (3, (), {}),
    (11, ("-k",), {"keep_unmatched": True}),
    (12, ("-d",), {"keep_duplicates": True}),
    (13, ("-q", "5", "-l", "400"), {"mapq_thresh": 5, "max_template_len": 400}),
])
def test_binary_matches_harness_chain(binary, reference, tmp_path, seed, extra, opts):
    from oracle.bindings import bcf_diff
    bam, n, tl, refs = bamgen.make_stream(seed, n_contigs=1, contig_len=30000)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    out = os.path.join(str(tmp_path), "cpu.bcf")
    run_binary(binary, fa, bf, out, extra=extra)
    text, got = hostio.read_bcf(out)
    assert "##contig=<ID=ctg0,length=%d>" % int(tl[0]) in text and text.rstrip().endswith("FORMAT\tS")
    want = chain_records(reference, bam, tl, refs, **opts)
    d = bcf_diff(got, want)
    assert d["records_a"] == d["records_b"] == d["identical"] and d["records_a"] > 100 and d["order_violations"] == 0, d


def _keyed(buf):
    from tests.util import split_bcf
    out = {}
    for r in split_bcf(buf):
        h = np.frombuffer(r[:32], dtype=np.uint32)
        out[(int(h[2]), int(h[3]) + 1)] = r
    return out


def contig_tail_race(got, want, blocks):
    """Several contigs: the reference binary has a race at contig ends.  process_thread fetches the next contig's sequence as
    soon as call_genotypes_ML has handed the last block of the old one to the calc threads (src/process_template.c:30 ->
    src/get_sequence.c:24 -> free_sequence, src/read_reference.c:35-42: end_pos = 0) while print_thread is still writing that
    block and clips every record against end_pos (src/print_vcf.c:157): records of the LAST block of a contig that is not the
    last one can be missing from the file, nothing else.  -> number of records lost that way; asserts everything else."""
    g, w = _keyed(got), _keyed(want)
    assert set(g) <= set(w)
    assert all(g[k] == w[k] for k in g)
    last = {}
    for b in blocks:
        last[int(b["tid"])] = (int(b["x"]), int(b["y"]))
    final_ctg = max(last)
    for (c, pos) in set(w) - set(g):
        assert c != final_ctg and last[c][0] <= pos <= last[c][1] + 2, (c, pos, last[c])
    return len(w) - len(g)


def test_binary_on_several_contigs(binary, reference, tmp_path):
    bam, n, tl, refs = bamgen.make_stream(3, n_contigs=3)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    out = os.path.join(str(tmp_path), "cpu.bcf")
    run_binary(binary, fa, bf, out)
    _, got = hostio.read_bcf(out)
    blocks = reference.read_input(bam, tl, refs)[0]
    want = chain_records(reference, bam, tl, refs)
    lost = contig_tail_race(got, want, blocks)
    print("records lost to the reference's end-of-contig race: %d of %d" % (lost, len(_keyed(want))))


def test_binary_all_positions_and_vcf_text(binary, reference, tmp_path):
    """-A writes every site; -O v is the same records as text (one line per BCF record, same positions)"""
    from oracle.bindings import bcf_diff
    bam, n, tl, refs = bamgen.make_stream(5, n_contigs=1, contig_len=20000)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    out_b, out_v = os.path.join(str(tmp_path), "a.bcf"), os.path.join(str(tmp_path), "a.vcf")
    run_binary(binary, fa, bf, out_b, extra=("-A",))
    run_binary(binary, fa, bf, out_v, otype="v", extra=("-A",))
    _, got = hostio.read_bcf(out_b)
    want = chain_records(reference, bam, tl, refs, all_positions=True)
    d = bcf_diff(got, want)
    assert d["records_a"] == d["records_b"] == d["identical"] and d["order_violations"] == 0, d
    lines = [l for l in open(out_v).read().split("\n") if l and not l.startswith("#")]
    assert len(lines) == d["records_a"]
    # positions and contigs of the text lines follow the BCF records
    from tests.util import split_bcf
    recs = split_bcf(got)
    for ln, r in list(zip(lines, recs))[::97]:
        f = ln.split("\t")
        hdr = np.frombuffer(r[:32], dtype=np.uint32)
        assert f[0] == names[int(hdr[2])] and int(f[1]) == int(hdr[3]) + 1
        assert f[8].startswith("GT:FT:") or f[8].startswith("GT:")


def test_compressed_bcf_is_the_same_stream(binary, tmp_path):
    """-O b (BGZF level 6) and -O u (BGZF level 0) hold the same bytes"""
    bam, n, tl, refs = bamgen.make_stream(7, n_contigs=1)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    a, b = os.path.join(str(tmp_path), "u.bcf"), os.path.join(str(tmp_path), "b.bcf")
    run_binary(binary, fa, bf, a, otype="u")
    run_binary(binary, fa, bf, b, otype="b")
    assert os.path.getsize(b) < os.path.getsize(a)
    ta, ra = hostio.read_bcf(a)
    tb, rb = hostio.read_bcf(b)
    assert ta == tb and ra.tobytes() == rb.tobytes() and len(ra) > 0


def synthetic_dbsnp(rng, tl, frac=0.02):
    """-> ({contig name: entries for hostio.write_dbsnp_index}, [per contig dbsnp_arrays(...)]): known sites on `frac` of the
    positions, a fifth of them flagged "always written" (src/dbSNP.c:278-279), ids rs + 6 or 8 digits, a second prefix"""
    from oracle.bindings import dbsnp_arrays
    files, arrays = {}, []
    for c, L in enumerate(tl):
        pos = np.unique(rng.integers(1, int(L) + 1, size=max(4, int(frac * int(L)))))
        ents, flat = [], []
        for p_ in pos:
            pfx = int(rng.integers(0, 2))
            digits = "%0*d" % (6 if rng.random() < 0.3 else 8, int(rng.integers(0, 999999)))
            always = bool(rng.random() < 0.2)
            ents.append((int(p_), pfx, digits, always))
            flat.append((int(p_), 3 if always else 1, (("rs", "ss")[pfx] + digits).encode()))
        files["ctg%d" % c] = ents
        arrays.append(dbsnp_arrays(flat))
    return files, arrays


def test_binary_with_dbsnp_index(binary, reference, tmp_path):
    """-D: the reference's own index reader (src/dbSNP.c, never run by the harness, which serves dbSNP_lookup_name from a
    table) over an index file written in its format; ids, "always written" sites and everything else as the harness chain
    gives them with the same table"""
    from oracle.bindings import bcf_diff
    rng = np.random.default_rng(77)
    bam, n, tl, refs = bamgen.make_stream(31, n_contigs=1, contig_len=30000)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    files, arrays = synthetic_dbsnp(rng, tl)
    idx = os.path.join(str(tmp_path), "db.idx")
    hostio.write_dbsnp_index(idx, files, prefixes=("rs", "ss"), bins_per_block=50)
    out = os.path.join(str(tmp_path), "cpu.bcf")
    run_binary(binary, fa, bf, out, extra=("-D", idx))
    _, got = hostio.read_bcf(out)
    want = chain_records(reference, bam, tl, refs, dbsnp=arrays)
    d = bcf_diff(got, want)
    assert d["records_a"] == d["records_b"] == d["identical"] and d["order_violations"] == 0, d
    plain = chain_records(reference, bam, tl, refs)
    assert len(want) > len(plain)                      # ids and always-written sites are in there


def test_bgzf_files_are_plain_multi_member_gzip(binary, tmp_path):
    """BGZF is a series of gzip members (SAM specification 4.1): what hostio writes as input and what the stand-in writes as
    output (-O b: deflated, -O u: stored blocks) must decompress with an independent gzip implementation to the same payload
    the BGZF-aware reader sees, and end with the 28-byte empty EOF block"""
    import gzip
    bam, n, tl, refs = bamgen.make_stream(9, n_contigs=1)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    raw = open(bf, "rb").read()
    payload = gzip.decompress(raw)
    assert payload == hostio.read_bgzf(bf) and payload.startswith(b"BAM\x01") and payload.endswith(bytes(bam[-64:]))
    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    assert raw.endswith(eof)
    for ot in ("u", "b"):
        out = os.path.join(str(tmp_path), "o_%s.bcf" % ot)
        run_binary(binary, fa, bf, out, otype=ot)
        r = open(out, "rb").read()
        assert r.endswith(eof)
        assert gzip.decompress(r) == hostio.read_bgzf(out)
    vz = os.path.join(str(tmp_path), "o.vcf.gz")
    run_binary(binary, fa, bf, vz, otype="z")
    vt = os.path.join(str(tmp_path), "o.vcf")
    run_binary(binary, fa, bf, vt, otype="v")
    assert gzip.decompress(open(vz, "rb").read()) == open(vt, "rb").read()


def test_binary_region_runs_concatenate_to_the_whole_run(binary, reference, tmp_path):
    """Level-2 sharding as the reference documents it (src/process_sam_header.c:52-70, 108-169: one run per -C region, outputs
    concatenated): a contig cut between two blocks, each half called by its own process through the region machinery
    (sam_itr_queryi / sam_itr_next, src/get_template_vector.c:69-74,93-97; the writer clips to the region, src/print_vcf.c:154-157)
    -- the two files hold exactly the records of the one-process run.  Also: regions asked for out of order, and a region whose
    first reads begin before it."""
    bam, n, tl, refs = bamgen.make_stream(41, n_contigs=1, contig_len=40000)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    blocks = reference.read_input(bam, tl, refs)[0]
    assert len(blocks) >= 4
    k = len(blocks) // 2
    assert int(blocks[k]["x"]) - int(blocks[k - 1]["y"]) > 6
    cut = (int(blocks[k - 1]["y"]) + int(blocks[k]["x"])) // 2
    L = int(tl[0])
    whole = os.path.join(str(tmp_path), "whole.bcf")
    run_binary(binary, fa, bf, whole)
    _, rw = hostio.read_bcf(whole)
    parts = []
    for tag, (a, b) in (("a", (0, cut)), ("b", (cut, L))):
        bed = os.path.join(str(tmp_path), tag + ".bed")
        open(bed, "w").write("ctg0\t%d\t%d\n" % (a, b))
        out = os.path.join(str(tmp_path), tag + ".bcf")
        err = run_binary(binary, fa, bf, out, extra=("-C", bed))
        assert "Processing region ctg0:%d-%d" % (a + 1, b) in err and "(Index)" in err
        parts.append(hostio.read_bcf(out)[1])
    assert len(parts[0]) > 0 and len(parts[1]) > 0
    assert np.concatenate(parts).tobytes() == rw.tobytes()
    # a region that begins inside a block: reads that start before it but reach into it are part of it (htslib's overlap rule),
    # so the sites of the region get the same records as in the whole run
    x, y = int(blocks[k]["x"]) + 40, int(blocks[k]["y"]) - 40
    bed = os.path.join(str(tmp_path), "in.bed")
    open(bed, "w").write("ctg0\t%d\t%d\n" % (x - 1, y))
    out = os.path.join(str(tmp_path), "in.bcf")
    run_binary(binary, fa, bf, out, extra=("-C", bed))
    gi, gw = _keyed(hostio.read_bcf(out)[1]), _keyed(rw)
    inside = {kk: v for kk, v in gw.items() if x <= kk[1] <= y}
    assert set(gi) == set(inside) and len(gi) > 20
    # interior sites (more than a read length from the region's edges, where every read that covers them is in the region) are identical
    deep = [kk for kk in gi if x + 150 <= kk[1] <= y - 150]
    assert deep and all(gi[kk] == gw[kk] for kk in deep)


def test_binary_regions_out_of_file_order(binary, reference, tmp_path):
    """-C with two contigs named in the opposite of their order in the file: the scanning stand-in for the index starts over
    from the first record for the second region; the third contig is never read into a block"""
    bam, n, tl, refs = bamgen.make_stream(3, n_contigs=3)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    bed = os.path.join(str(tmp_path), "r.bed")
    open(bed, "w").write("ctg2\t0\t%d\nctg0\t0\t%d\n" % (int(tl[2]), int(tl[0])))
    out = os.path.join(str(tmp_path), "r.bcf")
    err = run_binary(binary, fa, bf, out, extra=("-C", bed))
    assert err.index("Processing region ctg2") < err.index("Processing region ctg0")
    text, got = hostio.read_bcf(out)
    assert "##contig=<ID=ctg1" not in text
    # the header lists the wanted contigs in the file's order: CHROM 0 = ctg0, 1 = ctg2
    g = _keyed(got)
    blocks = reference.read_input(bam, tl, refs)[0]
    want = _keyed(chain_records(reference, bam, tl, refs))
    w0 = {(0, p): v for (c, p), v in want.items() if c == 0}
    w2 = {(1, p): v[:8] + np.array([1], dtype="<i4").tobytes() + v[12:] for (c, p), v in want.items() if c == 2}      # CHROM field renumbered
    g0 = {k: v for k, v in g.items() if k[0] == 0}
    g2 = {k: v for k, v in g.items() if k[0] == 1}
    assert len(g0) + len(g2) == len(g)
    assert g0 == w0                                            # processed last: complete
    last2 = [b for b in blocks if int(b["tid"]) == 2][-1]
    assert set(g2) <= set(w2) and all(g2[k] == w2[k] for k in g2)
    for (c, p) in set(w2) - set(g2):                           # processed first: its last block may lose records to the reference's race
        assert int(last2["x"]) <= p <= int(last2["y"]) + 2


def test_dbsnp_index_with_explicit_prefix_ids(binary, reference, tmp_path):
    """index entries whose name prefix does not fit the two bits of the entry byte carry a two-byte prefix id
    (src/dbSNP.c:235-243, 326-331): five prefixes, ids of 4 to 10 digits, several compressed blocks per contig"""
    from oracle.bindings import bcf_diff, dbsnp_arrays
    rng = np.random.default_rng(8)
    bam, n, tl, refs = bamgen.make_stream(33, n_contigs=1, contig_len=25000)
    names, fa, bf = write_case(str(tmp_path), bam, tl, refs)
    prefixes = ("rs", "ss", "esv", "nsv", "x")
    ents, flat = [], []
    for p_ in np.unique(rng.integers(1, int(tl[0]) + 1, size=900)):
        pfx = int(rng.integers(0, 5))
        digits = "%0*d" % (int(rng.choice([4, 6, 8, 10])), int(rng.integers(0, 9999)))
        always = bool(rng.random() < 0.25)
        ents.append((int(p_), pfx, digits, always))
        flat.append((int(p_), 3 if always else 1, (prefixes[pfx] + digits).encode()))
    idx = os.path.join(str(tmp_path), "db.idx")
    hostio.write_dbsnp_index(idx, {"ctg0": ents}, prefixes=prefixes, bins_per_block=7)
    out = os.path.join(str(tmp_path), "cpu.bcf")
    run_binary(binary, fa, bf, out, extra=("-D", idx))
    _, got = hostio.read_bcf(out)
    want = chain_records(reference, bam, tl, refs, dbsnp=[dbsnp_arrays(flat)])
    d = bcf_diff(got, want)
    assert d["records_a"] == d["records_b"] == d["identical"] and d["order_violations"] == 0, d
    assert any(b"esv" in r or b"nsv" in r for r in _keyed(got).values())
