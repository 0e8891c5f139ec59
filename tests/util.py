"""Shared comparison helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerances.  Integer / byte / index fields are bit-exact.  The doubles (gt_prob[], fisher_strand) go through
# log/exp/lgamma, where glibc (reference) and the CUDA math library may differ in the last ulp: BASELINE.json's
# north_star allows 1e-6 relative in log-likelihood; we hold the path to a much tighter 1e-9 relative with an
# absolute floor of 1e-12 (log10 units).
PROB_RTOL = 1e-9
PROB_ATOL = 1e-12

INT_FIELDS = ("counts", "qual", "mq", "aq")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def assert_pileup_equal(got, want):
    for f in ("counts", "n", "quality", "mapq2"):
        if got[f].tobytes() != want[f].tobytes():
            bad = np.nonzero(np.any((got[f] != want[f]).reshape(len(got), -1), axis=1))[0]
            raise AssertionError("pileup field %s differs at %d sites; first %d: got %r want %r"
                                 % (f, len(bad), bad[0], got[f][bad[0]], want[f][bad[0]]))


def near_tie(want_prob):
    """sites whose two best log10 posteriors are closer than 1e-9 (for reports; no comparison is waived on this ground)"""
    s = np.sort(want_prob, axis=1)
    return (s[:, -1] - s[:, -2]) < 1e-9


def assert_gt_meth_close(got, got_skip, want, want_skip, exact_doubles=False, flagged=None, report=None):
    """got/want: GT_METH arrays; skip flags compared exactly; called sites compared field by field.

    max_gt must be IDENTICAL -- except at sites the device itself flagged as standing inside its guard band (`flagged`:
    indices into got, from bsgpu_guard_read kind 1).  There the call may differ from the reference's, provided the
    reference's own two best posteriors are within the band too; the number of such sites goes to report["flagged_differ"]."""
    assert got_skip.tobytes() == np.asarray(want_skip, dtype=np.uint8).tobytes(), "skip flags differ"
    m = np.asarray(want_skip) == 0
    g, w = got[m], want[m]
    for f in INT_FIELDS:
        if g[f].tobytes() != w[f].tobytes():
            bad = np.nonzero(np.any((g[f] != w[f]).reshape(len(g), -1), axis=1))[0]
            raise AssertionError("field %s differs at %d sites; first: got %r want %r" % (f, len(bad), g[f][bad[0]], w[f][bad[0]]))
    fl = np.zeros(len(got), dtype=bool)
    if flagged is not None and len(flagged):
        fl[np.asarray(flagged, dtype=np.int64)] = True
    fl = fl[m]
    differ = g["max_gt"] != w["max_gt"]
    diff = differ & ~fl
    assert not diff.any(), "max_gt differs at %d sites the device did not flag, first got %r want %r (probs %r)" % (
        diff.sum(), g["max_gt"][diff][0], w["max_gt"][diff][0], w["gt_prob"][diff][0])
    both = differ & fl
    assert near_tie(w["gt_prob"][both]).all(), "a flagged site differs although the reference's two best posteriors are well apart"
    if report is not None:
        report["flagged"] = report.get("flagged", 0) + int(fl.sum())
        report["flagged_differ"] = report.get("flagged_differ", 0) + int(both.sum())
    if exact_doubles:
        assert g["gt_prob"].tobytes() == w["gt_prob"].tobytes()
        assert g["fisher_strand"].tobytes() == w["fisher_strand"].tobytes()
    else:
        np.testing.assert_allclose(g["gt_prob"], w["gt_prob"], rtol=PROB_RTOL, atol=PROB_ATOL)
        same_gt = g["max_gt"] == w["max_gt"]
        np.testing.assert_allclose(g["fisher_strand"][same_gt], w["fisher_strand"][same_gt], rtol=1e-8, atol=1e-10)
    # skipped sites carry an all-zero record
    assert not got[~m].tobytes().strip(b"\0"), "skipped sites must be zero records"
    return int(m.sum())


def assert_vcf_close(got, want):
    assert (got["ready"] == 1).all()
    return assert_gt_meth_close(got["gtm"], got["skip"], want["gtm"], want["skip"])


def writer_fields(gtm, skip):
    """The integer fields the reference's writer derives from a gt_meth record (src/print_vcf.c:140-217, 584-591):
    genotype (first strict maximum of gt_prob[]), QUAL / GQ phred, FS, QD and the FILTER bits q20 / qd2 / fs60 / mq40 /
    mac1.  numpy restatement used to check that ulp-level differences in the doubles never reach the VCF."""
    m = np.asarray(skip) == 0
    g = gtm[m]
    prob = g["gt_prob"]
    gt = np.argmax(prob, axis=1)                      # first maximum, like the strict '>' scan of the writer
    z = prob[np.arange(len(g)), gt]
    z1 = np.exp(z * np.log(10.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        ph = np.where(z1 >= 1.0, 255, np.minimum((-10.0 * np.log(1.0 - np.minimum(z1, 1.0 - 1e-300)) / np.log(10.0)), 255.0)).astype(np.int64)
    ph = np.where(z1 >= 1.0, 255, ph)
    fs = (-g["fisher_strand"] * 10.0 + 0.5).astype(np.int64)
    c = g["counts"].astype(np.int64)
    dp1 = c[:, :4].sum(axis=1)
    qd = np.where(dp1 > 0, ph // np.maximum(dp1, 1), ph)
    flt = (ph < 20) * 1 + (qd < 2) * 2 + (fs > 60) * 4 + (g["mq"] < 40) * 8
    a = {1: (c[:, 1] + c[:, 5] + c[:, 7], c[:, 0] + c[:, 4]), 2: (c[:, 2] + c[:, 6], c[:, 0]), 3: (c[:, 3] + c[:, 7], c[:, 0] + c[:, 4]),
         5: (c[:, 2] + c[:, 6] + c[:, 4], c[:, 1] + c[:, 5] + c[:, 7]), 6: (c[:, 3], c[:, 1] + c[:, 5]),
         8: (c[:, 3] + c[:, 7], c[:, 2] + c[:, 6] + c[:, 4])}
    mac1 = np.zeros(len(g), dtype=bool)
    for k, (u, v) in a.items():
        mac1 |= (gt == k) & ((u <= 1) | (v <= 1))
    flt = np.where((flt == 0) & mac1, 128, flt)
    return {"gt": gt, "phred": ph, "fs": fs, "qd": qd, "flt": flt}


def same_profile(a, b, what="", recycled_vectors=False):
    """a: restatement or product, b: reference.  recycled_vectors: b comes from read_input's recycled align_details, whose
    absent mates are often EMPTY rather than NULL vectors (src/al_utils.c:52-63); process_template_vector counts those
    as reads with no bases (src/process_template.c:49,61), so filter_cts[gt_flt_none] of the reference then depends on its
    allocation history and can only be bounded from below by the number of mates that exist.  filter_bases[gt_flt_none]
    is then also written by two threads without a lock (read_input at src/get_template_vector.c:363 and the process
    thread at src/process_template.c:62), so the reference may lose updates: bounded from above by the exact sum."""
    assert a["used"] == b["used"], (what, a["used"], b["used"])
    np.testing.assert_array_equal(a["conv"], b["conv"], err_msg=what)
    np.testing.assert_array_equal(a["base_filter"], b["base_filter"], err_msg=what)
    np.testing.assert_array_equal(a["filter_cts"][1:], b["filter_cts"][1:], err_msg=what)
    np.testing.assert_array_equal(a["filter_bases"][1:], b["filter_bases"][1:], err_msg=what)
    if recycled_vectors:
        assert a["filter_cts"][0] <= b["filter_cts"][0], what
        assert a["filter_bases"][0] >= b["filter_bases"][0] >= 0.98 * a["filter_bases"][0], what
    else:
        assert a["filter_cts"][0] == b["filter_cts"][0] and a["filter_bases"][0] == b["filter_bases"][0], what


def golden_profile(name):
    """the reference's --report-file side channels for golden `name` (tests/golden/profile_v1.npz)"""
    g = load_golden("profile_v1")
    return {k: g[name + "__" + k] for k in ("used", "conv", "base_filter", "filter_cts", "filter_bases")}, g.get(name + "__ref")


def random_gt_vcf(rng, n, skip_frac=0.25, deep_frac=0.05):
    """gt_vcf records that are not the output of any pileup but visit every branch of the writer: any call, any depth (one-
    and two-byte integer vectors), posteriors from certain to hopeless, strand bias, low MQ, empty count slots"""
    from bs_call_b200.records import GT_VCF
    v = np.zeros(n, dtype=GT_VCF)
    g = v["gtm"]
    depth = np.where(rng.random(n) < deep_frac, rng.integers(100, 70000, size=n), rng.integers(0, 40, size=n))
    for k in range(8):
        g["counts"][:, k] = (rng.random(n) < 0.45) * rng.integers(0, 1 + depth // 3 + 1, size=n)
    g["qual"] = rng.integers(0, 44, size=(n, 8))
    lp = -np.abs(rng.normal(0, 1, size=(n, 10))) * rng.choice([0.01, 1.0, 30.0, 200.0], size=(n, 1))
    best = rng.integers(0, 10, size=n)
    lp[np.arange(n), best] = -np.abs(rng.normal(0, 1, size=n)) * rng.choice([0.0, 1e-9, 1e-3, 0.3], size=n)
    g["gt_prob"] = lp
    g["fisher_strand"] = -np.abs(rng.normal(0, 1, size=n)) * rng.choice([0.0, 0.5, 4.0, 9.0], size=n)
    g["mq"] = rng.integers(0, 61, size=n)
    g["aq"] = rng.integers(0, 44, size=n)
    g["max_gt"] = best
    v["ready"] = 1
    v["skip"] = (rng.random(n) < skip_frac) | (g["counts"].sum(axis=1) == 0)
    return v


def split_bcf(buf):
    """BCF record bytes -> list of per-record byte strings"""
    out, at = [], 0
    while at < len(buf):
        l = 8 + int(buf[at:at + 4].view("<u4")[0]) + int(buf[at + 4:at + 8].view("<u4")[0])
        out.append(buf[at:at + l].tobytes())
        at += l
    assert at == len(buf)
    return out

