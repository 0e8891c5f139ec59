"""Small seeded generators of synthetic bisulfite read blocks for the parity tests (numpy, CPU only).

Produces the flat template records shared by the oracle, the reference harness and the product's
raw-template entry point: `templates` (TEMPLATE dtype), `bases` (one packed byte per read base,
base | qual<<2, src/input_sam.c:76-86) and `misms` (CIGAR-derived events, src/input_sam.c:90-136).
"""
import numpy as np

from bs_call_b200.records import TEMPLATE, MISMS

INS, DEL, SOFT = 1, 2, 3      # reference enum gt_misms_t: CIGAR D -> INS, CIGAR I -> DEL, S -> SOFT


def random_reference(rng, length, gc=0.41, n_runs=0):
    """codes 0=N 1=A 2=C 3=G 4=T for positions 1..length (index i <-> position i+1)."""
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    ref = rng.choice(np.arange(1, 5, dtype=np.uint8), size=length, p=p).astype(np.uint8)
    for _ in range(n_runs):
        a = int(rng.integers(0, max(1, length - 50)))
        ref[a:a + int(rng.integers(5, 50))] = 0
    return ref


def window_codes(ref, lo, hi):
    """codes of positions [lo, hi] as get_sequence_string() hands them out: N from the contig's last position on
    (src/get_sequence.c:41-48)"""
    pos = np.arange(lo, hi + 1)
    padded = np.concatenate([ref, np.zeros(max(0, hi - len(ref)) + 1, dtype=np.uint8)])
    return np.where(pos < len(ref), padded[pos - 1], 0).astype(np.uint8)


def _sample_genotypes(rng, ref, snp_rate):
    """two haplotypes as base indices 0..3 (N positions get A)."""
    base = np.where(ref > 0, ref - 1, 0).astype(np.uint8)
    hap = np.stack([base, base.copy()])
    snp = rng.random(len(ref)) < snp_rate
    idx = np.nonzero(snp)[0]
    for i in idx:
        alt = (hap[0, i] + int(rng.integers(1, 4))) % 4
        if rng.random() < 2 / 3:
            hap[int(rng.integers(0, 2)), i] = alt
        else:
            hap[:, i] = alt
    return hap


def _read_bases(rng, hap_row, refp, lo, hi, strand, meth_cpg, meth_other, conv, qual):
    """forward-strand bases of positions [lo,hi) (0-based) as seen through bisulfite conversion.
    refp is the reference codes padded with one 0 on each side."""
    b = hap_row[lo:hi].copy()
    n = len(b)
    prv = refp[lo:lo + n]
    nxt = refp[lo + 2:lo + 2 + n]
    if strand == 1:      # C2T: unmethylated C reads as T
        isc = b == 1
        cpg = isc & (nxt == 3)
        m = np.where(cpg, meth_cpg, meth_other)
        convert = isc & (rng.random(n) >= m) & (rng.random(n) < conv)
        b[convert] = 3
    elif strand == 2:    # G2A: unmethylated C on the bottom strand reads as A on the top strand
        isg = b == 2
        cpg = isg & (prv == 2)
        m = np.where(cpg, meth_cpg, meth_other)
        convert = isg & (rng.random(n) >= m) & (rng.random(n) < conv)
        b[convert] = 0
    err = rng.random(n) < 10.0 ** (-qual / 10.0)
    b[err] = (b[err] + rng.integers(1, 4, size=int(err.sum()))) % 4
    return b


def _quals(rng, n, hi_frac=0.85):
    q = np.where(rng.random(n) < hi_frac, 37, rng.integers(2, 44, size=n)).astype(np.uint8)
    return q


def make_block(rng, ref, start, end, depth=30, read_len=100, paired=True, frag_mean=220, frag_sd=50,
               snp_rate=0.002, meth_cpg=0.7, meth_other=0.01, conv=0.99, indel_frac=0.0, clip_frac=0.0,
               n_frac=0.002, lowmapq_frac=0.03, nonconv_frac=0.0, max_mapq=60, single_mate_frac=0.0):
    """Templates whose leftmost starts lie in [start, end) (1-based), sorted by leftmost start.

    Returns (templates, bases, misms, y) with y = max over mates of pos + reference_span, as the reference's
    block builder computes it (src/get_template_vector.c:209-218)."""
    L = len(ref)
    hap = _sample_genotypes(rng, ref, snp_rate)
    refp = np.concatenate([[0], ref, [0]]).astype(np.uint8)
    span = end - start
    ntemp = max(1, int(depth * span / (read_len * (2 if paired else 1))))
    starts = np.sort(rng.integers(start, end, size=ntemp))
    T = np.zeros(ntemp, dtype=TEMPLATE)
    bases, misms = [], []
    boff = moff = 0
    y = 0
    for i, fs in enumerate(starts):
        fs = int(fs)
        t = T[i]
        top = rng.random() < 0.5
        strand = 0 if rng.random() < nonconv_frac else (1 if top else 2)
        t["bs_strand"] = strand
        t["orientation"] = 0 if top else 1
        hrow = hap[int(rng.integers(0, 2))]
        mates = []
        if paired and rng.random() >= single_mate_frac:
            flen = int(np.clip(rng.normal(frag_mean, frag_sd), read_len // 2, 1000))
            rl0 = min(read_len, flen)
            rl1 = min(read_len, flen)
            mates = [(0, fs, rl0), (1, fs + flen - rl1, rl1)]
        else:
            k = 0 if (not paired or rng.random() < 0.5) else 1
            mates = [(k, fs, read_len)]
        for (k, pos, rl) in mates:
            if pos + rl + 8 >= L:
                rl = max(1, L - 8 - pos)
            # build the read through a random CIGAR: M [D/I] M, optional soft clips
            ev = []
            ref_lo = pos - 1
            q_parts, b_parts = [], []
            refspan = 0
            readpos = 0
            lclip = int(rng.integers(1, 8)) if rng.random() < clip_frac else 0
            rclip = int(rng.integers(1, 8)) if rng.random() < clip_frac else 0
            if lclip:
                ev.append((SOFT, 0, lclip))
                b_parts.append(rng.integers(0, 4, size=lclip).astype(np.uint8))
                readpos += lclip
            body = max(4, rl - lclip - rclip)
            if rng.random() < indel_frac and body > 20:
                cut = int(rng.integers(5, body - 5))
                sz = int(rng.integers(1, 4))
                qual_for_err = 37
                b_parts.append(_read_bases(rng, hrow, refp, ref_lo, ref_lo + cut, strand, meth_cpg, meth_other, conv, qual_for_err))
                readpos += cut
                refspan += cut
                if rng.random() < 0.5:     # deletion from the reference (CIGAR D) -> INS event, zero-filled later
                    ev.append((INS, readpos, sz))
                    refspan += sz
                    rest = body - cut
                    b_parts.append(_read_bases(rng, hrow, refp, ref_lo + refspan, ref_lo + refspan + rest, strand, meth_cpg, meth_other, conv, qual_for_err))
                    readpos += rest
                    refspan += rest
                else:                      # insertion to the reference (CIGAR I) -> DEL event, dropped later
                    ev.append((DEL, readpos, sz))
                    b_parts.append(rng.integers(0, 4, size=sz).astype(np.uint8))
                    readpos += sz
                    rest = max(1, body - cut - sz)
                    b_parts.append(_read_bases(rng, hrow, refp, ref_lo + refspan, ref_lo + refspan + rest, strand, meth_cpg, meth_other, conv, qual_for_err))
                    readpos += rest
                    refspan += rest
            else:
                b_parts.append(_read_bases(rng, hrow, refp, ref_lo, ref_lo + body, strand, meth_cpg, meth_other, conv, 37))
                readpos += body
                refspan += body
            if rclip:
                ev.append((SOFT, readpos, rclip))
                b_parts.append(rng.integers(0, 4, size=rclip).astype(np.uint8))
                readpos += rclip
            b = np.concatenate(b_parts).astype(np.uint8)
            q = _quals(rng, len(b))
            packed = (b | (q << 2)).astype(np.uint8)
            packed[rng.random(len(b)) < n_frac] = 0          # N -> 0x00
            t["present"][k] = 1
            t["read_off"][k] = boff
            t["read_len"][k] = len(packed)
            t["reference_span"][k] = refspan
            t["mapq"][k] = int(rng.integers(0, 20)) if rng.random() < lowmapq_frac else int(rng.integers(20, max_mapq + 1))
            t["mm_off"][k] = moff
            t["mm_n"][k] = len(ev)
            if k == 0:
                t["forward_position"] = pos
            else:
                t["reverse_position"] = pos
            bases.append(packed)
            boff += len(packed)
            misms.extend(ev)
            moff += len(ev)
            y = max(y, pos + refspan)
    M = np.zeros(len(misms), dtype=MISMS)
    for i, (ty, po, sz) in enumerate(misms):
        M[i] = (ty, po, sz)
    B = np.concatenate(bases) if bases else np.zeros(0, dtype=np.uint8)
    return T, B, M, y


def random_pileups(rng, n, depth=30, het_frac=0.05, empty_frac=0.03, max_depth=None):
    """Per-site pileup records with realistic structure (for the likelihood kernel tests)."""
    from bs_call_b200.records import PILEUP
    P = np.zeros(n, dtype=PILEUP)
    ref = rng.integers(0, 5, size=n).astype(np.uint8)
    d = rng.poisson(depth, size=n)
    if max_depth:
        d = np.minimum(d, max_depth)
    d[rng.random(n) < empty_frac] = 0
    for i in range(n):
        di = int(d[i])
        if di == 0:
            continue
        rb = (int(ref[i]) - 1) if ref[i] else int(rng.integers(0, 4))
        alts = [rb]
        if rng.random() < het_frac:
            alts.append(int(rng.integers(0, 4)))
        for _ in range(di):
            b = alts[int(rng.integers(0, len(alts)))]
            if rng.random() < 0.01:
                b = int(rng.integers(0, 4))
            st = int(rng.integers(0, 3)) if rng.random() < 0.1 else int(rng.integers(1, 3))
            if st == 1 and b == 1 and rng.random() < 0.6:
                b = 3
            if st == 2 and b == 2 and rng.random() < 0.6:
                b = 0
            c = [[0, 1, 2, 3], [0, 5, 2, 7], [4, 1, 6, 3]][st][b]
            o = int(rng.integers(0, 2))
            q = int(rng.integers(20, 44))
            P["counts"][i, o, c] += 1
            P["quality"][i, c] += q
            P["mapq2"][i] += float(int(rng.integers(20, 61)) ** 2)
            P["n"][i] += 1
    return P, ref


# the profile vector grows with the longest read seen so far and counts on its top entry are dropped (see
# k_profile_resolve): the order of the blocks matters, so each run feeds several blocks of different shapes
PROFILE_RUNS = [
    [dict(depth=20, read_len=50, paired=False, nonconv_frac=0.2), dict(depth=20, read_len=100, paired=True), dict(depth=10, read_len=75, paired=False)],
    [dict(depth=25, read_len=100, paired=True, indel_frac=0.4, clip_frac=0.4, frag_mean=130, frag_sd=40),
     dict(depth=25, read_len=100, paired=True, single_mate_frac=0.5, clip_frac=0.3)],
    [dict(depth=15, read_len=60, paired=True, frag_mean=90, frag_sd=20, indel_frac=0.5, clip_frac=0.5, n_frac=0.05, single_mate_frac=0.2)] * 3,
]
