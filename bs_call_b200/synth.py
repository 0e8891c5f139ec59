"""Host-side (numpy) synthetic workloads, used where no GPU is involved (the CPU reference arm of bench.py, CPU tests).

The device-side generators live in libbsgpu (bsgpu_synth_sites_dev / bsgpu_synth_block_dev); this module draws from
the same distributions but is not bit-identical to them.
"""
import numpy as np

from .records import PILEUP

# class of (bisulfite strand, base): reference table base_tab_st, src/call_genotypes.c:17-19
CLASS_OF = np.array([[0, 1, 2, 3], [0, 5, 2, 7], [4, 1, 6, 3]])


def synth_sites_numpy(seed, n, mean_depth=30.0):
    """n per-site count vectors: depth ~ Poisson(mean_depth) split over strand index and bisulfite strand,
    ref base A/C/G/T = .295/.205/.205/.295, 2 % methylated (m = 0.7) sites, 0.1 % het / 0.05 % hom-alt SNPs,
    0.3 % error bases, q in [20,43], MAPQ 60, 3 % empty sites.  Returns (pileup[n], ref[n])."""
    rng = np.random.default_rng(seed)
    rb = rng.choice(4, size=n, p=[0.295, 0.205, 0.205, 0.295])
    a0 = rb.copy()
    a1 = rb.copy()
    u = rng.random(n)
    alt = (rb + rng.integers(1, 4, size=n)) % 4
    het = u < 0.001
    hom = (u >= 0.001) & (u < 0.0015)
    a1[het] = alt[het]
    a0[hom] = alt[hom]
    a1[hom] = alt[hom]
    meth = np.where(rng.random(n) < 0.02, 0.7, 0.01)
    depth = rng.poisson(mean_depth, size=n)
    depth[rng.random(n) < 0.03] = 0
    # probability of each (ori, class) outcome per site
    pv = np.zeros((n, 2, 8))
    keep_c = meth + (1.0 - meth) * 0.01          # P(read C | allele C, converting strand)
    for allele in (a0, a1):
        for st in (1, 2):
            w = 0.25                               # allele 1/2 x strand 1/2
            # base actually read before errors, as a distribution over 4 bases
            pb = np.zeros((n, 4))
            pb[np.arange(n), allele] = 1.0
            if st == 1:
                isc = allele == 1
                pb[isc, 1] = keep_c[isc]
                pb[isc, 3] = 1.0 - keep_c[isc]
            else:
                isg = allele == 2
                pb[isg, 2] = keep_c[isg]
                pb[isg, 0] = 1.0 - keep_c[isg]
            pb = pb * (1.0 - 0.003) + (1.0 - pb) * (0.003 / 3.0)
            for b in range(4):
                pv[:, :, CLASS_OF[st, b]] += (w * 0.5 * pb[:, b])[:, None]
    pv = pv.reshape(n, 16)
    pv /= pv.sum(axis=1, keepdims=True)
    cnt = rng.multinomial(depth, pv).reshape(n, 2, 8).astype(np.uint32)
    tot = cnt.sum(axis=1)
    tot = tot.astype(np.int64)
    qsum = 20 * tot + rng.binomial(23 * tot, 0.5)
    p = np.zeros(n, dtype=PILEUP)
    p["counts"] = cnt
    p["n"] = depth
    p["quality"] = qsum.astype(np.float32)
    p["mapq2"] = (3600.0 * depth).astype(np.float32)
    return p, (rb + 1).astype(np.uint8)
