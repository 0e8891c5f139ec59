"""Drop-in check (-m gpu): the reference's own process_template_vector (unmodified objects, host side) calling the
product's link-compatible call_genotypes_ML / init_calc_threads / join_calc_threads
(bs_call_b200/csrc/bsgpu_dropin.c over libbsgpu.so) -- oracle/_ref/libbsref_gpu.so, prebuilt where the reference
tree exists.  The gt_vcf[] the print thread would consume must be what the all-CPU reference produced (goldens)."""
import numpy as np
import pytest

from tests import blockgen, util

pytestmark = pytest.mark.gpu

BLOCKS = ["block_pe_plain", "block_pe_indel_clip_trim", "block_se_deep", "block_mixed"]


@pytest.fixture(scope="module")
def dropin():
    from oracle.bindings import ReferenceWithGpuDropin, dropin_available
    if not dropin_available():
        pytest.skip("oracle/_ref/libbsref_gpu.so not built (reference tree absent at build time)")
    return ReferenceWithGpuDropin(calc_threads=1)


@pytest.mark.parametrize("name", BLOCKS)
def test_dropin_blocks_match_reference_goldens(dropin, name):
    from oracle.bindings import ReferenceWithGpuDropin
    g = util.load_golden(name)
    d = ReferenceWithGpuDropin(left_trim=tuple(int(v) for v in g["left_trim"]), right_trim=tuple(int(v) for v in g["right_trim"]))
    # process_block wants the whole contig; rebuild one that has the golden window at the right place
    x, y = int(g["x"]), int(g["y"])
    ctg = np.zeros(y + 16, dtype=np.uint8)
    ctg[x - 1:x - 1 + len(g["ref"])] = g["ref"]
    xo, pile, vcf, ref, nt, nb = d.process_block(g["templates"], g["bases"], g["misms"], ctg, y)
    assert xo == x
    assert nb.tobytes() == g["norm_bases"].tobytes()          # host normalisation is the reference's own
    util.assert_vcf_close(vcf, g["vcf"])
    ReferenceWithGpuDropin()


def test_dropin_many_blocks_back_to_back(dropin, oracle):
    """successive blocks of different sizes through the same hand-off (buffer growth, ref/ref1 swap each block)"""
    rng = np.random.default_rng(9)
    for i, span in enumerate((3000, 40000, 800, 15000)):
        ref = blockgen.random_reference(rng, span + 3000, n_runs=1)
        T, B, M, y = blockgen.make_block(rng, ref, 200, 200 + span, depth=20, read_len=100, paired=True, indel_frac=0.1, clip_frac=0.1)
        x, pile, vcf, refw, nt, nb = dropin.process_block(T, B, M, ref, y)
        xo, wpile, want = oracle.process_block(T, B, M, refw, y)
        assert xo == x
        util.assert_vcf_close(vcf, want)
