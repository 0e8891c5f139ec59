"""bs_call_b200: B200-native pileup + bisulfite genotype-likelihood path of bs_call (host mirror over libbsgpu)."""
from .records import PILEUP, GT_METH, GT_VCF, SEG, TEMPLATE, MISMS, GENOTYPES  # noqa: F401
