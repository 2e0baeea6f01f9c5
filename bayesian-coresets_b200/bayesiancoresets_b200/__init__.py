"""B200-native coreset construction behind the bayesian-coresets API.

    import bayesiancoresets_b200 as bc
    prj = bc.LogisticRegressionProjector(sampler, S)
    cs = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.GIGA)
    cs.build(200); wts, pts, idcs = cs.get(); cs.error()

Mirrors bayesiancoresets/__init__.py:1-2 for the accelerated path (HilbertCoreset, the GIGA /
FrankWolfe / OrthoPursuit solvers, Projector / BlackBoxProjector) and adds device projectors.
Requires the in-tree CUDA library (lib/libbcg_b200.so) and a B200; there is no CPU fallback.
"""
from . import util
from . import snnls
from .projector import (Projector, BlackBoxProjector, LogisticRegressionProjector, GaussianProjector,
                        PoissonProjector)
from .coreset import Coreset, HilbertCoreset, SparseVICoreset, BatchPSVICoreset, UniformSamplingCoreset
from ._native import DeviceVecs, Dataset, Context, BcgError, pinned_empty, pinned_copy
from . import comm
from .samplers import GaussianPosteriorSampler, LaplaceSampler
