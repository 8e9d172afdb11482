"""lowrankmodels.jl_b200 — B200-native engine for the prox-grad hot path of LowRankModels.jl.

Host-side mirror of the reference interface for that path (GLRM / Loss / Regularizer /
ProxGradParams / fit!), a ctypes binding of the C ABI (include/glrm_b200.h), and the CUDA engine
under csrc/.  Import through the root-level `lowrankmodels_b200` loader (the directory name carries
a dot, which Python's import statement cannot spell)."""
from .convergence import ConvergenceHistory, update_ch
from .fit import Engine, fit, fit_inplace
from .glrm import GLRM, ObsLists, Repeated, add_offset, scale_regularizer, sort_observations
from .losses import (BvSLoss, ClassificationLoss, DiffLoss, HingeLoss, HuberLoss, L1Loss, LogisticLoss,
                     Loss, MultinomialLoss, MultinomialOrdinalLoss, OrdinalHingeLoss, OrdisticLoss,
                     OvALoss, PeriodicLoss, PoissonLoss, QuadLoss, QuantileLoss, WeightedHingeLoss,
                     embedding_dim, get_yidxs)
from .params import AbstractParams, B200ProxGradParams, Params, ProxGradParams, SparseProxGradParams
from .regularizers import (KSparseConstraint, MNLOrdinalReg, NonNegConstraint, NonNegOneReg, OneReg,
                           OneSparseConstraint, OrdinalReg, QuadConstraint, QuadReg, Regularizer,
                           RemQuadReg, SimplexConstraint, UnitOneSparseConstraint, ZeroReg,
                           fixed_last_latent_features, fixed_latent_features, FixedLatentFeaturesConstraint,
                           FixedLastLatentFeaturesConstraint, lastentry1,
                           lastentry_unpenalized)
from .encode import encode_params, encode_problem, encode_sparse_params
from .domains import (BoolDomain, CategoricalDomain, CountDomain, Domain, OrdinalDomain, PeriodicDomain, RealDomain,
                      error_metric, impute, impute_missing, loss_domain)
from . import _abi, distributed, synth

__all__ = [n for n in dir() if not n.startswith("_")]
