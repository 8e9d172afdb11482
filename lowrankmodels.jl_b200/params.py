"""ProxGradParams — mirror of /root/reference/src/algorithms/proxgrad.jl:4-31."""
from __future__ import annotations


class AbstractParams:  # fit.jl:4
    pass


class ProxGradParams(AbstractParams):
    """Same seven fields, same keyword constructor and defaults as the reference
    (proxgrad.jl:13-31); `inner_iter` raises both inner counts (:22-23)."""

    def __init__(self, stepsize=1.0, *, max_iter=100, inner_iter_X=1, inner_iter_Y=1, inner_iter=1,
                 abs_tol=0.00001, rel_tol=0.0001, min_stepsize=None):
        stepsize = float(stepsize)
        self.stepsize = stepsize
        self.max_iter = int(max_iter)
        self.inner_iter_X = max(int(inner_iter_X), int(inner_iter))
        self.inner_iter_Y = max(int(inner_iter_Y), int(inner_iter))
        self.abs_tol = float(abs_tol)
        self.rel_tol = float(rel_tol)
        self.min_stepsize = float(0.01 * stepsize if min_stepsize is None else min_stepsize)

    def __repr__(self):
        return ("ProxGradParams(stepsize=%g, max_iter=%d, inner_iter_X=%d, inner_iter_Y=%d, abs_tol=%g, "
                "rel_tol=%g, min_stepsize=%g)" % (self.stepsize, self.max_iter, self.inner_iter_X,
                                                   self.inner_iter_Y, self.abs_tol, self.rel_tol,
                                                   self.min_stepsize))


class B200ProxGradParams(ProxGradParams):
    """The solver-selection type the Julia shim adds (`struct B200ProxGradParams <: AbstractParams`,
    julia/LowRankModelsB200.jl): ProxGradParams' fields + the device to run on."""

    def __init__(self, stepsize=1.0, *, device=0, **kw):
        super().__init__(stepsize, **kw)
        self.device = int(device)


def Params(*args, **kwargs):  # fit.jl:5
    return ProxGradParams(*args, **kwargs)


class SparseProxGradParams(AbstractParams):
    """Mirror of /root/reference/src/algorithms/sparse_proxgrad.jl:4-18 (five fields, same defaults)."""

    def __init__(self, stepsize=1.0, *, max_iter=100, inner_iter=1, abs_tol=0.00001, min_stepsize=None, device=0):
        stepsize = float(stepsize)
        self.stepsize = stepsize
        self.max_iter = int(max_iter)
        self.inner_iter = int(inner_iter)
        self.abs_tol = float(abs_tol)
        self.min_stepsize = float(0.01 * stepsize if min_stepsize is None else min_stepsize)
        self.device = int(device)
