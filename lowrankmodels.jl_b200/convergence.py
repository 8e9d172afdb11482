"""ConvergenceHistory / update_ch! — mirror of /root/reference/src/convergence.jl:3-27."""
from __future__ import annotations


class ConvergenceHistory:
    def __init__(self, name: str = "unnamed_convergence_history", optval=0):
        self.name = name
        self.objective = []
        self.dual_objective = []
        self.primal_residual = []
        self.dual_residual = []
        self.times = []
        self.stepsizes = []
        self.optval = optval


def update_ch(ch: ConvergenceHistory, dt: float, obj: float, stepsize=0, pr=0, dr=0):
    """update_ch!(ch, dt, obj, ...) (convergence.jl:16-27): append, times are cumulative."""
    ch.objective.append(float(obj))
    ch.primal_residual.append(pr)
    ch.dual_residual.append(dr)
    ch.stepsizes.append(stepsize)
    if not ch.times:
        ch.times.append(float(dt))
    else:
        ch.times.append(ch.times[-1] + float(dt))
    return ch
