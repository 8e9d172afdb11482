"""proxgrad_ref.py — TEST INFRASTRUCTURE, NOT PRODUCT.

A second, independent restatement of the reference's hot path in plain Python/numpy, written to read
like the Julia source line by line (small cases only; pure-Python loops):

    fit!(glrm, ::ProxGradParams)        /root/reference/src/algorithms/proxgrad.jl:34-220
    row_objective / col_objective       /root/reference/src/evaluate_fit.jl:24-55
    objective / calc_penalty            /root/reference/src/evaluate_fit.jl:4-23,91-104
    evaluate / grad of each loss        /root/reference/src/losses.jl:136-676
    evaluate / prox of each regularizer /root/reference/src/regularizers.jl:52-348

It exists so that the C oracle (oracle/glrm_oracle.c) is checked against something that shares no
code with it.  It works on the host-side mirror objects (lowrankmodels.jl_b200: GLRM, losses,
regularizers), materialises XY exactly like the reference, and is O(m*d*k) per iteration.

PARITY PINNING: scalar formulas are pinned by the reference's own known answers
(tests/golden/reference_known_answers.json); end-to-end trajectories are unpinned (no Julia here).
"""
from __future__ import annotations

import math
import time

import numpy as np

TOL = 1e-12  # regularizers.jl:25


# ------------------------------------------------------------------------------------------- losses
def myBool(a):  # losses.jl:104
    if a == 1:
        return True
    if a == -1 or a == 0:
        return False
    raise ValueError("InexactError")


def _exp(x):  # Julia's exp overflows to Inf instead of raising
    with np.errstate(over="ignore"):
        return float(np.exp(x))


def _sign(x):
    return float(int(x > 0) - int(x < 0))


def evaluate(l, u, a):
    name = type(l).__name__
    s = l.scale
    if name == "QuadLoss":
        return s * (u - a) ** 2                                                   # :144
    if name == "L1Loss":
        return s * abs(u - a)                                                     # :158
    if name == "HuberLoss":                                                      # :173-175
        c = l.crossover
        return (abs(u - a) - c + c ** 2) * s if abs(u - a) > c else (u - a) ** 2 * s
    if name == "QuantileLoss":                                                   # :193-196
        diff = a - u
        return s * l.quantile * diff if diff > 0 else -s * (1 - l.quantile) * diff
    if name == "PeriodicLoss":                                                   # :216
        return s * (1 - math.cos((a - u) * (2 * math.pi) / l.T))
    if name == "PoissonLoss":                                                    # :237-239
        return s * (_exp(u) - a * u + (0 if a == 0 else a * (math.log(a) - 1)))
    if name == "OrdinalHingeLoss":                                               # :258-278
        if u > l.max - 1:
            n = min(math.floor(u), l.max - 1) - a
            loss = n * (n + 1) / 2 + (n + 1) * (u - l.max + 1)
        elif u > a:
            n = min(math.floor(u), l.max) - a
            loss = n * (n + 1) / 2 + (n + 1) * (u - math.floor(u))
        elif u > l.min + 1:
            n = a - max(math.ceil(u), l.min + 1)
            loss = n * (n + 1) / 2 + (n + 1) * (math.ceil(u) - u)
        else:
            n = a - max(math.ceil(u), l.min + 1)
            loss = n * (n + 1) / 2 + (n + 1) * (l.min + 1 - u)
        return s * loss
    if name == "LogisticLoss":                                                   # :304
        b = myBool(a)
        with np.errstate(over="ignore"):
            return float(s * np.log(1 + np.exp(-(2 * b - 1) * u)))
    if name == "WeightedHingeLoss":                                              # :326-332
        b = myBool(a)
        loss = s * max(1 - (2 * b - 1) * u, 0)
        if l.case_weight_ratio != 1.0 and b:
            loss *= l.case_weight_ratio
        return loss
    # vector-valued: u is a 1-D array, a an integer level (1-based)
    u = np.array(u, dtype=float)
    a = int(a)
    if name == "MultinomialLoss":                                                # :369-380
        sumexp = 0
        M = np.max(u) - u[a - 1]
        for j in range(len(u)):
            sumexp += math.exp(u[j] - u[a - 1] - M)
        return s * (math.log(sumexp) + M)
    if name == "OvALoss":                                                        # :424-430
        loss = 0
        for j in range(len(u)):
            loss += evaluate(l.bin_loss, u[j], a == j + 1)
        return s * loss
    if name == "BvSLoss":                                                        # :461-467
        loss = 0
        for j in range(len(u)):
            loss += evaluate(l.bin_loss, u[j], a > j + 1)
        return s * loss
    if name == "OrdisticLoss":                                                   # :499-505
        diffusquared = u[a - 1] ** 2 - u ** 2
        M = np.max(diffusquared)
        invlik = np.sum(np.exp(diffusquared - M))
        return s * (M + math.log(invlik))
    if name == "MultinomialOrdinalLoss":                                         # :581-590
        enforce_MNLOrdRules(u)
        if a == 1:
            return -s * math.log(math.exp(0) - math.exp(u[0]))
        elif a == l.max:
            return -s * u[a - 2]
        return -s * math.log(math.exp(u[a - 2]) - math.exp(u[a - 1]))
    raise NotImplementedError(name)


def enforce_MNLOrdRules(u, TOL=1e-3):  # :572-578
    u[0] = min(-TOL, u[0])
    for j in range(1, len(u)):
        u[j] = min(u[j], u[j - 1] - TOL)
    return u


def grad(l, u, a):
    name = type(l).__name__
    s = l.scale
    if name == "QuadLoss":
        return 2 * (u - a) * s                                                    # :146
    if name == "L1Loss":
        return _sign(u - a) * s                                                   # :160
    if name == "HuberLoss":                                                      # :177
        return _sign(u - a) * s if abs(u - a) > l.crossover else (u - a) * s
    if name == "QuantileLoss":                                                   # :198-201
        diff = a - u
        return -s * l.quantile if diff > 0 else s * (1 - l.quantile)
    if name == "PeriodicLoss":                                                   # :218
        return -s * ((2 * math.pi) / l.T) * math.sin((a - u) * (2 * math.pi) / l.T)
    if name == "PoissonLoss":
        return s * (_exp(u) - a)                                                  # :241
    if name == "OrdinalHingeLoss":                                               # :280-292
        if u > a:
            g = min(math.ceil(u), l.max) - a
        else:
            g = -(a - max(math.floor(u), l.min))
        return s * g
    if name == "LogisticLoss":                                                   # :306
        aa = 2 * myBool(a) - 1
        with np.errstate(over="ignore"):
            return float(-aa * s / (1 + np.exp(aa * u)))
    if name == "WeightedHingeLoss":                                              # :334-341
        b = myBool(a)
        an = 2 * b - 1
        g = 0 if an * u >= 1 else -an * s
        if l.case_weight_ratio != 1.0 and b:
            g *= l.case_weight_ratio
        return g
    u = np.array(u, dtype=float)
    a = int(a)
    if name == "MultinomialLoss":                                                # :382-398
        g = np.zeros(len(u))
        g[a - 1] = -1
        for j in range(len(u)):
            M = np.max(u) - u[j]
            sumexp = 0
            for jp in range(len(u)):
                sumexp += math.exp(u[jp] - u[j] - M)
            g[j] += math.exp(-M) / sumexp
        return s * g
    if name == "OvALoss":                                                        # :432-438
        return s * np.array([grad(l.bin_loss, u[j], a == j + 1) for j in range(len(u))])
    if name == "BvSLoss":                                                        # :469-475
        return s * np.array([grad(l.bin_loss, u[j], a > j + 1) for j in range(len(u))])
    if name == "OrdisticLoss":                                                   # :507-519
        g = np.zeros(len(u))
        g[a - 1] = 2 * u[a - 1]
        for j in range(len(u)):
            diffusquared = u[j] ** 2 - u ** 2
            M = np.max(diffusquared)
            invlik = np.sum(np.exp(diffusquared - M))
            g[j] -= 2 * u[j] * math.exp(-M) / invlik
        return s * g
    if name == "MultinomialOrdinalLoss":                                         # :592-608
        enforce_MNLOrdRules(u)
        g = np.zeros(len(u))
        if a == 1:
            g[0] = -math.exp(u[0]) / (math.exp(0) - math.exp(u[0]))
        elif a == l.max:
            g[a - 2] = 1
        else:
            g[a - 1] = -math.exp(u[a - 1]) / (math.exp(u[a - 2]) - math.exp(u[a - 1]))
            g[a - 2] = math.exp(u[a - 2]) / (math.exp(u[a - 2]) - math.exp(u[a - 1]))
        return -s * g
    raise NotImplementedError(name)


# ------------------------------------------------------------------------------------ regularizers
def _argmax(u):  # Julia argmax: first maximal element (linear index)
    return int(np.argmax(u))


def reg_evaluate(r, a):
    name = type(r).__name__
    a = np.asarray(a, dtype=float)
    if name == "ZeroReg":
        return 0                                                                  # :95
    if name == "QuadReg":
        return r.scale * np.sum(a ** 2)                                           # :58
    if name == "QuadConstraint":
        return math.inf if np.linalg.norm(a) > r.max_2norm + TOL else 0           # :74
    if name == "OneReg":
        return r.scale * np.sum(np.abs(a))                                        # :88
    if name == "NonNegConstraint":                                               # :105-112
        return math.inf if (a < 0).any() else 0
    if name == "NonNegOneReg":                                                   # :129-136
        return math.inf if (a < 0).any() else r.scale * np.sum(a)
    if name == "OneSparseConstraint":                                            # :239-253
        oneflag = False
        for ai in a.ravel(order="F"):
            if oneflag:
                if ai != 0:
                    return math.inf
            elif ai != 0:
                oneflag = True
        return 0
    if name == "KSparseConstraint":                                              # :261-276
        nonzcount = 0
        for ai in a.ravel(order="F"):
            if nonzcount == r.k:
                if ai != 0:
                    return math.inf
            elif ai != 0:
                nonzcount += 1
        return 0
    if name == "UnitOneSparseConstraint":                                        # :300-316
        oneflag = False
        for ai in a.ravel(order="F"):
            if ai == 0:
                continue
            elif ai == 1:
                if oneflag:
                    return math.inf
                oneflag = True
            else:
                return math.inf
        return 0
    if name == "SimplexConstraint":                                              # :338-346
        if abs(np.sum(a) - 1) > TOL:
            return math.inf
        return math.inf if (a < 0).any() else 0
    if name in ("OrdinalReg", "MNLOrdinalReg"):                                  # :378,405
        return reg_evaluate(r.r, a[:-1] if a.ndim == 1 else a[:-1, 0])
    if name == "lastentry1":                                                     # :171-172
        if a.ndim == 1:
            return reg_evaluate(r.r, a[:-1]) if a[-1] == 1 else math.inf
        return reg_evaluate(r.r, a[:-1, :]) if (a[-1, :] == 1).all() else math.inf
    if name == "lastentry_unpenalized":                                          # :184,187
        return reg_evaluate(r.r, a[:-1] if a.ndim == 1 else a[:-1, :])
    if name == "fixed_latent_features":                                          # :208
        n = len(r.y)
        return reg_evaluate(r.r, a[n:]) if (a[:n] == r.y).all() else math.inf
    if name == "fixed_last_latent_features":                                     # :230
        n = len(r.y)
        return reg_evaluate(r.r, a[:len(a) - n]) if (a[len(a) - n:] == r.y).all() else math.inf
    if name == "RemQuadReg":                                                     # :423
        return r.scale * np.sum((a - r.m) ** 2)
    raise NotImplementedError(name)


def prox(r, u, alpha):
    name = type(r).__name__
    u = np.array(u, dtype=float)
    if name == "ZeroReg":
        return u                                                                  # :93
    if name == "QuadReg":
        return 1 / (1 + 2 * alpha * r.scale) * u                                  # :56
    if name == "QuadConstraint":
        with np.errstate(divide="ignore", invalid="ignore"):
            return (r.max_2norm) / np.linalg.norm(u) * u                          # :72
    if name == "OneReg":                                                         # :83-86
        t = r.scale * alpha
        return np.maximum(u - t, 0) + np.minimum(u + t, 0)
    if name == "NonNegConstraint":
        return np.maximum(u, 0)                                                   # :103
    if name == "NonNegOneReg":
        return np.maximum(u - alpha, 0)                                           # :122
    if name == "OneSparseConstraint":                                            # :237
        flat = u.ravel(order="F")
        idx = _argmax(flat)
        v = np.zeros_like(flat)
        v[idx] = flat[idx]
        return v.reshape(u.shape, order="F")
    if name == "KSparseConstraint":                                              # :277-283
        flat = u.ravel(order="F")
        ids = np.argsort(-np.abs(flat), kind="stable")[:r.k]
        v = np.zeros_like(flat)
        v[ids] = flat[ids]
        return v.reshape(u.shape, order="F")
    if name == "UnitOneSparseConstraint":                                        # :297
        flat = u.ravel(order="F")
        v = np.zeros_like(flat)
        v[_argmax(flat)] = 1
        return v.reshape(u.shape, order="F")
    if name == "SimplexConstraint":                                              # :325-337
        flat = u.ravel(order="F")
        n = len(flat)
        y = np.sort(flat)[::-1]
        ysum = np.cumsum(y)
        t = (ysum[-1] - 1) / n
        for i in range(n - 1):
            if (ysum[i] - 1) / (i + 1) >= y[i + 1]:
                t = (ysum[i] - 1) / (i + 1)
                break
        return np.maximum(u - t, 0)
    if name in ("OrdinalReg", "MNLOrdinalReg"):                                  # :361-377 / :390-404
        u2 = u.reshape(-1, 1) if u.ndim == 1 else u.copy()
        um = np.mean(u2[:-1, :], axis=1)
        um = prox(r.r, um, alpha)
        u2[:-1, :] = um[:, None]
        if name == "MNLOrdinalReg":
            u2[-1, 0] = min(-1e-3, u2[-1, 0])
            for j in range(1, u2.shape[1]):
                u2[-1, j] = min(u2[-1, j], u2[-1, j - 1] - 1e-3)
        return u2[:, 0] if u.ndim == 1 else u2
    if name == "lastentry1":                                                     # :167,169
        if u.ndim == 1:
            return np.concatenate([prox(r.r, u[:-1], alpha), [1.0]])
        return np.vstack([prox(r.r, u[:-1, :], alpha), np.ones((1, u.shape[1]))])
    if name == "lastentry_unpenalized":                                          # :182,185
        if u.ndim == 1:
            return np.concatenate([prox(r.r, u[:-1], alpha), [u[-1]]])
        return np.vstack([prox(r.r, u[:-1, :], alpha), u[-1:, :]])
    if name == "fixed_latent_features":                                          # :203  [r.y; prox(r.r, u[(r.n+1):end], alpha)]
        n = len(r.y)
        return np.concatenate([r.y, prox(r.r, u[n:], alpha)])
    if name == "fixed_last_latent_features":                                     # :223  [prox(r.r, u[(r.n+1):end], alpha); r.y]  (sic)
        n = len(r.y)
        return np.concatenate([prox(r.r, u[n:], alpha), r.y])
    if name == "RemQuadReg":                                                     # :417-418
        return (u + 2 * alpha * r.scale * r.m) / (1 + 2 * alpha * r.scale)
    raise NotImplementedError(name)


# ----------------------------------------------------------------------------------- evaluate_fit.jl
def _yslice(ystart, f):
    return slice(int(ystart[f]), int(ystart[f + 1]))


def row_objective(glrm, i, x, Y, ystart):                                         # evaluate_fit.jl:24-38
    err = 0.0
    XY = x @ Y                                                                    # :29
    for j in glrm.observed_features[i]:
        sl = _yslice(ystart, j)
        u = float(XY[sl.start]) if glrm.losses[j].code < 10 else XY[sl].copy()
        err += evaluate(glrm.losses[j], u, glrm.A[i, j])                          # :31
    err += reg_evaluate(glrm.rx[i], x)                                            # :35
    return err


def col_objective(glrm, j, y, X, ystart):                                         # evaluate_fit.jl:39-55
    err = 0.0
    XY = X.T @ y                                                                  # :45  (m,) or (m, d_f)
    scalar = glrm.losses[j].code < 10
    for i in glrm.observed_examples[j]:                                           # :46-49 (+ losses.jl:623-650)
        u = XY[i] if y.ndim == 1 else XY[i, :]
        err += evaluate(glrm.losses[j], float(u) if scalar else u, glrm.A[i, j])
    err += reg_evaluate(glrm.ry[j], y)                                            # :52
    return err


def objective(glrm, X, Y, ystart, include_regularization=True):                   # evaluate_fit.jl:4-23
    XY = X.T @ Y
    err = 0.0
    for j in range(glrm.A.shape[1]):
        sl = _yslice(ystart, j)
        scalar = glrm.losses[j].code < 10
        for i in glrm.observed_examples[j]:
            u = XY[i, sl]
            err += evaluate(glrm.losses[j], float(u[0]) if scalar else u, glrm.A[i, j])
    if include_regularization:                                                    # calc_penalty :91-104
        for i in range(X.shape[1]):
            err += reg_evaluate(glrm.rx[i], X[:, i])
        for f in range(glrm.A.shape[1]):
            sl = _yslice(ystart, f)
            err += reg_evaluate(glrm.ry[f], Y[:, sl.start] if glrm.losses[f].code < 10 else Y[:, sl])
    return err


# --------------------------------------------------------------------------------------- proxgrad.jl
def fit_reference(glrm, params, X=None, Y=None):
    """Restatement of fit!(glrm, params) (proxgrad.jl:34-220).  Returns (X, Y, objective list,
    alpharow, alphacol).  Works on copies unless X/Y are passed (then mutates them like the reference)."""
    from lowrankmodels_b200 import get_yidxs

    A = glrm.A
    losses, rx, ry = glrm.losses, glrm.rx, glrm.ry
    X = glrm.X.copy() if X is None else X
    Y = glrm.Y.copy() if Y is None else Y
    k = glrm.k
    m, n = A.shape
    ystart = get_yidxs(losses)                                                    # :52
    d = int(ystart[-1])                                                           # :53
    XY = X.T @ Y                                                                  # :65-66
    alpharow = params.stepsize * np.ones(m)                                       # :69
    alphacol = params.stepsize * np.ones(n)                                       # :70
    scaled_abs_tol = params.abs_tol * sum(len(glrm.observed_features[i]) for i in range(m))  # :72
    ch = [objective(glrm, X, Y, ystart)]                                          # :76
    g = np.zeros(k)                                                               # :80
    G = np.zeros((k, d))                                                          # :82
    obj_by_row = np.zeros(m)
    obj_by_col = np.zeros(n)
    scalar = [l.code < 10 for l in losses]

    def yview(M, f):  # vf[f] = view(Y,:,yidxs[f])  (:96): a vector for scalar losses, a block otherwise
        sl = _yslice(ystart, f)
        return M[:, sl.start] if scalar[f] else M[:, sl]

    for i in range(1, params.max_iter + 1):                                       # :107
        if params.inner_iter_X > 1 or params.inner_iter_Y > 1:                    # :112-115
            alpharow[:] = params.stepsize
            alphacol[:] = params.stepsize
        for _inneri in range(params.inner_iter_X):                                # :117
            for e in range(m):                                                    # :118
                g[:] = 0.0                                                        # :119
                for f in glrm.observed_features[e]:                               # :122
                    sl = _yslice(ystart, f)
                    u = float(XY[e, sl.start]) if scalar[f] else XY[e, sl].copy()
                    curgrad = grad(losses[f], u, A[e, f])                         # :125
                    if np.isscalar(curgrad) or np.ndim(curgrad) == 0:
                        g += curgrad * Y[:, sl.start]                             # :127
                    else:
                        g += Y[:, sl] @ curgrad                                   # :130
                l = len(glrm.observed_features[e]) + 1                            # :134
                obj_by_row[e] = row_objective(glrm, e, X[:, e], Y, ystart)        # :135
                newx = X[:, e].copy()
                while alpharow[e] > params.min_stepsize:                          # :136
                    stepsize = alpharow[e] / l                                    # :137
                    newx = newx + (-stepsize) * g                                 # :140
                    newx = prox(rx[e], newx, stepsize)                            # :142
                    if row_objective(glrm, e, newx, Y, ystart) < obj_by_row[e]:   # :143
                        X[:, e] = newx                                            # :144
                        alpharow[e] *= 1.05                                       # :145
                        break
                    else:
                        newx = X[:, e].copy()                                     # :148
                        alpharow[e] *= .7                                         # :149
                        if alpharow[e] < params.min_stepsize:                     # :150-153
                            alpharow[e] = params.min_stepsize * 1.1
                            break
            XY = X.T @ Y                                                          # :157
        for _inneri in range(params.inner_iter_Y):                                # :160
            G[:] = 0.0                                                            # :161
            for f in range(n):                                                    # :162
                sl = _yslice(ystart, f)
                for e in glrm.observed_examples[f]:                               # :165
                    u = float(XY[e, sl.start]) if scalar[f] else XY[e, sl].copy()
                    curgrad = grad(losses[f], u, A[e, f])                         # :168
                    if np.isscalar(curgrad) or np.ndim(curgrad) == 0:
                        G[:, sl.start] += curgrad * X[:, e]                       # :170
                    else:
                        G[:, sl] += np.outer(X[:, e], curgrad)                    # :173
                l = len(glrm.observed_examples[f]) + 1                            # :177
                obj_by_col[f] = col_objective(glrm, f, yview(Y, f), X, ystart)    # :178
                newy = yview(Y, f).copy()
                gf = yview(G, f)
                while alphacol[f] > params.min_stepsize:                          # :179
                    stepsize = alphacol[f] / l                                    # :180
                    newy = newy + (-stepsize) * gf                                # :183
                    newy = prox(ry[f], newy, stepsize)                            # :185
                    new_obj_by_col = col_objective(glrm, f, newy, X, ystart)      # :186
                    if new_obj_by_col < obj_by_col[f]:                            # :187
                        if scalar[f]:
                            Y[:, sl.start] = newy                                 # :188
                        else:
                            Y[:, sl] = newy
                        alphacol[f] *= 1.05                                       # :189
                        obj_by_col[f] = new_obj_by_col                            # :190
                        break
                    else:
                        newy = yview(Y, f).copy()                                 # :193
                        alphacol[f] *= .7                                         # :194
                        if alphacol[f] < params.min_stepsize:                     # :195-198
                            alphacol[f] = params.min_stepsize * 1.1
                            break
            XY = X.T @ Y                                                          # :202
        obj = float(np.sum(obj_by_col))                                           # :205
        ch.append(obj)                                                            # :207
        obj_decrease = ch[-2] - obj                                               # :210
        if i > 10 and (obj_decrease < scaled_abs_tol or obj_decrease / obj < params.rel_tol):  # :211
            break
    return X, Y, ch, alpharow, alphacol


# ---------------------------------------------------------------------------------- sparse_proxgrad.jl
def fit_sparse_reference(glrm, params):
    """Restatement of fit!(glrm, params::SparseProxGradParams) (sparse_proxgrad.jl:21-130).  Returns
    (Xbest, Ybest, objective list, alpha)."""
    from lowrankmodels_b200 import get_yidxs

    A = glrm.A
    losses, rx, ry = glrm.losses, glrm.rx, glrm.ry
    Xb, Yb = glrm.X.copy(), glrm.Y.copy()                                         # glrm.X / glrm.Y: best model yet
    X, Y = Xb.copy(), Yb.copy()                                                   # :33 working variables
    k = glrm.k
    m, n = A.shape
    ystart = get_yidxs(losses)
    alpha = params.stepsize                                                       # :44
    tol = params.abs_tol * sum(len(glrm.observed_features[i]) for i in range(m))  # :46
    ch = [objective(glrm, Xb, Yb, ystart)]                                        # :50
    steps_in_a_row = 0
    for i in range(1, params.max_iter + 1):                                       # :60
        for _ in range(params.inner_iter):                                        # :62
            for e in range(m):                                                    # :63
                g = np.zeros(k)                                                   # :64
                for f in glrm.observed_features[e]:                               # :67-72
                    g += grad(losses[f], float(X[:, e] @ Y[:, f]), A[e, f]) * Y[:, f]
                l = len(glrm.observed_features[e]) + 1                            # :74
                g *= -alpha / l                                                   # :75
                X[:, e] = X[:, e] + g                                             # :77
                X[:, e] = prox(rx[e], X[:, e], alpha / l)                         # :79
        for _ in range(params.inner_iter):                                        # :83
            for f in range(n):                                                    # :84
                g = np.zeros(k)                                                   # :85
                for e in glrm.observed_examples[f]:                               # :88-92
                    g += grad(losses[f], float(X[:, e] @ Y[:, f]), A[e, f]) * X[:, e]
                l = len(glrm.observed_examples[f]) + 1                            # :94
                g *= -alpha / l                                                   # :95
                Y[:, f] = Y[:, f] + g                                             # :97
                Y[:, f] = prox(ry[f], Y[:, f], alpha / l)                         # :99
        obj = objective(glrm, X, Y, ystart)                                       # :102
        if obj < ch[-1]:                                                          # :104
            ch.append(obj)                                                        # :106
            Xb[...] = X                                                           # :107
            Yb[...] = Y
            alpha = alpha * 1.05                                                  # :108
            steps_in_a_row = max(1, steps_in_a_row + 1)                           # :109
        else:
            alpha = alpha / max(1.5, -steps_in_a_row)                             # :113
            X[...] = Xb                                                           # :115
            Y[...] = Yb
            steps_in_a_row = min(0, steps_in_a_row - 1)                           # :116
        if (i > 10 and (steps_in_a_row > 3 and ch[-2] - obj < tol)) or alpha <= params.min_stepsize:   # :119
            break
    ch.append(ch[-1])                                                             # :126
    return Xb, Yb, ch, alpha
