"""Host-side restatement of the R set-up that precedes the native calls of seqFitNullGLMM_SPA.

The reference does this part in R (R/saige_main.r:356-387 QR transform, :480-497 and :535-583
glm() start values, SPAtest's null-model object).  There is no R in this image, so the Python
mirror of the R driver (saigegds_b200.api.seqFitNullGLMM_SPA) uses these numpy versions.  They are
not on the GPU hot path: they run once per fit on p <= ~20 columns.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np
from scipy.special import ndtri
from scipy.stats import rankdata

DBL_EPSILON = float(np.finfo(np.float64).eps)


@dataclass
class Fit0:
    """The fields of R's glm object that the native code reads (saige_fitnull.cpp:968-984)."""
    y: np.ndarray
    coefficients: np.ndarray
    linear_predictors: np.ndarray
    fitted_values: np.ndarray
    family: str                      # "binomial" or "gaussian"
    offset: np.ndarray | None = None
    X: np.ndarray | None = None      # model.matrix(fit0)
    residuals: np.ndarray | None = None


@dataclass
class ObjNoK:
    """SPAtest:::ScoreTest_wSaddleApprox_NULL_Model output (binary) / R/saige_main.r:560-570 (quant)."""
    y: np.ndarray
    mu: np.ndarray
    res: np.ndarray
    V: np.ndarray
    X1: np.ndarray
    XV: np.ndarray
    XXVX_inv: np.ndarray


def parse_formula(formula: str):
    """'y ~ x1 + x2' -> ('y', ['x1', 'x2'], intercept).  Additive numeric terms only."""
    lhs, rhs = formula.split("~")
    terms = [t.strip() for t in re.split(r"\+", rhs) if t.strip()]
    intercept = True
    out = []
    for t in terms:
        if t in ("-1", "0"):
            intercept = False
        elif t == "1":
            intercept = True
        elif t.endswith("-1") or t.endswith("- 1"):
            intercept = False
            out.append(t[:t.rfind("-")].strip())
        else:
            out.append(t)
    return lhs.strip(), out, intercept


def model_matrix(data: dict, terms, intercept=True) -> np.ndarray:
    cols = [np.ones(len(next(iter(data.values()))))] if intercept else []
    cols += [np.asarray(data[t], dtype=np.float64) for t in terms]
    return np.column_stack(cols)


def _lstsq(X, y):
    return np.linalg.lstsq(X, y, rcond=None)[0]


def linkinv_logit(eta):
    """R family.c logit_linkinv (thresholds +-30, DBL_EPSILON clamps)."""
    eta = np.asarray(eta, dtype=np.float64)
    tmp = np.where(eta < -30, DBL_EPSILON, np.where(eta > 30, 1.0 / DBL_EPSILON, np.exp(np.clip(eta, -700, 700))))
    return tmp / (1.0 + tmp)


def mu_eta_logit(eta):
    """R family.c logit_mu_eta."""
    eta = np.asarray(eta, dtype=np.float64)
    e = np.exp(np.clip(eta, -700, 700))
    return np.where(np.abs(eta) > 30, DBL_EPSILON, e / ((1.0 + e) * (1.0 + e)))


def glm_binomial(X: np.ndarray, y: np.ndarray, epsilon=1e-8, maxit=25) -> Fit0:
    """R's glm.fit for family=binomial(): IRLS from mustart=(y+0.5)/2, deviance stopping rule."""
    y = np.asarray(y, dtype=np.float64)
    mu = (y + 0.5) / 2.0
    eta = np.log(mu / (1.0 - mu))
    devold = _binom_dev(y, mu)
    coef = np.zeros(X.shape[1])
    for _ in range(maxit):
        me = mu_eta_logit(eta)
        z = eta + (y - mu) / me
        w = np.sqrt(me * me / (mu * (1.0 - mu)))
        coef = _lstsq(X * w[:, None], z * w)
        eta = X @ coef
        mu = linkinv_logit(eta)
        dev = _binom_dev(y, mu)
        if abs(dev - devold) / (abs(dev) + 0.1) < epsilon:
            break
        devold = dev
    return Fit0(y=y, coefficients=coef, linear_predictors=eta, fitted_values=mu, family="binomial",
                X=X, residuals=(y - mu) / mu_eta_logit(eta))


def _binom_dev(y, mu):
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.where(y > 0, y * np.log(y / mu), 0.0)
        b = np.where(y < 1, (1 - y) * np.log((1 - y) / (1 - mu)), 0.0)
    return 2.0 * float(np.sum(a + b))


def glm_gaussian(X: np.ndarray, y: np.ndarray) -> Fit0:
    y = np.asarray(y, dtype=np.float64)
    coef = _lstsq(X, y)
    eta = X @ coef
    return Fit0(y=y, coefficients=coef, linear_predictors=eta, fitted_values=eta.copy(), family="gaussian",
                X=X, residuals=y - eta)


def independent_columns(X: np.ndarray, tol: float = 1e-7) -> np.ndarray:
    """Indices of the design-matrix columns R's `lm(y ~ X - 1)` would keep: a column that is linearly dependent on the columns
    before it gets an NA coefficient (pivoted Householder QR of `lm.fit`, tolerance 1e-7) and is dropped before the QR
    transform (R/saige_main.r:362-376: `X <- X[, !is.na(fit$coefficients)]`)."""
    X = np.asarray(X, dtype=np.float64)
    keep, basis = [], []
    for j in range(X.shape[1]):
        v = X[:, j].copy()
        nrm0 = np.linalg.norm(v)
        for q in basis:                      # modified Gram-Schmidt against the kept columns
            v -= (q @ v) * q
        nrm = np.linalg.norm(v)
        if nrm0 > 0 and nrm > tol * nrm0:
            keep.append(j)
            basis.append(v / nrm)
    return np.asarray(keep, dtype=np.int64)


def qr_transform(X: np.ndarray):
    """R/saige_main.r:378-380: X_new = qr.Q(qr(X)) * sqrt(n), X_qrr = qr.R(qr(X))."""
    Q, R = np.linalg.qr(X)
    return Q * np.sqrt(X.shape[0]), R


def rank_norm(x: np.ndarray) -> np.ndarray:
    """.rank_norm, R/saige_main.r:64: qnorm((rank(x) - 0.5)/length(x)) (average ties)."""
    return ndtri((rankdata(x, method="average") - 0.5) / len(x))


def sd(x):
    return float(np.std(np.asarray(x, dtype=np.float64), ddof=1))


def _get_X1(X1: np.ndarray) -> np.ndarray:
    """SPAtest:::ScoreTest_wSaddleApprox_Get_X1: drop a duplicated 2nd column, reduce to full rank."""
    if X1.shape[1] >= 2 and np.sum(np.abs(X1[:, 0] - X1[:, 1])) == 0:
        X1 = np.delete(X1, 1, axis=1)
    rank = np.linalg.matrix_rank(X1)
    if rank < X1.shape[1]:
        u = np.linalg.svd(X1, full_matrices=False)[0]
        X1 = u[:, :rank]
    return X1


def null_model_binary(X: np.ndarray, fit0: Fit0) -> ObjNoK:
    """SPAtest:::ScoreTest_wSaddleApprox_NULL_Model (called at R/saige_main.r:488)."""
    X1 = _get_X1(X)
    mu = fit0.fitted_values
    V = mu * (1.0 - mu)
    XV = (X1 * V[:, None]).T
    XVX_inv = np.linalg.inv(X1.T @ (X1 * V[:, None]))
    return ObjNoK(y=fit0.y, mu=mu, res=fit0.y - mu, V=V, X1=X1, XV=XV, XXVX_inv=X1 @ XVX_inv)


def null_model_quant(X: np.ndarray, fit0: Fit0) -> ObjNoK:
    """R/saige_main.r:560-570."""
    X1 = _get_X1(X)
    mu = fit0.fitted_values
    return ObjNoK(y=fit0.y, mu=mu, res=fit0.y - mu, V=np.ones(len(mu)), X1=X1, XV=X1.T.copy(),
                  XXVX_inv=X1 @ np.linalg.inv(X1.T @ X1))


def initial_tau_binary(tau_init=(0.0, 0.0)):
    """R/saige_main.r:491-497."""
    tau = np.array([1.0, 0.0])
    ti = np.nan_to_num(np.asarray(tau_init, dtype=np.float64))
    ti[ti < 0] = 0
    tau[1] = 0.5 if ti[1] == 0 else ti[1]
    return tau


def initial_tau_quant(fit0: Fit0, tau_init=(0.0, 0.0)):
    """R/saige_main.r:573-583: tau = var(Y) * tau / sum(tau), Y the working response of the identity link."""
    ti = np.nan_to_num(np.asarray(tau_init, dtype=np.float64))
    ti[ti < 0] = 0
    tau = ti.copy()
    if tau.sum() == 0:
        tau = np.array([0.5, 0.5])
    offset = np.zeros(len(fit0.y)) if fit0.offset is None else fit0.offset
    Y = fit0.linear_predictors - offset + (fit0.y - fit0.fitted_values)
    return float(np.var(Y, ddof=1)) * tau / tau.sum()
