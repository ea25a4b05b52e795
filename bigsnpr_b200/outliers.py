"""Host-side outlier statistic of bed_autoSVD / snp_autoSVD (R/autoSVD.R:142-148, :295-302):

    S.col  <- sqrt(bigutilsr::dist_ogk(obj.svd$v))                 # robust Mahalanobis distance of the loadings
    S2.col <- bigutilsr::rollmean(S.col[ind], roll.size)           # per chromosome
    thr    <- bigutilsr::tukey_mc_up(S2.col, alpha = alpha.tukey)
    ind.col.excl <- which(S2.col > thr)

``bigutilsr`` is an un-vendored dependency of the reference (DESCRIPTION:30): its source is not in
the reference checkout and R is not installed, so the functions below are RESTATED FROM THE PUBLISHED ALGORITHMS that
package documents it implements, on the small (m x k) matrices the GPU engine returns -- pure NumPy / SciPy, nothing here
is on the GPU path:

* ``dist_ogk``: orthogonalised Gnanadesikan-Kettenring estimator (Maronna & Zamar 2002) with the tau-scale of Yohai & Zamar
  (c1 = 4.5, c2 = 3, consistency factor), two iterations, hard-rejection re-weighting at the chi-square 0.9 quantile -- the
  ``robustbase::covOGK(sigmamu = scaleTau2, weight.fn = hard.rejection)`` recipe -- then squared Mahalanobis distances to
  the re-weighted centre / covariance.
* ``rollmean``: Gaussian-weighted moving average of half-width ``size``; weights dnorm on an even grid between the normal
  quantiles of the first and last plotting position (``ppoints``), truncated and renormalised at the ends.
* ``tukey_mc_up``: upper fence of the skewness-adjusted boxplot (Hubert & Vandervieren 2008), Q3 + coef * IQR * exp(3 MC)
  (exp(4 MC) for MC < 0) with the medcouple MC, the coefficient chosen so that a Gaussian sample of this size exceeds the
  fence with probability ``alpha`` overall (Sidak correction).

**Parity unpinned**: the reference's tests pin this step only structurally (tests/testthat/test-2-autoSVD.R); with no copy
of bigutilsr to run, the numbers below are not checked against it.  The engine steps around it (counts, clumping, SVD) are.
"""
from __future__ import annotations

import numpy as np


def _erho(b):
    from scipy.stats import norm

    return 2 * ((1 - b * b) * norm.cdf(b) - b * norm.pdf(b) + b * b) - 1


def scale_tau2(x, c1=4.5, c2=3.0, mu_too=False):
    """tau-estimate of scale (and location) of Yohai & Zamar, as in robustbase::scaleTau2(consistency = TRUE)."""
    from scipy.stats import norm

    x = np.asarray(x, dtype=np.float64)
    n = x.size
    medx = np.median(x)
    ax = np.abs(x - medx)
    sigma0 = np.median(ax)
    if sigma0 <= 0:
        return (medx, 0.0) if mu_too else 0.0
    w = 1 - (ax / (sigma0 * c1)) ** 2
    w = ((np.abs(w) + w) / 2) ** 2
    mu = np.sum(x * w) / np.sum(w)
    r = (x - mu) / sigma0
    rho = np.minimum(r * r, c2 * c2)
    n_es2 = n * _erho(c2 * norm.ppf(0.75))
    s = sigma0 * np.sqrt(np.sum(rho) / n_es2)
    return (mu, s) if mu_too else s


def covrob_ogk(U, niter=2, beta=0.9):
    """OGK location / scatter with hard-rejection re-weighting; returns dict(center, cov, wcenter, wcov, weights)."""
    from scipy.stats import chi2

    X = np.asarray(U, dtype=np.float64)
    n, p = X.shape
    Z = X.copy()
    A = []
    for _ in range(niter):
        d = np.array([scale_tau2(Z[:, j]) for j in range(p)])
        d[d == 0] = 1.0
        Z = Z / d
        R = np.eye(p)
        for i in range(1, p):
            for j in range(i):
                sp, sm = scale_tau2(Z[:, i] + Z[:, j]), scale_tau2(Z[:, i] - Z[:, j])
                R[i, j] = R[j, i] = (sp * sp - sm * sm) / 4
        _, E = np.linalg.eigh(R)
        E = E[:, ::-1]  # eigen(): decreasing order
        A.append(d[:, None] * E)
        Z = Z @ E
    ms = np.array([scale_tau2(Z[:, j], mu_too=True) for j in range(p)])
    center, sg = ms[:, 0], ms[:, 1]
    sg[sg == 0] = 1.0
    Zs = (Z - center) / sg
    dist = np.sum(Zs * Zs, axis=1)
    cov = np.diag(sg * sg)
    for Ai in reversed(A):
        cov = Ai @ cov @ Ai.T
        center = Ai @ center
    d0 = np.median(dist) * chi2.ppf(beta, p) / chi2.ppf(0.5, p)
    w = (dist <= d0).astype(np.float64)
    sw = w.sum()
    wcenter = (X * w[:, None]).sum(axis=0) / sw
    Zw = (X - wcenter) * np.sqrt(w)[:, None]
    wcov = Zw.T @ Zw / sw
    return {"center": center, "cov": cov, "wcenter": wcenter, "wcov": wcov, "weights": w, "distances": dist}


def dist_ogk(U, niter=2):
    """Squared robust Mahalanobis distances of the rows of U (bigutilsr::dist_ogk)."""
    X = np.asarray(U, dtype=np.float64)
    ogk = covrob_ogk(X, niter=niter)
    D = X - ogk["wcenter"]
    return np.einsum("ij,ij->i", D @ np.linalg.inv(ogk["wcov"]), D)


def rollmean(x, size):
    """Gaussian-weighted rolling mean of half-width `size`, ends renormalised (bigutilsr::rollmean)."""
    from scipy.stats import norm

    x = np.asarray(x, dtype=np.float64)
    if size == 0:
        return x.copy()
    half = int(np.floor(size))
    ln = 2 * half + 1
    if ln > x.size:
        raise ValueError("Parameter 'size' is too large.")
    a = 3.0 / 8.0 if ln <= 10 else 0.5  # ppoints(): (1:n - a) / (n + 1 - 2a)
    lo, hi = norm.ppf((1 - a) / (ln + 1 - 2 * a)), norm.ppf((ln - a) / (ln + 1 - 2 * a))
    w = norm.pdf(np.linspace(lo, hi, ln))
    num = np.convolve(x, w[::-1], mode="same")
    den = np.convolve(np.ones_like(x), w[::-1], mode="same")
    return num / den


def medcouple(x, eps=1e-14):
    """Medcouple (Brys, Hubert & Struyf 2004): median of h(xi, xj) = ((xi - m) - (m - xj)) / (xi - xj) over xi >= m >= xj.
    Found by bisection on its value: for a trial t the number of kernel values <= t is a sum of searchsorted counts (h is
    monotone in both arguments), O(n log n) per trial -- no n^2 kernel matrix."""
    x = np.sort(np.asarray(x, dtype=np.float64))
    x = x[np.isfinite(x)]
    n = x.size
    if n < 3:
        return 0.0
    m = np.median(x)
    scale = 2 * max(abs(x[0] - m), abs(x[-1] - m))
    if scale == 0:
        return 0.0
    zp = (x[x > m] - m) / scale   # > 0, ascending
    zm = (x[x < m] - m) / scale   # < 0, ascending  (observations equal to the median carry no skewness information)
    if zp.size == 0 or zm.size == 0:
        return 0.0
    total = zp.size * zm.size
    target = (total + 1) // 2  # lower median rank; averaged with the next for even counts below

    def count_le(t):  # number of pairs with h <= t  <=>  zj <= zi * (t - 1) / (1 + t)
        if t >= 1:
            return total
        if t <= -1:
            return 0
        return int(np.searchsorted(zm, zp * (t - 1) / (1 + t), side="right").sum())

    def kth(k):
        lo, hi = -1.0, 1.0
        while hi - lo > eps:
            mid = 0.5 * (lo + hi)
            if count_le(mid) >= k:
                hi = mid
            else:
                lo = mid
        return hi

    if total % 2:
        return kth(target)
    return 0.5 * (kth(total // 2) + kth(total // 2 + 1))


def tukey_mc_up(x, coef=None, alpha=0.05):
    """Upper fence of the adjusted boxplot with a multiple-testing-aware coefficient (bigutilsr::tukey_mc_up)."""
    from scipy.stats import norm

    x = np.asarray(x, dtype=np.float64)
    x = x[~np.isnan(x)]
    if coef is None:
        alpha1 = 1 - (1 - alpha) ** (1.0 / x.size)  # Sidak: per-observation level
        # for N(0, 1): Q3 = 0.6745, IQR = 1.349; the fence Q3 + coef * IQR sits at the (1 - alpha1) quantile
        coef = (norm.isf(alpha1) - norm.ppf(0.75)) / (2 * norm.ppf(0.75))
    q1, q3 = np.quantile(x, [0.25, 0.75])
    mc = medcouple(x)
    return q3 + coef * (q3 - q1) * np.exp((3.0 if mc >= 0 else 4.0) * mc)


def autosvd_outlier_fun(roll_size=50, alpha_tukey=0.05):
    """The reference's default detector as the callable _auto_svd expects: (v, chromosomes of the kept variants) -> 0-based
    indices of the outlier variants (R/autoSVD.R:295-302)."""

    def fun(v, infos_chr_keep):
        S = np.sqrt(dist_ogk(np.asarray(v, dtype=np.float64)))
        S2 = np.full(S.size, np.nan)
        chrs = np.asarray(infos_chr_keep)
        for ch in np.unique(chrs):
            ind = np.nonzero(chrs == ch)[0]
            S2[ind] = rollmean(S[ind], roll_size)
        thr = tukey_mc_up(S2, alpha=alpha_tukey)
        return np.nonzero(S2 > thr)[0]

    return fun
