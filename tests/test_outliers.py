"""Host-side outlier statistic of autoSVD (bigsnpr_b200/outliers.py): properties of the restated bigutilsr functions.
bigutilsr is un-vendored and R is absent, so these are property checks of the published definitions, not parity pins."""
import numpy as np
import pytest

from bigsnpr_b200 import outliers as O


def test_medcouple_equals_the_kernel_median():
    rng = np.random.default_rng(3)
    for n, gen in ((51, rng.gamma), (200, rng.gamma), (333, lambda a, size: -rng.gamma(a, size=size))):
        x = gen(2.0, size=n)
        m = np.median(x)
        zp, zm = x[x > m], x[x < m]
        H = ((zp[:, None] - m) - (m - zm[None, :])) / (zp[:, None] - zm[None, :])
        assert abs(O.medcouple(x) - np.median(H)) < 1e-12
    assert abs(O.medcouple(rng.normal(size=100001))) < 0.02          # symmetric -> ~0
    assert O.medcouple(rng.exponential(size=20001)) > 0.25            # right-skewed -> positive
    assert O.medcouple(np.ones(10)) == 0.0


def test_tau_scale_and_ogk_on_contaminated_gaussians():
    rng = np.random.default_rng(4)
    x = 3.0 * rng.normal(size=200000) + 7.0
    mu, s = O.scale_tau2(x, mu_too=True)
    assert abs(mu - 7.0) < 0.05 and abs(s - 3.0) < 0.03               # consistent at the normal model
    x[:10000] = 1e4                                                    # 5 % gross outliers barely move it
    assert abs(O.scale_tau2(x) - 3.0) < 0.5                            # bounded rho: +11 % where the sd would be 700x
    C = np.array([[1, 0.6, 0], [0.6, 1, 0.3], [0, 0.3, 1.0]])
    X = rng.multivariate_normal([1, 2, 3], C, size=20000)
    X[:300] += 20
    og = O.covrob_ogk(X)
    assert np.allclose(og["wcenter"], [1, 2, 3], atol=0.05)
    corr = og["wcov"] / np.sqrt(np.outer(np.diag(og["wcov"]), np.diag(og["wcov"])))
    assert np.allclose(corr, C, atol=0.03)                             # shape recovered despite the contamination
    d = O.dist_ogk(X)
    assert d[:300].min() > 50 * np.median(d[300:])
    # affine equivariance of the distances up to the estimator's own (non-equivariant) coordinate choice: scaling columns
    d2 = O.dist_ogk(X * np.array([10.0, 0.1, 3.0]))
    assert np.allclose(d2, d, rtol=1e-6)


def test_rollmean_weights_and_edges():
    x = np.arange(100.0)
    r = O.rollmean(x, 5)
    assert np.allclose(r[5:-5], x[5:-5])                               # symmetric weights reproduce a linear trend
    assert x[0] < r[0] < x[5] and x[-6] < r[-1] < x[-1]                # truncated, renormalised ends
    assert np.array_equal(O.rollmean(x, 0), x)
    with pytest.raises(ValueError, match="too large"):
        O.rollmean(np.arange(5.0), 3)
    spike = np.zeros(101)
    spike[50] = 1.0
    w = O.rollmean(spike, 10)
    assert abs(w.sum() - 1) < 1e-12 and w[50] == w.max() and w[39] == 0 and w[40] > 0  # 21 Gaussian weights


def test_tukey_fence_controls_the_family_error_and_flags_a_region():
    rng = np.random.default_rng(5)
    hits = 0
    for _ in range(40):
        S = np.abs(rng.normal(size=20000))
        hits += int((S > O.tukey_mc_up(S, alpha=0.05)).any())
    assert hits <= 12                                                  # skew-adjusted fence: no flood of false positives
    v = rng.normal(size=(30000, 5)) / np.sqrt(30000)
    v[12000:12400, 2] += 6 / np.sqrt(30000)                            # a long-range-LD-like block loading on one PC
    idx = O.autosvd_outlier_fun(50, 0.05)(v, np.ones(30000, dtype=int))
    assert idx.size > 300 and idx.min() >= 11900 and idx.max() <= 12500
