"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's one-component von Mises-Fisher-Nakagami model
(mixturemodel.py:153-239; era/EMvMFNM.py: maximization 169-199, logvMFpdf 206-224, lognakagamipdf 229-236,
logbesseli 239-252, vMFNM_sample 274-285, vsamp 324-378, house 401-429).  Pinned against the live reference by
tests/test_vmfn.py::test_oracle_matches_reference_vectors (tests/golden/vmfn_vectors.npz).  Nothing in
rareeventestimation_b200 imports this."""
import numpy as np
from scipy.special import gammaln, iv


def fit(X, w=None):
    """Parameters (mu (d,), kappa, m, omega) of the K=1 fit: the EM's responsibilities stay 1, so its fixed point is
    one M-step with the (normalised or not) sample weights (EMvMFNM.py:60-75, 139-145, 169-199)."""
    n, d = X.shape
    w = np.full(n, 1.0 / n) if w is None else np.asarray(w, dtype=float)
    R = np.sqrt(np.sum(X * X, axis=1))
    Xn = X / R[:, None]
    nk = w.sum()
    mu_un = Xn.T @ w
    norm_mu = np.sqrt(np.sum(mu_un * mu_un))
    mu = mu_un / norm_mu
    xi = min(norm_mu / nk, 0.95)
    kappa = (xi * d - xi**3) / (1 - xi**2)
    omega = (w @ (R * R)) / nk
    mu4 = (w @ R**4) / nk
    m = omega**2 / (mu4 - omega**2)
    if m < 0 or m > 20 * d:
        m = d / 2
    return mu, kappa, m, omega


def log_bessel_i(nu, x):
    if nu == 0:
        return np.log(iv(nu, x))
    frac = x / nu
    square = 1 + frac**2
    root = np.sqrt(square)
    eta = root + np.log(frac) - np.log(1 + root)
    return -np.log(np.sqrt(2 * np.pi * nu)) + nu * eta - 0.25 * np.log(square)


def logpdf(X, mu, kappa, m, omega):
    n, d = X.shape
    R = np.sqrt(np.sum(X * X, axis=1))
    Xn = X / R[:, None]
    if kappa == 0:
        vmf = -(np.log(d) + np.log(np.pi ** (d / 2)) - gammaln(d / 2 + 1)) * np.ones(n)
    else:
        c = (d / 2 - 1) * np.log(kappa) - (d / 2) * np.log(2 * np.pi) - log_bessel_i(d / 2 - 1, kappa)
        vmf = kappa * (Xn @ mu) + c
    nak = np.log(2) + m * (np.log(m) - np.log(omega) - R * R / omega) + np.log(R) * (2 * m - 1) - gammaln(m)
    return vmf + nak - (d - 1) * np.log(R)


def house(x):
    n = len(x)
    s = x[: n - 1] @ x[: n - 1]
    v = np.concatenate([x[: n - 1], [1.0]])
    if s == 0:
        return v, 0.0
    nrm = np.sqrt(x[n - 1] ** 2 + s)
    v[n - 1] = x[n - 1] - nrm if x[n - 1] <= 0 else -s / (x[n - 1] + nrm)
    b = 2 * v[n - 1] ** 2 / (s + v[n - 1] ** 2)
    return v / v[n - 1], b


def variates(rng, n, d, kappa, m, omega):
    """The reference's draws for n samples, in its order: all radii first, then per sample the accept loop
    (beta, uniform) and the d-1 normals of the complement direction."""
    r = np.sqrt(rng.gamma(m, scale=omega / m, size=(n, 1))).reshape(-1)
    t1 = np.sqrt(4 * kappa * kappa + (d - 1) ** 2)
    b = (-2 * kappa + t1) / (d - 1)
    x0 = (1 - b) / (1 + b)
    mh = (d - 1) / 2
    c = kappa * x0 + (d - 1) * np.log(1 - x0 * x0)
    w, V = np.empty(n), np.empty((n, d - 1))
    for i in range(n):
        t, u = -1000, 1
        while t < np.log(u):
            z = rng.beta(mh, mh)
            u = rng.uniform()
            wi = (1 - (1 + b) * z) / (1 - (1 - b) * z)
            t = kappa * wi + (d - 1) * np.log(1 - x0 * wi) - c
        w[i] = wi
        V[i] = rng.normal(size=(1, d - 1))
    return r, w, V


def assemble(r, w, V, mu):
    """Samples from the primitive variates (EMvMFNM.py:349-367, 285)."""
    n, dm1 = V.shape
    X = np.empty((n, dm1 + 1))
    X[:, :dm1] = np.sqrt(1 - w * w)[:, None] * V / np.linalg.norm(V, axis=1, keepdims=True)
    X[:, dm1] = w
    v, b = house(np.asarray(mu, dtype=float))
    X = X - b * np.outer(X @ v, v)
    return r[:, None] * X


def sample(rng, n, mu, kappa, m, omega):
    r, w, V = variates(rng, n, len(mu), kappa, m, omega)
    return assemble(r, w, V, mu)
