"""GPU tests of the split DMMA path (ree_split.cu; ree_fused_path(d) == 2, 56 <= d <= 151): the transform-only step,
the one-pass double Gram and the transform-only Mahalanobis kernel against the fused sweeps (same C ABI, sums != NULL)
and against NumPy restatements of the reference formulas (solver.py:629-636, 665-674, 683-686)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from rareeventestimation_b200.engine import Engine

    return Engine.get()


def _setup(eng, n, d, seed):
    from rareeventestimation_b200 import _lib
    from rareeventestimation_b200.engine import DeviceLsf

    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)) * 1.1 + 0.3
    A = rng.standard_normal((d, d)) / np.sqrt(d)
    F = np.linalg.cholesky(A @ A.T + 0.5 * np.eye(d))
    src = eng.upload(X)
    lsf = DeviceLsf(_lib.LSF_AFFINE, d, None, 3.5)
    m_noise = rng.standard_normal(d) * 0.1
    return rng, X, F, src, lsf, m_noise


@pytest.mark.parametrize("n,d", [(1, 100), (7, 56), (96, 100), (1001, 100), (4099, 64), (3000, 103), (1000, 150),
                                 (777, 120), (65, 151), (70001, 100), (50003, 59)])
def test_split_step_matches_numpy_and_fused(eng, n, d):
    assert eng.path(d) == 2
    rng, X, F, src, lsf, m_noise = _setup(eng, n, d, n + d)
    t = 0.6
    Z = eng.philox_normal(n, d, 77, 5)
    Zh = eng.download(Z)
    want = t * X + (m_noise + Zh @ F.T)
    Fd, md = eng.to_device(F), eng.to_device(m_noise)
    slen = d + d * d + 2
    # split, Philox in registers (triangular factor)
    a = eng.alloc_ensemble(n, d)
    eng.cbree_step(src, a, t, md, Fd, lsf, None, None, seed=77, draw=5)
    got = eng.download(a)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-13 * np.abs(want).max())
    np.testing.assert_allclose(a.g[:n].cpu().numpy(), 3.5 - got.sum(axis=1) / np.sqrt(d), rtol=0, atol=1e-13)
    np.testing.assert_allclose(a.e[:n].cpu().numpy(), 0.5 * (got**2).sum(axis=1) + 0.5 * d * np.log(2 * np.pi),
                               rtol=1e-14)
    # split, injected noise (dense factor) and the fused sweep with the same noise: same X' to rounding
    b, c = eng.alloc_ensemble(n, d), eng.alloc_ensemble(n, d)
    eng.cbree_step(src, b, t, md, Fd, lsf, None, None, z=Z)
    sums = eng.empty(slen)
    eng.cbree_step(src, c, t, md, Fd, lsf, None, sums, z=Z)
    np.testing.assert_allclose(eng.download(b), got, rtol=0, atol=2e-13 * np.abs(want).max())
    np.testing.assert_allclose(eng.download(b), eng.download(c), rtol=0, atol=1e-15 * np.abs(want).max())
    np.testing.assert_allclose(b.g[:n].cpu().numpy(), c.g[:n].cpu().numpy(), rtol=0, atol=1e-14)
    # unweighted one-pass sums == fused sweep's sums
    shift = eng.to_device(want.mean(axis=0) + 0.01)
    sums_f = eng.empty(slen)
    eng.cbree_step(src, c, t, md, Fd, lsf, shift, sums_f, z=Z)
    su, _ = eng.cbree_moments(c, None, None, shift, eng.empty(2 * slen), weighted=False)
    sf, ss = sums_f.cpu().numpy(), su.cpu().numpy()
    np.testing.assert_allclose(ss[:d], sf[:d], rtol=0, atol=1e-12 * n)
    np.testing.assert_allclose(ss[d:d + d * d], sf[d:d + d * d], rtol=0, atol=1e-12 * np.abs(sf[d:d + d * d]).max())
    assert ss[d + d * d] == sf[d + d * d] == np.sum(c.g[:n].cpu().numpy() <= 0) and ss[d + d * d + 1] == n


@pytest.mark.parametrize("n,d", [(5, 100), (1500, 100), (2500, 72), (1200, 150), (900, 111)])
def test_split_moments_and_maha_match_numpy(eng, n, d):
    rng = np.random.default_rng(n * d)
    X = rng.standard_normal((n, d)) @ (np.eye(d) + 0.1 * rng.standard_normal((d, d))) + 0.2
    ens = eng.upload(X)
    g = 3.5 - X.sum(axis=1) / np.sqrt(d)
    e = 0.5 * (X**2).sum(axis=1) + 0.5 * d * np.log(2 * np.pi)
    ens.g[:n].copy_(eng.to_device(g))
    ens.e[:n].copy_(eng.to_device(e))
    sigma, beta = 1.7, 0.8
    slen = d + d * d + 2
    a_vec, ess4 = eng.empty(ens.ld), eng.empty(4)
    eng.reduce_ess(ens, sigma, beta, 0, out=ess4, a_out=a_vec)
    sg = sigma * g
    ell = -np.log(2) + np.log1p(-sg / np.sqrt(1 + sg * sg)) - e
    a = beta * ell
    np.testing.assert_allclose(a_vec[:n].cpu().numpy(), a, rtol=1e-13)
    w = np.exp(a - a.max())
    shift_h = X.mean(axis=0) * 0.9
    shift = eng.to_device(shift_h)
    su, sw = eng.cbree_moments(ens, a_vec, ess4[0:1], shift, eng.empty(2 * slen))
    su, sw = su.cpu().numpy(), sw.cpu().numpy()
    Y = X - shift_h
    np.testing.assert_allclose(su[:d], Y.sum(axis=0), rtol=0, atol=1e-12 * n)
    np.testing.assert_allclose(su[d:d + d * d].reshape(d, d), Y.T @ Y, rtol=0, atol=1e-12 * n)
    assert su[d + d * d] == np.sum(g <= 0) and su[d + d * d + 1] == n
    np.testing.assert_allclose(sw[:d], (Y * w[:, None]).sum(axis=0), rtol=0, atol=1e-12 * w.sum())
    np.testing.assert_allclose(sw[d:d + d * d].reshape(d, d), (Y * w[:, None]).T @ Y, rtol=0, atol=1e-11 * w.sum())
    np.testing.assert_allclose(sw[d + d * d], w.sum(), rtol=1e-13)
    # factorise + Mahalanobis / IS pass
    if n <= d:
        return
    mean, cov, L, Linv, stats = eng.empty(d), eng.empty(d * d), eng.empty(d * d), eng.empty(d * d), eng.empty(8)
    eng.moments_chol(eng.to_device(su), shift, d, 1, False, mean, cov, L, Linv, stats)
    np.testing.assert_allclose(cov.cpu().numpy().reshape(d, d), np.cov(X, rowvar=False), rtol=0, atol=1e-11)
    is4, logq = eng.empty(4), eng.empty(ens.ld)
    eng.cbree_maha(ens, mean, Linv, stats[0:1], sigma, beta, 0, None, None, is4, logq=logq)
    from scipy.stats import multivariate_normal

    want_lq = multivariate_normal.logpdf(X, mean=X.mean(axis=0), cov=np.cov(X, rowvar=False))
    np.testing.assert_allclose(logq[:n].cpu().numpy(), want_lq, rtol=0, atol=1e-9)
    r = -e - want_lq
    R = r.max()
    v = np.exp(r - R) * (g <= 0)
    t4 = is4.cpu().numpy()
    np.testing.assert_allclose(t4[0], R, rtol=0, atol=1e-9)
    np.testing.assert_allclose(t4[1:], [v.sum(), (v**2).sum(), n], rtol=1e-8, atol=1e-300)
    # and the fused sweep gives the same IS statistics / weighted sums about its own shift (the mean)
    sums_f, is4_f = eng.empty(slen), eng.empty(4)
    eng.cbree_maha(ens, mean, Linv, stats[0:1], sigma, beta, 0, ess4[0:1], sums_f, is4_f)
    np.testing.assert_allclose(is4_f.cpu().numpy(), t4, rtol=1e-12)
    m_w_f, C_w_f, m_w_s, C_w_s = eng.empty(d), eng.empty(d * d), eng.empty(d), eng.empty(d * d)
    eng.moments_chol(sums_f, mean, d, 0, True, m_w_f, C_w_f)
    eng.moments_chol(eng.to_device(sw), shift, d, 0, True, m_w_s, C_w_s)
    np.testing.assert_allclose(m_w_s.cpu().numpy(), m_w_f.cpu().numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(C_w_s.cpu().numpy(), C_w_f.cpu().numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(C_w_s.cpu().numpy().reshape(d, d), np.cov(X, rowvar=False, aweights=w, ddof=0),
                               rtol=0, atol=1e-11)


def test_split_kernels_are_deterministic(eng):
    n, d = 20000, 100
    rng, X, F, src, lsf, m_noise = _setup(eng, n, d, 3)
    Fd, md = eng.to_device(F), eng.to_device(m_noise)
    slen = d + d * d + 2
    outs = []
    for _ in range(2):
        a = eng.alloc_ensemble(n, d)
        eng.cbree_step(src, a, 0.5, md, Fd, lsf, None, None, seed=5, draw=9)
        a_vec, ess4 = eng.empty(a.ld), eng.empty(4)
        eng.reduce_ess(a, 2.0, 1.0, 0, out=ess4, a_out=a_vec)
        su, sw = eng.cbree_moments(a, a_vec, ess4[0:1], None, eng.empty(2 * slen))
        outs.append((eng.download(a), su.cpu().numpy().copy(), sw.cpu().numpy().copy()))
    for x, y in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("n,d", [(1, 100), (9471, 100), (9473, 64), (20001, 100), (37, 6), (1001, 4), (515, 20)])
def test_update_kernels_stay_inside_their_buffers(eng, n, d):
    """Guard bands around X', g', e' (all written by the update sweeps: warp-specialised transform for 56 <= d <= 103,
    DMMA-Gram thread-per-particle sweep for d <= 6, fused sweep in between) must survive ragged sizes untouched."""
    import torch

    from rareeventestimation_b200 import _lib
    from rareeventestimation_b200.engine import DeviceLsf, Ensemble

    rng = np.random.default_rng(n + d)
    src = eng.upload(rng.standard_normal((n, d)))
    ld = src.ld
    big = torch.full(((d + 2) * ld,), float("nan"), dtype=torch.float64, device=src.X.device)
    vec = torch.full((2, 3 * ld), float("nan"), dtype=torch.float64, device=src.X.device)
    dst = Ensemble(big[ld:(d + 1) * ld].view(d, ld), n, d, ld, 0, vec[0, ld:2 * ld], vec[1, ld:2 * ld])
    A = rng.standard_normal((d, d)) / np.sqrt(d)
    F = eng.to_device(np.linalg.cholesky(A @ A.T + 0.5 * np.eye(d)))
    lsf = DeviceLsf(_lib.LSF_AFFINE, d, None, 3.5)
    split = eng.path(d) == 2
    sums = None if split else eng.empty(d + d * d + 2)
    eng.cbree_step(src, dst, 0.7, eng.zeros(d), F, lsf, None, sums, seed=5, draw=9)
    torch.cuda.synchronize()
    assert torch.isnan(big[:ld]).all() and torch.isnan(big[(d + 1) * ld:]).all()
    assert torch.isnan(vec[:, :ld]).all() and torch.isnan(vec[:, 2 * ld:]).all()
    assert torch.isfinite(dst.X[:, :n]).all() and torch.isfinite(dst.g[:n]).all() and torch.isfinite(dst.e[:n]).all()
    # the redraw of the resampling branch writes the same buffers
    if d >= 2:
        model_mu = rng.standard_normal(d)
        model_mu /= np.linalg.norm(model_mu)
        from rareeventestimation_b200.mixture import _house

        v, b = _house(model_mu)
        big.fill_(float("nan"))
        eng.vmfn_sample(dst, eng.to_device(v), b, 4.0, d / 2, float(d), 11, 3)
        torch.cuda.synchronize()
        assert torch.isnan(big[:ld]).all() and torch.isnan(big[(d + 1) * ld:]).all()
        assert torch.isfinite(dst.X[:, :n]).all()
