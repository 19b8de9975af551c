"""Out-of-bounds and race evidence without compute-sanitizer (closed on the GPU pool, profiles/r2_sanitizer_unavailable.txt).

Every kernel family runs on an ensemble whose X / g / e arrays are windows into larger canary-filled allocations, at
ragged sizes (n not a multiple of 8 / 96 / 128, d in every kernel family: small-d, fused DMMA, split DMMA, modular).
Afterwards (1) every canary byte outside the documented outputs is untouched, (2) a second run gives the same bits:
the cross-CTA reductions are two-stage and fixed-order, so a shared-memory race or a missing barrier would show up as a
run-to-run difference."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CANARY = -1.2345678901234567e+123
GUARD = 4096  # doubles on either side of every array


@pytest.fixture(scope="module")
def eng():
    from rareeventestimation_b200.engine import Engine

    return Engine.get()


def guarded(eng, numel):
    buf = torch.full((numel + 2 * GUARD,), CANARY, dtype=torch.float64, device=eng.device)
    return buf, buf[GUARD:GUARD + numel]


def guarded_ensemble(eng, n, d, fill=None):
    from rareeventestimation_b200.engine import Ensemble, _pad

    ld = _pad(n)
    bx, x = guarded(eng, d * ld)
    bg, g = guarded(eng, ld)
    be, e = guarded(eng, ld)
    ens = Ensemble(x.view(d, ld), n, d, ld, 0, g, e)
    if fill is not None:
        ens.X[:, :n].copy_(torch.from_numpy(np.ascontiguousarray(fill.T)))
    return ens, (bx, bg, be)


def intact(bufs):
    for b in bufs:
        head, tail = b[:GUARD], b[-GUARD:]
        if not (bool((head == CANARY).all()) and bool((tail == CANARY).all())):
            return False
    return True


@pytest.mark.parametrize("n,d", [(1, 2), (13, 6), (1001, 4), (777, 37), (1237, 64), (1000, 100), (905, 103), (611, 150),
                                 (333, 152), (257, 200), (130, 256)])
def test_iteration_kernels_stay_inside_their_buffers_and_are_deterministic(eng, n, d):
    from rareeventestimation_b200 import _lib
    from rareeventestimation_b200.engine import DeviceLsf

    rng = np.random.default_rng(n * 1000 + d)
    X0 = rng.standard_normal((n, d))
    lsf = DeviceLsf(_lib.LSF_AFFINE, d, None, 3.5)
    slen = d + d * d + 2
    split = eng.path(d) == 2

    def run():
        src, bs = guarded_ensemble(eng, n, d, X0)
        dst, bd = guarded_ensemble(eng, n, d)
        bufs = list(bs) + list(bd)
        small = {}
        for name, numel in [("sums", 2 * slen), ("sums_w", slen), ("mean", d), ("cov", d * d), ("L", d * d), ("Linv", d * d),
                            ("stats", 8), ("ess4", 4), ("is4", 4), ("F", d * d), ("fst", 8), ("m_noise", d), ("a", src.ld),
                            ("w", src.ld), ("logq", src.ld), ("adapt", 16)]:
            b, v = guarded(eng, numel)
            bufs.append(b)
            small[name] = v
        eng.lsf_eval(lsf, src, out=src.g)
        eng.energy(src, out=src.e)
        small["m_noise"].fill_(0.01)
        C = torch.eye(d, dtype=torch.float64, device=eng.device) * 0.7 + 0.02
        eng.chol_scaled(C.reshape(-1).contiguous(), d, 1.3, small["F"], small["fst"])
        sums_u = small["sums"][:slen]
        # sweep 1 (Philox noise), ESS / log-weights, moments, factorisation, sweep 2, search kernel
        eng.cbree_step(src, dst, 0.6, small["m_noise"], small["F"], lsf, None, None if split else sums_u, seed=7, draw=3)
        eng.reduce_ess(dst, 1.5, 0.8, 0, out=small["ess4"], a_out=small["a"])
        if split:
            eng.cbree_moments(dst, small["a"], small["ess4"][0:1], None, small["sums"])
        else:
            eng.ensemble_sums(dst, None, sums_u)
        eng.moments_chol(sums_u, None, d, 1, False, small["mean"], small["cov"], small["L"], small["Linv"], small["stats"])
        if split:
            eng.cbree_maha(dst, small["mean"], small["Linv"], small["stats"][0:1], 1.5, 0.8, 0, None, None, small["is4"],
                           logq=small["logq"])
        else:
            eng.cbree_maha(dst, small["mean"], small["Linv"], small["stats"][0:1], 1.5, 0.8, 0, small["ess4"][0:1],
                           small["sums_w"], small["is4"], w=small["w"], logq=small["logq"])
        out = eng.adapt(dst, 0, 1.5, 2.5, 2.0, 0.8, True, 0.5 * n)
        torch.cuda.synchronize()
        assert intact(bufs), "a kernel wrote outside its arrays"
        res = {k: v.clone() for k, v in small.items() if k in ("mean", "cov", "Linv", "stats", "ess4", "is4", "logq")}
        res["X"], res["g"], res["e"] = dst.X[:, :n].clone(), dst.g[:n].clone(), dst.e[:n].clone()
        res["adapt"] = torch.from_numpy(np.array(out))
        return res

    a, b = run(), run()
    for k in a:
        va, vb = a[k].cpu().numpy(), b[k].cpu().numpy()
        if k == "logq":
            va, vb = va[:n], vb[:n]
        assert np.array_equal(va, vb, equal_nan=True), f"{k}: two runs differ"
    assert np.isfinite(a["X"].cpu().numpy()).all() and np.isfinite(a["mean"].cpu().numpy()).all()


@pytest.mark.parametrize("n,d", [(999, 5), (4097, 50), (1203, 100)])
def test_sampler_and_vmfn_kernels_stay_inside_their_buffers(eng, n, d):
    from rareeventestimation_b200.mixture import VMFNMixture

    ens, bufs = guarded_ensemble(eng, n, d)
    eng.philox_normal(n, d, 11, 5, first=123, out=ens)
    ens.X[:, :n].add_(0.8)
    model = VMFNMixture(1)
    model.fit_device(eng, ens)
    dst, bd = guarded_ensemble(eng, n, d)
    model.sample_device(eng, dst, 99, 4)
    bq, logq = guarded(eng, dst.ld)
    b4, is4 = guarded(eng, 4)
    eng.energy(dst, out=dst.e)
    dst.g.fill_(0.5)
    model.logpdf_device(eng, dst, logq=logq, is4=is4)
    torch.cuda.synchronize()
    assert intact(list(bufs) + list(bd) + [bq, b4])
    assert np.isfinite(dst.X[:, :n].cpu().numpy()).all() and np.isfinite(logq[:n].cpu().numpy()).all()
