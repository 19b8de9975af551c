"""Multi-GPU parity: the same global problem solved on 1 and on G GPUs gives the same trajectory --
Philox counters are global particle indices, only the order of the cross-rank sums differs.  Three paths (split DMMA
d = 100, small-d oscillator, vMFNM resampling) through tools/shard_check.py under torchrun; skipped where the box has
fewer devices (the driver's single-GPU tier)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _devices():
    import torch

    return torch.cuda.device_count()


def _run(world, out, n, port):
    tool = os.path.join(ROOT, "tools", "shard_check.py")
    if world == 1:
        cmd = [sys.executable, tool, "--out", out, "--particles", str(n)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), tool, "--out", out, "--particles", str(n)]
    env = dict(os.environ)
    env.pop("CUDA_VISIBLE_DEVICES", None) if world > 1 else None
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return json.load(open(out))


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_solve_matches_single_gpu(world, tmp_path):
    if _devices() < world:
        pytest.skip(f"needs {world} GPUs")
    n = 120_000 + 8  # not a multiple of every world size times the tile sizes: ragged shards
    one = _run(1, str(tmp_path / "s1.json"), n, 0)
    many = _run(world, str(tmp_path / f"s{world}.json"), n, 29540 + world)
    for cfg in one:
        assert one[cfg]["costs"] == many[cfg]["costs"] and one[cfg]["msg"] == many[cfg]["msg"], cfg
        for key in ("sigma", "beta", "t_step", "cvar", "estimate"):
            x, y = np.array(one[cfg][key], dtype=float), np.array(many[cfg][key], dtype=float)
            assert x.shape == y.shape, (cfg, key)
            np.testing.assert_allclose(y, x, rtol=1e-8, atol=1e-300, err_msg=f"{cfg} {key}")


@pytest.mark.parametrize("d,n,kw", [(4, 2000, {}), (100, 40_000, dict(num_steps=6)), (6, 30_001, dict(num_steps=8))])
def test_single_process_multi_device_solve(d, n, kw):
    """``CBREE(devices=2).solve(problem)`` is a normal call in ONE process: one host thread per device,
    NVLink mailboxes through peer access.  Same trajectory as one device (only the order of the cross-device sums
    differs) and one Solution over all particles.  README example (d = 4, N = 2000), the headline dimension, and the
    oscillator with a ragged split."""
    if _devices() < 2:
        pytest.skip("needs 2 GPUs")
    from copy import deepcopy

    import rareeventestimation_b200 as ree

    prob = deepcopy(ree.prob_nonlin_osc) if d == 6 else ree.make_linear_problem(d)
    prob.set_sample(n, seed=1)
    common = dict(seed=1, quiet=True, save_history=True, **kw)
    one = ree.CBREE(**common).solve(prob)
    two = ree.CBREE(devices=2, **common).solve(prob)
    assert two.msg == one.msg and two.costs == one.costs and two.num_steps == one.num_steps
    np.testing.assert_allclose(two.prob_fail_hist, one.prob_fail_hist, rtol=1e-8, atol=1e-300)
    np.testing.assert_allclose(two.temp_hist, one.temp_hist, rtol=1e-8, equal_nan=True)
    assert two.ensemble_hist.shape == one.ensemble_hist.shape == (len(one.prob_fail_hist), n, d)
    assert two.N == n
    np.testing.assert_allclose(np.asarray(two.ensemble_hist[-1]), np.asarray(one.ensemble_hist[-1]), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(np.asarray(two.lsf_eval_hist)[-1], np.asarray(one.lsf_eval_hist)[-1], rtol=1e-7, atol=1e-9)
    # the engines of both devices are usable by ordinary solves afterwards
    again = ree.CBREE(**common).solve(prob)
    np.testing.assert_array_equal(again.prob_fail_hist, one.prob_fail_hist)


@pytest.mark.parametrize("world", [2, 8])
def test_peer_memory_collectives(world):
    """NVLink mailbox exchange and the peer-memory all-reduce of the packed sums (csrc/ree_peer.cu) against NCCL and
    against the rank-order sum, bit for bit; both staging parities, lengths 1 ... the maximum (tools/peer_check.py)."""
    if _devices() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29580 + world), os.path.join(ROOT, "tools", "peer_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK (all-reduce through peer memory") == world
