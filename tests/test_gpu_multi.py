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
        cmd = [sys.executable, tool, "--out", out, "--n", str(n)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), tool, "--out", out, "--n", str(n)]
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
