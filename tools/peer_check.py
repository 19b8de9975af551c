"""2+ GPU check of the NVLink mailbox exchange (csrc/ree_peer.cu) against the NCCL all_gather + merge path:
torchrun --nproc-per-node 2 tools/peer_check.py   -> every rank prints OK and the latency of both paths."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rareeventestimation_b200.engine import Comm, Engine  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine.get(torch.device("cuda", local), Comm())
comm = eng.comm
rng = np.random.default_rng(100 + rank)
worst = 0.0
for it in range(200):
    t4 = np.array([rng.normal() * 50, rng.random() * 1e3 + 1, rng.random() * 1e3 + 1, float(rng.integers(1, 10**6))])
    if it % 17 == 3:
        t4 = np.array([-np.inf, 0.0, 0.0, 0.0])  # an empty rank
    a = eng.to_device(t4)
    b = a.clone()
    comm.peer = None if it == 0 else comm.peer
    comm.combine_expsums_(a)          # mailbox path (connects at the first call)
    assert comm.peer, "peer mailboxes did not connect"
    keep, comm.peer = comm.peer, False
    comm.combine_expsums_(b)          # NCCL all_gather + ree_merge_expsums
    comm.peer = keep
    x, y = a.cpu().numpy(), b.cpu().numpy()
    # two ranks: one merge either way, identical bits.  More ranks: the mailbox kernel merges pairwise in rank order, the
    # NCCL path's merge kernel rescales everything to the global shift once -- same value, last-bit differences.
    if dist.get_world_size() == 2:
        assert np.array_equal(x, y, equal_nan=True), (it, x, y)
    else:
        assert np.array_equal(np.isfinite(x), np.isfinite(y)), (it, x, y)
        np.testing.assert_allclose(x[np.isfinite(x)], y[np.isfinite(y)], rtol=1e-13)
    # ... and every rank holds the same bits (what the identical host control flow of all ranks relies on)
    everyone = [torch.empty_like(a) for _ in range(dist.get_world_size())]
    dist.all_gather(everyone, a)
    assert all(torch.equal(everyone[0], e) or (torch.isnan(e).any() and torch.isnan(everyone[0]).any()) for e in everyone), it
comm.check_peers()
out = eng.empty(4)
for name, peer in (("mailbox", True), ("nccl", False)):
    comm.peer = peer
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(300):
        out.copy_(a)
        comm.combine_expsums_(out)
        eng.fetch(out)
    torch.cuda.synchronize()
    print(f"rank {rank}: {name:8s} {1e6 * (time.perf_counter() - t0) / 300:7.1f} us per exchange + fetch", flush=True)
comm.peer = True
print(f"rank {rank}: OK (200 exchanges: identical bits on all ranks, equal to the NCCL path)", flush=True)

# ---- the packed-sums all-reduce through the staging areas behind the mailboxes (ree_peer_allreduce_sum) -------------
world = dist.get_world_size()
for n in (1, 7, 44, 2 * (100 * 100 + 100 + 2), int(eng.lib.ree_peer_allreduce_max())):
    for rep in range(6):  # more than two in a row: both staging parities are reused
        parts = [np.random.default_rng(1000 * r + 7 * rep + n).standard_normal(n) * 10.0 ** (r % 5) for r in range(world)]
        want = np.zeros(n)
        for r in range(world):  # rank order, like the kernel
            want = want + parts[r]
        v = eng.to_device(parts[rank])
        os.environ["REE_PEER_ALLREDUCE"] = "1"
        comm.allreduce_sum_(v)                 # peer kernel
        got = v.cpu().numpy()
        assert np.array_equal(got, want), (n, rep, np.abs(got - want).max())
        w = eng.to_device(parts[rank])
        os.environ["REE_PEER_ALLREDUCE"] = "0"
        comm.allreduce_sum_(w)                 # NCCL
        os.environ.pop("REE_PEER_ALLREDUCE")
        np.testing.assert_allclose(w.cpu().numpy(), want, rtol=1e-13, atol=1e-9)
    # mixed with tuple exchanges (same sequence counter)
    comm.combine_expsums_(eng.to_device(np.array([1.0, 2.0, 3.0, 4.0])))
comm.check_peers()
buf = eng.to_device(np.ones(2 * (100 * 100 + 100 + 2)))
for name, flag in (("peer kernel", "1"), ("nccl", "0")):
    os.environ["REE_PEER_ALLREDUCE"] = flag
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        comm.allreduce_sum_(buf)
    e1.record(); torch.cuda.synchronize()
    print(f"rank {rank}: all-reduce of 20204 doubles, {name:11s} {1e3 * e0.elapsed_time(e1) / 200:7.1f} us per call (device time)", flush=True)
os.environ.pop("REE_PEER_ALLREDUCE")
print(f"rank {rank}: OK (all-reduce through peer memory equals the rank-order sum bit for bit)", flush=True)
dist.destroy_process_group()
