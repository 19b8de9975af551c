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
    assert np.array_equal(x, y, equal_nan=True), (it, x, y)
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
print(f"rank {rank}: OK (200 exchanges bit-identical to the NCCL path)", flush=True)
dist.destroy_process_group()
