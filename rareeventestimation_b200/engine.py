"""Device engine: owns the HBM-resident particle ensemble (SoA) and drives the sm_100a kernels of
``libree_b200.so`` through the C ABI.  PyTorch is used for device memory, streams and
``torch.distributed`` only; every N-sized computation is a hand-written CUDA kernel.

Sharding: one process per GPU; each rank keeps a contiguous slice of the particles
for the whole solve.  Only per-iteration partial sums cross ranks (``all_reduce`` of the packed raw
sums, ``all_gather`` of the max-shifted 4-tuples), so every rank holds bit-identical consensus
statistics and runs the identical host control flow.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import ReeError, ReeLsf, check

F64 = torch.float64


# --------------------------------------------------------------------------- #
# cross-rank combination of partial statistics (pure tensor logic: testable on CPU/gloo)
# --------------------------------------------------------------------------- #
def merge_expsums(parts: torch.Tensor) -> torch.Tensor:
    """Merge (R,4) rows [M, S1, S2, cnt] of max-shifted exponential sums in fixed rank order.

    sum exp(a) = exp(M) S1, sum exp(2a) = exp(2M) S2; result is shifted to the global max."""
    M = parts[:, 0]
    top = torch.max(M)
    if not torch.isfinite(top):
        out = parts.sum(dim=0)
        out[0] = top
        return out
    f = torch.exp(M - top)
    f = torch.where(torch.isfinite(M), f, torch.zeros_like(f))
    return torch.stack([top, (parts[:, 1] * f).sum(), (parts[:, 2] * f * f).sum(), parts[:, 3].sum()])


class Comm:
    """Thin wrapper over a torch.distributed process group (or a single process)."""

    # The packed d x d sums: NCCL's all_reduce (19 us for 162 KB on 2 GPUs) beats the single-CTA peer-memory kernel
    # (41 us) under torchrun, so it stays the default there (REE_PEER_ALLREDUCE=1 selects the kernel); the threaded
    # communicator of CBREE(devices=N) has no NCCL and uses the kernel.
    _PEER_ALLREDUCE = "0"

    def __init__(self, group=None, enabled: Optional[bool] = None):
        import torch.distributed as dist

        self.dist = dist
        if enabled is None:
            enabled = dist.is_available() and dist.is_initialized()
        self.enabled = bool(enabled) and dist.is_initialized() and dist.get_world_size(group) > 1
        self.group = group
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.peer = None  # None: not tried yet; True / False: NVLink mailboxes connected / NCCL all_gather instead

    def _connect_peers(self, device) -> bool:
        """Map every rank's mailbox into this process (CUDA IPC).  All ranks decide together: peer mode only if every
        rank is on this host and every mapping succeeded; otherwise all of them keep the NCCL all_gather."""
        import os
        import socket

        lib = _lib.load()
        ok = os.environ.get("REE_PEER", "1") != "0" and self.world <= 16
        buf = (C.c_ubyte * 64)()
        with torch.cuda.device(device):
            if ok:
                ok = lib.ree_peer_create(self.rank, self.world, buf) == 0
            infos = [None] * self.world
            self.dist.all_gather_object(infos, (socket.gethostname(), ok, bytes(buf)), group=self.group)
            ok = all(i[1] for i in infos) and len({i[0] for i in infos}) == 1
            if ok:
                ok = lib.ree_peer_connect(b"".join(i[2] for i in infos)) == 0
            flags = [None] * self.world
            self.dist.all_gather_object(flags, ok, group=self.group)  # also the barrier before the first exchange
        return all(flags)

    def _peer_allreduce(self, t: torch.Tensor) -> bool:
        """Sum `t` over the ranks with the peer-memory kernel (``ree_peer_allreduce_sum``) if the NVLink mailboxes are
        connected and `t` qualifies (fp64, contiguous, on the device, short enough).  All ranks decide alike."""
        if not (t.is_cuda and t.dtype == F64 and t.is_contiguous() and t.numel() >= 1):
            return False
        lib = _lib.load()
        if t.numel() > lib.ree_peer_allreduce_max() or os.environ.get("REE_PEER_ALLREDUCE", self._PEER_ALLREDUCE) == "0":
            return False
        if self.peer is None:
            self.peer = self._connect_peers(t.device)
        if not self.peer:
            return False
        stream = C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
        check(lib.ree_peer_allreduce_sum(t.data_ptr(), t.numel(), stream), "ree_peer_allreduce_sum")
        return True

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.enabled and not self._peer_allreduce(t):
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def combine_expsums_(self, out4: torch.Tensor) -> torch.Tensor:
        """all_gather the local 4-tuple and merge in rank order (identical bits on every rank)."""
        if self.enabled and out4.is_cuda:
            if self.peer is None:
                self.peer = self._connect_peers(out4.device)
            if self.peer:  # one kernel: store to all mailboxes, wait, merge (csrc/ree_peer.cu)
                stream = C.c_void_p(torch.cuda.current_stream(out4.device).cuda_stream)
                check(_lib.load().ree_peer_combine_expsums(out4.data_ptr(), stream), "ree_peer_combine_expsums")
                return out4
        if self.enabled:
            gathered = torch.empty(self.world * 4, dtype=out4.dtype, device=out4.device)
            self.dist.all_gather_into_tensor(gathered, out4.contiguous(), group=self.group)
            if out4.is_cuda:  # one small kernel instead of a dozen elementwise launches
                stream = C.c_void_p(torch.cuda.current_stream(out4.device).cuda_stream)
                check(_lib.load().ree_merge_expsums(gathered.data_ptr(), self.world, out4.data_ptr(), stream),
                      "ree_merge_expsums")
            else:
                out4.copy_(merge_expsums(gathered.view(self.world, 4)))
        return out4

    def check_peers(self) -> None:
        """Raise if a peer exchange timed out since the last check (one 4-byte D2H copy)."""
        if self.peer:
            check(_lib.load().ree_peer_check(), "ree_peer_check")

    def broadcast_object(self, obj, src: int = 0):
        """The object rank `src` passes, on every rank."""
        if not self.enabled:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def all_gather_vec(self, t: torch.Tensor) -> torch.Tensor:
        """Concatenation over the ranks (rank order) of equally sized device vectors."""
        if not self.enabled:
            return t.clone()
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def shard(self, n_global: int):
        """Contiguous slice [lo, hi) of this rank."""
        base, rem = divmod(int(n_global), self.world)
        lo = self.rank * base + min(self.rank, rem)
        return lo, lo + base + (1 if self.rank < rem else 0)


class ThreadGroup:
    """Shared state of one process that drives `devices` with one host thread each (``CBREE(devices=N)``)."""

    def __init__(self, devices):
        import threading

        self.devices = [int(x) for x in devices]
        self.world = len(self.devices)
        self.barrier = threading.Barrier(self.world)
        self.slots = [None] * self.world
        self.failed = None


class ThreadComm(Comm):
    """The communicator of one thread of a :class:`ThreadGroup`: same interface as :class:`Comm` (one process per GPU
    under torchrun), no torch.distributed.  The 4-tuple merges run in the same NVLink mailbox kernels (peer access
    instead of CUDA IPC); the d x d sums are added in rank order from peer-to-peer copies, so every rank holds
    identical bits."""

    _PEER_ALLREDUCE = "1"

    def __init__(self, tgroup: ThreadGroup, rank: int):
        self.dist = None
        self.group = None
        self.tgroup = tgroup
        self.enabled = tgroup.world > 1
        self.rank, self.world = int(rank), tgroup.world
        self.peer = None

    def _sync_point(self):
        try:
            self.tgroup.barrier.wait(timeout=600)
        except Exception:
            raise ReeError("a device thread of this solve failed or stalled (see the first exception)") from None

    def _connect_peers(self, device) -> bool:
        lib = _lib.load()
        buf = (C.c_ubyte * 64)()
        with torch.cuda.device(device):
            check(lib.ree_peer_create(self.rank, self.world, buf), "ree_peer_create")
            self._sync_point()
            devs = (C.c_int * self.world)(*self.tgroup.devices)
            check(lib.ree_peer_connect_local(devs), "ree_peer_connect_local")
            self._sync_point()
        return True

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if not self.enabled or self._peer_allreduce(t):
            return t
        g = self.tgroup  # host-mediated fall-back (CPU tensors in the tests, other dtypes, long vectors)
        self._drain(t)  # this rank's contribution is final
        g.slots[self.rank] = t
        self._sync_point()
        acc = None
        for r in range(self.world):  # rank order on every rank: identical bits
            part = g.slots[r] if r == self.rank else g.slots[r].to(t.device)
            acc = part.clone() if acc is None else acc.add_(part)
        self._drain(t)       # all peer reads are done ...
        self._sync_point()   # ... on every rank, before anybody's buffer changes
        t.copy_(acc)
        return t

    @staticmethod
    def _drain(t: torch.Tensor) -> None:
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()

    def combine_expsums_(self, out4: torch.Tensor) -> torch.Tensor:
        if self.enabled:
            if self.peer is None:
                self.peer = self._connect_peers(out4.device)
            stream = C.c_void_p(torch.cuda.current_stream(out4.device).cuda_stream)
            check(_lib.load().ree_peer_combine_expsums(out4.data_ptr(), stream), "ree_peer_combine_expsums")
        return out4

    def broadcast_object(self, obj, src: int = 0):
        if not self.enabled:
            return obj
        g = self.tgroup
        if self.rank == src:
            g.slots[src] = obj
        self._sync_point()
        out = g.slots[src]
        self._sync_point()
        return out

    def all_gather_vec(self, t: torch.Tensor) -> torch.Tensor:
        if not self.enabled:
            return t.clone()
        g = self.tgroup
        self._drain(t)
        g.slots[self.rank] = t
        self._sync_point()
        out = torch.cat([g.slots[r] if r == self.rank else g.slots[r].to(t.device) for r in range(self.world)])
        self._drain(t)
        self._sync_point()
        return out


# --------------------------------------------------------------------------- #
@dataclass
class DeviceLsf:
    """Device descriptor of a limit-state function (struct ree_lsf) + the tensors it points to."""

    kind: int
    d: int
    coef: Optional[torch.Tensor] = None
    offset: float = 0.0
    p0: float = 0.0
    p1: float = 0.0
    n_ele: int = 0

    def struct(self) -> ReeLsf:
        return ReeLsf(self.kind, self.d, self.coef.data_ptr() if self.coef is not None else None, self.offset,
                      self.p0, self.p1, self.n_ele, 0)


class Ensemble:
    """An HBM-resident ensemble: X is SoA ``[d][ld]`` fp64, g / e are the per-particle LSF and energy."""

    __slots__ = ("X", "g", "e", "n", "d", "ld", "first")

    def __init__(self, X, n, d, ld, first=0, g=None, e=None):
        self.X, self.n, self.d, self.ld, self.first = X, int(n), int(d), int(ld), int(first)
        self.g, self.e = g, e


def _pad(n: int) -> int:
    return max(8, (int(n) + 7) // 8 * 8)


class Engine:
    """One engine per (process, device)."""

    _instances = {}

    @classmethod
    def get(cls, device=None, comm: Optional[Comm] = None) -> "Engine":
        if not torch.cuda.is_available():
            raise ReeError("rareeventestimation_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        key = (device.index if device.index is not None else torch.cuda.current_device())
        eng = cls._instances.get(key)
        if eng is None:
            eng = cls._instances[key] = Engine(torch.device("cuda", key))
            eng.comm = Comm(enabled=False)
            eng._comm_given = False
            eng._load_nvector_kernels()
        if comm is not None:  # an explicit communicator / group stays until another one is passed
            eng.comm, eng._comm_given = comm, True
        elif not eng._comm_given:
            # default: follow torch.distributed's state.  The communicator is only rebuilt when that state changed, so
            # callers inside a solve (DeviceLSF / StandardNormalEnergy / VMFNMixture on host arrays, user callbacks)
            # never drop the connected peer mailboxes or trigger a collective re-connection on one rank only.
            import torch.distributed as dist

            live = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            if live != eng.comm.enabled or (live and eng.comm.world != dist.get_world_size()):
                eng.comm = Comm()
        return eng

    def _load_nvector_kernels(self):
        """CUDA loads a kernel at its first launch; the sigma / beta searches reach some reduction kernels only after
        several iterations (the Brent search takes over from the closed-form sigma update around iteration 9), where
        the load showed up as a sporadic +50 ms.  One launch of each on an 8-element vector at engine creation."""
        with torch.cuda.device(self.device):
            ens = self.alloc_ensemble(8, 1)
            ens.g.fill_(0.5)
            ens.e.fill_(1.0)
            out = self.empty(4)
            ell = self.logtgt(ens.g, ens.e, 8, 1.0, 0)
            rng2 = self.vector_range(ell, 8, 1)
            self.reduce_ess(ens, 1.0, 1.0, 0, out=out, a_out=self.empty(ens.ld))
            self.reduce_scaled(ell, 8, 1, 0.5, out=out)
            self.reduce_scaled(ell, 8, 1, 0.5, out=out, range2=rng2)
            self.reduce_cvar(ens, 2.0, 1.0, 0, out=out)   # algebraic kernel (sigma grows)
            self.reduce_cvar(ens, 1.0, 2.0, 0, out=out)   # generic kernel
            self.count_failed(ens)
            self.adapt(ens, 0, 1.0, 1.5, 2.0, 1.0, True, 4.0)  # persistent search kernel, algebraic objective
            self.adapt(ens, 2, 1.0, 1.5, 2.0, 1.0, True, 4.0)  # ... generic objective
            self.fetch(out)

    def __init__(self, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        self.comm = Comm(enabled=False)
        self._ws = {}
        self._pool = {}
        self._scratch = {}
        self._staging = None
        self._pin_ring, self._pin_next, self._pin_down, self._side = None, 0, None, None
        self._pin_slots = [None, None, None]
        self._open_post = None
        self.adapt_launches = 0
        self._fetch_buf = np.empty(64, dtype=np.float64)
        with torch.cuda.device(self.device):
            self.num_ctas = self.lib.ree_num_ctas()

    # -- plumbing ------------------------------------------------------------ #
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def sync(self):
        torch.cuda.current_stream(self.device).synchronize()

    def workspace(self, d: int) -> torch.Tensor:
        ws = self._ws.get(d)
        if ws is None:
            with torch.cuda.device(self.device):
                nbytes = self.lib.ree_workspace_bytes(d)
            ws = self._ws[d] = torch.zeros(nbytes // 8 + 8, dtype=F64, device=self.device)
        return ws

    def staging(self, elems: int) -> torch.Tensor:
        if self._staging is None or self._staging.numel() < elems:
            self._staging = torch.empty(elems, dtype=F64, device=self.device)
        return self._staging

    def empty(self, *shape):
        return torch.empty(*shape, dtype=F64, device=self.device)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=F64, device=self.device)

    def to_device(self, arr) -> torch.Tensor:
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64))
        return torch.from_numpy(a).to(self.device, non_blocking=False)

    # small host <-> device traffic of the iteration loop goes through pinned memory: a pageable cudaMemcpy stalls the
    # host until the stream has drained AND costs ~50 us of driver staging per call
    _RING, _RING_DOUBLES = 16, 4096

    def to_device_small(self, arr) -> torch.Tensor:
        """Stream-ordered upload of a small fp64 vector (<= 4096 values) from a ring of pinned buffers; a slot is reused
        after 16 uploads (every iteration of a solve ends in a synchronising readback long before that)."""
        a = np.asarray(arr, dtype=np.float64).reshape(-1)
        if a.size > self._RING_DOUBLES:
            return self.to_device(a)
        if self._pin_ring is None:
            self._pin_ring = [torch.empty(self._RING_DOUBLES, dtype=F64, pin_memory=True) for _ in range(self._RING)]
            self._pin_next = 0
        slot = self._pin_ring[self._pin_next]
        self._pin_next = (self._pin_next + 1) % self._RING
        slot.numpy()[: a.size] = a
        out = torch.empty(a.size, dtype=F64, device=self.device)
        out.copy_(slot[: a.size], non_blocking=True)
        return out

    def to_host(self, t: torch.Tensor) -> np.ndarray:
        """Synchronising readback of a device vector through a pinned buffer (a view that the next call overwrites)."""
        n = t.numel()
        if self._pin_down is None or self._pin_down.numel() < n:
            self._pin_down = torch.empty(max(n, 1 << 16), dtype=F64, pin_memory=True)
        self._pin_down[:n].copy_(t.reshape(-1), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._pin_down[:n].numpy()

    def to_host_async(self, t: torch.Tensor, slot: int):
        """Stream-ordered readback into pinned buffer `slot` (0 or 1): returns (host view, event).  The view is valid once
        the event has completed and until the next readback into the same slot."""
        n = t.numel()
        if self._pin_slots[slot] is None or self._pin_slots[slot].numel() < n:
            self._pin_slots[slot] = torch.empty(max(n, 1 << 12), dtype=F64, pin_memory=True)
        buf = self._pin_slots[slot][:n]
        buf.copy_(t.reshape(-1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return buf.numpy(), ev

    def side_stream(self) -> "torch.cuda.Stream":
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    # -- ensemble I/O ---------------------------------------------------------- #
    def alloc_ensemble(self, n: int, d: int, first: int = 0, with_vectors=True) -> Ensemble:
        """Ensemble buffers come from a free list first: at the headline size one ensemble is 8 GB and a
        cudaMalloc/cudaFree pair per iteration would cost more than the sweeps."""
        free = self._pool.get((int(n), int(d)))
        if free:
            ens = free.pop()
            ens.first = int(first)
            return ens
        ld = _pad(n)
        X = self.empty(d, ld)
        g = self.empty(ld) if with_vectors else None
        e = self.empty(ld) if with_vectors else None
        return Ensemble(X, n, d, ld, first, g, e)

    def release(self, ens: Optional[Ensemble], keep: int = 8):
        """Return an ensemble nobody references any more to the free list."""
        if ens is None or ens.g is None or ens.e is None:
            return
        free = self._pool.setdefault((ens.n, ens.d), [])
        if len(free) < keep:
            free.append(ens)

    def drop_pool(self):
        self._pool.clear()
        self._scratch.clear()

    def reserve(self, n: int, d: int, count: int):
        """Make sure `count` ensembles of this shape sit in the free list (one cudaMalloc each, outside the
        iteration loop); the allocations of a solve are then reused by the next one."""
        free = self._pool.setdefault((int(n), int(d)), [])
        while len(free) < count:
            ld = _pad(n)
            free.append(Ensemble(self.empty(d, ld), n, d, ld, 0, self.empty(ld), self.empty(ld)))

    def scratch(self, name: str, numel: int) -> torch.Tensor:
        """Named N-sized scratch vector that persists across iterations (torch's caching allocator would otherwise
        carve it out of a freed 8 GB ensemble block and force a fresh cudaMalloc for the next ensemble)."""
        t = self._scratch.get(name)
        if t is None or t.numel() < numel:
            t = self._scratch[name] = self.empty(numel)
        return t

    def fetch(self, t: torch.Tensor, n: Optional[int] = None) -> np.ndarray:
        """Synchronising readback of a small fp64 device result (<= 64 values) through ``ree_fetch_small``."""
        n = t.numel() if n is None else n
        buf = self._fetch_buf
        check(self.lib.ree_fetch_small(t.data_ptr(), n, buf.ctypes.data, self.stream), "ree_fetch_small")
        return buf[:n].copy()

    def fetch_array(self, t: torch.Tensor) -> np.ndarray:
        """Host copy of a small device vector of any length (ree_fetch_small up to 64 values, D2H copy beyond)."""
        return self.fetch(t) if t.numel() <= 64 else t.cpu().numpy()

    def upload(self, host: np.ndarray, first: int = 0) -> Ensemble:
        """Host (n,d) C-order array -> SoA ensemble (chunked H2D + device transpose)."""
        host = np.ascontiguousarray(host, dtype=np.float64)
        if host.ndim != 2:
            raise ValueError("ensemble must be a 2-d array (n, d)")
        n, d = host.shape
        ens = self.alloc_ensemble(n, d, first)
        rows = max(1, min(n, (64 << 20) // (8 * d)))
        st = self.staging(rows * d)
        check(self.lib.ree_upload_ensemble(host.ctypes.data, ens.X.data_ptr(), n, d, ens.ld, st.data_ptr(),
                                           st.numel(), self.stream), "ree_upload_ensemble")
        return ens

    def download(self, ens: Ensemble, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((ens.n, ens.d), dtype=np.float64)
        rows = max(1, min(ens.n, (64 << 20) // (8 * ens.d)))
        st = self.staging(rows * ens.d)
        check(self.lib.ree_download_ensemble(ens.X.data_ptr(), out.ctypes.data, ens.n, ens.d, ens.ld, st.data_ptr(),
                                             st.numel(), self.stream), "ree_download_ensemble")
        return out

    def download_vector(self, t: torch.Tensor, n: int) -> np.ndarray:
        """First n values of a device N-vector as a new host array (``ree_download_vector``)."""
        out = np.empty(n, dtype=np.float64)
        check(self.lib.ree_download_vector(t.data_ptr(), out.ctypes.data, n, self.stream), "ree_download_vector")
        return out

    def upload_vector(self, v: np.ndarray, ld: int) -> torch.Tensor:
        t = self.empty(ld)
        t[: v.shape[0]].copy_(torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)))
        return t

    def philox_normal(self, n: int, d: int, seed: int, draw: int, first: int = 0, out: Optional[Ensemble] = None):
        """Device N(0,1) block with rank-invariant counters (global particle index)."""
        ens = out if out is not None else self.alloc_ensemble(n, d, first)
        check(self.lib.ree_philox_normal(ens.X.data_ptr(), n, d, ens.ld, seed & (2**64 - 1), draw, first,
                                         self.stream), "ree_philox_normal")
        return ens

    # -- evaluation ------------------------------------------------------------ #
    def lsf_eval(self, lsf: DeviceLsf, ens: Ensemble, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        g = out if out is not None else self.empty(ens.ld)
        s = lsf.struct()
        check(self.lib.ree_lsf_eval(C.byref(s), ens.X.data_ptr(), ens.n, ens.ld, g.data_ptr(), self.stream),
              "ree_lsf_eval")
        return g

    def energy(self, ens: Ensemble, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        e = out if out is not None else self.empty(ens.ld)
        check(self.lib.ree_energy(ens.X.data_ptr(), ens.n, ens.d, ens.ld, e.data_ptr(), self.stream), "ree_energy")
        return e

    def logtgt(self, g, e, n, sigma, kind, out=None) -> torch.Tensor:
        out = out if out is not None else self.empty(g.numel())
        check(self.lib.ree_logtgt(g.data_ptr(), e.data_ptr() if e is not None else None, n, float(sigma), kind,
                                  out.data_ptr(), self.stream), "ree_logtgt")
        return out

    # -- reductions (return device 4-vectors, already combined across ranks) ----- #
    def reduce_ess(self, ens: Ensemble, sigma, beta, kind, out=None, use_e=True, a_out=None) -> torch.Tensor:
        """`a_out` (ld doubles) receives the log-weights beta*l the split path's moments kernel consumes."""
        out = out if out is not None else self.empty(4)
        ws = self.workspace(ens.d)
        check(self.lib.ree_reduce_ess(ens.g.data_ptr(), ens.e.data_ptr() if use_e else None, ens.n, float(sigma),
                                      float(beta), kind, a_out.data_ptr() if a_out is not None else None,
                                      ws.data_ptr(), out.data_ptr(), self.stream), "ree_reduce_ess")
        return self.comm.combine_expsums_(out)

    def reduce_cvar(self, ens: Ensemble, sigma_new, sigma_old, kind, out=None) -> torch.Tensor:
        out = out if out is not None else self.empty(4)
        ws = self.workspace(ens.d)
        check(self.lib.ree_reduce_cvar(ens.g.data_ptr(), ens.n, float(sigma_new), float(sigma_old), kind,
                                       ws.data_ptr(), out.data_ptr(), self.stream), "ree_reduce_cvar")
        return self.comm.combine_expsums_(out)

    def reduce_scaled(self, ell: torch.Tensor, n: int, d: int, scale: float, out=None, range2=None) -> torch.Tensor:
        """[max, sum exp, sum exp^2, count] of scale*ell (a log-target vector from :meth:`logtgt`); ``range2`` (from
        :meth:`vector_range`) makes the shift known in advance."""
        out = out if out is not None else self.empty(4)
        ws = self.workspace(d)
        check(self.lib.ree_reduce_scaled(ell.data_ptr(), n, float(scale),
                                         range2.data_ptr() if range2 is not None else None, ws.data_ptr(),
                                         out.data_ptr(), self.stream), "ree_reduce_scaled")
        return self.comm.combine_expsums_(out)

    def peers_connected(self) -> bool:
        """True if cross-rank merges can run inside kernels (single rank, or NVLink mailboxes connected)."""
        comm = self.comm
        if not comm.enabled:
            return True
        if comm.peer is None:
            comm.peer = comm._connect_peers(self.device)
        return bool(comm.peer)

    def adapt(self, ens: Ensemble, tgt_kind: int, sigma: float, sigma_hi: Optional[float], cvar_tgt: float, beta: float,
              do_beta: bool, ess_tgt_n: float, overlap=None, wait: bool = True) -> Optional[np.ndarray]:
        """sigma / beta searches of one iteration in one persistent kernel (``ree_cbree_adapt``): bounded Brent over
        sigma in [sigma, sigma_hi] (skipped if ``sigma_hi`` is None), then the hybrd root find over beta (or only the ESS
        of (sigma, beta)).  Returns the 16 result doubles on the host; needs :meth:`peers_connected`."""
        p = _lib.ReeAdapt(tgt_kind, 1 if sigma_hi is not None else 0, 1 if do_beta else 0, 500, 400,
                          1 if self.comm.enabled else 0, float(sigma),
                          float(sigma_hi) if sigma_hi is not None else float(sigma), float(cvar_tgt), 1e-5, float(beta),
                          float(ess_tgt_n), 1.49012e-08, 100.0, float(np.finfo(np.float64).eps))
        ell = self.scratch("ell", ens.ld)
        out = self.scratch("adapt_out", 16)
        ws = self.workspace(ens.d)
        check(self.lib.ree_cbree_adapt(ens.g.data_ptr(), ens.e.data_ptr() if ens.e is not None else None, ens.n,
                                       C.byref(p), ell.data_ptr(), ws.data_ptr(), out.data_ptr(), self.stream),
              "ree_cbree_adapt")
        self.adapt_launches += 1  # (a search launched ahead is only adopted if no other search ran on this engine since)
        if not wait:
            # (read back behind the kernel: whoever adopts the result does not have to launch a fetch for it)
            self._adapt_async = self.to_host_async(out, 2) + (self.adapt_launches,)
            return None
        if overlap is not None:
            overlap()
        return self.fetch(out, 16)

    def adapt_result(self) -> np.ndarray:
        """The 16 result doubles of the last :meth:`adapt` launch on this device (synchronising)."""
        pend = getattr(self, "_adapt_async", None)
        if pend is not None and pend[2] == self.adapt_launches:
            pend[1].synchronize()
            return pend[0].copy()
        return self.fetch(self.scratch("adapt_out", 16), 16)

    def vector_range(self, v: torch.Tensor, n: int, d: int) -> torch.Tensor:
        """Device [max, min] over the finite entries of this rank's part of an N-vector."""
        out = self.empty(2)
        ws = self.workspace(d)
        check(self.lib.ree_vector_range(v.data_ptr(), n, ws.data_ptr(), out.data_ptr(), self.stream), "ree_vector_range")
        return out

    def count_failed(self, ens: Ensemble, out=None, reduce=True) -> torch.Tensor:
        out = out if out is not None else self.empty(1)
        ws = self.workspace(ens.d)
        check(self.lib.ree_count_failed(ens.g.data_ptr(), ens.n, ws.data_ptr(), out.data_ptr(), self.stream),
              "ree_count_failed")
        return self.comm.allreduce_sum_(out) if reduce else out

    # -- sweeps ------------------------------------------------------------------- #
    def ensemble_sums(self, ens: Ensemble, shift: Optional[torch.Tensor], sums: torch.Tensor, with_g=True):
        ws = self.workspace(ens.d)
        check(self.lib.ree_ensemble_sums(ens.X.data_ptr(), ens.n, ens.d, ens.ld,
                                         ens.g.data_ptr() if (with_g and ens.g is not None) else None,
                                         shift.data_ptr() if shift is not None else None, ws.data_ptr(),
                                         sums.data_ptr(), self.stream), "ree_ensemble_sums")
        return self.comm.allreduce_sum_(sums)

    def path(self, d: int) -> int:
        """1: fused sweeps, 2: split transform / moments / Mahalanobis kernels, 0: modular kernels."""
        return int(self.lib.ree_fused_path(int(d)))

    def cbree_moments(self, ens: Ensemble, a: Optional[torch.Tensor], wmax: Optional[torch.Tensor],
                      shift: Optional[torch.Tensor], sums_uw: torch.Tensor, weighted=True):
        """Split path: unweighted and weighted shifted sums of one ensemble in one pass.  `sums_uw` holds both
        (2 x sums_len) so that one all_reduce combines them across ranks."""
        ws = self.workspace(ens.d)
        slen = ens.d + ens.d * ens.d + 2
        su, sw = sums_uw[0:slen], sums_uw[slen:2 * slen]
        check(self.lib.ree_cbree_moments(ens.X.data_ptr(), ens.n, ens.d, ens.ld,
                                         ens.g.data_ptr() if ens.g is not None else None,
                                         a.data_ptr() if weighted else None, wmax.data_ptr() if weighted else None,
                                         shift.data_ptr() if shift is not None else None, ws.data_ptr(),
                                         su.data_ptr(), sw.data_ptr() if weighted else None, self.stream),
              "ree_cbree_moments")
        self.comm.allreduce_sum_(sums_uw if weighted else su)
        return su, sw

    def cbree_step(self, src: Ensemble, dst: Ensemble, t: float, m_noise: torch.Tensor, F: torch.Tensor,
                   lsf: Optional[DeviceLsf], shift: Optional[torch.Tensor], sums: Optional[torch.Tensor],
                   z: Optional[Ensemble] = None, seed: int = 0, draw: int = 0, reduce=True):
        ws = self.workspace(src.d)
        s = lsf.struct() if lsf is not None else None
        mode = _lib.NOISE_INJECT if z is not None else _lib.NOISE_PHILOX
        if z is not None and z.ld != src.ld:
            raise ValueError("noise block must share the ensemble's leading dimension")
        check(self.lib.ree_cbree_step(src.X.data_ptr(), dst.X.data_ptr(), src.n, src.d, src.ld, float(t),
                                      m_noise.data_ptr(), F.data_ptr(), mode,
                                      z.X.data_ptr() if z is not None else None, seed & (2**64 - 1), draw,
                                      src.first, C.byref(s) if s is not None else None,
                                      shift.data_ptr() if shift is not None else None, dst.g.data_ptr(),
                                      dst.e.data_ptr(), ws.data_ptr(), sums.data_ptr() if sums is not None else None,
                                      self.stream), "ree_cbree_step")
        if reduce and sums is not None:
            self.comm.allreduce_sum_(sums)
        return sums

    # -- von Mises-Fisher-Nakagami model (resampling branch) ------------------------------------ #
    def vmfn_fit_sums(self, ens: Ensemble, w: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[sum w x/|x| (d) | sum w |x|^2 | sum w |x|^4 | sum w] over all ranks (EMvMFNM.py:169-199)."""
        sums = self.empty(ens.d + 3)
        rinv = self.scratch("vmfn_rinv", ens.ld)
        ws = self.workspace(ens.d)
        check(self.lib.ree_vmfn_fit_sums(ens.X.data_ptr(), ens.n, ens.d, ens.ld, w.data_ptr() if w is not None else None,
                                         rinv.data_ptr(), ws.data_ptr(), sums.data_ptr(), self.stream),
              "ree_vmfn_fit_sums")
        return self.comm.allreduce_sum_(sums)

    def vmfn_sample(self, dst: Ensemble, hv: torch.Tensor, hb: float, kappa: float, m: float, omega: float, seed: int,
                    draw: int, r=None, w=None, v: Optional[Ensemble] = None):
        if v is not None and v.ld != dst.ld:
            raise ValueError("injected direction normals must share the ensemble's leading dimension")
        check(self.lib.ree_vmfn_sample(dst.X.data_ptr(), dst.n, dst.d, dst.ld, hv.data_ptr(), float(hb), float(kappa),
                                       float(m), float(omega), seed & (2**64 - 1), draw, dst.first,
                                       r.data_ptr() if r is not None else None,
                                       w.data_ptr() if w is not None else None,
                                       v.X.data_ptr() if v is not None else None, self.stream), "ree_vmfn_sample")
        return dst

    def vmfn_logpdf(self, ens: Ensemble, mu: torch.Tensor, kappa: float, m: float, omega: float, c0: float,
                    logq: Optional[torch.Tensor] = None, is4: Optional[torch.Tensor] = None,
                    fuse_lsf: Optional[DeviceLsf] = None):
        """``fuse_lsf`` (affine): the kernel also writes ``ens.g`` and ``ens.e`` of the points it reads."""
        ws = self.workspace(ens.d)
        s = fuse_lsf.struct() if fuse_lsf is not None else None
        need_ge = is4 is not None or fuse_lsf is not None
        check(self.lib.ree_vmfn_logpdf(ens.X.data_ptr(), ens.n, ens.d, ens.ld, mu.data_ptr(), float(kappa), float(m),
                                       float(omega), float(c0),
                                       ens.g.data_ptr() if need_ge else None,
                                       ens.e.data_ptr() if need_ge else None,
                                       logq.data_ptr() if logq is not None else None, ws.data_ptr(),
                                       is4.data_ptr() if is4 is not None else None,
                                       C.byref(s) if s is not None else None, self.stream), "ree_vmfn_logpdf")
        if is4 is not None:
            self.comm.combine_expsums_(is4)

    def moments_chol(self, sums, shift, d, ddof, weighted, mean, cov, L=None, Linv=None, stats=None):
        check(self.lib.ree_moments_chol(sums.data_ptr(), shift.data_ptr() if shift is not None else None, d, ddof,
                                        1 if weighted else 0, mean.data_ptr(), cov.data_ptr(),
                                        L.data_ptr() if L is not None else None,
                                        Linv.data_ptr() if Linv is not None else None,
                                        stats.data_ptr() if stats is not None else None, self.stream),
              "ree_moments_chol")

    def chol_scaled(self, Cmat: torch.Tensor, d: int, scale: float, L: torch.Tensor, stats: torch.Tensor):
        check(self.lib.ree_chol_scaled(Cmat.data_ptr(), d, float(scale), L.data_ptr(), stats.data_ptr(),
                                       self.stream), "ree_chol_scaled")

    def cbree_maha(self, ens: Ensemble, mean, Linv, logdet, sigma, beta, kind, wmax, sums, is4, w=None, logq=None):
        """Sweep 2.  The weighted sums come back shifted by ``mean``."""
        ws = self.workspace(ens.d)
        if w is None and sums is not None and ens.d > 103:  # two-sweep calling sequence on the modular kernels
            w = self.empty(ens.ld)
        check(self.lib.ree_cbree_maha(ens.X.data_ptr(), ens.n, ens.d, ens.ld, ens.g.data_ptr(), ens.e.data_ptr(),
                                      mean.data_ptr(), Linv.data_ptr(), logdet.data_ptr(), float(sigma),
                                      float(beta), kind, wmax.data_ptr() if wmax is not None else None,
                                      logq.data_ptr() if logq is not None else None,
                                      w.data_ptr() if w is not None else None, ws.data_ptr(),
                                      sums.data_ptr() if sums is not None else None, is4.data_ptr(), self.stream),
              "ree_cbree_maha")
        if sums is not None:  # split path (sums None): only the IS statistics cross ranks
            self.comm.allreduce_sum_(sums)
        self.comm.combine_expsums_(is4)


# --------------------------------------------------------------------------- #
# host-side formulas on the reduced 4-tuples
# --------------------------------------------------------------------------- #
def ess_from(t4) -> float:
    """(sum w)^2 / sum w^2 from [M, S1, S2, cnt]   (solver.py:523, 596)."""
    return float(t4[1]) ** 2 / float(t4[2])


def cvar_from(t4) -> float:
    """Coefficient of variation of exp(a)*mult from the shifted sums, with my_log_cvar's conventions
    (utilities.py:76-102): inf if the mean is zero or the result is NaN; mean/std over the finite entries."""
    nf = float(t4[3])
    if nf <= 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    mean = float(t4[1]) / nf
    if mean == 0:
        return math.inf
    var = float(t4[2]) / nf - mean * mean
    out = math.sqrt(max(var, 0.0)) / mean
    return math.inf if math.isnan(out) else out
