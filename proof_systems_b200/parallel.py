"""Multi-GPU MSM: the one exchange step of the path.

sum_i s_i P_i shards by points (the reference does the same on CPU with rayon::join,
poly-commitment/src/ipa.rs:652-662): every rank runs the full Pippenger on its slice and ends with one Jacobian point
(96 bytes).  NCCL has no user-defined reduction, so the final point-sum is ONE all_gather of world x 96 bytes over
NVLink followed by world-1 point additions on every rank (result replicated, identical bits on all ranks).
The message is latency-bound (~10 us); there is nothing to overlap or fuse with the kernels.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from ._lib import jacobian_sum


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of the n points owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PointSumAllGather:
    """Reusable buffers for the exchange (pinned host staging + device tensors): the per-call cost is two small async
    copies, one NCCL all_gather of world x 96 bytes and one stream synchronisation."""

    def __init__(self, curve: int, device: torch.device, group=None):
        self.curve, self.group, self.device = curve, group, device
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > 1:
            self.h_mine = torch.empty((1, 12), dtype=torch.int64).pin_memory()
            self.h_all = torch.empty((self.world, 12), dtype=torch.int64).pin_memory()
            self.d_mine = torch.empty((1, 12), dtype=torch.int64, device=device)
            self.d_all = torch.empty((self.world, 12), dtype=torch.int64, device=device)

    def __call__(self, partial_xyz: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return np.ascontiguousarray(partial_xyz, dtype=np.uint64).reshape(12)
        self.h_mine.numpy().view(np.uint64)[0] = partial_xyz
        self.d_mine.copy_(self.h_mine, non_blocking=True)
        dist.all_gather_into_tensor(self.d_all, self.d_mine, group=self.group)
        self.h_all.copy_(self.d_all, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return jacobian_sum(self.curve, self.h_all.numpy().view(np.uint64))


class ShardedMsm:
    """One MSM sharded by points over the ranks of a process group with the exchange kept on the device: every rank leaves the
    slice sums of its part in HBM (zk_msm_partial, nothing synchronised), ONE NCCL all_gather of world x c x 128 bytes is
    enqueued behind the kernels on the context's stream, and the gathered partials are added per slice on the device before
    the single read-back (zk_msm_finish_gathered).  No host round trip between the MSM and the collective.

    `stream` is the torch stream the context runs on (ctx.set_stream(stream.cuda_stream))."""

    MAX_POINTS = 4096

    def __init__(self, ctx, curve: int, device: torch.device, stream: "torch.cuda.Stream", group=None):
        self.ctx, self.curve, self.device, self.stream, self.group = ctx, curve, device, stream, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.d_mine = torch.empty((self.MAX_POINTS, 16), dtype=torch.int64, device=device)   # written by the library's stream only
        self.d_all = None

    def __call__(self, bases, scalars_ptr: int, n: int, off: int = 0, mont: bool = False, window_bits: int = 0) -> np.ndarray:
        """Every rank must end with the same (c, groups): bases with the same table window, or — without a table — an explicit
        window_bits (the default window depends on the slice length)."""
        c, groups = self.ctx.msm_partial(bases, scalars_ptr, n, self.d_mine.data_ptr(), self.MAX_POINTS, off=off, mont=mont, window_bits=window_bits)
        cnt = c * groups
        if self.world == 1:
            return self.ctx.msm_finish_gathered(self.curve, self.d_mine.data_ptr(), 1, c, groups)
        if self.d_all is None or self.d_all.shape[1] != cnt:
            self.d_all = torch.empty((self.world, cnt, 16), dtype=torch.int64, device=self.device)
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(self.d_all, self.d_mine[:cnt], group=self.group)
        return self.ctx.msm_finish_gathered(self.curve, self.d_all.data_ptr(), self.world, c, groups)


def all_gather_point_sum(curve: int, partial_xyz: np.ndarray, group=None, device: torch.device | None = None) -> np.ndarray:
    """Sum the per-rank Jacobian partials.  Works on any backend: pass device=cuda for NCCL, leave None for gloo."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.ascontiguousarray(partial_xyz, dtype=np.uint64).reshape(12)
    world = dist.get_world_size(group)
    mine = torch.from_numpy(np.ascontiguousarray(partial_xyz, dtype=np.uint64).view(np.int64).reshape(1, 12))
    if device is not None:
        mine = mine.to(device)
    gathered = torch.empty((world, 12), dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    return jacobian_sum(curve, gathered.cpu().numpy().view(np.uint64))


class LibraryComm:
    """The exchange owned by the library (csrc/comm.cu): one NCCL communicator per context, created from a 128-byte unique id that
    rank 0 obtains from zk_comm_unique_id and the ranks share through `torch.distributed` (plumbing only: a broadcast of 128
    bytes at start-up); afterwards zk_msm_sharded runs the MSM kernels, ncclAllGather and the cross-rank sum from C on the
    context's stream — no Python, no torch collective on the data path."""

    def __init__(self, ctx, group=None):
        import ctypes

        from ._lib import check, lib
        self.ctx = ctx
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            check(lib().zk_comm_unique_id(ctypes.c_void_p(uid.ctypes.data)))
        if self.world > 1:
            t = torch.from_numpy(uid).cuda()
            dist.broadcast(t, src=0, group=group)
            uid = t.cpu().numpy()
        self._h = ctypes.c_void_p()
        check(lib().zk_comm_init_rank(ctx._h, ctypes.c_void_p(uid.ctypes.data), self.world, self.rank, ctypes.byref(self._h)))

    def msm(self, bases, scalars_ptr: int, n: int, off: int = 0, mont: bool = False, window_bits: int = 0) -> np.ndarray:
        """one MSM over world x n points (this rank's slice); collective; Jacobian [12], identical on every rank"""
        import ctypes

        from ._lib import _u64p, check, lib
        out = np.empty(12, dtype=np.uint64)
        check(lib().zk_msm_sharded(self._h, bases._h, off, n, ctypes.c_void_p(scalars_ptr), int(mont), window_bits, out.ctypes.data_as(_u64p)))
        return out

    def close(self):
        from ._lib import lib
        if getattr(self, "_h", None):
            lib().zk_comm_destroy(self._h)
            self._h = None
