"""Locus sharding across the GPUs of one box: every rank owns a contiguous block
of loci, runs the identical graph view / proposal list, and the per-component and per-proposal
partial log-sums are combined with one fp64 sum all-reduce (NCCL over NVLink for CUDA engines;
the same code path runs over gloo with host tensors in the CPU tests)."""
from __future__ import annotations

import numpy as np

LOCUS_BLOCK = 50   # shard boundaries fall on multiples of this, so data generation can be blocked


def shard_range(n_loci: int, rank: int, world: int):
    nb = (n_loci + LOCUS_BLOCK - 1) // LOCUS_BLOCK
    b0, b1 = rank * nb // world, (rank + 1) * nb // world
    return min(b0 * LOCUS_BLOCK, n_loci), min(b1 * LOCUS_BLOCK, n_loci)


class ShardedEngine:
    """An engine on this rank's locus shard + the all-reduce that makes its results global.

    `engine` is any object with the CEngine interface. With `device_buffers=True` (CUDA engine,
    NCCL process group) results stay on the device between the kernel and the collective."""

    def __init__(self, engine, group=None, device_buffers: bool = False):
        import torch
        import torch.distributed as dist
        self.engine, self.group, self.dist, self.torch = engine, group, dist, torch
        self.device_buffers = device_buffers
        self._buf = None

    def _reduce_host(self, arr: np.ndarray) -> np.ndarray:
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
        if self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.numpy()

    def _dev_buf(self, n: int):
        if self._buf is None or self._buf.numel() < n:
            self._buf = self.torch.empty(max(n, 1024), dtype=self.torch.float64, device="cuda")
        return self._buf

    def step(self, view, kid: int, props: np.ndarray, chain: int = 0):
        """One Gibbs-step pass: sweep `view`, evaluate `props` for `kid`; returns (ll_cc, lsum),
        both global (summed over all locus shards) with a single collective."""
        n_cc, n_p = view.n_cc, len(props)
        if self.device_buffers:
            buf = self._dev_buf(n_cc + n_p)
            self.engine.sweep_dev(view, buf.data_ptr(), chain=chain)
            self.engine.eval_dev(kid, props, buf.data_ptr() + 8 * n_cc, chain=chain)
            seg = buf[: n_cc + n_p]
            if self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1:
                self.dist.all_reduce(seg, op=self.dist.ReduceOp.SUM, group=self.group)
            out = seg.cpu().numpy()
        else:
            part = np.concatenate([self.engine.sweep(view, chain=chain), self.engine.eval(kid, props, chain=chain)])
            out = self._reduce_host(part)
        return out[:n_cc], out[n_cc:]

    def sweep(self, view, chain: int = 0) -> np.ndarray:
        if self.device_buffers:
            buf = self._dev_buf(view.n_cc)
            self.engine.sweep_dev(view, buf.data_ptr(), chain=chain)
            seg = buf[: view.n_cc]
            if self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1:
                self.dist.all_reduce(seg, op=self.dist.ReduceOp.SUM, group=self.group)
            return seg.cpu().numpy()
        return self._reduce_host(self.engine.sweep(view, chain=chain))

    def eval_swaps(self, view, swaps, chain: int = 0) -> np.ndarray:
        """`pf_eval_swaps` over locus shards: per-shard partial log-sums, one all-reduce."""
        from .abi import make_swaps
        swaps = make_swaps(swaps)
        if self.device_buffers:
            buf = self._dev_buf(len(swaps))
            self.engine.eval_swaps_dev(view, swaps, buf.data_ptr(), chain=chain)
            seg = buf[: len(swaps)]
            if self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1:
                self.dist.all_reduce(seg, op=self.dist.ReduceOp.SUM, group=self.group)
            return seg.cpu().numpy()
        return self._reduce_host(self.engine.eval_swaps(view, swaps, chain=chain))

    def prefilter(self, kids, as_father, cands, cand_bias=None, top_k: int = 15, chain: int = 0):
        """`pf_prefilter` over locus shards: every rank computes its dense partial scores, one
        all-reduce sums them, then they are ranked (same result on every rank). The bias is added
        on rank 0 only."""
        torch, dist = self.torch, self.dist
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        rank0 = (not multi) or dist.get_rank(self.group) == 0
        n = len(kids) * len(cands)
        buf = torch.zeros(max(n, 1), dtype=torch.float64, device="cuda" if self.device_buffers else "cpu")
        self.engine.prefilter_scores_dev(kids, as_father, cands, cand_bias if rank0 else None, buf.data_ptr(), chain=chain)
        if multi:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return self.engine.prefilter_topk_dev(kids, cands, buf.data_ptr(), top_k=top_k, chain=chain)
