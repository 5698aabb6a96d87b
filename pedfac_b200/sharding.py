"""Locus sharding across the GPUs of one box: every rank owns a contiguous block
of loci and runs the identical graph view / proposal list; results that are sums over loci are
added up across the ranks.

A CUDA engine does that itself: after `pf_comm_connect` (libpedfac_cuda exchanges the partial vectors
through peer-mapped mailboxes inside one kernel, include/pedfac_cuda.h) every call already returns
the sums over all shards, and this class is a thin veneer. For the CPU tests (oracle engines, gloo)
the same interface is provided with a host-side fp64 all-reduce."""
from __future__ import annotations

import numpy as np

LOCUS_BLOCK = 50   # shard boundaries fall on multiples of this, so data generation can be blocked


def shard_range(n_loci: int, rank: int, world: int):
    nb = (n_loci + LOCUS_BLOCK - 1) // LOCUS_BLOCK
    if nb < world:
        raise ValueError(f"{n_loci} loci are {nb} blocks of {LOCUS_BLOCK}: too few to shard over {world} ranks")
    b0, b1 = rank * nb // world, (rank + 1) * nb // world
    return min(b0 * LOCUS_BLOCK, n_loci), min(b1 * LOCUS_BLOCK, n_loci)


class ShardedEngine:
    """An engine on this rank's locus shard whose results are global (summed over all shards).

    `engine` is any object with the CEngine interface. `native=True` (a CUDA engine inside an
    initialised torch.distributed job): the library's own cross-GPU sum is connected and used —
    nothing here touches the results. Otherwise partial results are summed with a host all-reduce."""

    def __init__(self, engine, group=None, native: bool = False):
        import torch
        import torch.distributed as dist
        self.engine, self.group, self.dist, self.torch = engine, group, dist, torch
        self.native = native
        if native:
            if not (dist.is_initialized() and dist.get_world_size(group) > 1):
                raise RuntimeError("ShardedEngine(native=True) needs an initialised torch.distributed job with more than one rank")
            engine.comm_init_from_torch(dist, dist.get_rank(group), dist.get_world_size(group))

    def _reduce_host(self, arr: np.ndarray) -> np.ndarray:
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
        if self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.numpy()

    def step(self, view, kid: int, props: np.ndarray, chain: int = 0):
        """One Gibbs-step pass: sweep `view`, evaluate `props` for `kid`; returns (ll_cc, lsum),
        both global, with a single exchange."""
        from .abi import make_proposals
        props = make_proposals(props)
        if self.native:
            return self.engine.sweep_eval(view, kid, props, chain=chain)
        n_cc = view.n_cc
        part = np.concatenate([self.engine.sweep(view, chain=chain), self.engine.eval(kid, props, chain=chain)])
        out = self._reduce_host(part)
        return out[:n_cc], out[n_cc:]

    def sweep(self, view, chain: int = 0) -> np.ndarray:
        if self.native:
            return self.engine.sweep(view, chain=chain)
        return self._reduce_host(self.engine.sweep(view, chain=chain))

    def eval_swaps(self, view, swaps, chain: int = 0) -> np.ndarray:
        """`pf_eval_swaps` over locus shards."""
        from .abi import make_swaps
        swaps = make_swaps(swaps)
        if self.native:
            return self.engine.eval_swaps(view, swaps, chain=chain)
        return self._reduce_host(self.engine.eval_swaps(view, swaps, chain=chain))

    def prefilter(self, kids, as_father, cands, cand_bias=None, top_k: int = 15, chain: int = 0):
        """`pf_prefilter` over locus shards: every rank computes its dense partial scores, they are
        summed, then ranked (same result on every rank). The bias is added on rank 0 only."""
        if self.native:
            return self.engine.prefilter(kids, as_father, cands, cand_bias, top_k=top_k, chain=chain)
        torch, dist = self.torch, self.dist
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        rank0 = (not multi) or dist.get_rank(self.group) == 0
        n = len(kids) * len(cands)
        buf = torch.zeros(max(n, 1), dtype=torch.float64)        # host engine: host buffer
        self.engine.prefilter_scores_dev(kids, as_father, cands, cand_bias if rank0 else None, buf.data_ptr(), chain=chain)
        if multi:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return self.engine.prefilter_topk_dev(kids, cands, buf.data_ptr(), top_k=top_k, chain=chain)
