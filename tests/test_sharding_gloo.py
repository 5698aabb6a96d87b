"""world_size-2 CPU test (gloo) of the N > 1 host path: locus shards + one all-reduce per
Gibbs-step pass must reproduce the unsharded result. The per-rank engine here is the CPU oracle
(test infrastructure standing in for the CUDA engine, which needs a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from pedfac_b200 import PF_PROP_SELFING, PF_PROP_TRIO, make_proposals
from pedfac_b200.sharding import ShardedEngine, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path[:0] = [str(root), str(root / "tests")]
    from oracle.binding import OracleEngine
    from scenarios import random_forest
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = random_forest(np.random.default_rng(3), n_indiv=40, S=230, max_kids=4)
    s0, s1 = shard_range(sc.S, rank, world)
    eng = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr, locus_range=(s0, s1))
    sc.upload(eng)
    sh = ShardedEngine(eng)
    view, _ = sc.ped.view()
    props = make_proposals([(PF_PROP_TRIO, 1, 2, -1), (PF_PROP_TRIO, -1, 7, -1), (PF_PROP_SELFING, 5, -1, -1)])
    ll_cc, lsum = sh.step(view, 0, props)
    kids, cands = np.arange(0, 12, dtype=np.int32), np.arange(5, 40, dtype=np.int32)
    role = (kids % 2).astype(np.uint8)
    bias = -np.linspace(0.0, 3.0, len(cands))
    pslot, pscore = sh.prefilter(kids, role, cands, bias, top_k=6)
    # a swap proposal on a pruned copy of the pedigree (sharded: partial sums + all-reduce)
    site = sc.ped.swap_sites()[0]
    x, z, mx, mz, m_of_x = site
    ped2 = sc.ped.copy()
    pa, ma, prev = ped2.detach_kid(x)
    ped2.founder[x] = False
    labels2, _ = ped2.components()
    sh.sweep(ped2.view(labels2)[0])
    view3, _ = ped2.view(labels2, only={int(labels2[x]), int(labels2[pa]), int(labels2[ma])})
    sw = sh.eval_swaps(view3, [(1, 0, x, z, mz, mx, pa, ma, -1 if prev is None else prev), (3, 1, x, z, mz, mx, pa, ma, -1 if prev is None else prev)])
    q.put((rank, (s0, s1), ll_cc.copy(), lsum.copy(), sh.sweep(view).copy(), pslot.copy(), pscore.copy(), sw.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_locus_sharding_matches_unsharded():
    from oracle.binding import OracleEngine
    from scenarios import random_forest
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1][0] == 0 and res[0][1][1] == res[1][1][0] and res[1][1][1] == 230
    sc = random_forest(np.random.default_rng(3), n_indiv=40, S=230, max_kids=4)
    full = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(full)
    view, _ = sc.ped.view()
    want_cc = full.sweep(view)
    want = full.eval(0, [(PF_PROP_TRIO, 1, 2, -1), (PF_PROP_TRIO, -1, 7, -1), (PF_PROP_SELFING, 5, -1, -1)])
    kids, cands = np.arange(0, 12, dtype=np.int32), np.arange(5, 40, dtype=np.int32)
    want_slot, want_score = full.prefilter(kids, (kids % 2).astype(np.uint8), cands, -np.linspace(0.0, 3.0, len(cands)), top_k=6)
    for r_ in res:
        np.testing.assert_array_equal(r_[5], want_slot)                    # sharded prefilter ranks like the unsharded one
        np.testing.assert_allclose(r_[6], want_score, rtol=1e-12)
    np.testing.assert_array_equal(res[0][6], res[1][6])
    x, z, mx, mz, m_of_x = sc.ped.swap_sites()[0]
    ped2 = sc.ped.copy()
    pa, ma, prev = ped2.detach_kid(x)
    ped2.founder[x] = False
    labels2, _ = ped2.components()
    full.sweep(ped2.view(labels2)[0])
    view3, _ = ped2.view(labels2, only={int(labels2[x]), int(labels2[pa]), int(labels2[ma])})
    want_sw = full.eval_swaps(view3, [(1, 0, x, z, mz, mx, pa, ma, -1 if prev is None else prev), (3, 1, x, z, mz, mx, pa, ma, -1 if prev is None else prev)])
    for r_ in res:
        np.testing.assert_allclose(r_[7], want_sw, rtol=1e-12)
    full.sweep(view)
    for rank, _, ll_cc, lsum, sw, _ps, _pc, _sw in res:
        np.testing.assert_allclose(ll_cc, want_cc, rtol=1e-13, atol=1e-12)
        np.testing.assert_allclose(sw, want_cc, rtol=1e-13, atol=1e-12)
        np.testing.assert_allclose(lsum, want, rtol=1e-13)
    # every rank holds bit-identical global results (the chain logic relies on that)
    np.testing.assert_array_equal(res[0][3], res[1][3])
    np.testing.assert_array_equal(res[0][2], res[1][2])


def test_shard_ranges_tile_the_loci():
    for S in (1, 49, 50, 500, 2000, 10000, 10007):
        for world in (1, 2, 4, 8):
            if (S + 49) // 50 < world:                     # fewer 50-locus blocks than ranks: refused, not an empty shard
                with pytest.raises(ValueError):
                    shard_range(S, 0, world)
                continue
            r = [shard_range(S, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == S
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
