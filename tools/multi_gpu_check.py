"""Locus shards on several GPUs, checked against one unsharded engine (run under torchrun, one rank per
GPU; the process group is gloo — it only carries the 64-byte mailbox handles, the sums themselves go
through libpedfac_cuda's own peer-memory exchange):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_check.py

Every result that is a sum over loci must equal the unsharded engine's (1e-12 relative) and be
bit-identical on all ranks (the ranks' host logic has to stay in lock step)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import torch.distributed as dist
from pedfac_b200 import Engine, synth, PF_PROP_TRIO, PF_PROP_PRONG
from pedfac_b200.abi import make_proposals
from pedfac_b200.sharding import ShardedEngine, shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
S = 800
for soft in (False, True):
    d = synth.simulate(260, S, soft=soft, seed=46)
    cap_i, cap_m = d.n_total + 8, d.n_total
    s0, s1 = shard_range(S, rank, world)
    eng = Engine(S, cap_i, cap_m, locus_range=(s0, s1), device=local)
    synth.upload(eng, d)
    sh = ShardedEngine(eng, native=True)
    full = Engine(S, cap_i, cap_m, device=local)
    synth.upload(full, d)
    ped = synth.truth_pedigree(d, cap_i, cap_m)
    labels, _ = ped.components()
    view, _ = ped.view(labels)
    ll_s, ll_f = sh.sweep(view), full.sweep(view)
    np.testing.assert_allclose(ll_s, ll_f, rtol=1e-12, atol=1e-11)
    kid = int(d.slot[int(np.flatnonzero(d.pa >= 0)[0])])
    others = [int(s) for s in view.indiv if labels[s] != labels[kid]][:40]
    props = [(PF_PROP_TRIO, -1, -1, -1)] + [(PF_PROP_TRIO, x, -1, -1) for x in others] + \
            [(PF_PROP_TRIO, others[i], others[i + 1], -1) for i in range(0, len(others) - 1, 2)] + \
            [(PF_PROP_PRONG, -1, -1, mm) for mm in list(ped.marriages)[:8] if mm != ped.parents_marriage_of(kid)]
    props = make_proposals(props)
    (cc_s, ls_s), (cc_f, ls_f) = sh.step(view, kid, props), full.sweep_eval(view, kid, props)
    np.testing.assert_allclose(cc_s, cc_f, rtol=1e-12, atol=1e-11); np.testing.assert_allclose(ls_s, ls_f, rtol=1e-12, atol=1e-11)
    # device-resident variants leave the global sums on the device as well
    buf = torch.zeros(len(props), dtype=torch.float64, device="cuda")
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.eval_dev(kid, props, buf.data_ptr()); torch.cuda.synchronize()
    np.testing.assert_allclose(buf.cpu().numpy(), ls_f, rtol=1e-12, atol=1e-11)
    # one submission, one cross-GPU sum, issued on the second stream and fenced
    eng.comm_set_async(True)
    buf2 = torch.zeros(view.n_cc + len(props), dtype=torch.float64, device="cuda")
    for _ in range(3):
        eng.sweep_eval_dev(view, kid, props, buf2.data_ptr())
    eng.comm_fence(); torch.cuda.synchronize()
    np.testing.assert_allclose(buf2.cpu().numpy(), np.concatenate([cc_f, ls_f]), rtol=1e-12, atol=1e-11)
    eng.comm_set_async(False)
    # prefilter: ranking needs the scores of all loci
    obs = np.flatnonzero(d.observed)
    kids = np.array([int(d.slot[i]) for i in obs if d.gen[i] == 0][:24], dtype=np.int32)
    cands = np.array([int(d.slot[i]) for i in obs if d.gen[i] == 1][:60], dtype=np.int32)
    role = np.zeros(len(kids), dtype=np.uint8)
    bias = np.linspace(-1.0, 0.0, len(cands))
    a, b = sh.prefilter(kids, role, cands, bias, top_k=10), full.prefilter(kids, role, cands, bias, top_k=10)
    assert (a[0] == b[0]).all(), "prefilter ranking differs"
    np.testing.assert_allclose(a[1], b[1], rtol=1e-12, atol=1e-11)
    # bit-identical on every rank
    mine = np.concatenate([ll_s, cc_s, ls_s, a[1].ravel()]).tobytes()
    allb = [None] * world
    dist.all_gather_object(allb, mine)
    assert all(x == allb[0] for x in allb), "ranks disagree bitwise"
    if rank == 0:
        print(f"multi-gpu check ok: soft={soft} world={world} components {len(ll_s)} proposals {len(props)} cross-GPU sums {eng.comm_reduce_count()}", flush=True)
    del sh, eng, full
dist.barrier()
dist.destroy_process_group()
