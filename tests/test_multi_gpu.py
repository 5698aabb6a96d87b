"""Locus shards on two GPUs (needs a box with at least two): libpedfac_cuda's own cross-GPU sum
(pf_comm_*, peer-memory exchange inside xreduce_kernel) against an unsharded engine, and the
one-process-per-GPU mode of the `pedigraph` executable against a single-GPU run of the same chain."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

from pedfac_b200 import frontend, synth

ROOT = Path(__file__).resolve().parent.parent
GPU = ROOT / "pedfac_b200" / "bin" / "pedigraph"


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs_two = pytest.mark.skipif(gpu_count() < 2, reason="needs two GPUs on one node")


@pytest.mark.gpu
@needs_two
def test_sharded_engines_sum_across_gpus_like_one_engine():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(ROOT / "tools" / "multi_gpu_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert r.stdout.count("multi-gpu check ok") == 2


@pytest.mark.gpu
@needs_two
def test_pedigraph_on_two_gpus_samples_the_same_chain(tmp_path):
    d = synth.simulate(220, 300, soft=True, seed=46)
    outs = []
    for tag, env in (("one", {}), ("two", {"PEDFAC_GPUS": "2"})):
        rd = tmp_path / tag
        frontend.write_run_dir(d, rd)
        r = subprocess.run([str(GPU), "-d", str(rd), "-n", "2", "-r", str(rd / "rand")], capture_output=True, text=True, timeout=900,
                           env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr[-1500:]
        outs.append(((rd / "out" / "ped.txt").read_text(), (rd / "out" / "ll.txt").read_text()))
        if tag == "two":       # every rank made the same moves
            assert (rd / "out" / ".shard1" / "ped.txt").read_text() == outs[-1][0]
    assert outs[0] == outs[1]
