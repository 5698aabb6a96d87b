#!/bin/bash
# Measurement aid: flat gprof profile of the sampler's host code on a config-B chain (GPU box).
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/gprof_gpu
g++ -O2 -pg -g -fno-inline-functions -std=c++17 -o gpurun_out/gprof_gpu/pedigraph_pg pedfac_b200/csrc/host/pedigraph.cpp -Iinclude -Lpedfac_b200/lib -lpedfac_cuda -Wl,-rpath,$PWD/pedfac_b200/lib -lm -pthread
python - <<'PY'
import sys, subprocess, os
from pathlib import Path
sys.path.insert(0, '.')
import bench
from pedfac_b200 import frontend, synth
n, S, soft, _, _ = bench.WORKLOADS["B"]
d = synth.simulate(n, S, soft=soft, seed=46)
rd = Path('gpurun_out/gprof_gpu/run')
frontend.write_run_dir(d, rd)
r = subprocess.run([str(Path('gpurun_out/gprof_gpu/pedigraph_pg').resolve()), '-d', str(rd.resolve()), '-n', '6', '-r', str((rd / 'rand').resolve())],
                   capture_output=True, text=True, cwd='gpurun_out/gprof_gpu', env=dict(os.environ, PEDFAC_STATS='1'))
print(r.stderr[-1500:])
PY
cd gpurun_out/gprof_gpu && gprof -b -p pedigraph_pg gmon.out | head -45 > gprof_flat.txt; cat gprof_flat.txt; rm -rf run
