import struct, os
import numpy as np
import bench
from pedfac_b200 import synth
n, S, soft, _, _ = bench.WORKLOADS["B"]
d = synth.simulate(n, S, soft=soft, seed=46, locus_range=(0, 64))
tr = bench.Trace(d, 48)
os.makedirs("/tmp/planbench", exist_ok=True)
with open("/tmp/planbench/views.bin", "wb") as f:
    f.write(struct.pack("<iii", len(tr.steps), tr.cap_i, tr.cap_m))
    for v, k, p in tr.steps:
        arrs = [v.indiv, v.indiv_cc, v.indiv_founder.astype(np.int32), v.marr, v.marr_pa, v.marr_ma,
                v.marr_selfing.astype(np.int32), v.marr_kid_begin, v.marr_kids]
        f.write(struct.pack("<i", v.n_cc))
        for a in arrs:
            a = np.ascontiguousarray(a, dtype=np.int32)
            f.write(struct.pack("<i", len(a))); f.write(a.tobytes())
print("wrote", len(tr.steps), "views")
