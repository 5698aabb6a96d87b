#!/bin/bash
# Measurement aid: config-B bench (device-timed replay + the real sampler) for every way of cutting
# sibships into sweep tasks (PF_INLINE_KIDS kids handled by the family task, PF_CHUNK_KIDS per chunk).
for v in 4:3 4:2 3:3 3:2 2:3 2:2 2:1 1:1; do ik=${v%:*}; ck=${v#*:}
  PF_INLINE_KIDS=$ik PF_CHUNK_KIDS=$ck python bench.py --no-anchor --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=j['roofline'].get('per_kernel',{})
print('inline $ik chunk $ck value %.3fM ms %.3f e2e %.3fM sweep_us %s chain_s %s' % (j['value']/1e6, j['ms_per_step'], j['e2e']['value']/1e6, r.get('sweep',{}).get('avg_launch_us'), j.get('chain',{}).get('seconds')))"
done
