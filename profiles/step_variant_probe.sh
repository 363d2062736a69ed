#!/bin/bash
# usage: profiles/step_variant_probe.sh lib1.so lib2.so ...   (prints single-step timings from bench.py extras)
for lib in "$@"; do
  SUPPLYNET_LIB=$lib python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); a=d['single_step_65536']; b=d['single_step_4M']
print('$lib', 'us/launch 65536: state', round(a['state_only']['us_per_launch'],2), 'obs', round(a['with_obs']['us_per_launch'],2), '| 4M: state', round(b['state_only']['us_per_launch'],1), 'obs', round(b['with_obs']['us_per_launch'],1))"
done
