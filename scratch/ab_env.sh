#!/bin/bash
# A/B of an environment switch on ONE box: alternate the two settings of $1 over short bench runs.
VAR=${1:-SPB_OVERLAP_FEEDBACK}
B="python bench.py --images 2048 --steps 4 --warmup 2 --no-cpu-baseline --no-e2e --no-extras"
for rep in 1 2 3; do
  for v in 0 1; do
    env $VAR=$v $B 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$VAR=$v rep$rep value %.0f ms_per_step %.1f sm_mhz %s' % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz']))"
  done
done
