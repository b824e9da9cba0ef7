#!/bin/bash
# A/B of one environment switch: per-kernel roofline table of the short bench with and without it, plus ncu launch lists
SW=$1
SHORT="python bench.py --steps 6 --warmup 3 --no-configs --no-cpu-baseline --apm-iters 2"
for v in on off; do
  if [ $v = off ]; then export $SW=1; fi
  $SHORT > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
print("$v: value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(x["ms_total"]/d["steps"],3) for k,x in d["roofline"]["kernels"].items()})
PY
  ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/ab_launches_$v.csv $SHORT > /dev/null 2>&1
done
