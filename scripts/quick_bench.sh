#!/bin/bash
# quick on-box check of a build: parity tests, the headline bench without the configs / CPU legs, native sampler statistics
TAG=$1
python -m pytest tests -m gpu -x -q -k "${2:-parity or dropin or native}" > gpurun_out/qt_$TAG.log 2>&1; tail -3 gpurun_out/qt_$TAG.log
python bench.py --no-configs --no-cpu-baseline > gpurun_out/qb_$TAG.json 2> gpurun_out/qb_$TAG.err || tail -5 gpurun_out/qb_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/qb_$TAG.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f ms/step %.3f apm %.0f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["apm_iters_per_s"]["value"], d["roofline"]["frac"]))
for k,v in d["roofline"]["kernels"].items(): print("  %-14s %6.3f ms/step share %.3f frac %s" % (k, v["ms_total"]/d["steps"], v["share_of_step"], v.get("frac")))
PY
B=256 FRACS=0.5 ITERS=60 python scripts/dev_native_sampler.py 2>&1 | tail -2
