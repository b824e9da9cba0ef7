#!/bin/bash
# per-kernel time table of the FULL step at several batch sizes (roofline pass of bench.py: overlap off, event per launch)
for m in "$@"; do
python bench.py --chains $m --steps 8 --warmup 3 --no-configs --no-cpu-baseline --apm-iters 2 > gpurun_out/ct_$m.json 2> gpurun_out/ct_$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ct_$m.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("m=$m value %.0f ms/step %.3f |" % (d["value"], d["ms_per_step"]), " ".join("%s %.3f" % (n.replace("k_",""), x["ms_total"]/d["steps"]) for n,x in k.items()), "| sum %.3f" % sum(x["ms_total"]/d["steps"] for x in k.values()))
PY
done
