#!/bin/bash
# Evidence run on the GPU box (through gpurun): bench lines of both arms, then -- only after the same command has exited 0
# without ncu -- the ncu launch list of a short bench run and one `ncu --set full` capture of the dominant kernel.
# Usage (on the box): bash scripts/evidence.sh TAG      -> gpurun_out/{bench_TAG.json, bench_TAG_reference.json,
#                                                           launches_TAG.csv, prof_TAG_full_raw.csv, prof_TAG_source.csv}
set -u
TAG=$1
O=gpurun_out
mkdir -p $O
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err || { echo "bench failed"; tail -5 $O/bench_$TAG.err; exit 1; }
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err
SHORT="python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --apm-iters 2"
$SHORT > $O/short_$TAG.json 2> $O/short_$TAG.err || { echo "short bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_$TAG.csv $SHORT > $O/ncu1_$TAG.log 2>&1
# the Cholesky launches of one FULL step (skip the warm-up step's): chol(K), Newton rounds, chol(M')
ncu --set full --clock-control none --import-source on -k regex:^k_chol_flow$ -s 10 -c 10 -o $O/prof_${TAG}_full -f $SHORT > $O/ncu2_$TAG.log 2>&1
ncu -i $O/prof_${TAG}_full.ncu-rep --page raw --csv > $O/prof_${TAG}_full_raw.csv 2>/dev/null
ls -la $O | tail -12
