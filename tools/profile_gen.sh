#!/bin/bash
# ncu evidence for the generator (run under gpurun, one GPU):
#   1. launch list of a short bench run (per-launch gpu__time_duration, cold-cache, serialised)
#   2. two --set full captures of the generator kernel (549 and 2199 sample steps) -> steady-state DRAM / L2 traffic per step
# The persistent kernel is launched without the cooperative attribute under the profiler (QPNET_GEN_NOCOOP).
set -x
out=gpurun_out
export QPNET_GEN_NOCOOP=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r01h_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --frames 100 --no-cpu-baseline --no-train > $out/r01h_bench_under_ncu.log 2>&1
for f in 5 20; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:f2_gen_kernel -c 1 -f -o $out/r01h_fold2_f$f \
      python tools/gen_once.py --frames $f > $out/r01h_ncu_f$f.log 2>&1
  ncu -i $out/r01h_fold2_f$f.ncu-rep --page raw --csv > $out/r01h_fold2_f$f.raw.csv 2>/dev/null
done
ls -la $out | tail -12
