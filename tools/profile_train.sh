#!/bin/bash
# ncu evidence for the training step (run under gpurun, one GPU): launch list of one step + full captures of the tcgen05 kernels
out=gpurun_out; tag=${1:-r01j}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches_train_step.csv \
    python tools/train_once.py > $out/${tag}_train_once.log 2>&1
for k in tc_wgrad_kernel tc_gemm_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 6 -f -o $out/${tag}_$k \
      python tools/train_once.py > $out/${tag}_ncu_$k.log 2>&1
  ncu -i $out/${tag}_$k.ncu-rep --page raw --csv > $out/${tag}_$k.raw.csv 2>/dev/null
done
