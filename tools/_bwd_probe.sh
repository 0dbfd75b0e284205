mkdir -p gpurun_out/p
QPNET_BWD_TC=0 timeout 300 python tools/bwd_tc_probe.py dump gpurun_out/p/m0.pt 2>&1 | tail -3
for m in 1 7; do
  QPNET_BWD_TC=$m timeout 300 python tools/bwd_tc_probe.py dump gpurun_out/p/m$m.pt 2>&1 | tail -3
  python tools/bwd_tc_probe.py cmp gpurun_out/p/m0.pt gpurun_out/p/m$m.pt
done
rm -rf gpurun_out/p
for m in 0 1 3 7; do echo "mask $m"; QPNET_BWD_TC=$m timeout 300 python tools/train_breakdown.py 2>&1 | grep bf16; done
