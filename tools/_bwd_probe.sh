mkdir -p gpurun_out/p
QPNET_BWD_TC=0 timeout 300 python tools/bwd_tc_probe.py dump gpurun_out/p/m0.pt 2>&1 | tail -3
for m in 7; do
  QPNET_BWD_TC=$m timeout 300 python tools/bwd_tc_probe.py dump gpurun_out/p/m$m.pt 2>&1 | tail -3
  python tools/bwd_tc_probe.py cmp gpurun_out/p/m0.pt gpurun_out/p/m$m.pt
done
rm -rf gpurun_out/p
for m in 7; do echo "mask $m"; QPNET_BWD_TC=$m timeout 300 python tools/train_breakdown.py 2>&1 | grep -v Warn; done
timeout 600 python -m pytest tests -x -q -m gpu -k "forward or backward or train or grad" 2>&1 | tail -5
