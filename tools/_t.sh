timeout 300 python tools/train_breakdown.py 2>&1 | grep -v "^fp32"
QPNET_WGRAD_TMA=0 timeout 300 python tools/train_breakdown.py 2>&1 | grep -v "^fp32"
