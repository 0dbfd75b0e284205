timeout 600 python -m pytest tests -m gpu -q -k "large_batch or (large_group and 256)" 2>&1 | tail -2
for u in 256 128; do QPNET_GEN_KERNEL=f3x2 timeout 200 python tools/ab_kernels.py --utts $u --frames 60 --kernels f3x2 --reps 1 2>&1 | tail -1; done
