timeout 900 python -m pytest tests -m gpu -q -k "large_batch or large_group or every_kernel or wide_aux" 2>&1 | tail -2
for u in 128 64 33; do QPNET_GEN_KERNEL=f3 timeout 200 python tools/ab_kernels.py --utts $u --frames 60 --kernels f3 --reps 1 2>&1 | tail -1; done
timeout 200 python tools/ab_kernels.py --utts 32 --frames 60 --kernels f3,fold2 --reps 1 2>&1 | tail -2
