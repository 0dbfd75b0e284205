for u in 32 64 128; do QPNET_GEN_KERNEL=f3 timeout 200 python tools/ab_kernels.py --utts $u --frames 60 --kernels f3 --reps 1; done
