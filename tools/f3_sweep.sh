timeout 1500 python -m pytest tests -m gpu -q -k "generator or deep_preset_full or decode" 2>&1 | tail -6
for u in 256 192 160 128; do QPNET_GEN_KERNEL=f3 timeout 200 python tools/ab_kernels.py --utts $u --frames 60 --kernels f3 --reps 1; done
