#!/bin/bash
# Quick check + A/B timing of the tcgen05 generators after a kernel edit (run under gpurun, one GPU, ~1 minute):
#   1. the symbol-identity test (a 259-utterance batch on f3x2 + f3 against solo launches on f3) and the 256-utterance
#      forced-logit case against the oracle
#   2. us per sample step at 256 / 192 utterances (f3x2), 128 / 64 (f3) and 32 (fold2)
timeout 600 python -m pytest tests -m gpu -q -k "large_batch or (large_group and 256)" 2>&1 | tail -2
for u in 256 192; do QPNET_GEN_KERNEL=f3x2 timeout 200 python tools/ab_kernels.py --utts $u --frames 60 --kernels f3x2 --reps 1; done
for u in 128 64; do QPNET_GEN_KERNEL=f3 timeout 200 python tools/ab_kernels.py --utts $u --frames 60 --kernels f3 --reps 1; done
timeout 200 python tools/ab_kernels.py --utts 32 --frames 60 --kernels fold2 --reps 1
