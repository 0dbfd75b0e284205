for w in 0 1 2 3; do echo "== wpol $w"; QPNET_F3_WPOL=$w QPNET_GEN_KERNEL=f3x2 timeout 200 python tools/ab_kernels.py --utts 256 --frames 60 --kernels f3x2 --reps 1; done
