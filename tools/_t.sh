timeout 600 python -m pytest tests -x -q -m gpu -k "train or checkpoint or backward or grad" 2>&1 | tail -4
timeout 300 python tools/train_breakdown.py 2>&1 | grep bf16
timeout 300 python tools/bench_train.py --steps 10 --warmup 3 2>/dev/null | cut -c1-200
