timeout 300 python -m pytest tests -m gpu -q -x -k "free_running and full_sampling or teacher_forced and full" -s 2>&1 | grep -E "full_sampling|assert|passed|failed" | head -20
