for cfg in "1 0" "1 100" "1 300" "3 0" "3 200" "0 0" "2 0"; do set -- $cfg; echo "== poll_all=$1 backoff=$2"; for u in 32 128; do QPNET_F3_POLL_ALL=$1 QPNET_F3_BACKOFF=$2 timeout 200 python tools/ab_kernels.py --utts $u --frames 60 --kernels f3 --reps 1; done; done
echo "== B sweep (default)"; for u in 64 96; do timeout 200 python tools/ab_kernels.py --utts $u --frames 60 --kernels f3 --reps 1; done
QPNET_F3_POLL_ALL=1 timeout 200 python tools/f3_trace.py --utts 32 --frames 20 | tail -64
