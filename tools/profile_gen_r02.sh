#!/bin/bash
# ncu evidence for the generators, round 2 (run under gpurun, one GPU):
#   1. launch list of a short bench run (per-launch gpu__time_duration, cold-cache, serialised)
#   2. per kernel (tcgen05 f3x2 at 256 utterances, f3 at 128, mma.sync fold2 at 32): two light captures (5 and 20 frames) ->
#      steady-state DRAM traffic per sample step; one --set full capture of the longer run
# The persistent kernels are launched without the cooperative attribute under the profiler (QPNET_GEN_NOCOOP).
out=gpurun_out
tag=${1:-r02}
export QPNET_GEN_NOCOOP=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --frames 60 --no-cpu-baseline --no-train --no-extras > $out/${tag}_bench_under_ncu.log 2>&1
for spec in "f3x2 256 f3x2_gen_kernel" "f3 128 f3_gen_kernel" "fold2 32 f2_gen_kernel"; do
  set -- $spec
  for f in 5 20; do
    QPNET_GEN_KERNEL=$1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none -k regex:$3 -c 1 --csv --log-file $out/${tag}_traffic_$1_f$f.csv python tools/gen_once.py --utts $2 --frames $f > /dev/null 2>&1
  done
  QPNET_GEN_KERNEL=$1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -c 1 -f -o $out/${tag}_$1_full \
      python tools/gen_once.py --utts $2 --frames 20 > $out/${tag}_ncu_$1.log 2>&1
  ncu -i $out/${tag}_$1_full.ncu-rep --page raw --csv > $out/${tag}_$1_full.raw.csv 2>/dev/null
done
python - "$tag" <<'PY'
import csv, json, sys
tag = sys.argv[1]
def load(p):
    d = {}
    for r in csv.reader(open(p)):
        if len(r) > 14 and r[0].isdigit():
            try:
                d[r[12]] = (float(r[14].replace(",", "")), r[13])
            except ValueError:      # "n/a": the metric does not exist on this chip
                pass
    return d
def val(x):
    v, u = x
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
res = {"kernels": {}}
for name, kern, utts in (("f3x2", "qp::f3x2::f3x2_gen_kernel", 256), ("f3", "qp::f3::f3_gen_kernel", 128), ("fold2", "qp::f2::f2_gen_kernel", 32)):
    try:
        a, b = load(f"gpurun_out/{tag}_traffic_{name}_f5.csv"), load(f"gpurun_out/{tag}_traffic_{name}_f20.csv")
        steps = (20 - 5) * 110
        rd = (val(b["dram__bytes_read.sum"]) - val(a["dram__bytes_read.sum"])) / steps
        wr = (val(b["dram__bytes_write.sum"]) - val(a["dram__bytes_write.sum"])) / steps
        res["kernels"][kern] = {"utterances": utts, "dram_bytes_per_step": rd + wr, "dram_read_per_step": rd, "dram_write_per_step": wr,
                                "l2_sector_hit_rate_pct": b["lts__t_sector_hit_rate.pct"][0],
                                "us_per_step_under_ncu": (val(b["gpu__time_duration.sum"]) - val(a["gpu__time_duration.sum"])) / steps / 1e3,
                                "tensor_pipe_inst_per_step": (b["sm__inst_executed_pipe_tensor.sum"][0] - a["sm__inst_executed_pipe_tensor.sum"][0]) / steps,
                                "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of two captures (tools/profile_gen_r02.sh: 5 and 20 frames), difference per sample step"}
        print(name, res["kernels"][kern])
    except Exception as e:
        print(name, "failed:", repr(e))
json.dump(res, open(f"gpurun_out/{tag}_gen_kernel_traffic.json", "w"), indent=1)
PY
ls -la $out | tail -14
