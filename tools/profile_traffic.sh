#!/bin/bash
# steady-state DRAM traffic of the generator: two light ncu captures (549 and 2199 sample steps), difference per step
out=gpurun_out
export QPNET_GEN_NOCOOP=1
for f in 5 20; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none \
      -k regex:f2_gen_kernel -c 1 --csv --log-file $out/traffic_f$f.csv python tools/gen_once.py --frames $f > /dev/null 2>&1
done
python - <<'PY'
import csv
def load(p):
    d = {}
    for r in csv.reader(open(p)):
        if len(r) > 14 and r[0].isdigit():
            d[r[12]] = (float(r[14].replace(",", "")), r[13])
    return d
a, b = load("gpurun_out/traffic_f5.csv"), load("gpurun_out/traffic_f20.csv")
def gb(x):
    v, u = x
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
steps = (20 - 5) * 110
rd = (gb(b["dram__bytes_read.sum"]) - gb(a["dram__bytes_read.sum"])) / steps
wr = (gb(b["dram__bytes_write.sum"]) - gb(a["dram__bytes_write.sum"])) / steps
print("per sample step: dram read %.1f MB, write %.1f MB; L2 hit rate %.1f %% ; kernel %s" % (rd / 1e6, wr / 1e6, b["lts__t_sector_hit_rate.pct"][0], b["gpu__time_duration.sum"]))
PY
