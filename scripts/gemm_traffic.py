"""DRAM traffic of the tcgen05 GEMM launches of one training step, from an ncu metrics CSV
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_bf16_tc_kernel --csv`
around `bench.py --steps 1 --warmup 1 --no-graph`).  Writes the per-launch averages bench.py reports as
`roofline.traffic`.   usage: gemm_traffic.py launches.csv out.json [launches_per_step]"""
import collections
import csv
import json
import re
import sys

path, out = sys.argv[1], sys.argv[2]
per_step = int(sys.argv[3]) if len(sys.argv) > 3 else 54
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
rows = collections.OrderedDict()
for row in csv.DictReader(l for l in open(path) if not l.startswith("==")):
    d = rows.setdefault(int(row["ID"]), {"kernel": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void <unnamed>::", "")})
    d[row["Metric Name"]] = float(row["Metric Value"].replace(",", "")) * UNIT[row["Metric Unit"]]
launches = list(rows.values())
# whole steps only, counted from the end (the first launches belong to set-up / capture-free warm-up of the same step mix)
n = (len(launches) // per_step) * per_step
launches = launches[len(launches) - n:]
tot_r = sum(l["dram__bytes_read.sum"] for l in launches)
tot_w = sum(l["dram__bytes_write.sum"] for l in launches)
tot_t = sum(l["gpu__time_duration.sum"] for l in launches)
by = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for l in launches:
    b = by[l["kernel"]]
    b[0] += 1
    b[1] += l["dram__bytes_read.sum"]
    b[2] += l["dram__bytes_write.sum"]
    b[3] += l["gpu__time_duration.sum"]
res = {"source": path, "launches": n, "launches_per_step": per_step,
       "dram_bytes_per_launch": (tot_r + tot_w) / n, "dram_read_bytes_per_launch": tot_r / n,
       "dram_write_bytes_per_launch": tot_w / n, "us_per_launch_under_ncu": tot_t / n,
       "dram_gbs_under_ncu": (tot_r + tot_w) / (tot_t * 1e-6) / 1e9,
       "by_kernel": {k: {"launches": v[0], "read_MB_per_launch": v[1] / v[0] / 1e6, "write_MB_per_launch": v[2] / v[0] / 1e6,
                         "us_per_launch": v[3] / v[0]} for k, v in by.items()}}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
