#!/bin/bash
# the other BASELINE.json configs, one bench line each (1 GPU)
mkdir -p gpurun_out
python bench.py --mode infer --batch 512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer512.json 2>/dev/null; echo "infer512 rc=$?"
python bench.py --mode train --batch 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train256.json 2>/dev/null; echo "train256 rc=$?"
python bench.py --mode train --batch 64 --nodes 126 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stress126.json 2>/dev/null; echo "stress rc=$?"
python bench.py --mode train --batch 64 --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2>/dev/null; echo "fp32 rc=$?"
python - <<'PY'
import json
for f in ("infer512", "train256", "stress126", "fp32"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % f))
        print(f, round(d["value"]), "samples/s", round(d["ms_per_step"], 2), "ms/step  e2e", round(d["e2e"]["value"]), d["roofline"]["step_frac_of_tensor_peak"])
    except Exception as e:
        print(f, "FAILED", e)
PY
