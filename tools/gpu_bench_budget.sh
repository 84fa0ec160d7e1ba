#!/bin/bash
# GPU check of bench.py's wall-clock budget (1 GPU, ~4 min): our arm with the budget already spent (every optional
# leg must yield, the line must still be complete), the reference arm with a budget that forces the smaller sample.
mkdir -p gpurun_out
DPE_BENCH_BUDGET_S=1 timeout 260 python bench.py --steps 1 --warmup 1 > gpurun_out/budget_ours.txt 2> gpurun_out/budget_ours.err; echo "rc=$?" >> gpurun_out/budget_ours.txt
tail -c 600 gpurun_out/budget_ours.err; python - <<'PY'
import json
for f in ("gpurun_out/budget_ours.txt",):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print({k: d[k] for k in ("value", "ms_per_step", "value_fast_arithmetic")}, d["e2e"]["statistic"], d["e2e"]["value"], d["cpu_baseline"]["kind"], d["roofline"]["frac"])
PY
DPE_BENCH_BUDGET_S=85 timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/budget_ref.txt 2> gpurun_out/budget_ref.err; echo "rc=$?" >> gpurun_out/budget_ref.txt
tail -c 400 gpurun_out/budget_ref.err; cut -c1-700 gpurun_out/budget_ref.txt
