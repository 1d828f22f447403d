#!/bin/bash
# what the driver runs at round end, in one GPU call: the GPU suite, smoke(), both bench arms (default flags)
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/final_ref.json 2>gpurun_out/final_ref.err; tail -c 600 gpurun_out/final_ref.json; echo
python bench.py > gpurun_out/final_b200.json 2>gpurun_out/final_b200.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_b200.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, d["roofline"]["frac"], d["e2e"]["value"],
      d["e2e"].get("ingest_yuyv", {}).get("frames_per_s"), d["cpu_baseline"]["value"], d["stage_ms"])
PY
