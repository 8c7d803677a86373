#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8
for c in 1 2; do
  python bench.py --steps 2048 --warmup 64 --no-cpu --e2e-steps 400 --e2e-chunks $c > gpurun_out/e2e_c$c.json 2>gpurun_out/e2e_c$c.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_c$c.json").read().strip().splitlines()[-1])
print("chunks $c", "e2e %.2fM" % (d["e2e"]["value"]/1e6), "ms/step %.3f" % d["e2e"]["ms_per_step"], "value %.3fG" % (d["value"]/1e9), d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"])
PY
done
