#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # label, env, args
  env $2 timeout 300 python bench.py ${@:3} --no-cpu --e2e-steps 8 > gpurun_out/v.json 2>gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/v.json").read().strip().splitlines()[-1])
print("$1", "step us %.2f" % (d["ms_per_step"]*1e3), "post us %.2f" % d["roofline"]["launch_us"], "frac %.3f" % d["roofline"]["frac"], "value %.3fG" % (d["value"]/1e9))
PY
}
for v in base new varA varB varC; do
  cp ab/$v.so leibnizgym_b200/libleibniz_b200.so
  run c2_${v} X=0 --steps 8192 --warmup 256
  run c5_${v} X=0 --workload c5 --steps 2048 --warmup 64
  run c4_${v} X=0 --workload c4 --steps 2048 --warmup 64
done
cp ab/new.so leibnizgym_b200/libleibniz_b200.so
