#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -8
run() { # label, env, args
  env $2 timeout 300 python bench.py ${@:3} --no-cpu --e2e-steps 8 > gpurun_out/v.json 2>gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/v.json").read().strip().splitlines()[-1])
print("$1", "step us %.2f" % (d["ms_per_step"]*1e3), "post us %.2f" % d["roofline"]["launch_us"], "frac %.3f" % d["roofline"]["frac"], "value %.3fG" % (d["value"]/1e9))
PY
}
run c2 X=0 --steps 4096 --warmup 128
run c2kp X=0 --workload c2kp --steps 4096 --warmup 128
run c5 X=0 --workload c5 --steps 2048 --warmup 128
run c3 X=0 --workload c3 --steps 1024 --warmup 64
run c3ref X=0 --workload c3ref --steps 1024 --warmup 64
run c3reset X=0 --workload c3reset --steps 1024 --warmup 64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pre_physics" -s 40 -c 2 -f -o gpurun_out/prof_c5b \
    python bench.py --workload c5 --steps 128 --warmup 32 --no-cpu --e2e-steps 8 > gpurun_out/ncu_c5b.log 2>&1
