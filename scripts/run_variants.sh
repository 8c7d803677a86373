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
run c2_tma LG_TMA=1 --steps 4096 --warmup 128
run c2_ldg LG_TMA=0 --steps 4096 --warmup 128
run c2_tma LG_TMA=1 --steps 4096 --warmup 128
run c2_ldg LG_TMA=0 --steps 4096 --warmup 128
run big_tma2 LG_TMA=1 --envs 262144 --steps 512 --warmup 64 --ring 4
run big_tma1 LG_STREAM=0 --envs 262144 --steps 512 --warmup 64 --ring 4
run big_ldg LG_TMA=0 --envs 262144 --steps 512 --warmup 64 --ring 4
run c3_tma2 LG_TMA=1 --workload c3 --steps 1024 --warmup 64
run c3_tma1 LG_STREAM=0 --workload c3 --steps 1024 --warmup 64
run c3_ldg LG_TMA=0 --workload c3 --steps 1024 --warmup 64
run c2sym_tma LG_TMA=1 --workload c2sym --steps 4096 --warmup 128
run c2sym_ldg LG_TMA=0 --workload c2sym --steps 4096 --warmup 128
