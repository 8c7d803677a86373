#!/bin/bash
# Same-box A/B of two builds of libleibniz_b200.so (run under gpurun; box-to-box noise is ~2 %, same-box ~0.2 %).
#   mkdir ab; cp leibnizgym_b200/libleibniz_b200.so ab/new.so; git stash; python -m leibnizgym_b200.build;
#   cp leibnizgym_b200/libleibniz_b200.so ab/base.so; git stash pop; gpurun -- 'bash scripts/run_variants.sh'
# ab/ is git-ignored but travels to the GPU box.  The library under test is restored to ab/new.so at the end.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # label, args
  timeout 300 python bench.py ${@:2} --no-cpu --e2e-steps 8 > gpurun_out/v.json 2>gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/v.json").read().strip().splitlines()[-1])
print("$1", "step us %.2f" % (d["ms_per_step"]*1e3), "post us %.2f" % d["roofline"]["launch_us"], "frac %.3f" % d["roofline"]["frac"], "value %.3fG" % (d["value"]/1e9))
PY
}
for rep in 1 2; do
  for v in base new; do
    cp ab/$v.so leibnizgym_b200/libleibniz_b200.so
    run c2_$v --steps 8192 --warmup 256
  done
done
for v in base new; do
  cp ab/$v.so leibnizgym_b200/libleibniz_b200.so
  run c5_$v --workload c5 --steps 2048 --warmup 64
  run c4_$v --workload c4 --steps 2048 --warmup 64
  run c3ref_$v --workload c3ref --steps 1024 --warmup 64
  run big_$v --envs 262144 --steps 512 --warmup 64 --ring 4
done
cp ab/new.so leibnizgym_b200/libleibniz_b200.so
