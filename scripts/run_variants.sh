python -m pytest tests -m gpu -q 2>&1 | tail -3
run() { tag=$1; shift; env "$@" python bench.py --steps 8192 --warmup 256 --no-cpu --e2e-steps 20 $EXTRA > gpurun_out/bench_$tag.json 2>gpurun_out/err_$tag.log; python -c "
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('$tag', 'step_us', round(d['ms_per_step']*1e3,2), 'post_us', round(d['roofline']['launch_us'],2), 'frac', round(d['roofline']['frac'],3), 'value', round(d['value']/1e9,3))"; tail -2 gpurun_out/err_$tag.log; }
EXTRA="" run v11 LG_X=1
EXTRA="--workload c5" run v11_c5 LG_X=1
EXTRA="--workload c4" run v11_c4 LG_X=1
EXTRA="--workload c3 --ring 16" run v11_c3 LG_X=1
EXTRA="--workload c2sym" run v11_c2sym LG_X=1
ncu --set full --clock-control none --import-source on -k regex:"post_physics|pre_physics" -s 40 -c 4 -f -o gpurun_out/prof_v11 python bench.py --steps 128 --warmup 32 --no-cpu --e2e-steps 8 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
