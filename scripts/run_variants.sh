timeout 300 python -m pytest tests -m gpu -q --tb=line -x 2>&1 | tail -4
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8192 --warmup 256 --no-cpu --e2e-steps 20 $EXTRA > gpurun_out/bench_$tag.json 2>gpurun_out/err_$tag.log; python -c "
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('$tag', 'step_us', round(d['ms_per_step']*1e3,2), 'post_us', round(d['roofline']['launch_us'],2), 'frac', round(d['roofline']['frac'],3), 'value', round(d['value']/1e9,3))"; tail -2 gpurun_out/err_$tag.log; }
EXTRA="" run v16 LG_X=1
EXTRA="--workload c5" run v16_c5 LG_X=1
EXTRA="--workload c4" run v16_c4 LG_X=1
