run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 4096 --warmup 256 --no-cpu --e2e-steps 20 $EXTRA > gpurun_out/bench_$tag.json 2>gpurun_out/err_$tag.log; python -c "
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('$tag', 'step_us', round(d['ms_per_step']*1e3,2), 'post_us', round(d['roofline']['launch_us'],2), 'frac', round(d['roofline']['frac'],3), 'value', round(d['value']/1e9,3))"; tail -2 gpurun_out/err_$tag.log; }
EXTRA="" run l264 LG_X=1
EXTRA="--envs 262144 --ring 8 --steps 1024" run l264_big LG_X=1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:post_physics -s 10 -c 2 python bench.py --steps 128 --warmup 32 --no-cpu --e2e-steps 8 --envs 262144 --ring 8 2>&1 | grep -E "dram__bytes|gpu__time" | head -6
