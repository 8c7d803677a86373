# Round profile: default bench line, reference arm, ncu launch list, ncu full capture (run under gpurun).
set -x
R=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python bench.py --impl reference --steps 2048 --warmup 8 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 128 --warmup 32 --no-cpu --e2e-steps 8 > gpurun_out/ncu_list_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"post_physics|pre_physics" -s 40 -c 4 -f -o gpurun_out/prof_$R \
    python bench.py --steps 128 --warmup 32 --no-cpu --e2e-steps 8 > gpurun_out/ncu_full_$R.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_$R.csv
tail -c 2500 gpurun_out/bench_$R.json; tail -c 1200 gpurun_out/bench_ref_$R.json; tail -3 gpurun_out/bench_$R.err
