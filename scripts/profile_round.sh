# Round profile (run under gpurun, 1 GPU, ~6 GPU-minutes): the bench line as the driver invokes it and with the default
# flags, the reference arm, one line per BASELINE config, the shard-size sweep, the ncu launch list and full captures.
# Afterwards, in the build container: python scripts/summarize_profiles.py <round>
set -x
R=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_driver_$R.json 2> gpurun_out/bench_driver_$R.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
python bench.py --all-workloads --no-e2e > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python bench.py --envs 65536 --steps 512 --warmup 64 --no-cpu --no-e2e > gpurun_out/bench_c2_65536_$R.json 2>/dev/null
python bench.py --envs 262144 --steps 256 --warmup 32 --no-cpu --no-e2e > gpurun_out/bench_c2_262144_$R.json 2>/dev/null
python bench.py --envs 1048576 --steps 64 --warmup 8 --no-cpu --no-e2e > gpurun_out/bench_c2_1048576_$R.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 64 --warmup 16 --no-cpu --no-e2e --min-timed-ms 2 > gpurun_out/ncu_list_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"post_physics|pre_physics" -s 40 -c 4 -f -o gpurun_out/prof_$R \
    python bench.py --steps 64 --warmup 16 --no-cpu --no-e2e --min-timed-ms 2 > gpurun_out/ncu_full_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"post_physics" -s 8 -c 1 -f -o gpurun_out/prof_big_$R \
    python bench.py --envs 262144 --steps 16 --warmup 4 --no-cpu --no-e2e --min-timed-ms 2 > gpurun_out/ncu_big_$R.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_$R.csv
tail -c 3500 gpurun_out/bench_driver_$R.json; tail -c 1500 gpurun_out/bench_ref_$R.json; tail -3 gpurun_out/bench_$R.err
