# Round profile (run under gpurun, 1 GPU): default bench line, reference arm, the other workloads, the shard-size sweep,
# ncu launch list and full captures.  Afterwards, in the build container: python scripts/summarize_profiles.py <round>
set -x
R=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python bench.py --impl reference --steps 2048 --warmup 8 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
for wl in c2sym c2kp c3 c3ref c3reset c4 c5; do
  python bench.py --workload $wl --steps 2048 --warmup 128 --no-cpu --e2e-steps 50 > gpurun_out/bench_${wl}_$R.json 2> gpurun_out/bench_${wl}_$R.err
done
python bench.py --envs 65536 --steps 2048 --warmup 128 --no-cpu --e2e-steps 8 --ring 8 > gpurun_out/bench_c2_65536_$R.json 2>/dev/null
python bench.py --envs 262144 --steps 512 --warmup 64 --no-cpu --e2e-steps 8 --ring 4 > gpurun_out/bench_c2_262144_$R.json 2>/dev/null
python bench.py --envs 1048576 --steps 128 --warmup 16 --no-cpu --e2e-steps 4 --ring 2 > gpurun_out/bench_c2_1048576_$R.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 128 --warmup 32 --no-cpu --e2e-steps 8 > gpurun_out/ncu_list_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"post_physics|pre_physics" -s 40 -c 4 -f -o gpurun_out/prof_$R \
    python bench.py --steps 128 --warmup 32 --no-cpu --e2e-steps 8 > gpurun_out/ncu_full_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"post_physics" -s 8 -c 1 -f -o gpurun_out/prof_big_$R \
    python bench.py --envs 262144 --steps 32 --warmup 8 --no-cpu --e2e-steps 4 --ring 4 > gpurun_out/ncu_big_$R.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_$R.csv
tail -c 2500 gpurun_out/bench_$R.json; tail -c 1200 gpurun_out/bench_ref_$R.json; tail -3 gpurun_out/bench_$R.err
