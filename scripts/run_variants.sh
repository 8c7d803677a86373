for m in "--e2e-mode copy" "--e2e-mode copy --e2e-full-upload"; do python bench.py --steps 2048 --warmup 128 --no-cpu --e2e-steps 200 $m > gpurun_out/bench_e2e.json 2>gpurun_out/err_e2e.log; python -c "
import json
d=json.load(open('gpurun_out/bench_e2e.json'))
print('$m', d['e2e'])"; tail -2 gpurun_out/err_e2e.log; done
python -m pytest tests -m gpu -q 2>&1 | tail -3
