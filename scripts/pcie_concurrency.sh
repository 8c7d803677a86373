#!/bin/bash
# Host-link probe under concurrency (run under `gpurun --gpus 8`): scripts/pcie_probe on 1, 2, 4 and 8 GPUs at the same
# time — contiguous vs 52-byte-row uploads, the contiguous download, and both directions together — to see what the
# end-to-end (host-buffer) step of N ranks on one host is bound by.  Output: gpurun_out/pcie_concurrency.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/pcie_probe.cu -o gpurun_out/pcie_probe || exit 1
OUT=gpurun_out/pcie_concurrency.txt
: > $OUT
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  for i in $(seq 0 $((n - 1))); do CUDA_VISIBLE_DEVICES=$i gpurun_out/pcie_probe 16384 > gpurun_out/pcie_$n.$i.txt 2>&1 & done
  wait
  echo "== $n GPUs at once (per-GPU figures, us; min .. max over the GPUs)" >> $OUT
  python - $n >> $OUT <<'PY'
import glob, re, sys
n = int(sys.argv[1])
rows = {}
for f in sorted(glob.glob(f"gpurun_out/pcie_{n}.*.txt")):
    for line in open(f):
        m = re.match(r"(.*?): ([0-9.]+) us", line.strip())
        if m:
            rows.setdefault(m.group(1), []).append(float(m.group(2)))
keep = ("small (dof", "root 2D obj row", "body full", "body 3x 2D tips", "D2H contiguous", "upload as shipped", "duplex upload+download", "lean upload, one stream", "duplex lean")
for k, v in rows.items():
    if k.startswith(keep):
        print(f"  {k:<48s} {min(v):8.1f} .. {max(v):8.1f}")
PY
done
cat $OUT
