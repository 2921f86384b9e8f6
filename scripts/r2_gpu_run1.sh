#!/bin/bash
# Round 2, GPU call 1: parity of the closed-form Kepler path, CTA-size and variant sweeps.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
: > gpurun_out/r2a_quick.jsonl
for s in 32 16 8 0; do
  echo "{\"cta_sats\": $s}" >> gpurun_out/r2a_quick.jsonl
  PERTURB_GPU_CTA_SATS=$s python bench.py --quick --workload c2 --steps 20 --warmup 3 2>&1 | tail -1 >> gpurun_out/r2a_quick.jsonl
done
for w in c3s c4s c5s; do
  python bench.py --quick --workload $w --steps 10 --warmup 3 2>&1 | tail -1 >> gpurun_out/r2a_quick.jsonl
done
bash scripts/sweep_variants.sh "c2 c3s c4s" 10 >> gpurun_out/r2a_quick.jsonl 2>&1
cat gpurun_out/r2a_quick.jsonl | cut -c1-220
