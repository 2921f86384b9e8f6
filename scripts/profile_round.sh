#!/bin/bash
# One GPU box, one GPU: the bench line, the ncu launch list of the same command, ncu --set full
# of the class kernels (config 2 and the resonant profiling cut), and the full-size quick lines.
# usage (from the repo root, on the GPU box): bash scripts/profile_round.sh <tag>
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err || exit 1
python bench.py --steps 2 --warmup 1 > $O/${TAG}_bench_short.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${TAG}_c2_launches.csv python bench.py --steps 2 --warmup 1 > $O/${TAG}_ncu_launches.log 2>&1
for w in c2 c4p c3p; do
  python bench.py --quick --workload $w --steps 2 --warmup 1 > /dev/null 2>&1 || exit 1
  n=2; [ $w = c3p ] && n=5
  ncu --set full --clock-control none --import-source on -k regex:sgp4_prop_kernel -c $n -f \
      -o $O/${TAG}_${w}_full python bench.py --quick --workload $w --steps 1 --warmup 1 > $O/${TAG}_ncu_${w}.log 2>&1
  ncu -i $O/${TAG}_${w}_full.ncu-rep --page raw --csv > $O/${TAG}_${w}_ncu_full_raw.csv 2>/dev/null
done
for w in c3 c4 c5; do
  python bench.py --quick --workload $w --steps 5 --warmup 3 2>/dev/null | tail -1
done > $O/${TAG}_bench_quick_c3_c4_c5_full.jsonl
ls -la $O | tail -20
