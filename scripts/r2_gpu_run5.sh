#!/bin/bash
# Round 2, GPU call 5 (re-run of call 4 after the container was re-created): bench line, ncu launch list + full captures, chunk sweep of the host path, tests.
O=gpurun_out
TAG=r2e
mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
echo "bench rc=$?"; tail -2 $O/${TAG}_bench_c2.err
timeout 300 python scripts/e2e_chunks.py > $O/${TAG}_e2e_chunks.jsonl 2>&1; cat $O/${TAG}_e2e_chunks.jsonl
timeout 300 python bench.py --steps 2 --warmup 1 --no-extra > $O/${TAG}_bench_short.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${TAG}_c2_launches.csv python bench.py --steps 2 --warmup 1 --no-extra > $O/${TAG}_ncu_launches.log 2>&1
for w in c2 c4p c3p; do
  n=2; [ $w = c3p ] && n=5
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgp4_prop_kernel -c $n -f \
      -o $O/${TAG}_${w}_full python bench.py --quick --workload $w --steps 1 --warmup 1 > $O/${TAG}_ncu_${w}.log 2>&1
  ncu -i $O/${TAG}_${w}_full.ncu-rep --page raw --csv > $O/${TAG}_${w}_ncu_full_raw.csv 2>/dev/null
done
ls -la $O | grep $TAG
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -rf --tb=short > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -60 $O/${TAG}_pytest.log | cut -c1-300
