#!/bin/bash
# Round 2, last GPU call (one GPU): bench line, ncu launch list and full captures of the shipped kernels.
O=gpurun_out
TAG=r2t
mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
echo "bench rc=$?"
for w in c4p c3p c2; do
  n=2; [ $w = c3p ] && n=5
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgp4_prop_kernel -c $n -f \
      -o /tmp/${TAG}_${w}_full python bench.py --quick --workload $w --steps 1 --warmup 1 > $O/${TAG}_ncu_${w}.log 2>&1
  ncu -i /tmp/${TAG}_${w}_full.ncu-rep --page raw --csv > $O/${TAG}_${w}_ncu_full_raw.csv 2>/dev/null
done
python scripts/ncu_summary.py /tmp/${TAG}_c4p_full.ncu-rep $((25000*1024)) $((25000*1024)) > $O/${TAG}_c4p_ncu_summary.txt 2>&1
python scripts/ncu_summary.py /tmp/${TAG}_c3p_full.ncu-rep $((5000*512)) $((10000*512)) $((15000*512)) $((68600*512)) $((1400*512)) > $O/${TAG}_c3p_ncu_summary.txt 2>&1
python scripts/ncu_summary.py /tmp/${TAG}_c2_full.ncu-rep $((29382*1440)) $((618*1440)) > $O/${TAG}_c2_ncu_summary.txt 2>&1
grep -E "^==|per evaluation|duration|fp64_pipe" $O/${TAG}_c3p_ncu_summary.txt $O/${TAG}_c4p_ncu_summary.txt | cut -c1-160
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${TAG}_c2_launches.csv python bench.py --steps 2 --warmup 1 --no-extra > $O/${TAG}_ncu_launches.log 2>&1
python - <<PY
import json
r = json.loads(open("$O/${TAG}_bench_c2.json").read().strip().splitlines()[-1])
print("value", r["value"], "ms", r["ms_per_step"], "e2e", r["e2e"]["value"], r["e2e"].get("frac_of_ceiling"))
for k, v in r["workloads"].items():
    print(k, v["value"], v["ms_per_step"])
PY
