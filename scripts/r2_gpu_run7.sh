#!/bin/bash
# Round 2, GPU call 7 (one GPU): tests, A/B of the deep-space kernels (solar / lunar phase split),
# bench line, ncu full captures of the deep classes.
O=gpurun_out
TAG=r2h
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -rf --tb=short > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log | cut -c1-600
q() { # lib workload [env...]
  local lib=$1 w=$2; shift 2
  env "$@" PERTURB_GPU_LIB=$lib python bench.py --quick --workload $w --steps 10 --warmup 3 2>/dev/null | tail -1 | \
    python -c "import sys,json; r=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$(basename $lib)','env':'$*','workload':r['workload'],'value':r['value'],'ms':r['ms_per_step'],'mhz':r['clocks']['sm_mhz']}))"
}
{
  for w in c4s c3s c2 c5s; do q $PWD/perturb_b200/libperturb_gpu.so $w; q $PWD/perturb_b200/variants/base.so $w; done
  q $PWD/perturb_b200/libperturb_gpu.so c2 PERTURB_GPU_CTA_SATS=16
  q $PWD/perturb_b200/libperturb_gpu.so c4s PERTURB_GPU_CTA_SATS=16
  q $PWD/perturb_b200/libperturb_gpu.so c4s PERTURB_GPU_CTA_SATS=32
} > $O/${TAG}_variants.jsonl 2>&1
cat $O/${TAG}_variants.jsonl
SECONDS=0
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
echo "bench rc=$? in ${SECONDS}s"; tail -2 $O/${TAG}_bench_c2.err
for w in c4p c3p c2; do
  n=2; [ $w = c3p ] && n=5
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgp4_prop_kernel -c $n -f \
      -o $O/${TAG}_${w}_full python bench.py --quick --workload $w --steps 1 --warmup 1 > $O/${TAG}_ncu_${w}.log 2>&1
  ncu -i $O/${TAG}_${w}_full.ncu-rep --page raw --csv > $O/${TAG}_${w}_ncu_full_raw.csv 2>/dev/null
done
python - <<PY
import json
r = json.loads(open("$O/${TAG}_bench_c2.json").read().strip().splitlines()[-1])
print("value", r["value"], "ms", r["ms_per_step"], "pipe", r["roofline"]["fp64_pipe"]["frac_of_fp64_issue_cycles"])
print("e2e", json.dumps(r["e2e"])[:900])
for k, v in r["workloads"].items():
    print(k, v["value"], v["ms_per_step"], v["fp64_pipe"]["frac_of_fp64_issue_cycles"])
PY
