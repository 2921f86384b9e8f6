#!/bin/bash
# Round 2, GPU call 6 (one GPU): tests, A/B of kernel-2 variants (rotation-based sin/cos of u, sqrt
# series), the bench line, ncu launch list + full capture of the near-Earth kernel.
O=gpurun_out
TAG=r2g
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -rf --tb=short > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log | cut -c1-300
q() { # lib workload [env...]
  local lib=$1 w=$2; shift 2
  env "$@" PERTURB_GPU_LIB=$lib python bench.py --quick --workload $w --steps 10 --warmup 3 2>/dev/null | tail -1 | \
    python -c "import sys,json; r=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$(basename $lib)','env':'$*','workload':r['workload'],'value':r['value'],'ms':r['ms_per_step'],'mhz':r['clocks']['sm_mhz']}))"
}
{
  for w in c2 c3s c5s c4s; do q $PWD/perturb_b200/libperturb_gpu.so $w; q $PWD/perturb_b200/variants/base.so $w; done
  for v in near6 norot pair5 pair4; do [ -f perturb_b200/variants/$v.so ] && q $PWD/perturb_b200/variants/$v.so c2; done
  for v in near6 norot; do [ -f perturb_b200/variants/$v.so ] && q $PWD/perturb_b200/variants/$v.so c3s; done
  q $PWD/perturb_b200/libperturb_gpu.so c2 PERTURB_GPU_CTA_SATS=16
  q $PWD/perturb_b200/libperturb_gpu.so c2 PERTURB_GPU_CTA_SATS=32
  q $PWD/perturb_b200/libperturb_gpu.so c2
  q $PWD/perturb_b200/variants/base.so c2
} > $O/${TAG}_variants.jsonl 2>&1
cat $O/${TAG}_variants.jsonl
SECONDS=0
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
echo "bench rc=$? in ${SECONDS}s"; tail -2 $O/${TAG}_bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${TAG}_c2_launches.csv python bench.py --steps 2 --warmup 1 --no-extra > $O/${TAG}_ncu_launches.log 2>&1
for w in c2 c3p; do
  n=2; [ $w = c3p ] && n=5
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgp4_prop_kernel -c $n -f \
      -o $O/${TAG}_${w}_full python bench.py --quick --workload $w --steps 1 --warmup 1 > $O/${TAG}_ncu_${w}.log 2>&1
  ncu -i $O/${TAG}_${w}_full.ncu-rep --page raw --csv > $O/${TAG}_${w}_ncu_full_raw.csv 2>/dev/null
done
ls -la $O | grep $TAG
python - <<PY
import json
r = json.loads(open("$O/${TAG}_bench_c2.json").read().strip().splitlines()[-1])
print("value", r["value"], "ms", r["ms_per_step"], "pipe", r["roofline"]["fp64_pipe"]["frac_of_fp64_issue_cycles"])
print("e2e", json.dumps(r["e2e"])[:700])
print("dropin", json.dumps(r["e2e_dropin"])[:500])
for k, v in r["workloads"].items():
    print(k, v["value"], v["ms_per_step"])
PY
