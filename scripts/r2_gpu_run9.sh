#!/bin/bash
# Round 2, GPU call 9 (one GPU): tests with the third-order Kepler iteration for eccentric orbits,
# A/B against the same library without it.
O=gpurun_out
TAG=r2l
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -rf --tb=short > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -12 $O/${TAG}_pytest.log | cut -c1-600
q() { # lib workload [env...]
  local lib=$1 w=$2; shift 2
  env "$@" PERTURB_GPU_LIB=$lib python bench.py --quick --workload $w --steps 10 --warmup 3 2>/dev/null | tail -1 | \
    python -c "import sys,json; r=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$(basename $lib)','env':'$*','workload':r['workload'],'value':r['value'],'ms':r['ms_per_step'],'mhz':r['clocks']['sm_mhz']}))"
}
{
  for w in c4s c3s c2; do q $PWD/perturb_b200/libperturb_gpu.so $w; q $PWD/perturb_b200/variants/nohalley.so $w; done
  q $PWD/perturb_b200/libperturb_gpu.so c4s; q $PWD/perturb_b200/variants/nohalley.so c4s
} > $O/${TAG}_variants.jsonl 2>&1
cat $O/${TAG}_variants.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgp4_prop_kernel -c 2 -f \
    -o /tmp/${TAG}_c4p_full python bench.py --quick --workload c4p --steps 1 --warmup 1 > $O/${TAG}_ncu_c4p.log 2>&1
python scripts/ncu_summary.py /tmp/${TAG}_c4p_full.ncu-rep $((25000*1024)) $((25000*1024)) > $O/${TAG}_c4p_ncu_summary.txt 2>&1
grep -E "^==|per evaluation|duration|fp64_pipe|regs|local" $O/${TAG}_c4p_ncu_summary.txt
