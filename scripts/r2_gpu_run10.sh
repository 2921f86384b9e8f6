#!/bin/bash
# Round 2, GPU call 10 (one GPU): resident-CTA variants of the deep-space kernels.
O=gpurun_out
TAG=r2m
mkdir -p $O
q() { # lib workload [env...]
  local lib=$1 w=$2; shift 2
  env "$@" PERTURB_GPU_LIB=$lib python bench.py --quick --workload $w --steps 10 --warmup 3 2>/dev/null | tail -1 | \
    python -c "import sys,json; r=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$(basename $lib)','env':'$*','workload':r['workload'],'value':r['value'],'ms':r['ms_per_step'],'mhz':r['clocks']['sm_mhz']}))"
}
{
  for v in ../libperturb_gpu r2at5 deep5 r2at4; do q $PWD/perturb_b200/variants/$v.so c4s; done
  for v in ../libperturb_gpu deep5; do q $PWD/perturb_b200/variants/$v.so c3s; done
  q $PWD/perturb_b200/libperturb_gpu.so c4s
} > $O/${TAG}_variants.jsonl 2>&1
cat $O/${TAG}_variants.jsonl
