#!/bin/bash
# Round 2, multi-GPU call: host-ingest diagnostics (topology, D2H ceiling one process / one process
# per GPU, NUMA binding on / off, chunked copies) and the complete bench line at N GPUs.
#   gpurun --gpus N -- 'bash scripts/r2_gpu_multi.sh N TAG'
N=${1:-2}
TAG=${2:-r2f_n$N}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
{
  nvidia-smi topo -m
  nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv
  for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do
    p=/sys/bus/pci/devices/$(echo ${b: -12} | tr A-Z a-z)/numa_node; echo "$b numa_node=$(cat $p 2>/dev/null)"
  done
  lscpu | grep -i -E 'numa|socket|^CPU\(s\)|model name|hypervisor'
  nproc; free -g | head -2
} > $O/${TAG}_topology.txt 2>&1
nvidia-smi nvlink -gt d -i 0 > $O/${TAG}_nvlink_before.txt 2>&1
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
rc=$?
nvidia-smi nvlink -gt d -i 0 > $O/${TAG}_nvlink_after.txt 2>&1
python - <<PY > $O/${TAG}_nvlink_delta.txt 2>&1
import re
def tot(path, key):
    return sum(int(x) for x in re.findall(key + r":\s*(\d+)\s*KiB", open(path).read()))
for key in ("Data Rx", "Data Tx"):
    a, b = tot("$O/${TAG}_nvlink_before.txt", key), tot("$O/${TAG}_nvlink_after.txt", key)
    print("GPU 0 NVLink %s over the bench run: %d KiB = %.3f GB (all links)" % (key, b - a, (b - a) * 1024e-9))
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for scope in (0xFFFFFFFF, 0):
    try:
        v = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, scope)])[0]
        print("NVML field DATA_RX scope %x: nvmlReturn=%d value=%s" % (scope, v.nvmlReturn, v.value.ullVal))
    except Exception as exc:
        print("NVML field DATA_RX scope %x: %s" % (scope, exc))
PY
cat $O/${TAG}_nvlink_delta.txt
echo "bench rc=$rc"; tail -3 $O/${TAG}_bench.err | cut -c1-300
PERTURB_BENCH_NUMA=off timeout 300 $TR bench.py --gpus $N --steps 5 --e2e-only 2>/dev/null | tail -1 > $O/${TAG}_e2e_nobind.json
PERTURB_GPU_HOST_CHUNKS=4 timeout 300 $TR bench.py --gpus $N --steps 5 --e2e-only 2>/dev/null | tail -1 > $O/${TAG}_e2e_chunks4.json
timeout 300 python scripts/d2h_ceiling.py 1 4 > $O/${TAG}_d2h_one_process.jsonl 2>&1
cat $O/${TAG}_topology.txt
for f in $O/${TAG}_e2e_nobind.json $O/${TAG}_e2e_chunks4.json $O/${TAG}_d2h_one_process.jsonl; do echo "== $f"; cut -c1-1200 $f; done
python - <<PY
import json
try:
    r = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "e2e", "e2e_states", "inproc_sharded"):
        print(k, json.dumps(r.get(k))[:900])
    print("e2e_dropin", json.dumps(r.get("e2e_dropin"))[:600])
    print("workloads", json.dumps(r.get("workloads"))[:900])
except Exception as exc:
    print("no bench line:", exc)
PY
