#!/bin/bash
# Round 2 (N GPUs): the host-gathered arm with equal satellite ranges and with ranges proportional to
# each rank's measured D2H rate, then the complete bench line.
N=${1:-4}
TAG=${2:-r2q_n$N}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus $N --steps 5 --e2e-only 2> $O/${TAG}_e2e_only.err | tail -1 > $O/${TAG}_e2e_only.json
echo "e2e-only rc=$?"; tail -3 $O/${TAG}_e2e_only.err | cut -c1-300
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench rc=$?"; tail -3 $O/${TAG}_bench.err | cut -c1-300
python - <<PY
import json
r = json.loads(open("$O/${TAG}_e2e_only.json").read().strip().splitlines()[-1])
print("e2e-only: equal", r["e2e"]["value"], r["e2e"]["d2h_gbs"], "balanced", json.dumps(r["e2e"].get("rate_balanced"))[:700])
r = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print("bench: value", r["value"], "e2e", r["e2e"]["value"], r["e2e"]["d2h_gbs"], r["e2e"].get("frac_of_ceiling"))
print("balanced", json.dumps(r["e2e_rate_balanced"])[:900])
PY
