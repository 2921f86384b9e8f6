#!/bin/bash
# Round 2 (two GPUs, ONE process): NVLink byte counters of the device-resident gather -- shard 1's
# kernel 2 stores its rows into GPU 0 over peer access -- read with ncu around the kernel
# launches (nvidia-smi / NVML link counters are N/A in these guests).
O=gpurun_out
TAG=${1:-r2k}
mkdir -p $O
ncu --query-metrics 2>/dev/null | grep -i -E "nvl[rt]x__bytes" | cut -c1-150 > $O/${TAG}_nvlink_metric_names.txt
M=nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum,$M \
    --clock-control none -k regex:sgp4_prop_kernel -c 40 --csv --log-file $O/${TAG}_p2p_ncu.csv \
    python scripts/p2p_gather_bench.py --steps 1 --times 512 > $O/${TAG}_p2p_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $O/${TAG}_p2p_ncu.log | cut -c1-400
grep -c sgp4_prop $O/${TAG}_p2p_ncu.csv
timeout 300 python scripts/p2p_gather_bench.py --steps 10 > $O/${TAG}_p2p_gather.json 2>&1; cat $O/${TAG}_p2p_gather.json | cut -c1-700
python - <<PY
import csv
rows = list(csv.reader(open("$O/${TAG}_p2p_ncu.csv")))
hdr = None
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if "<0, 0" in d["Kernel Name"] and d["Metric Name"] in ("nvltx__bytes.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
            print(d["ID"], "device", d["Device"], d["Metric Name"], d["Metric Value"], d["Metric Unit"])
PY
