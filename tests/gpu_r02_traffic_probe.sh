#!/bin/bash
# does pair mode's DRAM over-read (2x at 12.5M rows, B = 1024) come from the query super-tiles drifting apart?
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02_traffic_probe.log
for rows in 1000000 4000000 12500000; do
  P="env PROBE_ROWS=$rows PROBE_BATCHES=1024 PROBE_PATHS=mma python tests/quick_probe.py"
  ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k_mma_topk -s 3 -c 1 --csv $P 2>/dev/null | grep -E "dram__bytes_read|hit_rate|time_duration" | awk -F'","' -v r=$rows '{print r, $(NF-2), $(NF-1), $NF}' | tee -a $O/r02_traffic_probe.log
done
for b in 512 2048; do
  P="env PROBE_ROWS=12500000 PROBE_BATCHES=$b PROBE_PATHS=mma python tests/quick_probe.py"
  ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k_mma_topk -s 3 -c 1 --csv $P 2>/dev/null | grep -E "dram__bytes_read|hit_rate|time_duration" | awk -F'","' -v r="12.5M B=$b" '{print r, $(NF-2), $(NF-1), $NF}' | tee -a $O/r02_traffic_probe.log
done
