#!/bin/bash
# Per-launch issue / lane / DRAM counters of the bench frame's launches (serialised under ncu): where the frame's instructions go.
mkdir -p gpurun_out
TAG=${1:-r2_frame_metrics}
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
   --cache-control none --clock-control none -k "regex:^(?!.*(ploc|Ploc|flatten|morton|initLeaves|collapse|DeviceRadixSort|DeviceScan))" -s 150 -c 150 --csv --log-file gpurun_out/$TAG.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/$TAG.log 2>&1
python tools/frame_metrics_summary.py gpurun_out/$TAG.csv | tee gpurun_out/${TAG}_summary.txt
