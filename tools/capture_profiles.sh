#!/bin/bash
# Runs ON the GPU box (gpurun): the ncu captures behind profiles/ (launch list of the bench command, per-launch DRAM
# traffic and one full-set capture of the conv kernel).  Output goes to gpurun_out/; summarise with
# tools/summarize_ncu.py, tools/conv_traffic_json.py, tools/conv_layers_report.py.
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/bench_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:conv_tc --csv --log-file gpurun_out/conv_traffic.csv python tools/profile_image.py --net-only 640 > /dev/null 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/net640.csv \
  python tools/profile_image.py --net-only 640 > /dev/null 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_tc -c 12 -f -o gpurun_out/conv_full \
  python tools/profile_image.py --net-only 640 > /dev/null 2>&1
ls -la gpurun_out
