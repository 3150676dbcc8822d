#!/bin/bash
# Runs ON the GPU box (gpurun): the ncu captures behind profiles/ (round 2).  CSV output goes to gpurun_out/ (the .ncu-rep files stay
# in /tmp: gpurun copies back at most 64 MiB); summarise with
# tools/summarize_ncu.py, tools/ncu_layers.py, tools/hbm_kernels_report.py, tools/ncu_key.py.
set -x
mkdir -p gpurun_out
M3=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
# 1. launch list of the bench command (every kernel of 1 warm-up + 2 timed steps of every bench leg)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_bench_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
# 2. one GuidedResUnet forward on 8 padded 12 MP frames: time + DRAM bytes per launch
timeout 300 ncu --profile-from-start off --metrics $M3 --clock-control none --csv --log-file gpurun_out/r02_net_frames8.csv \
  python tools/profile_image.py --net-only 8 --frame 1536x2016 > /dev/null 2>&1
# 3. one pipeline step over 8 frames (both rounds): time + DRAM bytes of every kernel
timeout 300 ncu --profile-from-start off --metrics $M3 --clock-control none --csv --log-file gpurun_out/r02_step_frames8.csv \
  python tools/profile_image.py --frames 8 > /dev/null 2>&1
# 4. full-set captures: conv kernels (first 12 launches of a forward: c1.conv1 paired ... ), the last layer with the fused output conv,
#    and the HBM-side kernels of the estimator / VST
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_tc -c 12 -f -o /tmp/r02_conv_full \
  python tools/profile_image.py --net-only 8 --frame 1536x2016 > /dev/null 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_tc --launch-skip 27 -c 1 -f -o /tmp/r02_conv_last \
  python tools/profile_image.py --net-only 8 --frame 1536x2016 > /dev/null 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k "regex:box_fused|vst_fwd|vst_inv|hist0|hist1|hist2|masked_sums|head_conv" -c 14 -f -o /tmp/r02_hbm_full \
  python tools/profile_image.py --frames 4 > /dev/null 2>&1
for f in r02_conv_full r02_conv_last r02_hbm_full; do ncu -i /tmp/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv; done
ls -la gpurun_out | grep r02_
