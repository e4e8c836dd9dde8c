#!/bin/bash
# round-1 capture k: the render-to-texture second pass (full-screen quad at 3840x2160, bilinear/clamp from a render target)
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 30 --csv --log-file gpurun_out/launches_r1k_rtt.csv python profiles/scripts/rtt_pass.py 1 0 6 > gpurun_out/prof_r1k_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_tile_opaque|k_bin_small' -s 6 -c 2 -o gpurun_out/prof_r1k_rtt -f python profiles/scripts/rtt_pass.py 1 0 6 > gpurun_out/prof_r1k_full.log 2>&1
ls -la gpurun_out/ | tail -5
