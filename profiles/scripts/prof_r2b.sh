#!/bin/bash
# round-2 final captures: launch list + full capture of the config-3 frame; the ordered kernel on the same mesh with alpha_over; timing tables
# (every step under its own timeout: a capture that stalls must not eat the GPU budget)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --quick --steps 6 --warmup 3 --in-flight 1 > gpurun_out/prof_r2b_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_micro|k_tile_opaque|k_vertex|k_tile_offsets|k_large_fill' -s 15 -c 5 -o gpurun_out/prof_r2b -f python profiles/scripts/grid_frame.py grid10m 6 > gpurun_out/prof_r2b_full.log 2>&1
echo "full capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile_ordered -s 2 -c 1 -o gpurun_out/prof_r2b_ordered -f python profiles/scripts/ordered_grid.py 1250x1000 > gpurun_out/prof_r2b_ordered.log 2>&1
echo "ordered capture rc=$?"
SR_STAGES=0 timeout 200 python profiles/scripts/ordered_perf.py > gpurun_out/r2b_ordered_perf.txt 2>&1
timeout 200 python profiles/scripts/rtt.py > gpurun_out/r2b_rtt_pass_times.txt 2>&1
tail -3 gpurun_out/prof_r2b_full.log; ls -la gpurun_out/ | grep r2b
