#!/bin/bash
# round-2 captures: launch list + full capture of the config-3 frame on one GPU; the range-sharded frame's kernels (two ranks on one device)
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --quick --steps 6 --warmup 3 --in-flight 1 > gpurun_out/prof_r2a_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_micro|k_tile_opaque|k_vertex|k_tile_offsets|k_large_fill' -s 15 -c 5 -o gpurun_out/prof_r2a -f python bench.py --quick --steps 2 --in-flight 1 > gpurun_out/prof_r2a_full.log 2>&1
SR_SHARD_TIMEOUT_MS=5 FRAMES=2 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2a_shard_local.csv python profiles/scripts/shard_local.py > gpurun_out/prof_r2a_shard_launch.log 2>&1
SR_SHARD_TIMEOUT_MS=5 FRAMES=2 ncu --set full --clock-control none --import-source on -k regex:'k_shard_merge|k_vertex_marked|k_vis_rows_touched|k_vis_clear_foreign|k_fb_fill_foreign|k_tile_opaque' -s 8 -c 12 -o gpurun_out/prof_r2a_shard -f python profiles/scripts/shard_local.py > gpurun_out/prof_r2a_shard_full.log 2>&1
tail -2 gpurun_out/prof_r2a_shard_full.log; ls -la gpurun_out/ | grep r2a
