#!/bin/bash
# round-1 capture j (final build of this round): launch list + full capture of the config-3 frame, launch list of a Suzanne frame
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 6 --warmup 3 --in-flight 1 --no-cpu-baseline > gpurun_out/prof_r1j_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_micro|k_tile_opaque|k_vertex|k_tile_offsets|k_large_fill' -s 15 -c 5 -o gpurun_out/prof_r1j -f python bench.py --steps 2 --no-cpu-baseline --in-flight 1 > gpurun_out/prof_r1j_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/launches_r1j_suzanne.csv python profiles/scripts/turntable.py > gpurun_out/prof_r1j_tt.log 2>&1
ls -la gpurun_out/ | tail -8
