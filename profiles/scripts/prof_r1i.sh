#!/bin/bash
# round-1 capture i: the paths beside config 3 -- blended full_example frame (ordered path), face-normal lines (opaque
# line path), Suzanne frame (small-draw path)
ncu --set full --clock-control none --import-source on -k regex:'k_tile_ordered|k_bin_setup|k_bin_fill' -s 12 -c 6 -o gpurun_out/prof_r1i_ordered -f python profiles/scripts/fe_frame.py 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_lines_vis|k_tile_opaque|k_geo_normals' -s 6 -c 6 -o gpurun_out/prof_r1i_lines -f python profiles/scripts/lines_stage.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_bin_small|k_tile_opaque|k_clip' -s 40 -c 6 -o gpurun_out/prof_r1i_suzanne -f python profiles/scripts/turntable.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 24 --csv --log-file gpurun_out/launches_r1i_full_example_blend.csv python profiles/scripts/fe_frame.py 12 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/launches_r1i_suzanne.csv python profiles/scripts/turntable.py > /dev/null 2>&1
ls -la gpurun_out/*r1i*
