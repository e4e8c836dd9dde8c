#!/bin/bash
# sanitizer over the last additions of round 2: k_tile_few, the lane-per-triangle sweep's own tests, the single-launch clipper
export SR_UNDER_SANITIZER=1
K='few_triangles_direct or full_screen_pass_direct or (pixel_sized and 5.0) or (blended_in_order and 125) or suzanne_stages or two_colour or turntable_frames'
timeout 700 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > gpurun_out/r2c_memcheck.log 2>&1
tail -3 gpurun_out/r2c_memcheck.log
timeout 700 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "few_triangles_direct or (full_screen_pass_direct and 0-0) or (pixel_sized and 5.0) or (blended_in_order and 125)" > gpurun_out/r2c_racecheck.log 2>&1
tail -3 gpurun_out/r2c_racecheck.log
