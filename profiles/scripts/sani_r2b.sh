#!/bin/bash
# sanitizer over the round-2b code: ordered kernel's lane-per-triangle sweep + interleaved rows, two-plane texture buffers
export SR_UNDER_SANITIZER=1
K='alpha_over or user_blend or texture_buffer or test_stencil or discard or render_to_texture or full_example_scene'
timeout 800 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > gpurun_out/r2b_memcheck.log 2>&1
tail -4 gpurun_out/r2b_memcheck.log
timeout 800 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "alpha_over or user_blend or test_texture_buffer_with or test_stencil_wider" > gpurun_out/r2b_racecheck.log 2>&1
tail -4 gpurun_out/r2b_racecheck.log
grep -c "Hazard\|hazard" gpurun_out/r2b_racecheck.log
