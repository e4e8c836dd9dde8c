"""Campaign over points / lines / normal-visualisation geometry shaders (tests/fuzz_scenarios.py::run_geometry_scenario).
usage: python profiles/scripts/fuzz_geometry.py [iterations] [first_seed]"""
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from softrender_b200 import pipeline as P, scenes
import oracle_binding as ob
from fuzz_scenarios import run_geometry_scenario
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ctx = P.Context(0)
bad, t0 = 0, time.time()
for it in range(iters):
    msg = run_geometry_scenario(P, ctx, ob, scenes, seed0 + it)
    if msg:
        bad += 1
        print("MISMATCH", msg)
print(f"{iters} geometry scenarios, {bad} mismatches, {time.time() - t0:.1f} s")
