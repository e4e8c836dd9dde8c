"""Long randomised differential campaign against the oracle (scenario generator: tests/fuzz_scenarios.py).
usage: python profiles/scripts/fuzz.py [iterations] [first_seed] [big]"""
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from softrender_b200 import pipeline as P
import test_gpu_parity as T
from fuzz_scenarios import run_scenario

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
big = len(sys.argv) > 3 and sys.argv[3] == "big"
ctx = P.Context(0)
bad, t0 = 0, time.time()
for it in range(iters):
    msg = run_scenario(P, ctx, T.run_both_screen, seed0 + it, big)
    if msg:
        bad += 1
        print("MISMATCH", msg)
print(f"{iters} scenarios, {bad} mismatches, {time.time() - t0:.1f} s")
