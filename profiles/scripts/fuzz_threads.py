"""Concurrency stress: several host threads, one context each, run randomised pipeline and screen-space scenarios at the
same time on one GPU (frames in flight on independent contexts must not interfere).
usage: python profiles/scripts/fuzz_threads.py [threads] [iterations per thread]"""
import sys, os, time, threading
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from softrender_b200 import pipeline as P, scenes
import oracle_binding as ob
import test_gpu_parity as T
from fuzz_scenarios import run_pipeline_scenario, run_scenario

nthreads = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bad = []

def work(k):
    ctx = P.Context(0)
    for it in range(iters):
        seed = 100000 * (k + 1) + it
        msg = run_pipeline_scenario(P, ctx, ob, scenes, seed) if it % 2 else run_scenario(P, ctx, T.run_both_screen, seed)
        if msg:
            bad.append(msg)
    ctx.close()

t0 = time.time()
ts = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
for t in ts: t.start()
for t in ts: t.join()
for m in bad[:20]: print("MISMATCH", m)
print(f"{nthreads} threads x {iters} scenarios, {len(bad)} mismatches, {time.time() - t0:.1f} s")
