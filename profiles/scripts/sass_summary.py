"""cuobjdump -sass of libsoftrender_b200.so -> per-kernel instruction summary (profiles/r2_sass_summary.txt):
instruction count, registers / shared memory from -res-usage, and the counts of the mnemonics that matter for this design
(UBLKCP = cp.async.bulk / TMA bulk copies, SYNCS = mbarrier ops, REDG / RED / ATOM(G|S) = the key reductions, MUFU = SFU,
LDG / STG / LDS / STS, FFMA / FMUL / FADD -- FFMA appears only where FMA contraction is allowed, i.e. sr_div_exact and the
lit-shading interpolation; coverage arithmetic is FMUL + FADD)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "rust-softrender_b200/csrc/libsoftrender_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for ln in res.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in ln:
        usage[cur] = ln.strip()
        cur = None
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
WATCH = ["UBLKCP", "SYNCS", "REDG", "RED", "ATOMG", "ATOMS", "ATOM", "MUFU", "LDG", "STG", "LDS", "STS", "FFMA", "FMUL", "FADD", "SHFL", "BAR", "LDGSTS", "UTMALDG", "HMMA", "UTCHMMA"]
kernels = []
name, counts, total = None, None, 0
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        if name:
            kernels.append((name, total, counts))
        name, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and name:
        total += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[w] += 1
                break
        else:
            counts[op.split(".")[0]] += 0
if name:
    kernels.append((name, total, counts))
print(f"# {lib}: {len(kernels)} kernels, sm_100a SASS (cuobjdump -sass), counts of selected mnemonics per kernel")
print("# no tensor-core (HMMA / UTCHMMA) and no tensor-map TMA (UTMALDG) instructions anywhere: by design (no dense contraction; bulk copies are 1-D)")
for n, total, c in sorted(kernels, key=lambda k: -k[1]):
    d = demangle(n)
    d = re.sub(r"\(.*", "", d)
    print(f"{d}\n    {total} instructions; {usage.get(n, '')}\n    " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
