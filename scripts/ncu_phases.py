"""Bucket the SASS lines of the first kernel in an ncu report by execution count, to see where
warp time (stall samples) and issue slots go.  python scripts/ncu_phases.py <rep> [warps]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, body = rows[h0], []
for r in rows[h0 + 1:]:
    if r and r[0] == "Address":
        break
    if len(r) > 10:
        body.append(r)
f = lambda x: float(x) if x not in ("", "-") else 0.0
si, ei, ci = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Source")
tot_s = sum(f(r[si]) for r in body)
tot_e = sum(f(r[ei]) for r in body)
execs = sorted({f(r[ei]) for r in body}, reverse=True)
base = max(set(f(r[ei]) for r in body if "EXIT" in r[ci]) or [1.0])
print(f"warps (EXIT count) = {base:.0f}; total warp-instr = {tot_e:.3e}; samples = {tot_s:.0f}")
buckets = collections.OrderedDict()
for r in body:
    e = f(r[ei]) / base
    key = "<=1x per warp" if e <= 1.001 else "1-4x" if e <= 4 else "4-12x" if e <= 12 else "12-40x" if e <= 40 else ">40x"
    b = buckets.setdefault(key, [0, 0.0, 0.0, 0])
    b[0] += 1; b[1] += f(r[ei]); b[2] += f(r[si])
    if any(op in r[ci] for op in ("DFMA", "DMUL", "DADD", "DSETP")): b[3] += f(r[ei])
for k, b in buckets.items():
    print(f"{k:14s} lines={b[0]:5d} warp-instr={b[1]:.3e} ({b[1]/tot_e*100:5.1f}%) fp64-instr={b[3]:.3e} samples={b[2]/tot_s*100:5.1f}%")
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter()
for r in body:
    for n in names:
        agg[n] += f(r[hdr.index(n)])
print("stalls:", [(n, round(v / tot_s * 100, 1)) for n, v in agg.most_common(8)])
print("hottest lines:")
for r in sorted(body, key=lambda r: -f(r[si]))[:18]:
    print(f"  {f(r[si])/tot_s*100:5.1f}% x{f(r[ei])/base:7.1f}  {r[ci].strip()[:80]}")
