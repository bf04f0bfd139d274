"""Bucket executed warp-instructions of one kernel by runs of SASS lines with equal execution count.
usage: python tools/ncu_regions.py rep.ncu-rep kernel_regex [warps]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-count", "1", "--launch-skip", (sys.argv[4] if len(sys.argv) > 4 else "0")],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
c = {h: i for i, h in enumerate(hdr)}
seen, body = set(), []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[0] == "Address" or r[0] in seen:
        continue
    seen.add(r[0]); body.append(r)
f = lambda r, h: float(r[c[h]] or 0)
tot = sum(f(r, "Instructions Executed") for r in body)
warps = float(sys.argv[3]) if len(sys.argv) > 3 else max(f(r, "Instructions Executed") for r in body[:5])
print(f"{len(body)} SASS lines, {tot:.0f} warp-instr, {tot / warps:.0f} per warp ({warps:.0f} warps)")
grp = []
for i, r in enumerate(body):
    e = f(r, "Instructions Executed")
    if grp and grp[-1][2] == e:
        grp[-1][1] = i; grp[-1][3] += e; grp[-1][4] += f(r, "# Samples")
    else:
        grp.append([i, i, e, e, f(r, "# Samples"), r[c["Source"]].strip()[:46]])
ts = sum(g[4] for g in grp)
for g in grp:
    if g[3] / tot > 0.006 or g[4] / ts > 0.01:
        print(f"{g[0]:5d}-{g[1]:5d} n={g[1]-g[0]+1:4d} exec/warp={g[2]/warps:6.2f} inst={g[3]/tot*100:5.1f}% stall-samples={g[4]/ts*100:5.1f}%  {g[5]}")
