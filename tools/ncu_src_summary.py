"""Summarise an ncu --page source --csv dump: stall-reason totals and the hottest SASS lines.
usage: ncu -i rep --page source --csv --kernel-name K > f.csv ; python tools/ncu_src_summary.py f.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
def f(r, h):
    try: return float(r[col[h]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in body)
inst = sum(f(r, "Instructions Executed") for r in body)
print(f"SASS lines {len(body)}  samples {tot:.0f}  warp-instructions {inst:.0f}")
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = sorted(((sum(f(r, h) for r in body), h) for h in st), reverse=True)
print("stall totals:", ", ".join(f"{h[6:]}={v/tot*100:.1f}%" for v, h in agg[:8]))
print("hottest lines:")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:top]:
    reasons = sorted(((f(r, h), h[6:]) for h in st), reverse=True)[:2]
    print(f"{f(r,'# Samples'):7.0f} {f(r,'Instructions Executed'):9.0f}  {r[col['Source']].strip()[:80]:80s} {reasons[0][1]}:{reasons[0][0]:.0f} {reasons[1][1]}:{reasons[1][0]:.0f}")
