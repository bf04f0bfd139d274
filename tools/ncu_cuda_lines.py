"""Per CUDA-C source line: executed warp-instructions and stall samples (ncu --page source --print-source cuda --csv)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda", "--csv", "--kernel-name", "regex:" + kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
files = {}
cur = None
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] in ("#", "Line", "Address") or (len(r) > 1 and r[1] == "Source"):
        hdr = r; continue
    if hdr and cur and len(r) == len(hdr):
        files.setdefault(cur, []).append(dict(zip(hdr, r)))
allr = [(f, d) for f, rs in files.items() for d in rs]
def g(d, k):
    try: return float(d.get(k, 0) or 0)
    except ValueError: return 0.0
ti = sum(g(d, "Instructions Executed") for _, d in allr); ts = sum(g(d, "# Samples") for _, d in allr)
print(f"total warp-instr {ti:.0f}  samples {ts:.0f}")
for f, d in sorted(allr, key=lambda fd: -g(fd[1], "# Samples"))[:top]:
    ln = d.get("#") or d.get("Line") or "?"
    print(f"{g(d,'# Samples')/max(ts,1)*100:5.1f}% smp {g(d,'Instructions Executed')/max(ti,1)*100:5.1f}% inst  {f.split('/')[-1]}:{ln}  {d.get('Source','').strip()[:90]}")
