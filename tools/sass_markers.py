"""profiles/r2_sass_markers.txt: per-kernel counts of the SASS mnemonics that show what the shipped library uses
(cuobjdump -sass of mpmavatar_b200/libmpm_b200.so; runs without a GPU)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mpmavatar_b200", "libmpm_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MARKS = [("UBLKCP", r"\bUBLKCP"), ("REDG.F32x4", r"REDG\.E\.ADD\.F32x4"), ("FFMA2", r"\bFFMA2"), ("LDGSTS", r"\bLDGSTS"),
         ("MATCH.ANY", r"MATCH\.ANY"), ("SYNCS", r"\bSYNCS"), ("STG.128.SYS", r"STG\.E\.128\.STRONG\.SYS"),
         ("LDG.128.SYS", r"LDG\.E\.128\.STRONG\.SYS")]
print("# SASS markers of the shipped mpmavatar_b200/libmpm_b200.so (cuobjdump -sass, sm_100a), per kernel (tools/sass_markers.py):")
print("# UBLKCP = cp.async.bulk (TMA), REDG.E.ADD.F32x4 = 16-byte vector atomic, FFMA2 = packed fp32 FMA, LDGSTS = cp.async, MATCH = match.any,")
print("# SYNCS = mbarrier, STG/LDG.128.SYS = the flagged-data stores / polls of the peer-to-peer exchange in k_grid_update<true>")
cur, counts = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = {k: 0 for k, _ in MARKS}
        continue
    if cur:
        for k, rx in MARKS:
            if re.search(rx, line):
                counts[cur][k] += 1
for fn in sorted(counts):
    if "3mpm" not in fn or not any(counts[fn].values()):
        continue
    print(f"{fn:<100}" + "  ".join(f"{k} {v:3d}" for k, v in counts[fn].items()))
arch = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
print("# arch:")
print(arch.strip())
