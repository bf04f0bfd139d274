"""Print the lines of a tools/sweep_scene.sh output: value, e2e, parity and the sharded timeline."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        tl = (d.get("roofline") or {}).get("timeline_us") or d.get("timeline_us") or {}
        par = d.get("parity_vs_single_gpu") or {}
        print(f, "N=%d value=%.0f e2e=%.0f ms/step=%.2f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]),
              "parity x=%.2g v=%.2g" % (par.get("x", 0), par.get("v", 0)) if par else "")
        if tl.get("kernels"):
            print("   ", {k: (round(v["start_us"], 1), round(v["dur_us"], 1)) for k, v in tl["kernels"].items()},
                  {k: v for k, v in tl.items() if k != "kernels"})
