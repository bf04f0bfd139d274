#!/usr/bin/env python
"""Records HOW the reference's two callers call the hot-path API: every call on the solver / state / model objects in
/root/reference/train_material_params.py and run_demo.py (method name, number of positional arguments, keyword names,
attribute reads such as `mpm_solver.mesh.id` / `mpm_model.n_grid` / `mpm_state.particle_x`), extracted from the source
with `ast` -- the callers themselves cannot run offline (dataset, SMPL-X weights).  Output: tests/golden/caller_trace.json,
replayed by tests/test_boundary_cpu.py against this repo's mirror (signature binding on CPU, execution on the GPU box).

    python tests/golden/make_caller_trace.py
"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["/root/reference/train_material_params.py", "/root/reference/run_demo.py"]
OBJECTS = {"mpm_state": "MPMStateStruct", "mpm_model": "MPMModelStruct", "mpm_solver": "MPMWARP"}
CTORS = set(OBJECTS.values())


def receiver(node):
    """'mpm_state' for `mpm_state` / `self.mpm_state`, else None."""
    if isinstance(node, ast.Name) and node.id in OBJECTS:
        return node.id
    if isinstance(node, ast.Attribute) and node.attr in OBJECTS and isinstance(node.value, ast.Name) and node.value.id == "self":
        return node.attr
    return None


def main():
    calls, reads = [], []
    for path in FILES:
        tree = ast.parse(open(path).read())
        for n in ast.walk(tree):
            if isinstance(n, ast.Call):
                f = n.func
                rec = None
                if isinstance(f, ast.Name) and f.id in CTORS:
                    rec = dict(cls=f.id, method="__init__")
                elif isinstance(f, ast.Attribute) and receiver(f.value):
                    rec = dict(cls=OBJECTS[receiver(f.value)], method=f.attr)
                elif isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "wp":
                    rec = dict(cls="wp", method=f.attr)
                if rec:
                    rec.update(file=os.path.basename(path), line=n.lineno, n_pos=len(n.args),
                               keywords=[k.arg for k in n.keywords if k.arg is not None])
                    calls.append(rec)
            elif isinstance(n, ast.Attribute) and isinstance(n.ctx, ast.Load) and receiver(n.value):
                reads.append(dict(cls=OBJECTS[receiver(n.value)], attr=n.attr, file=os.path.basename(path), line=n.lineno))
    calls.sort(key=lambda c: (c["file"], c["line"]))
    called = {(c["cls"], c["method"]) for c in calls}
    reads = sorted({(r["cls"], r["attr"]) for r in reads if (r["cls"], r["attr"]) not in called})
    out = dict(source="ast of " + ", ".join(FILES), calls=calls, attribute_reads=[list(r) for r in reads])
    with open(os.path.join(HERE, "caller_trace.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(f"{len(calls)} calls, {len(reads)} attribute reads")


if __name__ == "__main__":
    main()
