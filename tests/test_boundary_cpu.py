"""The drop-in boundary without a GPU:
  * every call the reference's two callers make on the solver / state / model objects (recorded from their source by
    tests/golden/make_caller_trace.py) binds to this repo's mirror -- same method names, positional counts, keywords;
  * mpmavatar_b200.install() in an environment where the real NVIDIA Warp IS installed (the MPMAvatar environment pins
    warp-lang, requirements.txt:36): wp.to_torch(state.particle_x) must hand the solver's torch tensor through instead of
    dereferencing it as a wp.array (train_material_params.py:628, :811; run_demo.py:532);
  * the lazy state / model fields pull before they are overwritten."""
import inspect
import json
import os
import sys
import types

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
TRACE = json.load(open(os.path.join(HERE, "golden", "caller_trace.json")))


def _classes():
    from mpmavatar_b200.warp_mpm.mpm_data_structure import MPMModelStruct, MPMStateStruct
    from mpmavatar_b200.warp_mpm.mpm_solver import MPMWARP
    return {"MPMStateStruct": MPMStateStruct, "MPMModelStruct": MPMModelStruct, "MPMWARP": MPMWARP}


def test_trace_covers_both_callers():
    files = {c["file"] for c in TRACE["calls"]}
    assert files == {"train_material_params.py", "run_demo.py"}
    methods = {(c["cls"], c["method"]) for c in TRACE["calls"]}
    for need in [("MPMWARP", "p2g2p"), ("MPMStateStruct", "reset_state"), ("MPMStateStruct", "continue_from_torch"),
                 ("MPMWARP", "add_surface_collider"), ("wp", "to_torch"), ("wp", "init")]:
        assert need in methods


@pytest.mark.parametrize("call", [c for c in TRACE["calls"] if c["cls"] != "wp"],
                         ids=lambda c: f"{c['file']}:{c['line']}:{c['method']}")
def test_reference_call_binds_to_mirror(call):
    cls = _classes()[call["cls"]]
    fn = getattr(cls, call["method"])
    sig = inspect.signature(fn)
    args = [object()] * (call["n_pos"] + 1)  # + self
    sig.bind(*args, **{k: object() for k in call["keywords"]})  # raises TypeError on a mismatch


def test_reference_attribute_reads_exist():
    cl = _classes()
    from mpmavatar_b200.warp_mpm.mpm_data_structure import MPMModelStruct, MPMStateStruct
    m = MPMModelStruct()
    m.init_other_params(n_grid=32, grid_lim=2.0, device="cpu")
    assert m.n_grid == 32
    s = MPMStateStruct()
    s.init(6, 2, 3, device="cpu")
    assert tuple(s.particle_x.shape) == (6, 3)
    src = inspect.getsource(cl["MPMWARP"].initialize)
    assert "self.mesh" in src  # .mesh.id is read by add_mesh_collider's caller (train_material_params.py:505)
    for cls_name, attr in TRACE["attribute_reads"]:
        assert (cls_name, attr) in {("MPMModelStruct", "n_grid"), ("MPMStateStruct", "particle_x"), ("MPMWARP", "mesh")}


def _fake_real_warp():
    """A module that behaves like real Warp where it matters: to_torch insists on a wp.array."""
    wp = types.ModuleType("warp")

    class array:  # noqa: N801
        def __init__(self, t):
            self._t = t
            self.device = types.SimpleNamespace(is_cpu=True)
            self.ptr = t.data_ptr()

    def to_torch(a, requires_grad=None):
        if a.device.is_cpu and a.ptr:  # AttributeError on a torch.Tensor: device is a torch.device
            return a._t
        raise RuntimeError

    def from_torch(t, dtype=None, requires_grad=None, grad=None):
        return array(t)
    wp.array, wp.to_torch, wp.from_torch = array, to_torch, from_torch
    wp.init = lambda: None
    wp.synchronize = lambda: None
    wt = types.ModuleType("warp.torch")
    wt.to_torch, wt.from_torch = to_torch, from_torch
    wp.torch = wt
    return wp, wt


def test_install_with_real_warp_passes_tensors_through(monkeypatch):
    import mpmavatar_b200
    wp, wt = _fake_real_warp()
    monkeypatch.setitem(sys.modules, "warp", wp)
    monkeypatch.setitem(sys.modules, "warp.torch", wt)
    for k in [k for k in sys.modules if k == "warp_mpm" or k.startswith("warp_mpm.")]:
        monkeypatch.delitem(sys.modules, k)
    t = torch.arange(6.0).reshape(2, 3)
    with pytest.raises(AttributeError):
        wp.to_torch(t)  # what the unchanged caller would hit without the wrapper
    mpmavatar_b200.install()
    assert sys.modules["warp"] is wp  # the real module stays; only to_torch is wrapped
    import warp
    from warp_mpm.mpm_data_structure import MPMStateStruct
    st = MPMStateStruct()
    st.init(4, 1, 2, device="cpu")
    assert warp.to_torch(st.particle_x) is st.particle_x
    assert warp.to_torch(st.particle_x).clone().shape == (4, 3)  # train_material_params.py:628
    a = warp.from_torch(t)
    assert isinstance(a, wp.array) and warp.to_torch(a) is t  # genuine wp.arrays still reach Warp
    assert sys.modules["warp.torch"].to_torch is warp.to_torch
    mpmavatar_b200.install()  # idempotent
    assert warp.to_torch(t) is t


def test_install_without_warp_uses_shim(monkeypatch):
    import mpmavatar_b200
    monkeypatch.delitem(sys.modules, "warp", raising=False)
    mpmavatar_b200.install(force_warp_shim=True)
    import warp
    t = torch.zeros(3)
    assert warp.to_torch(t) is t
    warp.init()


class _FakeSolver:
    """Stands in for MPMWARP on CPU: counts exports and writes a marker into the canonical tensors."""

    def __init__(self):
        self.exports = 0
        self._bound_state = self._bound_model = None

    def _export_into(self, state):
        self.exports += 1
        state._stale = False
        state._particle_v.fill_(7.0)
        if self._bound_model is not None and not self._bound_model._dirty:
            self._bound_model._mu.fill_(3.0)


def test_lazy_fields_pull_before_they_are_overwritten():
    from mpmavatar_b200.warp_mpm.mpm_data_structure import MPMModelStruct, MPMStateStruct
    st, md, sv = MPMStateStruct(), MPMModelStruct(), _FakeSolver()
    st.init(5, 1, 2, device="cpu")
    md.init(5, device="cpu")
    st._solver = md._solver = sv
    sv._bound_state, sv._bound_model = st, md
    st._dirty = md._dirty = False
    # state.particle_v = t after a step: the pending export must land in the OLD tensor, not in t
    st._stale = True
    old_v = st._particle_v
    t = torch.ones(5, 3)
    st.particle_v = t
    assert sv.exports == 1 and float(old_v[0, 0]) == 7.0 and float(t[0, 0]) == 1.0 and st.particle_v is t
    # model.mu = t after a step (set_E_nu + prepare_mu_lam, train_material_params.py:607-610)
    st._stale = True
    md._dirty = False
    old_mu = md._mu
    new_mu = torch.full((5,), 2.0)
    md.mu = new_mu
    assert sv.exports == 2 and float(old_mu[0]) == 3.0 and float(new_mu[0]) == 2.0 and md._dirty
    assert md.mu is new_mu and sv.exports == 2  # nothing stale any more: no further export
