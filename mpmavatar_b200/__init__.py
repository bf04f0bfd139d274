"""mpmavatar_b200: B200-native (sm_100a) implementation of MPMAvatar's per-substep MPM garment
simulator behind the reference's own warp_mpm Python API.

    import mpmavatar_b200; mpmavatar_b200.install()
    from warp_mpm.mpm_solver import MPMWARP            # unchanged caller imports
    from warp_mpm.mpm_data_structure import MPMStateStruct, MPMModelStruct
"""
import functools
import importlib
import sys

__all__ = ["install"]


def _is_tensor(a) -> bool:
    try:
        import torch
        return isinstance(a, torch.Tensor)
    except Exception:
        return False


def _patch_real_warp(wp) -> None:
    """The MPMAvatar environment pins warp-lang (requirements.txt:36), and the callers hand solver state to
    `wp.to_torch` (train_material_params.py:628, :811; run_demo.py:532).  With the B200 solver that state IS a
    torch.Tensor, which real Warp's to_torch would dereference as a wp.array (a.device.is_cpu, a.ptr) and fail on: wrap
    to_torch / from_torch so that tensors pass through unchanged and everything else still reaches Warp."""
    if getattr(wp, "_mpmavatar_b200_patched", False):
        return
    real_to, real_from = getattr(wp, "to_torch", None), getattr(wp, "from_torch", None)

    def to_torch(a, *args, **kwargs):
        if _is_tensor(a):
            return a
        return real_to(a, *args, **kwargs)

    def from_torch(t, *args, **kwargs):
        return real_from(t, *args, **kwargs)

    if real_to is not None:
        wp.to_torch = functools.wraps(real_to)(to_torch)
        wt = sys.modules.get("warp.torch")
        if wt is not None and getattr(wt, "to_torch", None) is real_to:
            wt.to_torch = wp.to_torch
    if real_from is not None:
        wp.from_torch = functools.wraps(real_from)(from_torch)
    wp._mpmavatar_b200_patched = True


def install(force_warp_shim: bool = False) -> None:
    """Register this package's warp_mpm mirror under the reference's module paths
    (train_material_params.py:29-33).  `warp`: when NVIDIA Warp is importable (the real MPMAvatar environment) its
    to_torch is wrapped to pass the solver's torch tensors through; when it is absent a minimal shim provides
    wp.init() / wp.to_torch() / wp.synchronize()."""
    pkg = importlib.import_module(".warp_mpm", __name__)
    sys.modules["warp_mpm"] = pkg
    for sub in ("mpm_solver", "mpm_data_structure", "warp_utils"):
        sys.modules[f"warp_mpm.{sub}"] = importlib.import_module(f".warp_mpm.{sub}", __name__)
    wp = None
    if not force_warp_shim:
        try:
            import warp as wp  # noqa: F401
        except Exception:
            wp = None
    if wp is None:
        sys.modules["warp"] = importlib.import_module(".warp_shim", __name__)
    elif wp.__name__ != __name__ + ".warp_shim":
        _patch_real_warp(wp)
