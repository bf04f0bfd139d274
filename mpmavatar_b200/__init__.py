"""mpmavatar_b200: B200-native (sm_100a) implementation of MPMAvatar's per-substep MPM garment
simulator behind the reference's own warp_mpm Python API.

    import mpmavatar_b200; mpmavatar_b200.install()
    from warp_mpm.mpm_solver import MPMWARP            # unchanged caller imports
    from warp_mpm.mpm_data_structure import MPMStateStruct, MPMModelStruct
"""
import importlib
import sys

__all__ = ["install"]


def install(force_warp_shim: bool = False) -> None:
    """Register this package's warp_mpm mirror under the reference's module paths
    (train_material_params.py:29-33) and, when NVIDIA Warp is absent, a minimal `warp` shim
    for wp.init()/wp.to_torch()."""
    pkg = importlib.import_module(".warp_mpm", __name__)
    sys.modules["warp_mpm"] = pkg
    for sub in ("mpm_solver", "mpm_data_structure", "warp_utils"):
        sys.modules[f"warp_mpm.{sub}"] = importlib.import_module(f".warp_mpm.{sub}", __name__)
    have_warp = False
    if not force_warp_shim:
        try:
            import warp  # noqa: F401
            have_warp = True
        except Exception:
            have_warp = False
    if not have_warp:
        sys.modules["warp"] = importlib.import_module(".warp_shim", __name__)
