"""Drop-in mirror of the reference's warp_mpm package (module paths and public names of
/root/reference/warp_mpm), backed by libmpm_b200.so instead of NVIDIA Warp kernels."""
from .mpm_data_structure import MPMModelStruct, MPMStateStruct  # noqa: F401
from .mpm_solver import MPMWARP, MPMSolverWarp  # noqa: F401
