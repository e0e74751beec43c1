"""adtomo.jl_b200 -- B200-native Eikonal hot path (forward fast-sweeping solve + adjoint) of ADTomo.jl.

The product is the C-ABI CUDA library `libadtomo_b200.so` (sources in csrc/, ABI in
include/adtomo_b200.h).  This package is the thin host-side mirror of the reference's operator
interface (src/eikonal_op.jl: `eikonal`, `eikonal3d`) plus ctypes access to the batched and
fused entry points.  There is no CPU fallback: calls raise if the library or a GPU is missing.

The directory name contains a dot, so import it through the top-level shim `adtomo_jl_b200`.
"""
from .capi import (AdtomoError, Context, LIB_PATH, build_library, load_library, exported_symbols,  # noqa: F401
                   HOST, DEVICE)
from .eikonal_op import (eikonal, eikonal3d, eikonal_forward, eikonal_backward, eikonal3d_forward,  # noqa: F401
                         eikonal3d_backward)
from .inversion import (corner_sources, InversionProblem, shard_sources)  # noqa: F401
from .optimize import gpu_optimize, gpu_optimize_model, VelocityModel, DeviceVelocityModel, box_filter_periodic, save_checkpoint  # noqa: F401
