"""project-marshmallow_b200 -- B200-native cloud ray-march pass (sm_100a CUDA behind a C-ABI).

The directory name carries a hyphen (it mirrors the upstream repository name), so it is imported
through `load_package()` in the repo-root helper `_pkg.py`, under the module name
`project_marshmallow_b200`.

Layout:
    csrc/   hand-written CUDA kernels + the C-ABI implementation (include/marshmallow.h)
    host/   C++ host-side mirrors of the reference's SkyManager / Camera value producers and of its
            ComputeShader pipeline wrapper (CloudComputePass.h)
    capi.py ctypes binding of the C-ABI (no torch types cross it)
    host.py Python mirror of the reference's ComputeShader interface, used by tests and bench
"""
from .capi import (  # noqa: F401
    MM_FULL, MM_PHASE16, MM_ROWS_SNAKE, MM_FILTER_EXACT, MM_FILTER_HW, MM_FILTER_HYBRID,
    MM_SCHED_AUTO, MM_SCHED_STATIC, MM_SCHED_PERSISTENT, MM_SCHED_PACKED, MM_ARITH_IEEE, MM_ARITH_FMA,
    MM_TEX_PLACEMENT, MM_TEX_NIGHTSKY, MM_TEX_CURL, MM_TEX_LOWRES, MM_TEX_HIRES,
    MarshmallowError, library_path, load_library, exported_symbols, host_sky, host_camera, plan_block_rows,
)
from .host import ComputeShader, SkyManager, Camera, dispatchMulti  # noqa: F401
from . import multigpu  # noqa: F401
