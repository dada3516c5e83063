"""compressor-mpc on B200: the per-sample-time MPC control step of katie-jones/compressor-mpc,
batched over independent plant scenarios and run as hand-written sm_100a CUDA kernels behind
the C ABI in include/cmpc.h.  Load with `__graft_entry__.load_package()` (the directory name
carries a hyphen) or put the repo root on sys.path and import `compressor_mpc_b200`."""
from . import capi, scenarios, setupfile, workflow  # noqa: F401
from .capi import CmpcError, measure_fp64_peak, plant_defaults  # noqa: F401
from .controller import Configuration, InputConstraints, NerveCenter, SubController, from_setup  # noqa: F401
from .setupfile import MODE_CENT, MODE_COOP, MODE_NCOOP, PLANT_PARALLEL, PLANT_SERIAL  # noqa: F401
