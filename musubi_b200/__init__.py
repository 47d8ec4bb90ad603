"""musubi_b200 -- B200-native per-level LBM time step behind Musubi's plugin
surface (libmusb200.so, sm_100a).  Host-side mirror of the reference interface
for this path; see DESIGN.md and INTEGRATION.md."""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .scheme import (Scheme, compute_host, get_unique_id, multilevel_tables, mus_finalize,  # noqa: F401
                     mus_init, select_kernel, step_schemes)
from .treelm import DeviceCube, LevelDesc  # noqa: F401
from ._lib import Musb200Error  # noqa: F401
