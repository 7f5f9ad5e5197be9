"""hydrograd.jl_b200 -- B200-native (sm_100a) replacement of Hydrograd.jl's 2-D shallow-water RHS / VJP hot path.

The directory name contains a dot, so it is loaded through `_pkg.load()` at the repo root (module name
`hydrograd_jl_b200`).  Everything that computes lives in csrc/ (CUDA) behind include/hydrograd_b200.h."""
from . import _lib  # noqa: F401
from .api import (Context, plan_stats, plan_pipeline, plan_tables, HydrogradError, SWE2D_Extra_Parameters, custom_ODE_solve, swe_2D_consts,  # noqa: F401
                  swe_2d_rhs, swe_2d_rhs_pullback)
