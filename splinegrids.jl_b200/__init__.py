"""splinegrids.jl_b200 -- B200-native grid-evaluation hot path of SplineGrids.jl.

Host-side mirror of the reference's Julia API over ``libsplinegrids_b200.so`` (hand-written sm_100a
CUDA kernels behind the C ABI in ``include/splinegrids_b200.h``).  Julia's ``f!`` is spelled ``f_``.
The directory name contains a dot, so import it through ``__graft_entry__.load_package()`` (or
``importlib``) under the module name ``splinegrids_jl_b200``.

There is no CPU fallback: every compute path needs the CUDA library and a CUDA device.
"""
from . import _lib
from ._lib import SplineGridsB200Error, last_variant, launch_count, launch_count_reset, set_kernel_policy
from .arrays import (as_colmajor, is_colmajor, jl_empty, jl_ones, jl_zeros, reshape_colmajor, to_device, to_numpy)
from .config import adjoint_plans, asynchronous, is_synchronous, set_adjoint_plans, set_synchronous
from .control_points import (DefaultControlPoints, LocallyRefinedControlPoints, LocalRefinement,
                             activate_local_control_point_range_, activate_local_refinement_, copyto_,
                             deactivate_overwritten_control_points_, get_n_control_points, obtain)
from .distributed import (PeerGradientExchange, SlabShardedGrid, allgather_support_planes_, allreduce_gradient_, owned_planes,
                          slab_bounds, slab_supports)
from .graphs import CapturedCalls
from .knot_vector import KnotVector
from .linear_map import SplineGridLinearMap
from .refinement import (add_default_local_refinement, boehm_refinement_matrix, error_informed_local_refinement_,
                         insert_knot, refine)
from .refinement_matrix import (RefinementMatrix, mult_, mult_adjoint_, refinement_matrix_from_dense, rmeye)
from .spline_dimension import SplineDimension, build_, decompress, set_sample_indices_
from .spline_grid import NURBSGrid, SplineGrid, evaluate_, evaluate_adjoint_, evaluate_multi_
from .validation import SplineGridsError
from .autograd import evaluate_autograd, make_zero_

__all__ = [
    "SplineDimension", "SplineGrid", "NURBSGrid", "KnotVector", "RefinementMatrix", "DefaultControlPoints",
    "LocallyRefinedControlPoints", "LocalRefinement", "evaluate_", "evaluate_adjoint_", "evaluate_multi_", "mult_", "mult_adjoint_",
    "rmeye", "refinement_matrix_from_dense", "decompress", "set_sample_indices_", "build_", "insert_knot", "refine",
    "add_default_local_refinement", "activate_local_refinement_", "activate_local_control_point_range_",
    "deactivate_overwritten_control_points_", "error_informed_local_refinement_", "get_n_control_points", "obtain",
    "copyto_", "SplineGridLinearMap", "SlabShardedGrid", "PeerGradientExchange", "allreduce_gradient_", "slab_bounds", "slab_supports", "owned_planes", "allgather_support_planes_", "to_device", "to_numpy",
    "jl_zeros", "jl_ones", "jl_empty", "reshape_colmajor", "is_colmajor", "as_colmajor", "set_synchronous",
    "is_synchronous", "asynchronous", "set_kernel_policy", "last_variant", "launch_count", "launch_count_reset",
    "SplineGridsError", "SplineGridsB200Error", "boehm_refinement_matrix", "CapturedCalls", "set_adjoint_plans",
    "adjoint_plans", "evaluate_autograd", "make_zero_",
]
