"""``solvers`` -- the reference's solver wrappers over the B200 ``lap`` drop-in, plus the instance generators."""
from .lap_solver import LAPSolver, SeededLAPSolver  # noqa: F401
from .generators import (generate_uniform_costs, generate_sparse_costs, generate_metric_costs,  # noqa: F401
                         generate_clustered_costs, make_instance, mixed_batch, snap_to_fp32_grid)
from .advanced_dual import project_feasible, reduce_costs, check_dual_feasible  # noqa: F401,E402
from .dual_computation import compute_oracle_duals, compute_oracle_duals_batch, dual_from_matching_diff_constraints  # noqa: F401,E402
