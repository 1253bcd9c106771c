"""b200lap -- host-side runtime of the B200 (sm_100a) warm-start LAP hot path.

    Context, Model        device contexts / packed OneGNN weights (runtime.py)
    GNNPredictor          features -> OneGNN -> min-trick, the reference's inference glue (predictor.py)
    shard_bounds, solve_sharded, WorkQueue, drain_queue   instance-level sharding across the GPUs of one box:
                          static blocks or a dynamically drained queue (sharding.py)

The compute lives in ``libb200lap.so`` (C ABI in include/b200lap.h, sources in ../csrc); importing
this package does not require a GPU, calling into it does.
"""
from ._lib import B200LapError, LIB_PATH, ROW_FEAT_DIM, TRACE_NAMES, load  # noqa: F401
from .runtime import Context, HostPipeline, Model, default_context, pack_state_dict, state_dict_order, trace_dict  # noqa: F401
from .predictor import GNNPredictor  # noqa: F401
from .sharding import WorkQueue, bind_to_device_numa_node, drain_queue, shard_bounds, solve_sharded  # noqa: F401
