"""``gnn`` -- host-side mirror of the reference's OneGNN interface (features + model), device-backed."""
from .features import compute_row_features, compute_row_features_torch, ROW_FEAT_DIM  # noqa: F401
from .one_gnn import OneGNN, ResidualBlock  # noqa: F401
