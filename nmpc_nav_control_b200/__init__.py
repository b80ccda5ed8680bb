"""B200-native batched SQP-RTI NMPC solver for the omni4 / diff / tric models of
JorgeDFR/nmpc_nav_control (hot path only)."""
from .problem import MODELS, ModelSpec, get_model  # noqa: F401
