"""Alias of caduceus_b200.modeling_caduceus (same module path as ref:caduceus/modeling_caduceus.py)."""
from caduceus_b200.modeling_caduceus import *  # noqa: F401,F403
from caduceus_b200 import modeling_caduceus as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
