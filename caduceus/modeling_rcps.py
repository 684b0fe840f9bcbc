"""Alias of caduceus_b200.modeling_rcps (same module path as ref:caduceus/modeling_rcps.py)."""
from caduceus_b200.modeling_rcps import *  # noqa: F401,F403
from caduceus_b200 import modeling_rcps as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
