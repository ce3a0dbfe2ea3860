"""hypha-gateway role (implementation in `hypha_b200.roles`)."""
from ..roles import run_gateway  # noqa: F401
