"""hypha-data role (implementation in `hypha_b200.roles`)."""
from ..roles import run_data  # noqa: F401
