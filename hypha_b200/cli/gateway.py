"""`python -m hypha_b200.cli.gateway` == `hypha-gateway`."""
import sys
from typing import List, Optional

from . import main


def run(argv: Optional[List[str]] = None) -> int:
    """Console-script entry point (`[project.scripts] hypha-gateway`)."""
    return main("gateway", argv)


if __name__ == "__main__":
    sys.exit(run())
