"""`python -m hypha_b200.cli.worker` == `hypha-worker`."""
import sys
from typing import List, Optional

from . import main


def run(argv: Optional[List[str]] = None) -> int:
    """Console-script entry point (`[project.scripts] hypha-worker`)."""
    return main("worker", argv)


if __name__ == "__main__":
    sys.exit(run())
