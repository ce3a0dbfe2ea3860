"""`python -m hypha_b200.cli.scheduler` == `hypha-scheduler`."""
import sys
from typing import List, Optional

from . import main


def run(argv: Optional[List[str]] = None) -> int:
    """Console-script entry point (`[project.scripts] hypha-scheduler`)."""
    return main("scheduler", argv)


if __name__ == "__main__":
    sys.exit(run())
