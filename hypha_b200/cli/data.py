"""`python -m hypha_b200.cli.data` == `hypha-data`."""
import sys
from typing import List, Optional

from . import main


def run(argv: Optional[List[str]] = None) -> int:
    """Console-script entry point (`[project.scripts] hypha-data`)."""
    return main("data", argv)


if __name__ == "__main__":
    sys.exit(run())
