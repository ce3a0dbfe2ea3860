"""`python -m hypha_b200.cli.scheduler` == `hypha-scheduler`."""
import sys

from . import main

if __name__ == "__main__":
    sys.exit(main("scheduler"))
