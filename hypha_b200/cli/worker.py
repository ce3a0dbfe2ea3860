"""`python -m hypha_b200.cli.worker` == `hypha-worker`."""
import sys

from . import main

if __name__ == "__main__":
    sys.exit(main("worker"))
