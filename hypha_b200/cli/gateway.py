"""`python -m hypha_b200.cli.gateway` == `hypha-gateway`."""
import sys

from . import main

if __name__ == "__main__":
    sys.exit(main("gateway"))
