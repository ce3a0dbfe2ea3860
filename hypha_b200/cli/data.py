"""`python -m hypha_b200.cli.data` == `hypha-data`."""
import sys

from . import main

if __name__ == "__main__":
    sys.exit(main("data"))
