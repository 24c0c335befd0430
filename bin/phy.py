#!/usr/bin/env python
"""phy - likelihood of an alignment given a tree and a model, on the GPU (see phylo_utils_b200/cli.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phylo_utils_b200.cli import main  # noqa: E402

if __name__ == '__main__':
    sys.exit(main())
