#!/usr/bin/env python
"""Drop-in entry point: same command line as the reference's after.py, B200 engine underneath."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from afterqc_b200.cli import main  # noqa: E402

if __name__ == "__main__":
    main()
