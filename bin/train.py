#!/usr/bin/env python
"""bin/train.py of the reference (same flags, same data.npz / meta inputs, same model_<epoch>.bin outputs) on the
B200-native models; the driver lives in sert_b200/training.py."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from sert_b200 import training  # noqa: E402
from sert_b200.training import sparse_to_one_hot_multiple, train  # noqa: E402,F401  (names the reference's script exposes)

if __name__ == '__main__':
    sys.exit(training.main())
