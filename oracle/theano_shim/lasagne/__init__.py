"""Minimal restatement of the Lasagne 0.1 API surface used by sert/models.py and bin/train.py
(test infrastructure; see ../README.md)."""
from lasagne import init, layers, nonlinearities, objectives, updates  # noqa: F401

__version__ = '0.1-shim'
