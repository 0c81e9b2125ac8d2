"""Drop-in for the reference's sert/math_utils.py."""
from sert_b200.math_utils import entropy  # noqa: F401
