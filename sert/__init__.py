"""`sert` import surface of the reference (sert/__init__.py:1-4), served by the B200-native package."""
import sys

assert sys.version_info >= (3, 5)
