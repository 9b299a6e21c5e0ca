"""Puts tests/golden on sys.path so that `import make_golden` works from any test."""
import os
import sys

_p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
if _p not in sys.path:
    sys.path.insert(0, _p)
