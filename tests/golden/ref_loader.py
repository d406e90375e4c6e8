"""The stub loader for the unmodified reference lives in baseline/ref_loader.py (bench.py's reference arm uses it
too); this shim keeps `import ref_loader` working for make_golden.py and tests/test_oracle_vs_reference.py."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from baseline.ref_loader import *  # noqa: F401,F403,E402
from baseline.ref_loader import REFERENCE_ROOT, load_reference_synthesize, reference_available  # noqa: F401,E402
