"""Re-exports the case table of tests/golden/make_golden.py without running it."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py")
_spec = importlib.util.spec_from_file_location("make_golden", _p)
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
CASES, build_case, tweak = _m.CASES, _m.build_case, _m.tweak
