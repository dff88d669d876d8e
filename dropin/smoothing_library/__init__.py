"""`import smoothing_library as SL` resolving to the B200-native implementation.

Put this directory (dropin/) on PYTHONPATH ahead of the Pylians3 install; see INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from pylians3_b200.smoothing_library import *  # noqa: E402,F401,F403
