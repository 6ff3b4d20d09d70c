from pailliercryptolib_python_b200.bindings import ipcl_bindings  # noqa: F401
from pailliercryptolib_python_b200 import fixedpoint  # noqa: F401
import sys as _sys

_sys.modules[__name__ + ".ipcl_bindings"] = ipcl_bindings
_sys.modules[__name__ + ".fixedpoint"] = fixedpoint
