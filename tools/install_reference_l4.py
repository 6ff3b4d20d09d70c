"""Copies the reference's pure-Python layer, UNMODIFIED, out of /root/reference into baseline/_ref/ipcl_python_ref/
(git-ignored, like the base contract's `pip install --target baseline/_ref`; it travels to the GPU box, where
/root/reference does not exist).  tests/test_l4_flows.py::test_reference_l4_runs_over_the_real_shim loads it over this
repo's pybind11 shim.  Nothing in the product imports it.

    python tools/install_reference_l4.py
"""
import os
import shutil
import sys

SRC = "/root/reference/src/ipcl_python"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref", "ipcl_python_ref")


def install():
    if not os.path.exists(os.path.join(SRC, "ipcl_python.py")):
        return None
    os.makedirs(os.path.join(DST, "bindings"), exist_ok=True)
    shutil.copyfile(os.path.join(SRC, "ipcl_python.py"), os.path.join(DST, "ipcl_python.py"))
    shutil.copyfile(os.path.join(SRC, "bindings", "fixedpoint.py"), os.path.join(DST, "bindings", "fixedpoint.py"))
    return DST


if __name__ == "__main__":
    d = install()
    print(d or "no /root/reference here: nothing installed")
    sys.exit(0)
