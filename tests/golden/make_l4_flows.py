"""Generates tests/golden/l4_flows.json by running the UNMODIFIED reference Python layer
(/root/reference/src/ipcl_python/ipcl_python.py + bindings/fixedpoint.py, loaded by oracle/ref_l4.py) over the
Python-int mock of ipcl_bindings, through the reference's own test flows (oracle/l4_flows.py).

    python tests/golden/make_l4_flows.py            # needs /root/reference; rewrites l4_flows.json

The file pins the L4 semantics of this repo's ipcl_python.py (exponent alignment, negative-plaintext inversion,
matmul index maps, add-tree padding) to the reference's code: tests/test_l4_flows.py replays the flows on the GPU and
compares every ciphertext integer, exponent and decoded value.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import l4_flows  # noqa: E402
import ref_l4  # noqa: E402


def generate():
    api = ref_l4.load_reference_l4(ref_l4.mock_bindings_module(), l4_dir="/root/reference/src/ipcl_python")
    return {"source": "reference ipcl_python.py (unmodified) over oracle/ref_l4.py MockBindings",
            "flows": {k: l4_flows.run_flows(api, k) for k in l4_flows.keyspecs()}}


if __name__ == "__main__":
    data = generate()
    path = os.path.join(HERE, "l4_flows.json")
    with open(path, "w") as f:
        json.dump(data, f, indent=0, separators=(",", ":"))
        f.write("\n")
    steps = sum(len(fl) for k in data["flows"].values() for fl in k.values())
    print("wrote %s: %d steps, %.1f KB" % (path, steps, os.path.getsize(path) / 1024))
