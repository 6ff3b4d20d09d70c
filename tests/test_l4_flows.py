"""L4 parity against REFERENCE-GENERATED golden data.

tests/golden/l4_flows.json was produced by the unmodified reference Python layer (ipcl_python.py + fixedpoint.py) run
over a Python-int mock of its compiled bindings (oracle/ref_l4.py, tests/golden/make_l4_flows.py): the reference's own
test flows (tests/ipcl_python_test.py:21-119) and operator variants, obfuscator off, seeded inputs.

CPU (-m "not gpu"): the fixture regenerates identically where /root/reference exists, and its plain encrypt steps are
re-derived from the Python-int oracle alone.
GPU (-m gpu): the same flows through this repo's ipcl_python + CUDA must give the same ciphertext integers, exponents
and decoded values, bit for bit; and so must the reference's ipcl_python.py itself when it runs over this repo's
pybind11 shim (the drop-in claim of INTEGRATION.md section 1), where a copy of the reference L4 is available.
"""
import json
import os

import numpy as np
import pytest

import l4_flows
import paillier_oracle as O
import ref_l4

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "l4_flows.json")) as f:
    GOLDEN = json.load(f)["flows"]


def _compare(got, want, where):
    assert list(got) == list(want), "%s: step names differ" % where
    for step, w in want.items():
        g = got[step]
        for key in ("count", "expo", "ct_first", "ct_last", "ct_sha256"):
            assert g[key] == w[key], "%s / %s: %s differs" % (where, step, key)
        assert len(g["dec"]) == len(w["dec"])
        for a, b in zip(g["dec"], w["dec"]):
            assert type(a) is type(b) and a == b, "%s / %s: decoded value %r != %r" % (where, step, a, b)


@pytest.mark.skipif(not os.path.exists("/root/reference/src/ipcl_python/ipcl_python.py"), reason="needs /root/reference")
def test_fixture_regenerates_from_the_reference():
    api = ref_l4.load_reference_l4(ref_l4.mock_bindings_module(), l4_dir="/root/reference/src/ipcl_python")
    got = l4_flows.run_flows(api, "bench2048")
    for flow, want in GOLDEN["bench2048"].items():
        _compare(got[flow], want, "bench2048/" + flow)


def test_fixture_encrypt_steps_match_the_oracle():
    """The plain encrypt steps of the fixture re-derived without the reference: fp_encode + (1 + m n)."""
    rs = np.random.RandomState(11)
    for keyname, count in (("seeded1024", 100), ("bench2048", 24)):
        n, p, q, bits = l4_flows.keyspecs()[keyname]
        pk = O.PubKey(n, bits, False)
        rs = np.random.RandomState(11)
        x = np.ones(count) * rs.randint(100)
        enc = [O.fp_encode(float(v), n, n // 3 - 1) for v in x]
        want = GOLDEN[keyname]["add"]["en_x"]
        assert [e for _, e in enc] == want["expo"]
        assert l4_flows.digest([O.raw_encrypt(pk, m) for m, _ in enc])["ct_sha256"] == want["ct_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("keyname", ["seeded1024", "bench2048"])
def test_replay_through_this_repo(keyname):
    import pailliercryptolib_python_b200 as api
    got = l4_flows.run_flows(api, keyname)
    for flow, want in GOLDEN[keyname].items():
        _compare(got[flow], want, "%s/%s" % (keyname, flow))


@pytest.mark.gpu
@pytest.mark.skipif(ref_l4.reference_l4_dir() is None, reason="no copy of the reference L4 (tools/install_reference_l4.py)")
def test_reference_l4_runs_over_the_real_shim():
    from pailliercryptolib_python_b200.bindings import ipcl_bindings
    api = ref_l4.load_reference_l4(ipcl_bindings, name="_ref_ipcl_python_shim")
    got = l4_flows.run_flows(api, "bench2048")
    for flow, want in GOLDEN["bench2048"].items():
        _compare(got[flow], want, "reference L4 over shim: " + flow)
