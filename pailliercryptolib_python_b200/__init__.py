"""B200-native Paillier engine with the ipcl_python API surface.

    from pailliercryptolib_python_b200 import PaillierKeypair, PaillierPublicKey, PaillierPrivateKey, \
        PaillierEncryptedNumber, context, hybridControl, hybridMode      # same names as ipcl_python/__init__.py:4-11

The native pieces are built in-tree by `python -m pailliercryptolib_python_b200.build`; names resolve lazily so
`pailliercryptolib_python_b200.capi` (ctypes view of the C ABI) and `.build` import without the pybind11 module.
"""
__version__ = "0.1.0"

_API = {
    "PaillierKeypair": "ipcl_python", "PaillierPublicKey": "ipcl_python", "PaillierPrivateKey": "ipcl_python",
    "PaillierEncryptedNumber": "ipcl_python", "BNUtils": "ipcl_python",
    "FixedPointNumber": "fixedpoint", "FixedPointEndec": "fixedpoint",
    "context": "bindings.ipcl_bindings", "hybridControl": "bindings.ipcl_bindings", "hybridMode": "bindings.ipcl_bindings",
    "ipclKeypair": "bindings.ipcl_bindings", "ipclPublicKey": "bindings.ipcl_bindings", "ipclPrivateKey": "bindings.ipcl_bindings",
    "ipclPlainText": "bindings.ipcl_bindings", "ipclCipherText": "bindings.ipcl_bindings", "ipclBigNumber": "bindings.ipcl_bindings",
}
__all__ = sorted(_API)


def __getattr__(name):
    if name in _API:
        import importlib
        try:
            mod = importlib.import_module("." + _API[name], __name__)
        except ImportError as e:
            raise ImportError("%s needs the native modules: run `python -m pailliercryptolib_python_b200.build` (%s)" % (name, e))
        return getattr(mod, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
