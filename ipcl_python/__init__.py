"""Drop-in alias: `import ipcl_python` resolves to the B200 engine (pailliercryptolib_python_b200) with the reference
package's names (/root/reference/src/ipcl_python/__init__.py:4-11)."""
from pailliercryptolib_python_b200.ipcl_python import (  # noqa: F401
    BNUtils,
    PaillierEncryptedNumber,
    PaillierKeypair,
    PaillierPrivateKey,
    PaillierPublicKey,
)
from pailliercryptolib_python_b200.bindings.ipcl_bindings import context, hybridControl, hybridMode  # noqa: F401

__all__ = ["PaillierKeypair", "PaillierPublicKey", "PaillierPrivateKey", "PaillierEncryptedNumber", "context",
           "hybridControl", "hybridMode"]
