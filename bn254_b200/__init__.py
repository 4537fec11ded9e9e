"""bn254_b200 -- B200-native batch engine for the BN254 hash / sign / aggregate / pairing-verify path.

`bn254_b200.api` mirrors the reference crate's public API; `bn254_b200.engine` holds the batch entry points;
both call the CUDA library (bn254_b200/libbn254_b200.so) through the C ABI of include/bn254_b200.h.
Importing this package does not need a GPU; creating a context (first call) does, and fails loudly without one.
"""
from ._native import Context, EngineError, LIB_PATH  # noqa: F401
from .api import (ECDSA, Error, PrivateKey, PublicKey, PublicKeyG1, Signature, check_public_keys,  # noqa: F401
                  format_pairing_check_uncompressed_values, format_pairing_check_values)
from . import engine  # noqa: F401
