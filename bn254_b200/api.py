"""Host-side mirror of the reference crate's public API (/root/reference/src/lib.rs:60-63) over the CUDA engine.

Same names, argument meaning and error behaviour as the Rust crate, so the parity tests read like the
reference's own tests:

    PrivateKey / PublicKey / PublicKeyG1 / Signature      /root/reference/src/types.rs:13-286
    ECDSA.sign / ECDSA.verify, check_public_keys          /root/reference/src/ecdsa.rs:26-93
    Error (11 variants), raised instead of Result::Err    /root/reference/src/error.rs:5-29
    serde forms of PrivateKey / PublicKey                 /root/reference/src/serde.rs:10-56

Every arithmetic operation runs on the GPU through the C ABI (batch size 1 here; the batch entry points live in
bn254_b200.engine).  Points are held as the crate's uncompressed bytes (all-zero = the point at infinity, which
the crate can hold but not serialise).
"""
from . import engine as _engine
from ._native import Context

ERROR_NAMES = {
    1: "HashToPointError", 2: "IndexOutOfBounds", 3: "InvalidEncoding", 4: "InvalidGroupPoint", 5: "InvalidLength",
    6: "NotMemberError", 7: "ToAffineConversion", 8: "PointInJacobian", 9: "VerificationFailed", 10: "SerializationError",
    11: "HexDecodeFailed",
    255: "EngineFault",  # not a variant of the crate: the engine did not evaluate the item (include/bn254_b200.h BN254_ENGINE_FAULT)
}


class Error(Exception):
    """One instance per Error variant of /root/reference/src/error.rs:6-29; `.code` is the C-ABI status byte."""

    def __init__(self, code):
        self.code = int(code)
        self.variant = ERROR_NAMES.get(self.code, "Unknown(%d)" % self.code)
        super().__init__(self.variant)


class _TypedEngine:
    """The engine module bound to a context of its own with BN254_INPUTS_TYPED: the objects below ARE values of the crate's
    types (validated by their constructors, possibly infinity after + / -), which is exactly what that policy means."""

    def __init__(self):
        self._ctx = None

    def ctx(self):
        if self._ctx is None:
            self._ctx = Context(0)
            _engine.set_input_policy(_engine.INPUTS_TYPED, ctx=self._ctx)
        return self._ctx

    def __getattr__(self, name):
        fn = getattr(_engine, name)
        return lambda *a, **k: fn(*a, ctx=self.ctx(), **k)


E = _TypedEngine()


def _check(st):
    if st:
        raise Error(st)


def _hex(s):
    try:
        return bytes.fromhex(s)
    except ValueError:
        raise Error(11)


R_ORDER = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


class PrivateKey:
    """PrivateKey(Fr) -- /root/reference/src/types.rs:13-77.  Any 32 bytes are accepted and reduced mod r (Fr::from_slice)."""

    def __init__(self, data):
        if isinstance(data, str):  # TryFrom<&str>: hex
            data = _hex(data)
        data = bytes(data)
        if len(data) != 32:
            raise Error(5)  # InvalidLength (/root/reference/src/types_test.rs:29-46)
        self._k = int.from_bytes(data, "big") % R_ORDER

    @classmethod
    def random(cls, rng=None):
        import secrets
        k = (rng.randrange(R_ORDER) if rng is not None else secrets.randbelow(R_ORDER))
        return cls(k.to_bytes(32, "big"))

    def to_bytes(self):  # /root/reference/src/utils.rs:66-72
        return self._k.to_bytes(32, "big")

    def to_hex(self):  # String: TryFrom<PrivateKey>
        return self.to_bytes().hex()

    # serde: sequence of 32 u8 (/root/reference/src/serde.rs:10-35)
    def serialize(self):
        return list(self.to_bytes())

    @classmethod
    def deserialize(cls, seq):
        return cls(bytes(seq))


class _Point:
    SIZE = 0
    CSIZE = 0

    def __init__(self, raw):
        self.raw = bytes(raw)

    # --- byte formats
    @classmethod
    def from_uncompressed(cls, data):
        data = bytes(data)
        if len(data) != cls.SIZE:
            raise Error(5)
        _check(cls._validate(data)[0])
        return cls(data)

    @classmethod
    def from_compressed(cls, data):
        data = bytes(data)
        if len(data) != cls.CSIZE:
            raise Error(3)  # InvalidEncoding from bn::G*::from_compressed
        raw, st = cls._decompress(data)
        _check(st[0])
        return cls(raw)

    def to_uncompressed(self):
        if not any(self.raw):
            raise Error(8)  # PointInJacobian: infinity has no affine form (/root/reference/src/utils.rs:163,184)
        return self.raw

    def to_compressed(self):
        c, st = self._compress(self.raw)
        _check(st[0])
        return c

    # --- aggregation operators (/root/reference/src/types.rs:126-148,196-218,264-286)
    def __add__(self, other):
        r, st = self._sum(self.raw + other.raw, None)
        _check(st)
        return type(self)(r)

    def __sub__(self, other):
        r, st = self._sum(self.raw + other.raw, bytes([0, 1]))
        _check(st)
        return type(self)(r)

    def __neg__(self):
        r, st = self._sum(self.raw, bytes([1]))
        _check(st)
        return type(self)(r)

    def __eq__(self, other):
        return type(self) is type(other) and self.raw == other.raw

    def __hash__(self):
        return hash(self.raw)


class _G1Point(_Point):
    SIZE, CSIZE = 64, 33
    _validate = staticmethod(lambda d: E.g1_validate_batch(d))
    _decompress = staticmethod(lambda d: E.g1_decompress_batch(d))
    _compress = staticmethod(lambda d: E.g1_compress_batch(d))
    _sum = staticmethod(lambda d, neg: E.g1_sum(d, neg))


class _G2Point(_Point):
    SIZE, CSIZE = 128, 65
    _validate = staticmethod(lambda d: E.g2_validate_batch(d))
    _decompress = staticmethod(lambda d: E.g2_decompress_batch(d))
    _compress = staticmethod(lambda d: E.g2_compress_batch(d))
    _sum = staticmethod(lambda d, neg: E.g2_sum(d, neg))


class PublicKey(_G2Point):
    """PublicKey(G2) -- /root/reference/src/types.rs:81-148."""

    @classmethod
    def from_private_key(cls, sk):
        return cls(E.derive_pk_g2_batch(sk.to_bytes()))

    # serde: sequence of 65 u8, the compressed form (/root/reference/src/serde.rs:37-56)
    def serialize(self):
        return list(self.to_compressed())

    @classmethod
    def deserialize(cls, seq):
        return cls.from_compressed(bytes(seq))


class PublicKeyG1(_G1Point):
    """PublicKeyG1(G1) -- /root/reference/src/types.rs:151-218."""

    @classmethod
    def from_private_key(cls, sk):
        return cls(E.derive_pk_g1_batch(sk.to_bytes()))


class Signature(_G1Point):
    """Signature(G1) -- /root/reference/src/types.rs:221-286."""


class ECDSA:
    """/root/reference/src/ecdsa.rs:13-64."""

    @staticmethod
    def sign(message, private_key):
        message = bytes(message)
        sig, st = E.sign_batch(message if message else None, len(message), private_key.to_bytes())
        _check(st[0])
        return Signature(sig)

    @staticmethod
    def verify(message, signature, public_key):
        """Returns None on success, raises Error(VerificationFailed) otherwise (Result<()> in the crate)."""
        message = bytes(message)
        st = E.verify_batch(message if message else None, len(message), signature.raw, public_key.raw)
        _check(st[0])


def check_public_keys(public_key_g2, public_key_g1):
    """/root/reference/src/ecdsa.rs:78-93."""
    st = E.check_public_keys_batch(public_key_g2.raw, public_key_g1.raw)
    _check(st[0])


# ---------------------------------------------------------------------------------------------- precompile input formatters
# /root/reference/src/utils.rs:197-239.  Output = [(G1 64 B, G2 128 B); 2] in the dependency's Borsh form: every 32-byte
# coordinate little-endian, G1 = x || y, G2 = x.re || x.im || y.re || y.im (this is exactly what the reference's
# `_uncompressed_` variant produces by reversing each 32-byte chunk of the crate's big-endian uncompressed bytes,
# :223-229).  No reference test covers these functions: parity is unpinned.
NEG_G2_UNCOMPRESSED = bytes.fromhex(
    "1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed"
    "198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2"
    "1d9befcd05a5323e6da4d435f3b617cdb3af83285c2df711ef39c01571827f9d"
    "275dc4a288d1afb3cbb1ac09187524c7db36395df7be3b99e673b13a075a65ec")


def _pairs(blob):
    return [(blob[0:64], blob[64:192]), (blob[192:256], blob[256:384])]


def format_pairing_check_values(message, signature, public_key):
    """(message, 33-byte compressed signature, 65-byte compressed public key) -> [(H(m), pk), (sig, -G2)]"""
    message, signature, public_key = bytes(message), bytes(signature), bytes(public_key)
    if len(public_key) != 65 or len(signature) != 33:
        raise Error(3)  # InvalidEncoding from bn::G2 / G1::from_compressed (public key first, like the reference)
    blob, st = E.format_pairing_check_batch(message if message else None, len(message), signature, public_key, True)
    _check(st[0])
    return _pairs(blob)


def format_pairing_check_uncompressed_values(message, signature, public_key):
    """(message, 64-byte uncompressed signature, 128-byte uncompressed public key); like the reference, the point bytes
    are re-ordered without being validated (/root/reference/src/utils.rs:218-239)."""
    message, signature, public_key = bytes(message), bytes(signature), bytes(public_key)
    if len(signature) != 64 or len(public_key) != 128:
        # over-long inputs: the reference's Vec<u8> -> [u8; N] try_into fails, which maps to SerializationError
        # (/root/reference/src/error.rs:64-68); shorter ones make its slice indexing panic -- reported the same way here
        raise Error(10)
    blob, st = E.format_pairing_check_batch(message if message else None, len(message), signature, public_key, False)
    _check(st[0])
    return _pairs(blob)
