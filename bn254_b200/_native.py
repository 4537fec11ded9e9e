"""ctypes binding of the C ABI declared in include/bn254_b200.h.  No CPU fallback: if the CUDA library or a CUDA
device is missing, loading / context creation raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BN254_B200_LIB") or os.path.join(HERE, "libbn254_b200.so")  # override: tuning builds only

# every symbol include/bn254_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "bn254_ctx_create", "bn254_ctx_destroy", "bn254_last_error", "bn254_sync", "bn254_stream", "bn254_sm_count", "bn254_launch_count",
    "bn254_set_profiling", "bn254_phase_ms", "bn254_set_pairing_mode", "bn254_trim", "bn254_set_input_policy", "bn254_get_input_policy", "bn254_set_hash_try_limit", "bn254_set_test_fault",
    "bn254_hash_to_g1_batch", "bn254_hash_to_g1_batch_dev", "bn254_hash_to_g1_var",
    "bn254_sign_batch", "bn254_sign_batch_dev", "bn254_verify_batch", "bn254_verify_batch_dev",
    "bn254_verify_batch_rlc", "bn254_verify_batch_rlc_dev",
    "bn254_key_lines_bytes", "bn254_key_lines_prepare_dev", "bn254_verify_batch_cached_dev",
    "bn254_check_public_keys_batch", "bn254_pairing_check_batch", "bn254_pairing_check_batch_dev",
    "bn254_g1_sum", "bn254_g2_sum", "bn254_g1_sum_dev", "bn254_g2_sum_dev",
    "bn254_derive_pk_g2_batch", "bn254_derive_pk_g1_batch", "bn254_g1_mul_batch", "bn254_g2_mul_batch",
    "bn254_g1_compress_batch", "bn254_g1_decompress_batch", "bn254_g2_compress_batch", "bn254_g2_decompress_batch",
    "bn254_g1_validate_batch", "bn254_g2_validate_batch",
    "bn254_aggregate_verify_same_msg", "bn254_aggregate_verify_same_msg_dev", "bn254_aggregate_verify_distinct",
    "bn254_distinct_payload_dev", "bn254_finish_distinct_dev", "bn254_format_pairing_check_batch",
    "bn254_miller_partial_distinct", "bn254_miller_partial_distinct_dev", "bn254_finish_distinct",
    "bn254_miller_loop_batch", "bn254_final_exp_batch", "bn254_fq_op_batch", "bn254_fq12_op_batch", "bn254_layer_op_batch",
]

_lib = None


class EngineError(RuntimeError):
    """Engine-level failure (CUDA error, bad argument) -- not a per-item verdict."""


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError("%s is missing: build it with `python -m bn254_b200.build` (there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.bn254_last_error.restype = ctypes.c_char_p
        lib.bn254_last_error.argtypes = [ctypes.c_void_p]
        lib.bn254_stream.restype = ctypes.c_void_p
        lib.bn254_stream.argtypes = [ctypes.c_void_p]
        lib.bn254_launch_count.restype = ctypes.c_uint64
        lib.bn254_launch_count.argtypes = [ctypes.c_void_p]
        lib.bn254_sm_count.argtypes = [ctypes.c_void_p]
        lib.bn254_ctx_destroy.argtypes = [ctypes.c_void_p]
        lib.bn254_ctx_destroy.restype = None
        lib.bn254_sync.argtypes = [ctypes.c_void_p]
        lib.bn254_get_input_policy.argtypes = [ctypes.c_void_p]
        lib.bn254_key_lines_bytes.restype = ctypes.c_size_t
        lib.bn254_key_lines_bytes.argtypes = [ctypes.c_size_t]
        _lib = lib
    return _lib


def _ptr(x):
    """bytes-like, ctypes buffer, int device pointer or None -> c_void_p argument."""
    if x is None:
        return ctypes.c_void_p(0)
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if isinstance(x, bytes):
        return ctypes.cast(ctypes.c_char_p(x), ctypes.c_void_p)
    if isinstance(x, bytearray):
        return ctypes.cast((ctypes.c_char * len(x)).from_buffer(x), ctypes.c_void_p) if len(x) else ctypes.c_void_p(0)
    if isinstance(x, ctypes.Array):
        return ctypes.cast(x, ctypes.c_void_p)
    if isinstance(x, ctypes._SimpleCData):  # an out-parameter such as c_int
        return ctypes.cast(ctypes.pointer(x), ctypes.c_void_p)
    if hasattr(x, "ctypes"):  # numpy array
        return ctypes.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):  # torch tensor (host or device)
        return ctypes.c_void_p(x.data_ptr())
    raise TypeError("unsupported buffer type %r" % type(x))


class Context:
    """One engine context = one GPU (bn254_ctx)."""

    def __init__(self, device=0):
        self.lib = load()
        h = ctypes.c_void_p()
        rc = self.lib.bn254_ctx_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise EngineError("bn254_ctx_create(%d) failed (%d): %s" % (device, rc, self.lib.bn254_last_error(None).decode()))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.bn254_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name, *args):
        fn = getattr(self.lib, name)
        conv = []
        for a in args:
            if isinstance(a, _Size):
                conv.append(ctypes.c_size_t(a.v))
            elif isinstance(a, _Int):
                conv.append(ctypes.c_int(a.v))
            else:
                conv.append(_ptr(a))
        rc = fn(self.h, *conv)
        if rc != 0:
            raise EngineError("%s failed (%d): %s" % (name, rc, self.lib.bn254_last_error(self.h).decode()))

    def sync(self):
        rc = self.lib.bn254_sync(self.h)
        if rc != 0:
            raise EngineError("bn254_sync failed: %s" % self.lib.bn254_last_error(self.h).decode())

    @property
    def input_policy(self):
        return int(self.lib.bn254_get_input_policy(self.h))

    @property
    def stream(self):
        return self.lib.bn254_stream(self.h)

    @property
    def sm_count(self):
        return self.lib.bn254_sm_count(self.h)

    @property
    def launch_count(self):
        return int(self.lib.bn254_launch_count(self.h))


class _Size:
    def __init__(self, v):
        self.v = int(v)


class _Int:
    def __init__(self, v):
        self.v = int(v)


def S(v):
    return _Size(v)


def I(v):
    return _Int(v)


def out(n):
    return ctypes.create_string_buffer(max(int(n), 1))
