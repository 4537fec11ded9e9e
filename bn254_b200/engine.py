"""Batch entry points over host byte buffers (thin, typed wrappers of the C ABI).

Buffers are `bytes` (or anything exposing the buffer protocol through numpy / torch); results are `bytes`.
Statuses are one byte per item: 0 = Ok, otherwise the Error variant of /root/reference/src/error.rs:6-29.
"""
from ._native import Context, S, I, out

_default = {}
# bn254_set_input_policy (include/bn254_b200.h): how verify / check_public_keys / the aggregate checks read their point bytes
INPUTS_UNTRUSTED, INPUTS_TYPED = 0, 1
DISTINCT_PAYLOAD_BYTES = 448


def context(device=0):
    """Process-wide context per device (created on first use; raises EngineError without a GPU)."""
    c = _default.get(device)
    if c is None:
        c = _default[device] = Context(device)
    return c


def set_input_policy(policy, ctx=None):
    """INPUTS_UNTRUSTED (default): signatures / keys are decoded like from_uncompressed (no infinity, G2 r-torsion test);
    INPUTS_TYPED: values of the crate's types (all-zero bytes = infinity, membership trusted)."""
    (ctx or context()).call("bn254_set_input_policy", I(policy))


def set_hash_try_limit(max_tries, ctx=None):
    """Test hook: counters tried by hash_to_try_and_increment (255 in the reference)."""
    (ctx or context()).call("bn254_set_hash_try_limit", I(max_tries))


def _n(buf, size):
    assert len(buf) % size == 0, "buffer length %d is not a multiple of %d" % (len(buf), size)
    return len(buf) // size


def hash_to_g1_batch(msgs, msg_len, n, ctx=None):
    ctx = ctx or context()
    o, st = out(64 * n), out(n)
    ctx.call("bn254_hash_to_g1_batch", msgs, S(msg_len), S(n), o, st)
    return o.raw[:64 * n], st.raw[:n]


def hash_to_g1_var(msgs_list, ctx=None):
    import struct
    ctx = ctx or context()
    n = len(msgs_list)
    offs = [0]
    for m in msgs_list:
        offs.append(offs[-1] + len(m))
    blob = b"".join(msgs_list)
    o, st, tr = out(64 * n), out(n), out(n)
    ctx.call("bn254_hash_to_g1_var", blob if blob else None, struct.pack("<%dQ" % (n + 1), *offs), S(n), o, st, tr)
    return o.raw[:64 * n], st.raw[:n], tr.raw[:n]


def _msgs_ok(msgs, msg_len, n):
    assert (len(msgs) if msgs is not None else 0) == msg_len * n, "messages: %d bytes, expected %d x %d" % (
        len(msgs) if msgs is not None else 0, n, msg_len)


def sign_batch(msgs, msg_len, sks, ctx=None):
    ctx = ctx or context()
    n = _n(sks, 32)
    _msgs_ok(msgs, msg_len, n)
    o, st = out(64 * n), out(n)
    ctx.call("bn254_sign_batch", msgs, S(msg_len), sks, S(n), o, st)
    return o.raw[:64 * n], st.raw[:n]


def verify_batch_rlc(msgs, msg_len, sigs, pks, coeffs16=None, pks_in_g2=False, ctx=None):
    """Randomised batch form of verify_batch (bn254_verify_batch_rlc): -> (statuses, took_fast_path).  The statuses are
    verify_batch's; `coeffs16` (16 secret random bytes per item) defaults to the engine drawing them from /dev/urandom."""
    import ctypes
    ctx = ctx or context()
    n = _n(sigs, 64)
    assert _n(pks, 128) == n and (coeffs16 is None or len(coeffs16) == 16 * n)
    _msgs_ok(msgs, msg_len, n)
    st, fast = out(n), ctypes.c_int(0)
    ctx.call("bn254_verify_batch_rlc", msgs, S(msg_len), sigs, pks, S(n), coeffs16, I(1 if pks_in_g2 else 0), st, fast)
    return st.raw[:n], bool(fast.value)


def verify_batch(msgs, msg_len, sigs, pks, ctx=None):
    ctx = ctx or context()
    n = _n(sigs, 64)
    assert _n(pks, 128) == n
    _msgs_ok(msgs, msg_len, n)
    st = out(n)
    ctx.call("bn254_verify_batch", msgs, S(msg_len), sigs, pks, S(n), st)
    return st.raw[:n]


def check_public_keys_batch(pk_g2, pk_g1, ctx=None):
    ctx = ctx or context()
    n = _n(pk_g1, 64)
    st = out(n)
    ctx.call("bn254_check_public_keys_batch", pk_g2, pk_g1, S(n), st)
    return st.raw[:n]


def pairing_check_batch(g1s, g2s, k, n, ctx=None):
    ctx = ctx or context()
    st = out(n)
    ctx.call("bn254_pairing_check_batch", g1s if k else None, g2s if k else None, S(k), S(n), st)
    return st.raw[:n]


def miller_loop_batch(g1s, g2s, k, n, ctx=None):
    ctx = ctx or context()
    o, st = out(384 * n), out(n)
    ctx.call("bn254_miller_loop_batch", g1s if k else None, g2s if k else None, S(k), S(n), o, st)
    return o.raw[:384 * n], st.raw[:n]


def final_exp_batch(f, ctx=None):
    ctx = ctx or context()
    n = _n(f, 384)
    o, st = out(384 * n), out(n)
    ctx.call("bn254_final_exp_batch", f, S(n), o, st)
    return o.raw[:384 * n], st.raw[:n]


def fq_op_batch(op, a, b=None, ctx=None):
    ctx = ctx or context()
    n = _n(a, 32)
    o, st = out(32 * n), out(n)
    ctx.call("bn254_fq_op_batch", I(op), a, b, S(n), o, st)
    return o.raw[:32 * n], st.raw[:n]


def fq12_op_batch(op, a, b=None, ctx=None):
    ctx = ctx or context()
    n = _n(a, 384)
    o, st = out(384 * n), out(n)
    ctx.call("bn254_fq12_op_batch", I(op), a, b, S(n), o, st)
    return o.raw[:384 * n], st.raw[:n]


def g1_sum(pts, neg=None, ctx=None):
    ctx = ctx or context()
    n = _n(pts, 64)
    o, st = out(64), out(1)
    ctx.call("bn254_g1_sum", pts if n else None, neg, S(n), o, st)
    return o.raw[:64], st.raw[0]


def g2_sum(pts, neg=None, ctx=None):
    ctx = ctx or context()
    n = _n(pts, 128)
    o, st = out(128), out(1)
    ctx.call("bn254_g2_sum", pts if n else None, neg, S(n), o, st)
    return o.raw[:128], st.raw[0]


def derive_pk_g2_batch(sks, ctx=None):
    ctx = ctx or context()
    n = _n(sks, 32)
    o = out(128 * n)
    ctx.call("bn254_derive_pk_g2_batch", sks, S(n), o)
    return o.raw[:128 * n]


def derive_pk_g1_batch(sks, ctx=None):
    ctx = ctx or context()
    n = _n(sks, 32)
    o = out(64 * n)
    ctx.call("bn254_derive_pk_g1_batch", sks, S(n), o)
    return o.raw[:64 * n]


def g1_mul_batch(pts, scalars, ctx=None):
    ctx = ctx or context()
    n = _n(pts, 64)
    o, st = out(64 * n), out(n)
    ctx.call("bn254_g1_mul_batch", pts, scalars, S(n), o, st)
    return o.raw[:64 * n], st.raw[:n]


def g2_mul_batch(pts, scalars, ctx=None):
    ctx = ctx or context()
    n = _n(pts, 128)
    o, st = out(128 * n), out(n)
    ctx.call("bn254_g2_mul_batch", pts, scalars, S(n), o, st)
    return o.raw[:128 * n], st.raw[:n]


def _codec(name, data, in_size, out_size, ctx):
    ctx = ctx or context()
    n = _n(data, in_size)
    o, st = out(out_size * n), out(n)
    ctx.call(name, data, S(n), o, st)
    return o.raw[:out_size * n], st.raw[:n]


def g1_compress_batch(raw, ctx=None):
    return _codec("bn254_g1_compress_batch", raw, 64, 33, ctx)


def g1_decompress_batch(comp, ctx=None):
    return _codec("bn254_g1_decompress_batch", comp, 33, 64, ctx)


def g2_compress_batch(raw, ctx=None):
    return _codec("bn254_g2_compress_batch", raw, 128, 65, ctx)


def g2_decompress_batch(comp, ctx=None):
    return _codec("bn254_g2_decompress_batch", comp, 65, 128, ctx)


def g1_validate_batch(raw, ctx=None):
    ctx = ctx or context()
    n = _n(raw, 64)
    st = out(n)
    ctx.call("bn254_g1_validate_batch", raw, S(n), st)
    return st.raw[:n]


def g2_validate_batch(raw, ctx=None):
    ctx = ctx or context()
    n = _n(raw, 128)
    st = out(n)
    ctx.call("bn254_g2_validate_batch", raw, S(n), st)
    return st.raw[:n]


def aggregate_verify_same_msg(msg, sigs, pks, ctx=None):
    ctx = ctx or context()
    n = _n(sigs, 64)
    st = out(1)
    ctx.call("bn254_aggregate_verify_same_msg", msg if len(msg) else None, S(len(msg)), sigs if n else None, pks if n else None, S(n), st)
    return st.raw[0]


def aggregate_verify_distinct(msgs, msg_len, pks, agg_sig, ctx=None):
    ctx = ctx or context()
    n = _n(pks, 128)
    st = out(1)
    ctx.call("bn254_aggregate_verify_distinct", msgs if len(msgs) else None, S(msg_len), pks if n else None, S(n), agg_sig, st)
    return st.raw[0]


def miller_partial_distinct(msgs, msg_len, pks, ctx=None):
    ctx = ctx or context()
    n = _n(pks, 128)
    o, st = out(384), out(1)
    ctx.call("bn254_miller_partial_distinct", msgs if len(msgs) else None, S(msg_len), pks if n else None, S(n), o, st)
    return o.raw[:384], st.raw[0]


def finish_distinct(partials, agg_sig, ctx=None):
    ctx = ctx or context()
    n = _n(partials, 384)
    st = out(1)
    ctx.call("bn254_finish_distinct", partials if n else None, S(n), agg_sig, st)
    return st.raw[0]


def format_pairing_check_batch(msgs, msg_len, sigs, pks, compressed, ctx=None):
    """format_pairing_check_values (compressed=True: 33-byte sigs, 65-byte pks) / _uncompressed_values (64 / 128 bytes):
    -> (n x 384 bytes [(H(m), pk), (sig, -G2)] with little-endian coordinates, statuses)"""
    ctx = ctx or context()
    n = _n(sigs, 33 if compressed else 64)
    assert _n(pks, 65 if compressed else 128) == n
    _msgs_ok(msgs, msg_len, n)
    o, st = out(384 * n), out(n)
    ctx.call("bn254_format_pairing_check_batch", msgs, S(msg_len), sigs, pks, S(n), I(1 if compressed else 0), o, st)
    return o.raw[:384 * n], st.raw[:n]


class KeyLineCache:
    """Cached line coefficients of a fixed key set (bn254_key_lines_prepare_dev): `verify(msgs, msg_len, sigs, key_index)` gives
    verify_batch's statuses for triples (msg_i, sig_i, keys[key_index[i]]) without walking the keys again.  Device buffers are
    torch CUDA tensors owned by this object (16 704 bytes per key)."""

    def __init__(self, pks, ctx=None):
        import torch
        self.ctx = ctx or context()
        self.n_keys = _n(pks, 128)
        dev = torch.device("cuda", self.ctx.device)
        self._torch = torch
        nbytes = int(self.ctx.lib.bn254_key_lines_bytes(self.n_keys))
        self.lines = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.status = torch.zeros(max(self.n_keys, 1), dtype=torch.uint8, device=dev)
        d_pks = torch.frombuffer(bytearray(pks), dtype=torch.uint8).to(dev)
        self.ctx.call("bn254_key_lines_prepare_dev", d_pks, S(self.n_keys), self.lines, self.status)
        self.ctx.sync()

    def key_status(self):
        return bytes(self.status[:self.n_keys].cpu().numpy().tobytes())

    def verify_dev(self, d_msgs, msg_len, d_sigs, n, d_status, d_key_index=None):
        self.ctx.call("bn254_verify_batch_cached_dev", d_msgs, S(msg_len), d_sigs, self.lines, self.status, S(self.n_keys), d_key_index, S(n), d_status)

    def verify(self, msgs, msg_len, sigs, key_index=None):
        torch = self._torch
        dev = self.lines.device
        n = _n(sigs, 64)
        _msgs_ok(msgs, msg_len, n)
        up = lambda b: torch.frombuffer(bytearray(b if b else b"\0"), dtype=torch.uint8).to(dev)
        d_idx = None
        if key_index is not None:
            import numpy as np
            d_idx = torch.from_numpy(np.asarray(key_index, dtype=np.uint32).view(np.int32).copy()).to(dev)
        st = torch.zeros(max(n, 1), dtype=torch.uint8, device=dev)
        self.verify_dev(up(msgs), msg_len, up(sigs), n, st, d_idx)
        self.ctx.sync()
        return bytes(st[:n].cpu().numpy().tobytes())


def layer_op_batch(op, data, n_in, n_out, ctx=None):
    ctx = ctx or context()
    n = _n(data, 32 * n_in)
    o = out(32 * n_out * n)
    ctx.call("bn254_layer_op_batch", I(op), data, S(n_in), S(n), o, S(n_out))
    return o.raw[:32 * n_out * n]
