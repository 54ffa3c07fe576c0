"""Host-side mirror of `ethereum_consensus::crypto` (BLS), backed by the CUDA library — no CPU fallback.

Same names, argument meaning and error behaviour as /root/reference/ethereum-consensus/src/crypto/bls.rs:
`verify_signature` (:64-77), `aggregate` (:79-93), `aggregate_verify` (:95-112), `fast_aggregate_verify` (:114-132),
`eth_aggregate_public_keys` (:135-148), `eth_fast_aggregate_verify` (:150-160), `hash` (:12-20), the byte newtypes
`PublicKey` (:227-239) / `Signature` (:287-290) with their length checks (:257-266, :318-327) and
`Signature.is_infinity` (:343-347); errors mirror `Error` / `BLSTError` (:27-62).
`SecretKey` (key generation, signing) is not on the verification hot path and is not re-implemented here.

Rust `Result<(), Error>` becomes: return None on Ok, raise on Err.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from .ssz import hash  # noqa: F401,A004  (crypto::hash is the same one-shot SHA-256)

BLS_DST = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_"
BLS_PUBLIC_KEY_BYTES_LEN = 48
BLS_SIGNATURE_BYTES_LEN = 96
INFINITY_COMPRESSED_SIGNATURE = bytes([0xC0]) + bytes(95)

_BLST_TEXT = {1: "bad encoding", 2: "point not on curve", 3: "point not in group", 4: "aggregation type mismatch",
              5: "verification failed", 6: "public key is infinity", 7: "bad scalar input"}

_lib.register_protos({
    "b200_verify_signature": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200_fast_aggregate_verify": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200_eth_fast_aggregate_verify": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200_aggregate_verify": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200_aggregate": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200_eth_aggregate_public_keys": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200_fast_aggregate_verify_batch": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200_registry_load": (C.c_int32, [C.c_void_p, C.c_size_t]),
    "b200_registry_key_codes": (C.c_int32, [C.c_void_p, C.c_size_t]),
    "b200_fast_aggregate_verify_batch_indexed": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200_fast_aggregate_verify_batch_all": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int32)]),
    "b200_fast_aggregate_verify_batch_indexed_all": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int32)]),
    "b200_fast_aggregate_verify_batch_all_sharded": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int32)]),
    "b200_last_dominant_kernel_ms": (C.c_float, []),
    "b200_fp_selftest": (C.c_int32, [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]),
    "b200_measure_int_peak": (C.c_int32, [C.c_int32, C.POINTER(C.c_double)]),
})


class Error(Exception):
    """`crypto::Error` (crypto/bls.rs:27-42)."""


class EmptyAggregate(Error):
    def __init__(self):
        super().__init__("inputs required for aggregation but none were provided")


class SimpleSerializeError(Error):
    """Wrong byte length for a `ByteVector<N>` newtype (crypto/bls.rs:257-266, 318-327)."""


class BLSTError(Error):
    """`Error::BLST(BLSTError)` — text from crypto/bls.rs:48-62."""

    def __init__(self, code: int):
        self.code = code
        super().__init__(f"blst error: {_BLST_TEXT.get(code, code)}")


class InvalidSignature(Error):
    def __init__(self):
        super().__init__("invalid signature")


class PublicKey(bytes):
    """`PublicKey(ByteVector<48>)`: any 48 bytes are accepted here; curve checks happen at use (bls.rs:279-285)."""

    def __new__(cls, data=bytes(48)):
        b = bytes(data)
        if len(b) != BLS_PUBLIC_KEY_BYTES_LEN:
            raise SimpleSerializeError(f"expected {BLS_PUBLIC_KEY_BYTES_LEN} bytes, got {len(b)}")
        return super().__new__(cls, b)


class Signature(bytes):
    """`Signature(ByteVector<96>)`."""

    def __new__(cls, data=bytes(96)):
        b = bytes(data)
        if len(b) != BLS_SIGNATURE_BYTES_LEN:
            raise SimpleSerializeError(f"expected {BLS_SIGNATURE_BYTES_LEN} bytes, got {len(b)}")
        return super().__new__(cls, b)

    def is_infinity(self) -> bool:
        return bytes(self) == INFINITY_COMPRESSED_SIGNATURE


def _result(code: int, where: str) -> None:
    _lib.check(code, where)
    if code == _lib.SUCCESS:
        return None
    if code == _lib.VERIFY_FAIL:
        raise InvalidSignature()
    if code == _lib.EMPTY_AGGREGATE:
        raise EmptyAggregate()
    raise BLSTError(code)


def _ptr_array(items: Sequence[bytes]):
    keep = [bytes(x) for x in items]
    arr = (C.c_char_p * max(len(keep), 1))(*keep) if keep else (C.c_char_p * 1)()
    return arr, keep


def verify_signature(public_key: PublicKey, msg: bytes, signature: Signature) -> None:
    pk, sig, m = PublicKey(public_key), Signature(signature), bytes(msg)
    _result(_lib.lib().b200_verify_signature(_lib.ptr(pk), _lib.ptr(m), len(m), _lib.ptr(sig)), "verify_signature")


def fast_aggregate_verify(public_keys: Sequence[PublicKey], msg: bytes, signature: Signature) -> None:
    pks = [PublicKey(p) for p in public_keys]
    sig, m = Signature(signature), bytes(msg)
    arr, _keep = _ptr_array(pks)
    _result(_lib.lib().b200_fast_aggregate_verify(C.cast(arr, C.c_void_p), len(pks), _lib.ptr(m), len(m), _lib.ptr(sig)),
            "fast_aggregate_verify")


def eth_fast_aggregate_verify(public_keys: Sequence[PublicKey], message: bytes, signature: Signature) -> None:
    pks = [PublicKey(p) for p in public_keys]
    sig, m = Signature(signature), bytes(message)
    arr, _keep = _ptr_array(pks)
    _result(_lib.lib().b200_eth_fast_aggregate_verify(C.cast(arr, C.c_void_p), len(pks), _lib.ptr(m), len(m), _lib.ptr(sig)),
            "eth_fast_aggregate_verify")


def aggregate_verify(public_keys: Sequence[PublicKey], msgs: Sequence[bytes], signature: Signature) -> None:
    pks = b"".join(PublicKey(p) for p in public_keys)
    sig = Signature(signature)
    arr, keep = _ptr_array(msgs)
    lens = (C.c_size_t * max(len(keep), 1))(*[len(m) for m in keep])
    _result(_lib.lib().b200_aggregate_verify(_lib.ptr(pks), len(public_keys), C.cast(arr, C.c_void_p),
                                             C.cast(lens, C.c_void_p), len(keep), _lib.ptr(sig)), "aggregate_verify")


def aggregate(signatures: Sequence[Signature]) -> Signature:
    if len(signatures) == 0:
        raise EmptyAggregate()
    flat = b"".join(Signature(s) for s in signatures)
    out = (C.c_uint8 * 96)()
    _result(_lib.lib().b200_aggregate(_lib.ptr(flat), len(signatures), out), "aggregate")
    return Signature(bytes(out))


def eth_aggregate_public_keys(public_keys: Sequence[PublicKey]) -> PublicKey:
    if len(public_keys) == 0:
        raise EmptyAggregate()
    flat = b"".join(PublicKey(p) for p in public_keys)
    out = (C.c_uint8 * 48)()
    _result(_lib.lib().b200_eth_aggregate_public_keys(_lib.ptr(flat), len(public_keys), out), "eth_aggregate_public_keys")
    return PublicKey(bytes(out))


# ---- the throughput path ----------------------------------------------------------------------------------------
def _nbytes(buf) -> int:
    if hasattr(buf, "nbytes"):
        return int(buf.nbytes)
    if hasattr(buf, "numel"):
        return int(buf.numel() * buf.element_size())
    return len(buf)


def fast_aggregate_verify_batch(pks_flat, pk_offsets, msgs32, sigs) -> np.ndarray:
    """T tuples at once -> int32 code per tuple (0 Ok, 5 InvalidSignature, 1/2/3/6 BLST decode errors).
    pks_flat: (sum K) x 48 bytes; pk_offsets: uint32[T+1]; msgs32: T x 32; sigs: T x 96 (host buffers)."""
    off = np.ascontiguousarray(pk_offsets, dtype=np.uint32)
    t = len(off) - 1
    if t < 0 or (t >= 0 and int(off[0]) != 0):
        raise ValueError("pk_offsets must hold T+1 entries starting at 0")
    # the C side reads raw pointers: refuse buffers whose sizes disagree with the offsets
    if _nbytes(pks_flat) != 48 * int(off[-1]) or _nbytes(msgs32) != 32 * t or _nbytes(sigs) != 96 * t:
        raise ValueError(f"buffer sizes do not match the offsets: keys {_nbytes(pks_flat)} B for {int(off[-1])} keys, "
                         f"msgs {_nbytes(msgs32)} B, sigs {_nbytes(sigs)} B for {t} tuples")
    out = np.empty(max(t, 1), dtype=np.int32)
    _lib.check(_lib.lib().b200_fast_aggregate_verify_batch(_lib.ptr(pks_flat), _lib.ptr(off), _lib.ptr(msgs32), _lib.ptr(sigs),
                                                           t, _lib.ptr(out)), "fast_aggregate_verify_batch")
    return out[:t]


def fast_aggregate_verify_batch_all(pks_flat, pk_offsets, msgs32, sigs, seed: bytes = None, sharded: bool = False) -> bool:
    """Optimistic whole-batch check by random linear combination (T Miller loops + ONE final exponentiation): True iff
    every tuple verifies (false accept probability <= 2^-64 over `seed`; None = drawn by the library).  On False ask
    `fast_aggregate_verify_batch` for the per-tuple codes.  `sharded`: all ranks of the library's communicator share the
    batch (same arguments and seed on every rank); the Gt / G2 partials travel in one ncclAllGather."""
    off = np.ascontiguousarray(pk_offsets, dtype=np.uint32)
    t = len(off) - 1
    if t < 0 or _nbytes(pks_flat) != 48 * int(off[-1]) or _nbytes(msgs32) != 32 * t or _nbytes(sigs) != 96 * t:
        raise ValueError("buffer sizes do not match the offsets")
    if seed is not None and len(seed) != 32:
        raise ValueError("seed must be 32 bytes")
    if sharded and seed is None:
        raise ValueError("the ranks must share the seed")
    sd = np.frombuffer(bytes(seed), dtype=np.uint8) if seed is not None else None
    ok = C.c_int32(0)
    fn = _lib.lib().b200_fast_aggregate_verify_batch_all_sharded if sharded else _lib.lib().b200_fast_aggregate_verify_batch_all
    _lib.check(fn(_lib.ptr(pks_flat), _lib.ptr(off), _lib.ptr(msgs32), _lib.ptr(sigs), t, _lib.ptr(sd) if sd is not None else 0, C.byref(ok)),
               "fast_aggregate_verify_batch_all")
    return bool(ok.value)


class Registry:
    """Validated validator public keys resident in HBM (`state.validators[i].public_key` is immutable,
    phase0/validator.rs:10-13): `load` runs key_validate once per key, `verify_batch` names signers by index."""

    def __init__(self, pks_flat):
        n = (pks_flat.nbytes if hasattr(pks_flat, "nbytes") else len(pks_flat)) // 48
        _lib.check(_lib.lib().b200_registry_load(_lib.ptr(pks_flat), n), "registry_load")
        self.n = n

    def key_codes(self) -> np.ndarray:
        out = np.empty(max(self.n, 1), dtype=np.int32)
        _lib.check(_lib.lib().b200_registry_key_codes(_lib.ptr(out), self.n), "registry_key_codes")
        return out[:self.n]

    def verify_batch(self, indices, offsets, msgs32, sigs, extra_keys=None) -> np.ndarray:
        """`extra_keys` (flat 48-byte keys): keys that arrive with the block (deposits, bls-to-execution changes); index
        self.n + j names extra key j, which this call validates like the strict path would (`..._batch_mixed`)."""
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        off = np.ascontiguousarray(offsets, dtype=np.uint32)
        t = len(off) - 1
        if t < 0 or len(idx) != int(off[-1]) or _nbytes(msgs32) != 32 * t or _nbytes(sigs) != 96 * t:
            raise ValueError("indices / offsets / msgs / sigs sizes are inconsistent")
        out = np.empty(max(t, 1), dtype=np.int32)
        if extra_keys is not None and _nbytes(extra_keys):
            if _nbytes(extra_keys) % 48:
                raise ValueError("extra keys must be 48 bytes each")
            _lib.check(_lib.lib().b200_fast_aggregate_verify_batch_mixed(_lib.ptr(extra_keys), _nbytes(extra_keys) // 48, _lib.ptr(idx), _lib.ptr(off),
                                                                         _lib.ptr(msgs32), _lib.ptr(sigs), t, _lib.ptr(out)), "verify_batch_mixed")
        else:
            _lib.check(_lib.lib().b200_fast_aggregate_verify_batch_indexed(_lib.ptr(idx), _lib.ptr(off), _lib.ptr(msgs32),
                                                                           _lib.ptr(sigs), t, _lib.ptr(out)), "verify_batch_indexed")
        return out[:t]


def _registry_verify_batch_all(self, indices, offsets, msgs32, sigs, seed: bytes = None) -> bool:
    idx = np.ascontiguousarray(indices, dtype=np.uint32)
    off = np.ascontiguousarray(offsets, dtype=np.uint32)
    t = len(off) - 1
    if t < 0 or len(idx) != int(off[-1]) or _nbytes(msgs32) != 32 * t or _nbytes(sigs) != 96 * t:
        raise ValueError("indices / offsets / msgs / sigs sizes are inconsistent")
    sd = np.frombuffer(bytes(seed), dtype=np.uint8) if seed is not None else None
    ok = C.c_int32(0)
    _lib.check(_lib.lib().b200_fast_aggregate_verify_batch_indexed_all(_lib.ptr(idx), _lib.ptr(off), _lib.ptr(msgs32), _lib.ptr(sigs), t,
                                                                       _lib.ptr(sd) if sd is not None else 0, C.byref(ok)), "verify_batch_indexed_all")
    return bool(ok.value)


Registry.verify_batch_all = _registry_verify_batch_all


def last_kernel_ms() -> float:
    return float(_lib.lib().b200_last_kernel_ms())


def last_dominant_kernel_ms() -> float:
    return float(_lib.lib().b200_last_dominant_kernel_ms())


def fp_selftest(n: int = 1 << 16, seed: int = 1) -> int:
    m = C.c_uint32(0)
    _lib.check(_lib.lib().b200_fp_selftest(n, seed, C.byref(m)), "fp_selftest")
    return int(m.value)


def tune(knob: str, value: int) -> None:
    """Launch-shape knobs of the batch pipeline (include/b200_consensus.h, b200_tune): never change a result."""
    _lib.check(_lib.lib().b200_tune(knob.encode(), int(value)), f"tune({knob})")


def vm_load_programs(blob) -> None:
    """Swap in another schedule of the pairing programs (uint32 words written by tools/gen_pairing_vm.py --blob)."""
    a = np.ascontiguousarray(blob, dtype=np.uint32)
    _lib.check(_lib.lib().b200_vm_load_programs(_lib.ptr(a), a.size), "vm_load_programs")


def measure_int_peak(kind: int) -> float:
    """1e9 ops/s of IMAD.WIDE.U32 (0), IMAD.U32 (1) or the LOP3/SHF/IADD3 mix (2) measured on this device."""
    g = C.c_double(0)
    _lib.check(_lib.lib().b200_measure_int_peak(kind, C.byref(g)), "measure_int_peak")
    return float(g.value)
