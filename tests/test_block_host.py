"""Host-side logic of the signature-set collector (ethereum_consensus_b200/block.py) — no GPU: the checks
`is_valid_indexed_attestation` makes before the BLS call (phase0/helpers.rs:94-131), the first-failure / deposit-tolerance replay, and the packed-at-collection-time validator indices that registry mode hands to
`b200_fast_aggregate_verify_batch_mixed`."""
import hashlib

import numpy as np
import pytest

from ethereum_consensus_b200 import block

KEYS = [bytes([0x80 | (i & 0x1f)]) + hashlib.sha256(b"k%d" % i).digest() + bytes(15) for i in range(16)]
ROOT = hashlib.sha256(b"root").digest()
SIG = bytes([0xC0]) + bytes(95)


def test_indexed_attestation_host_checks():
    s = block.SignatureSet()
    with pytest.raises(block.InvalidIndexedAttestation, match="Empty"):
        s.add_indexed_attestation("attestation", KEYS, [], ROOT, SIG)
    with pytest.raises(block.InvalidIndexedAttestation, match="NotSorted"):
        s.add_indexed_attestation("attestation", KEYS, [3, 2], ROOT, SIG)
    with pytest.raises(block.InvalidIndexedAttestation, match="Duplicate"):
        s.add_indexed_attestation("attestation", KEYS, [1, 1, 2], ROOT, SIG)
    with pytest.raises(block.InvalidIndexedAttestation, match="InvalidIndex"):
        s.add_indexed_attestation("attestation", KEYS, [1, 16], ROOT, SIG)
    assert not s.entries
    s.add_indexed_attestation("attestation", KEYS, [1, 5, 9], ROOT, SIG)
    e = s.entries[0]
    assert e.pubkeys == [KEYS[1], KEYS[5], KEYS[9]] and e.indices == (1, 5, 9)


def test_indices_are_packed_once_and_cannot_go_stale():
    s = block.SignatureSet()
    s.add_by_index("voluntary_exit", KEYS, [7], ROOT, SIG)
    e = s.entries[0]
    assert isinstance(e.indices, tuple) and e._indices_np.dtype == np.uint32 and e._indices_np.tolist() == [7]
    with pytest.raises((TypeError, AttributeError)):
        e.indices[0] = 3                       # a tuple: no in-place edit behind the packed copy
    with pytest.raises(ValueError):
        e._indices_np[0] = 3                   # read-only array
    e.indices = [2, 4]                         # assignment repacks
    assert e.indices == (2, 4) and e._indices_np.tolist() == [2, 4]
    e.indices = None
    assert e.indices is None and e._indices_np is None


def test_byte_newtypes_guard_the_joined_buffers():
    s = block.SignatureSet()
    with pytest.raises(Exception):
        s.add("deposit", [KEYS[0][:47]], ROOT, SIG)
    with pytest.raises(Exception):
        s.add("deposit", [KEYS[0]], ROOT, SIG[:95])
    with pytest.raises(ValueError):
        s.add("deposit", [KEYS[0]], ROOT[:31], SIG)
    with pytest.raises(ValueError):
        s.add("no_such_site", [KEYS[0]], ROOT, SIG)
    assert not s.entries and s.verify().tolist() == []


def test_first_failure_and_deposit_tolerance_replay():
    """The reference's control flow over a code vector (phase0/block_processing.rs:97-99, 387-392): tolerant entries (deposits)
    never abort, the first failing abort-on-failure entry is the block's error.  (The sync-aggregate gather and `get_domain` hash
    on the device: tests/test_config_scale_gpu.py::test_get_domain_and_sync_aggregate_signing_root.)"""
    s2 = block.SignatureSet()
    s2.add("deposit", [KEYS[0]], ROOT, SIG, tolerant=True)
    s2.add("voluntary_exit", [KEYS[1]], ROOT, SIG)
    s2.add("voluntary_exit", [KEYS[2]], ROOT, SIG)
    assert s2.first_failure(np.array([5, 0, 0])) is None and s2.skipped_deposits(np.array([5, 0, 0])) == [0]
    assert s2.first_failure(np.array([5, 0, 6])) == (2, "voluntary_exit", 6)
    assert s2.first_failure(np.array([0, 4, 6])) == (1, "voluntary_exit", 4)
