"""Mirror of the reference's signing helpers, hashing on the device.

`compute_signing_root` / `verify_signed_data`  — /root/reference/ethereum-consensus/src/signing.rs:14-41
`compute_domain` / `compute_fork_data_root`    — /root/reference/ethereum-consensus/src/phase0/helpers.rs:506-529
`DomainType`                                   — /root/reference/ethereum-consensus/src/domains.rs:1-30
"""
from __future__ import annotations

import enum

from . import crypto, ssz


class DomainType(enum.IntEnum):
    BeaconProposer = 0
    BeaconAttester = 1
    Randao = 2
    Deposit = 3
    VoluntaryExit = 4
    SelectionProof = 5
    AggregateAndProof = 6
    SyncCommittee = 7
    SyncCommitteeSelectionProof = 8
    ContributionAndProof = 9
    BlsToExecutionChange = 10

    def as_bytes(self) -> bytes:
        return int(self).to_bytes(4, "little")


def compute_fork_data_root(current_version: bytes, genesis_validators_root: bytes) -> bytes:
    """hash_tree_root(ForkData{current_version, genesis_validators_root}) — two leaves."""
    return ssz.merkleize(bytes(current_version).ljust(32, b"\x00") + bytes(genesis_validators_root), 2)


def compute_domain(domain_type: DomainType, fork_version: bytes = b"\x00" * 4, genesis_validators_root: bytes = b"\x00" * 32) -> bytes:
    return DomainType(domain_type).as_bytes() + compute_fork_data_root(fork_version, genesis_validators_root)[:28]


def compute_signing_root(object_root: bytes, domain: bytes) -> bytes:
    """hash_tree_root(SigningData{object_root, domain}); `object_root` is the signed object's hash_tree_root."""
    return ssz.merkleize(bytes(object_root) + bytes(domain), 2)


def verify_signed_data(object_root: bytes, signature: bytes, public_key: bytes, domain: bytes) -> None:
    crypto.verify_signature(public_key, compute_signing_root(object_root, domain), signature)
