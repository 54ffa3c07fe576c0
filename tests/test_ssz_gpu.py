"""GPU parity: CUDA SSZ path (through the C ABI) vs the oracles, bit-exact."""
import ctypes
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import ssz_oracle as so
from ethereum_consensus_b200 import ssz, state as S
from tests.test_oracle_ssz import B3_BODY_ROOT, B3_BRANCH, B3_COMMITMENT, B3_DEPTH, B3_INDEX

pytestmark = pytest.mark.gpu
GOLDEN = json.loads((Path(__file__).parent / "golden" / "ssz_roots.json").read_text())


def test_sha256_device(engine):
    rng = np.random.default_rng(3)
    for n in [0, 1, 3, 55, 56, 63, 64, 65, 119, 120, 121, 128, 1000]:
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert ssz.hash(d) == so.sha256(d), n


@pytest.mark.parametrize("n,limit", [(0, None), (0, 16), (1, None), (1, 1), (2, None), (3, None), (5, 8), (5, 1 << 40),
                                     (63, None), (64, None), (65, None), (513, 1 << 24), (1000, None), (4097, 1 << 13),
                                     (100_000, None), (1 << 17, 1 << 40)])
def test_merkleize(engine, n, limit):
    d = np.random.default_rng(n).integers(0, 256, 32 * n, dtype=np.uint8)
    want = so.merkleize_bytes(d.tobytes(), limit) if n else so.merkleize_chunks([], limit)
    assert ssz.merkleize(d, limit) == want


def test_merkleize_limit_error(engine):
    with pytest.raises(ssz.MerkleizationError):
        ssz.merkleize(bytes(96), 2)


def test_mix_in_length_and_branch_kat_b3(engine):
    r = bytes(range(32))
    assert ssz.mix_in_length(r, 12345) == so.mix_in_length(r, 12345)
    # reference KAT B-3 (deneb/blob_sidecar.rs:70-132) through the device
    leaf = ssz.merkleize(B3_COMMITMENT + bytes(16), 2)
    assert leaf == so.Bytes48.htr(B3_COMMITMENT)
    assert ssz.is_valid_merkle_branch(leaf, B3_BRANCH, B3_DEPTH, B3_INDEX, B3_BODY_ROOT)
    assert not ssz.is_valid_merkle_branch(leaf, B3_BRANCH, B3_DEPTH, B3_INDEX ^ 1, B3_BODY_ROOT)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 5, 255, 256, 257, 1000, 70_001])
def test_validators_root(engine, oracle_ssz_c, n):
    st = S.synth_state(n, "mainnet")
    vb = st.validators.tobytes()
    out = ctypes.create_string_buffer(32)
    assert oracle_ssz_c.orc_htr_validators(vb, n, 1 << 40, 4, out) == 0
    assert ssz.hash_tree_root_validators(vb, n) == out.raw
    if n <= 1000:
        T = so.List(so.Validator, 1 << 40)
        assert out.raw == T.htr(S.to_oracle_value(st)["validators"])


@pytest.mark.parametrize("nbytes,limit,is_list", [(0, 1 << 38, True), (8, 1 << 38, True), (40, 1 << 38, True), (8 * 1000, 1 << 38, True),
                                                  (1000, 1 << 35, True), (33, 1 << 35, True), (8 * 8192, 2048, False), (8 * 300_001, 1 << 38, True)])
def test_packed(engine, oracle_ssz_c, nbytes, limit, is_list):
    d = np.random.default_rng(nbytes).integers(0, 256, nbytes, dtype=np.uint8)
    out = ctypes.create_string_buffer(32)
    assert oracle_ssz_c.orc_htr_packed(d.ctypes.data, nbytes, limit, int(is_list), nbytes // 8, 2, out) == 0
    assert ssz.hash_tree_root_packed(d, limit, is_list, nbytes // 8) == out.raw


@pytest.mark.parametrize("key", sorted(k for k in GOLDEN if k.endswith("hs3:hr2")))
def test_beacon_state_small_golden(engine, key):
    preset, n = key.split(":")[0], int(key.split(":")[1])
    st = S.synth_state(n, preset, n_historical_summaries=3, n_historical_roots=2)
    ssz_bytes = S.serialize(st)
    assert ssz.hash_tree_root_beacon_state(ssz_bytes, preset).hex() == GOLDEN[key]


def test_beacon_state_malformed(engine):
    ssz_bytes = S.serialize(S.synth_state(5, "minimal"))
    with pytest.raises(ssz.MerkleizationError):
        ssz.hash_tree_root_beacon_state(ssz_bytes[:-1], "minimal")
    with pytest.raises(ssz.MerkleizationError):
        ssz.hash_tree_root_beacon_state(ssz_bytes[:100], "minimal")


def test_beacon_state_full_mainnet(engine, oracle_ssz_c):
    """BASELINE.json config 3: 2**20 validators, bit-exact vs golden (hashlib) and vs the C oracle; resident + sharded paths."""
    st = S.synth_state(1 << 20, "mainnet")
    b = S.serialize(st)
    want = GOLDEN["mainnet:1048576:default"]
    assert ssz.hash_tree_root_beacon_state(b, "mainnet").hex() == want
    dev = ssz.DeviceBeaconState(b, "mainnet")
    assert dev.hash_tree_root().hex() == want
    assert dev.hash_tree_root().hex() == want
    dev.close()
    for world in (1, 2, 8):
        roots = b"".join(ssz.shard_roots(b, "mainnet", r, world) for r in range(world))
        assert ssz.combine_roots(b, "mainnet", world, roots).hex() == want
    # property: flipping one byte of one validator changes the root; restoring it restores the root
    pos = len(b) // 2
    b[pos] ^= 1
    assert ssz.hash_tree_root_beacon_state(b, "mainnet").hex() != want
    b[pos] ^= 1
    assert ssz.hash_tree_root_beacon_state(b, "mainnet").hex() == want


@pytest.mark.parametrize("n,world", [(5, 2), (1000, 4), (4097, 8), (0, 2)])
def test_sharded_state_small(engine, n, world):
    st = S.synth_state(n, "mainnet", n_historical_summaries=3, n_historical_roots=2)
    b = S.serialize(st)
    want = so.beacon_state_type("mainnet").htr(S.to_oracle_value(st))
    roots = b"".join(ssz.shard_roots(b, "mainnet", r, world) for r in range(world))
    assert ssz.combine_roots(b, "mainnet", world, roots) == want
