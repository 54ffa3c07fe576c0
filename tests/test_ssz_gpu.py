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


@pytest.mark.parametrize("n,preset,oracle_check", [(5, "minimal", True), (70, "minimal", True), (333, "mainnet", True),
                                                   (5000, "mainnet", False), (70001, "mainnet", False)])
def test_incremental_state_root(engine, n, preset, oracle_check):
    """patch a device-resident state, re-hash only the dirty paths.  After every round the
    incremental root must equal (a) a from-scratch GPU hash of the patched serialization and (b), for the small states,
    the Python oracle's root of the identically patched value."""
    from ethereum_consensus_b200._lib import EngineError
    st = S.synth_state(n, preset, n_historical_summaries=3, n_historical_roots=2)
    b = S.serialize(st).copy()
    lay = S.layout(st)
    dev = ssz.DeviceBeaconState(b, preset)
    rng = np.random.default_rng(n)
    typ = so.beacon_state_type(preset)
    assert dev.hash_tree_root_incremental() == dev.hash_tree_root() == ssz.hash_tree_root_beacon_state(b, preset)

    def check(full_first=False):
        want = ssz.hash_tree_root_beacon_state(b, preset)
        if full_first:      # the full O(N) re-hash of the resident state must honour pending updates too
            assert dev.hash_tree_root() == want
        got = dev.hash_tree_root_incremental()
        assert got == want
        assert bytes(S.serialize(st)) == bytes(b)
        if oracle_check:
            assert got == typ.htr(S.to_oracle_value(st))
        assert dev.hash_tree_root_incremental() == got      # nothing dirty: same root
        assert dev.hash_tree_root() == got                  # and the full re-hash agrees

    vbytes = st.validators.view(np.uint8).reshape(n, 121)
    for rnd in range(3):
        # validators: a few records change (effective balance, slashed flag, epochs, credentials)
        k = min(n, 1 + 8 * rnd + (n > 1000) * 300)
        idx = np.sort(rng.choice(n, k, replace=False)).astype(np.uint64)
        if rnd == 0:
            idx[-1] = n - 1                                    # the last record (ragged tail of the tree)
        recs = vbytes[idx.astype(np.int64)].copy()
        recs[:, 48:80] = rng.integers(0, 256, (len(idx), 32), dtype=np.uint8)
        recs[:, 80:88] = np.frombuffer(rng.integers(0, 32 * 10**9, len(idx), dtype=np.uint64).astype("<u8").tobytes(), dtype=np.uint8).reshape(-1, 8)
        recs[:, 88] ^= 1
        dev.update_elements("validators", idx, recs)
        vbytes[idx.astype(np.int64)] = recs
        o = lay["validators"][0]
        for i, r in zip(idx, recs):
            b[o + 121 * int(i): o + 121 * int(i) + 121] = r
        # packed lists: scattered elements, duplicates allowed (last write wins only if values agree: use unique)
        for name, arr in (("balances", st.balances), ("inactivity_scores", st.inactivity_scores),
                          ("previous_epoch_participation", st.previous_epoch_participation),
                          ("current_epoch_participation", st.current_epoch_participation)):
            m = min(n, 3 + 40 * rnd + (n > 1000) * 2000)
            ii = np.unique(rng.choice(n, m)).astype(np.uint64)
            if rnd == 1:
                ii = np.unique(np.append(ii, [0, n - 1])).astype(np.uint64)
            vals = rng.integers(0, 8 if arr.dtype == np.uint8 else 2**40, len(ii)).astype(arr.dtype)
            dev.update_elements(name, ii, vals)
            arr[ii.astype(np.int64)] = vals
            o, ln = lay[name]
            b[o:o + ln] = np.frombuffer(arr.tobytes(), dtype=np.uint8)
        # small fields through the byte interface: slot, one block root, one randao mix
        slot = int(rng.integers(1, 2**40)).to_bytes(8, "little")
        dev.update_bytes(lay["slot"][0], slot)
        st.fixed["slot"] = slot
        b[lay["slot"][0]: lay["slot"][0] + 8] = np.frombuffer(slot, dtype=np.uint8)
        for name, arr in (("block_roots", st.block_roots), ("randao_mixes", st.randao_mixes)):
            j = int(rng.integers(0, arr.shape[0]))
            v = rng.integers(0, 256, 32, dtype=np.uint8)
            dev.update_bytes(lay[name][0] + 32 * j, v.tobytes())
            arr[j] = v
            b[lay[name][0] + 32 * j: lay[name][0] + 32 * j + 32] = v
        if rnd == 2 and n >= 5:
            # one byte range straddling two big lists: the last two validators and the first three balances
            vo, vl = lay["validators"]
            bo, _ = lay["balances"]
            assert vo + vl == bo
            lo, hi = bo - 2 * 121, bo + 3 * 8
            patch = rng.integers(0, 256, hi - lo, dtype=np.uint8)
            patch[88] &= 1; patch[121 + 88] &= 1             # `slashed` stays a boolean
            dev.update_bytes(lo, patch.tobytes())
            b[lo:hi] = patch
            vbytes[n - 2:] = patch[:242].reshape(2, 121)
            st.balances[:3] = np.frombuffer(patch[242:].tobytes(), dtype="<u8")
        check(full_first=(rnd == 1))

    # rejected updates leave the state untouched
    root = dev.hash_tree_root_incremental()
    with pytest.raises(EngineError):
        dev.update_elements("balances", np.array([n], dtype=np.uint64), np.zeros(1, dtype="<u8"))
    with pytest.raises(EngineError):
        dev.update_bytes(lay["offset:balances"][0], (123).to_bytes(4, "little"))   # would move a variable-size field
    with pytest.raises(EngineError):
        dev.update_bytes(len(b) - 2, b"abcd")
    assert dev.hash_tree_root_incremental() == root
    dev.close()
