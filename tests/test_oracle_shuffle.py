"""The shuffling oracle pinned against itself: the reference's two formulations (per-index map, phase0/helpers.rs:249-283;
whole-list walk, :287-360) restated independently must agree — the same cross-check the reference's own runner performs
against the spec vectors (spec-tests/runners/shuffling.rs:35-45)."""
import hashlib

import numpy as np
import pytest

from oracle import shuffle_oracle as sh


@pytest.mark.parametrize("rounds", [10, 90])
def test_two_formulations_agree_small(rounds):
    for n in list(range(0, 40)) + [255, 256, 257, 300, 511, 513]:
        seed = hashlib.sha256(b"seed%d" % n).digest()
        by_index = [sh.compute_shuffled_index(i, n, seed, rounds) for i in range(n)]
        assert sorted(by_index) == list(range(n))                       # a permutation
        assert sh.compute_shuffled_indices(list(range(n)), seed, rounds) == by_index
        assert sh.shuffled_indices_numpy(n, seed, rounds).tolist() == by_index


def test_non_identity_input_and_trivial_sizes():
    seed = hashlib.sha256(b"x").digest()
    assert sh.compute_shuffled_indices([], seed) == []
    assert sh.compute_shuffled_indices([42], seed) == [42] and sh.compute_shuffled_index(0, 1, seed) == 0
    with pytest.raises(ValueError):
        sh.compute_shuffled_index(3, 3, seed)
    vals = [1000 + 7 * i for i in range(100)]
    want = [vals[sh.compute_shuffled_index(i, 100, seed)] for i in range(100)]
    assert sh.compute_shuffled_indices(vals, seed) == want
    assert sh.shuffled_indices_numpy(vals, seed).tolist() == want


def test_numpy_form_matches_list_walk_mid_size():
    seed = hashlib.sha256(b"mid").digest()
    n = 5000
    assert sh.shuffled_indices_numpy(n, seed, 90).tolist() == sh.compute_shuffled_indices(list(range(n)), seed, 90)


def test_active_indices_oracle():
    from ethereum_consensus_b200 import state as S
    st = S.synth_state(500, "minimal")
    lay = S.layout(st)
    ser = S.serialize(st)
    vo, vl = lay["validators"]
    recs = bytes(ser[vo: vo + vl])
    for epoch in (0, 1 << 10, 1 << 17, 1 << 18, 2**64 - 1):
        act = sh.get_active_validator_indices(recs, epoch)
        assert act == sorted(act) and all(0 <= i < 500 for i in act)
    assert 0 < len(sh.get_active_validator_indices(recs, 1 << 18)) <= 500
