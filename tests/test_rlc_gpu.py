"""The RLC whole-batch check (bls_rlc.cu, north_star's fused multi-pairing) against the per-tuple path: its boolean must be
the AND of the per-tuple verdicts — for every seed tried — and the per-tuple codes remain the source of truth."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from ethereum_consensus_b200 import crypto
from tests import workloads

pytestmark = pytest.mark.gpu
GOLDEN = json.loads((Path(__file__).parent / "golden" / "bls_cases.json").read_text())["fast_aggregate_verify"]


def _pack(cases):
    pks = np.frombuffer(b"".join(bytes.fromhex(p) for c in cases for p in c["pks"]), dtype=np.uint8)
    off = np.cumsum([0] + [len(c["pks"]) for c in cases]).astype(np.uint32)
    msgs = np.frombuffer(b"".join(bytes.fromhex(c["msg"]) for c in cases), dtype=np.uint8)
    sigs = np.frombuffer(b"".join(bytes.fromhex(c["sig"]) for c in cases), dtype=np.uint8)
    return pks, off, msgs, sigs


def test_rlc_on_golden_cases(engine):
    cases = [c for c in GOLDEN if len(c["msg"]) == 64]
    good = [c for c in cases if c["code"] == 0]
    assert len(good) >= 5
    seeds = [hashlib.sha256(b"rlc seed %d" % i).digest() for i in range(3)] + [None]
    for reps in (1, 7):           # 7 x good cases: more than one warp of tuples, ragged tail
        args = _pack(good * reps)
        assert (crypto.fast_aggregate_verify_batch(*args) == 0).all()
        for sd in seeds:
            assert crypto.fast_aggregate_verify_batch_all(*args, seed=sd) is True
    # every single reject class, alone among valid tuples, flips the batch to False (and the per-tuple path names it)
    for bad in [c for c in cases if c["code"] != 0]:
        mixed = good[:3] + [bad] + good[3:]
        args = _pack(mixed)
        codes = crypto.fast_aggregate_verify_batch(*args)
        assert codes.tolist() == [c["code"] for c in mixed]
        for sd in seeds[:2]:
            assert crypto.fast_aggregate_verify_batch_all(*args, seed=sd) is False, bad["name"]
    assert crypto.fast_aggregate_verify_batch_all(*_pack([]), seed=seeds[0]) is True   # empty batch: vacuously valid


def test_rlc_swapped_signatures_do_not_cancel(engine):
    """Two valid tuples with their signatures exchanged: each pairing check fails, and so must the combination — the
    plain (unweighted) product of the two checks would NOT notice if the defects cancelled; the random weights do."""
    good = [c for c in GOLDEN if c["code"] == 0 and len(c["msg"]) == 64 and c["pks"]]
    a, b = dict(good[0]), dict(good[1])
    a["sig"], b["sig"] = b["sig"], a["sig"]
    args = _pack([a, b] + good[2:5])
    assert crypto.fast_aggregate_verify_batch(*args).tolist()[:2] == [5, 5]
    for i in range(4):
        assert crypto.fast_aggregate_verify_batch_all(*args, seed=hashlib.sha256(b"x%d" % i).digest()) is False


def test_rlc_at_config_scale(engine, oracle_bls_c):
    """T = 2048 x K = 512 (one epoch's batch): all-valid -> True; the workload's adversarial mix -> False; registry mode too."""
    w = workloads.make_bls_workload(oracle_bls_c, 2048, 512, 3, threads=16)
    seed = hashlib.sha256(b"config scale").digest()
    assert crypto.fast_aggregate_verify_batch_all(w["pks"], w["off"], w["msgs"], w["sigs"], seed=seed) is False
    ok = np.nonzero(w["kind"] == 0)[0]
    K = w["K"]
    pk = w["pks"].reshape(-1, K * 48)[ok].reshape(-1).copy()
    off = (np.arange(len(ok) + 1, dtype=np.uint64) * K).astype(np.uint32)
    ms = w["msgs"].reshape(-1, 32)[ok].reshape(-1).copy()
    sg = w["sigs"].reshape(-1, 96)[ok].reshape(-1).copy()
    assert (crypto.fast_aggregate_verify_batch(pk, off, ms, sg) == 0).all()
    assert crypto.fast_aggregate_verify_batch_all(pk, off, ms, sg, seed=seed) is True
    assert crypto.fast_aggregate_verify_batch_all(pk, off, ms, sg) is True                    # library-drawn seed
    reg = crypto.Registry(w["registry"])
    idx = w["idx"].reshape(-1, K)[ok].reshape(-1).copy()
    assert reg.verify_batch_all(idx, off, ms, sg, seed=seed) is True
    sg2 = sg.copy(); sg2[96 * 100: 96 * 101] = sg[96 * 101: 96 * 102]                          # one wrong signature
    assert reg.verify_batch_all(idx, off, ms, sg2, seed=seed) is False
