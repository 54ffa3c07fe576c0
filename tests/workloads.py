"""Synthetic workloads at BASELINE.json's config sizes, shared by bench.py and the config-scale parity tests
(tests/test_config_scale_gpu.py): the generator lives here so that the numbers bench.py times and the verdicts the
tests pin against the CPU oracle are the SAME inputs.  Test infrastructure: signing
uses the C oracle (`orc`, oracle/c/bls_oracle.c through ctypes)."""
from __future__ import annotations

import hashlib
import json
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
SEED = 0xB200


def make_bls_workload(orc, T: int, K: int, rank: int, n_distinct: int = 1 << 15, n_registry: int = 1 << 20, threads: int = 8):
    """registry of 2**20 validators (n_distinct distinct keys tiled), T tuples of K signers drawn
    by a seeded permutation, 32-byte signing roots, aggregate signatures; ~3 % adversarial tuples."""
    sk0 = int.from_bytes(hashlib.sha256(b"b200/sk0" + SEED.to_bytes(8, "little")).digest(), "big") % R_ORDER
    delta = int.from_bytes(hashlib.sha256(b"b200/delta" + SEED.to_bytes(8, "little")).digest(), "big") % R_ORDER
    keys = np.empty((n_distinct, 48), dtype=np.uint8)
    orc.orc_pk_sequence(sk0.to_bytes(32, "big"), delta.to_bytes(32, "big"), n_distinct, keys.ctypes.data)
    rng = np.random.default_rng(SEED + 1000 * rank)
    perm = rng.permutation(n_registry).astype(np.uint32)
    need = T * K
    idx = np.resize(perm, need).astype(np.uint32)  # every validator attests need / n_registry times
    off = (np.arange(T + 1, dtype=np.uint64) * K).astype(np.uint32)
    msgs = np.frombuffer(b"".join(hashlib.sha256(b"b200/msg" + SEED.to_bytes(8, "little") + (rank << 32 | t).to_bytes(8, "little")).digest()
                                  for t in range(T)), dtype=np.uint8).copy().reshape(T, 32)
    kd = (idx % n_distinct).reshape(T, K).astype(object)
    kind = np.zeros(T, dtype=np.int32)  # 0 valid, 1 wrong msg, 2 signer missing, 3 sig not in group, 4 pk infinity, 5 keys cancel
    u = rng.random(T)
    kind[u < 0.01] = 1
    kind[(u >= 0.01) & (u < 0.02)] = 2
    kind[(u >= 0.02) & (u < 0.025)] = 3
    kind[(u >= 0.025) & (u < 0.0275)] = 4
    kind[(u >= 0.0275) & (u < 0.03)] = 5
    if K < 2:
        kind[kind == 5] = 1  # a cancelling pair needs two keys
    sks = np.empty((T, 32), dtype=np.uint8)
    for t in range(T):
        row = kd[t]
        s = (K * sk0 + delta * int(row.sum())) % R_ORDER
        if kind[t] == 2:
            s = (s - (sk0 + delta * int(row[-1]))) % R_ORDER
        sks[t] = np.frombuffer((s if s else 1).to_bytes(32, "big"), dtype=np.uint8)
    sign_msgs = msgs.copy()
    sign_msgs[kind == 1] ^= 0x55  # signature made over a different root
    sigs = np.empty((T, 96), dtype=np.uint8)
    orc.orc_sign_batch(sks.ctypes.data, sign_msgs.ctypes.data, T, sigs.ctypes.data, threads)
    cases = json.loads((ROOT / "tests" / "golden" / "bls_cases.json").read_text())["fast_aggregate_verify"]
    bad_sig = next(bytes.fromhex(c["sig"]) for c in cases if c["name"] == "signature not in subgroup")
    sigs[kind == 3] = np.frombuffer(bad_sig, dtype=np.uint8)
    flat = keys[idx % n_distinct].reshape(T, K, 48).copy()
    inf_pk = np.zeros(48, dtype=np.uint8); inf_pk[0] = 0xC0
    for t in np.nonzero(kind == 4)[0]:
        flat[t, K // 3] = inf_pk
    for t in np.nonzero(kind == 5)[0]:  # K/2 pairs (P, -P): the aggregate key is the point at infinity
        half = flat[t, : K // 2].copy()
        neg = half.copy(); neg[:, 0] ^= 0x20
        flat[t, : K // 2] = half; flat[t, K // 2: 2 * (K // 2)] = neg
    expect = np.select([kind == 0, kind == 4], [0, 6], default=5).astype(np.int32)
    registry = np.tile(keys, (n_registry // n_distinct, 1))
    return {"pks": flat.reshape(-1), "off": off, "msgs": msgs.reshape(-1), "sigs": sigs.reshape(-1), "expect": expect, "kind": kind,
            "idx": idx, "registry": registry.reshape(-1), "T": T, "K": K}


def make_deneb_block_plan(orc, n_registry: int = 1 << 20, n_distinct: int = 1 << 15, k_att: int = 512, k_slash: int = 2048,
                          n_att: int = 128, n_ps: int = 16, n_as: int = 2, n_dep: int = 16, n_exit: int = 16, n_chg: int = 16,
                          sync_size: int = 512, threads: int = 8, seed: int = SEED):
    """BASELINE configs[3] / SURVEY.md Appendix C at spec shape: the <= 215 signature checks of one deneb block in
    execution order.  Returns (registry_keys[n_registry x 48], plan) where each plan row is a dict
    {site, indices | None, pubkeys | None, root, sig, tolerant, eth, expect}: rows 1-5, 7, 9 name their signers by
    validator index (registry mode applies), rows 6 and 8 carry the public key in the message (strict path always).
    A few rows are adversarial on purpose (wrong signature, undecodable deposit key); `expect` is the code by construction."""
    sk0 = int.from_bytes(hashlib.sha256(b"b200/sk0" + seed.to_bytes(8, "little")).digest(), "big") % R_ORDER
    delta = int.from_bytes(hashlib.sha256(b"b200/delta" + seed.to_bytes(8, "little")).digest(), "big") % R_ORDER
    keys = np.empty((n_distinct, 48), dtype=np.uint8)
    orc.orc_pk_sequence(sk0.to_bytes(32, "big"), delta.to_bytes(32, "big"), n_distinct, keys.ctypes.data)
    registry = np.tile(keys, (n_registry // n_distinct, 1))
    rng = np.random.default_rng(seed + 77)

    def sk_of(indices):
        d = np.asarray(indices, dtype=np.int64) % n_distinct
        return (len(d) * sk0 + delta * int(d.astype(object).sum())) % R_ORDER

    def root(tag: bytes) -> bytes:
        return hashlib.sha256(b"b200/block/" + tag).digest()

    rows = []

    def by_index(site, indices, r, eth=False):
        rows.append({"site": site, "indices": [int(i) for i in indices], "pubkeys": None, "root": r, "tolerant": False, "eth": eth,
                     "sk": sk_of(indices), "expect": 0})

    proposer = int(rng.integers(n_registry))
    by_index("block_signature", [proposer], root(b"block"))
    by_index("randao", [proposer], root(b"randao"))
    for i in range(n_ps):
        who = int(rng.integers(n_registry))
        for h in range(2):
            by_index("proposer_slashing", [who], root(b"ps%d/%d" % (i, h)))
    for i in range(n_as):
        for h in range(2):
            by_index("attester_slashing", np.sort(rng.choice(n_registry, k_slash, replace=False)), root(b"as%d/%d" % (i, h)))
    for a in range(n_att):
        by_index("attestation", np.sort(rng.choice(n_registry, k_att, replace=False)), root(b"att%d" % a))
    for d in range(n_dep):   # new validators: keys outside the registry's index space, carried by the message
        j = int(rng.integers(n_distinct))
        rows.append({"site": "deposit", "indices": None, "pubkeys": [keys[j].tobytes()], "root": root(b"dep%d" % d), "tolerant": True,
                     "eth": False, "sk": sk_of([j]), "expect": 0})
    for e in range(n_exit):
        by_index("voluntary_exit", [int(rng.integers(n_registry))], root(b"exit%d" % e))
    for c in range(n_chg):
        j = int(rng.integers(n_distinct))
        rows.append({"site": "bls_to_execution_change", "indices": None, "pubkeys": [keys[j].tobytes()], "root": root(b"chg%d" % c),
                     "tolerant": False, "eth": False, "sk": sk_of([j]), "expect": 0})
    committee = rng.choice(n_registry, sync_size, replace=True)          # duplicates are legal in a sync committee
    bits = rng.random(sync_size) < 0.99
    by_index("sync_aggregate", committee[bits], root(b"sync"), eth=True)
    rows[-1]["committee"] = [int(i) for i in committee]
    rows[-1]["bits"] = [bool(b) for b in bits]

    sks = np.frombuffer(b"".join(int(r["sk"] if r["sk"] else 1).to_bytes(32, "big") for r in rows), dtype=np.uint8).copy()
    msgs = np.frombuffer(b"".join(r["root"] for r in rows), dtype=np.uint8).copy()
    sigs = np.empty((len(rows), 96), dtype=np.uint8)
    orc.orc_sign_batch(sks.ctypes.data, msgs.ctypes.data, len(rows), sigs.ctypes.data, threads)
    for r, s in zip(rows, sigs):
        r["sig"] = s.tobytes()
    # adversarial rows: two tolerated deposits (wrong signature; undecodable key) — not block failures
    dep = [i for i, r in enumerate(rows) if r["site"] == "deposit"]
    if len(dep) >= 8:
        rows[dep[3]]["sig"] = rows[dep[4]]["sig"]; rows[dep[3]]["expect"] = 5
        rows[dep[7]]["pubkeys"] = [bytes(48)]; rows[dep[7]]["expect"] = 1
    return registry, rows


class LazyKeys:
    """`state.validators[i].public_key` over the registry array, without materialising 2^20 bytes objects."""
    def __init__(self, registry):
        self.registry = registry

    def __len__(self):
        return len(self.registry)

    def __getitem__(self, i):
        return self.registry[i].tobytes()


def collect_block_signature_set(registry, rows):
    """Feeds a make_deneb_block_plan() plan through the host-side collector (ethereum_consensus_b200.block.SignatureSet)
    in execution order, the way process_block's sites would."""
    from ethereum_consensus_b200 import block, signing
    keys = LazyKeys(registry)
    fork = signing.Fork(bytes.fromhex("03000000"), bytes.fromhex("04000000"), 269568)
    gvr = hashlib.sha256(b"gvr").digest()
    s = block.SignatureSet()
    for r in rows:
        if r["site"] in ("attestation", "attester_slashing"):
            s.add_indexed_attestation(r["site"], keys, r["indices"], r["root"], r["sig"])
        elif r["site"] == "sync_aggregate":
            # through the collector's own gather (altair/block_processing.rs:216-243); the row was signed over r["root"], so
            # hand the pre-computed signing root in by entry surgery after checking the gather picked the same signers
            s.add_sync_aggregate([keys[i] for i in r["committee"]], r["bits"], r["sig"], 8_626_177, hashlib.sha256(b"prev").digest(), fork, gvr,
                                 committee_indices=r["committee"])
            assert list(s.entries[-1].indices) == r["indices"]
            s.entries[-1].signing_root = r["root"]
        elif r["indices"] is not None:
            s.add_by_index(r["site"], keys, r["indices"], r["root"], r["sig"])
        else:
            s.add(r["site"], r["pubkeys"], r["root"], r["sig"], tolerant=r["tolerant"], eth_variant=r["eth"])
    return s
