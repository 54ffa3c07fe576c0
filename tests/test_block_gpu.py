"""GPU parity for BASELINE.json configs[3] (deneb process_block full signature set, bit-exact vs the CPU oracle) and
configs[4] (epoch-scale attestation batch, sharded), plus the signing-root helpers."""
import ctypes
import hashlib

import numpy as np
import pytest

from ethereum_consensus_b200 import block, crypto, parallel, signing
from oracle import ssz_oracle as so

pytestmark = pytest.mark.gpu
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
SK0, DELTA = 0x1234567890abcdef1234567890abcdef, 0xfedcba0987654321


def _registry(orc, n):
    keys = np.empty((n, 48), dtype=np.uint8)
    orc.orc_pk_sequence(SK0.to_bytes(32, "big"), DELTA.to_bytes(32, "big"), n, keys.ctypes.data)
    return keys


def _sk(i): return (SK0 + i * DELTA) % R


def _sign_many(orc, sks, roots):
    n = len(sks)
    sk_b = np.frombuffer(b"".join(int(s if s else 1).to_bytes(32, "big") for s in sks), dtype=np.uint8).copy()
    m = np.frombuffer(b"".join(roots), dtype=np.uint8).copy()
    out = np.empty((n, 96), dtype=np.uint8)
    orc.orc_sign_batch(sk_b.ctypes.data, m.ctypes.data, n, out.ctypes.data, 8)
    return [out[i].tobytes() for i in range(n)]


def test_signing_helpers_match_oracle(engine):
    gvr = hashlib.sha256(b"gvr").digest()
    for dt in (signing.DomainType.BeaconAttester, signing.DomainType.SyncCommittee, signing.DomainType.Deposit):
        d = signing.compute_domain(dt, bytes.fromhex("04000000"), gvr)
        assert d == so.compute_domain(dt.as_bytes(), bytes.fromhex("04000000"), gvr)
        obj = hashlib.sha256(b"obj%d" % dt).digest()
        assert signing.compute_signing_root(obj, d) == so.compute_signing_root(obj, d)
    att = {"slot": 8_626_176, "index": 5, "beacon_block_root": hashlib.sha256(b"bbr").digest(),
           "source": {"epoch": 269566, "root": hashlib.sha256(b"s").digest()}, "target": {"epoch": 269567, "root": hashlib.sha256(b"t").digest()}}
    d = so.compute_domain(bytes.fromhex("01000000"), bytes.fromhex("04000000"), gvr)
    assert signing.compute_signing_root(so.AttestationData.htr(att), d) == so.compute_signing_root(so.AttestationData.htr(att), d)


def test_deneb_block_signature_set_bit_exact(engine, oracle_bls_c):
    """SURVEY.md Appendix C rows 1-9 in execution order; verdict vector identical to the CPU oracle's, first-failure
    and deposit-tolerance semantics as the reference."""
    orc = oracle_bls_c
    n_val = 2048
    keys = _registry(orc, n_val)
    pk = [keys[i].tobytes() for i in range(n_val)]
    rng = np.random.default_rng(4)
    root = lambda tag: hashlib.sha256(tag).digest()  # noqa: E731
    plan = []  # (site, signer indices, root, tolerant, eth)
    plan.append(("block_signature", [17], root(b"block"), False, False))
    plan.append(("randao", [17], root(b"randao"), False, False))
    for i in range(2):
        for h in range(2):
            plan.append(("proposer_slashing", [40 + i], root(b"ps%d%d" % (i, h)), False, False))
    for h in range(2):
        plan.append(("attester_slashing", sorted(rng.choice(n_val, 300, replace=False).tolist()), root(b"as%d" % h), False, False))
    for a in range(128):
        plan.append(("attestation", sorted(rng.choice(n_val, 96, replace=False).tolist()), root(b"att%d" % a), False, False))
    for dpt in range(16):
        plan.append(("deposit", [1000 + dpt], root(b"dep%d" % dpt), True, False))
    for e in range(16):
        plan.append(("voluntary_exit", [300 + e], root(b"exit%d" % e), False, False))
    for c in range(16):
        plan.append(("bls_to_execution_change", [500 + c], root(b"chg%d" % c), False, False))
    sync = sorted(rng.choice(512, 500, replace=True).tolist())  # duplicates allowed in a sync committee
    plan.append(("sync_aggregate", sync, root(b"sync"), False, True))

    sks = [sum(_sk(i) for i in signers) % R for _, signers, _, _, _ in plan]
    sigs = _sign_many(orc, sks, [p[2] for p in plan])
    # two deposits carry garbage signatures / keys: tolerated, not block failures
    dep = [i for i, p in enumerate(plan) if p[0] == "deposit"]
    sigs[dep[3]] = sigs[dep[4]]                       # valid point, wrong signature
    bad_key_dep = dep[7]

    s = block.SignatureSet()
    for i, (site, signers, r, tol, eth) in enumerate(plan):
        pks = [pk[j] for j in signers]
        if i == bad_key_dep:
            pks = [bytes(48)]                          # deposit pubkey from the message: any 48 bytes
        if site in ("attestation", "attester_slashing"):
            s.add_indexed_attestation(site, pk, signers, r, sigs[i])
        else:
            s.add(site, pks, r, sigs[i], tolerant=tol, eth_variant=eth)
    codes = s.verify()

    # CPU oracle on the identical tuples
    want = np.empty(len(plan), dtype=np.int32)
    flat = np.frombuffer(b"".join(p for e in s.entries for p in e.pubkeys), dtype=np.uint8)
    off = np.cumsum([0] + [len(e.pubkeys) for e in s.entries]).astype(np.uint32)
    msgs = np.frombuffer(b"".join(e.signing_root for e in s.entries), dtype=np.uint8)
    sg = np.frombuffer(b"".join(e.signature for e in s.entries), dtype=np.uint8)
    orc.orc_fast_aggregate_verify_batch(flat.ctypes.data, off.ctypes.data, msgs.ctypes.data, sg.ctypes.data, len(plan), want.ctypes.data, 8)
    assert codes.tolist() == want.tolist()
    assert codes[dep[3]] == 5 and codes[bad_key_dep] == 1 and (np.delete(codes, [dep[3], bad_key_dep]) == 0).all()
    assert s.first_failure(codes) is None and s.skipped_deposits(codes) == [dep[3], bad_key_dep]

    # a bad attestation in the middle aborts the block at that attestation, later failures are not reported
    att = [i for i, p in enumerate(plan) if p[0] == "attestation"]
    s.entries[att[50]].signature = s.entries[att[51]].signature
    s.entries[-1].signature = s.entries[0].signature
    codes2 = s.verify()
    assert s.first_failure(codes2) == (att[50], "attestation", 5)
    # host-side indexed-attestation checks fire before any BLS work
    for bad, msg in (([], "Empty"), ([5, 3], "NotSorted"), ([3, 3, 4], "Duplicate"), ([1, n_val], "InvalidIndex")):
        with pytest.raises(block.InvalidIndexedAttestation, match=msg):
            block.SignatureSet().add_indexed_attestation("attestation", pk, bad, root(b"x"), sigs[0])
    # sync aggregate with no participants and the infinity signature is valid (crypto/bls.rs:150-160)
    e = block.SignatureSet()
    e.add("sync_aggregate", [], root(b"sync"), crypto.INFINITY_COMPRESSED_SIGNATURE, eth_variant=True)
    assert e.verify().tolist() == [0]


def test_epoch_scale_batch_sharded(engine, oracle_bls_c):
    """configs[4] in miniature: slots x committees tuples, every validator attests once; contiguous shards verified
    independently reproduce the unsharded verdict vector (the NCCL exchange itself is covered by the 2-rank gloo test
    and the multi-GPU bench run)."""
    orc = oracle_bls_c
    slots, committees, k = 8, 8, 32
    t = slots * committees
    keys = _registry(orc, t * k)
    perm = np.random.default_rng(5).permutation(t * k)
    roots = [hashlib.sha256(b"epoch%d" % i).digest() for i in range(t)]
    sks = [sum(_sk(int(j)) for j in perm[i * k:(i + 1) * k]) % R for i in range(t)]
    sigs = _sign_many(orc, sks, roots)
    sigs[10] = sigs[11]
    flat = np.ascontiguousarray(keys[perm]).reshape(-1)
    off = (np.arange(t + 1) * k).astype(np.uint32)
    msgs = np.frombuffer(b"".join(roots), dtype=np.uint8)
    sg = np.frombuffer(b"".join(sigs), dtype=np.uint8)
    whole = crypto.fast_aggregate_verify_batch(flat, off, msgs, sg)
    assert whole.tolist() == [0] * 10 + [5] + [0] * (t - 11)
    for world in (2, 8):
        parts = []
        for r in range(world):
            lo, hi = parallel.tuple_shard(t, world, r)
            parts.append(crypto.fast_aggregate_verify_batch(flat[48 * k * lo:48 * k * hi], (off[lo:hi + 1] - off[lo]).astype(np.uint32),
                                                            msgs[32 * lo:32 * hi], sg[96 * lo:96 * hi]))
        assert np.concatenate(parts).tolist() == whole.tolist()
