"""Generates tests/golden/bls_cases.json with the big-int oracle (oracle/bls_oracle.py).

Run from the repo root:  python tests/golden/make_bls_golden.py      (~2-3 minutes of Python big-int time)
Every expected code/value in the file is the oracle's output; the oracle itself is pinned to the reference's KATs
B-1/B-2 in tests/test_oracle_bls.py.
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import bls_oracle as bo  # noqa: E402

SEED = 0xB200
F1, F2 = bo.F1, bo.F2


def sk(i: int) -> int:
    """SURVEY.md §8d key schedule."""
    h = hashlib.sha256(b"b200/sk" + SEED.to_bytes(8, "little") + i.to_bytes(8, "little")).digest()
    return int.from_bytes(h, "big") % (bo.R - 1) + 1


def msg(t: int) -> bytes:
    return hashlib.sha256(b"b200/msg" + SEED.to_bytes(8, "little") + t.to_bytes(8, "little")).digest()


_PK = {}


def pk(i: int) -> bytes:
    if i not in _PK:
        _PK[i] = bo.sk_to_pk(sk(i))
    return _PK[i]


def agg_sig(signers, m: bytes) -> bytes:
    H = bo.pt_from_affine(F2, bo.hash_to_g2(m))
    return bo.g2_compress(bo.pt_to_affine(F2, bo.pt_mul(F2, H, sum(sk(i) for i in signers) % bo.R)))


def main():
    fav = []

    def add(name, pks, m, sig):
        code = bo.fast_aggregate_verify(pks, m, sig)
        fav.append({"name": name, "pks": [p.hex() for p in pks], "msg": m.hex(), "sig": sig.hex(), "code": code})
        print(f"{name:32s} K={len(pks):3d} -> {code}", flush=True)

    t = 0
    for K in (1, 2, 3, 5, 31, 32, 33, 64, 70):
        s = list(range(t, t + K))
        add(f"valid K={K}", [pk(i) for i in s], msg(t), agg_sig(s, msg(t)))
        t += K
    s = list(range(8))
    good = agg_sig(s, msg(1000))
    add("wrong message", [pk(i) for i in s], msg(1001), good)
    add("signer missing from aggregate", [pk(i) for i in s], msg(1000), agg_sig(s[:-1], msg(1000)))
    add("extra signer in aggregate", [pk(i) for i in s[:-1]], msg(1000), good)
    add("duplicate key, signed twice", [pk(i) for i in s] + [pk(0)], msg(1000), agg_sig(s + [0], msg(1000)))
    add("duplicate key, signed once", [pk(i) for i in s] + [pk(0)], msg(1000), good)
    add("permuted keys", [pk(i) for i in reversed(s)], msg(1000), good)
    add("no keys", [], msg(1000), good)
    add("infinity signature", [pk(i) for i in s], msg(1000), bytes([0xC0]) + bytes(95))
    add("no keys + infinity signature", [], msg(1000), bytes([0xC0]) + bytes(95))
    inf_pk = bytes([0xC0]) + bytes(47)
    add("infinity pubkey in the middle", [pk(0), pk(1), inf_pk, pk(2)], msg(1000), good)
    add("infinity pubkey only", [inf_pk], msg(1000), good)
    neg0 = bytes([pk(0)[0] ^ 0x20]) + pk(0)[1:]
    add("pk + (-pk) sums to infinity", [pk(0), neg0], msg(1000), good)
    add("pk + (-pk) + others", [pk(0), neg0, pk(1), pk(2)], msg(1000), agg_sig([1, 2], msg(1000)))
    add("compression bit clear", [bytes([pk(0)[0] & 0x7F]) + pk(0)[1:]], msg(1000), good)
    add("infinity flag with payload", [bytes([0xC0]) + bytes(46) + b"\x01"], msg(1000), good)
    add("infinity flag with sign bit", [bytes([0xE0]) + bytes(47)], msg(1000), good)
    add("x >= p", [bytes([0x9F]) + b"\xff" * 47], msg(1000), good)
    add("x = p", [bytes([0x80 | (bo.P >> 376)]) + (bo.P & ((1 << 376) - 1)).to_bytes(47, "big")], msg(1000), good)
    # first x whose x^3+4 is a non-residue / first curve point outside the subgroup
    x = 1
    off_curve = in_curve_not_group = None
    while off_curve is None or in_curve_not_group is None:
        b = bytes([0x80]) + x.to_bytes(48, "big")[1:]
        c, pt = bo.g1_uncompress(b)
        if c == bo.POINT_NOT_ON_CURVE and off_curve is None:
            off_curve = b
        if c == 0 and not bo.in_subgroup(F1, pt) and in_curve_not_group is None:
            in_curve_not_group = b
        x += 1
    add("pubkey not on curve", [pk(0), off_curve], msg(1000), good)
    add("pubkey not in subgroup", [pk(0), in_curve_not_group, pk(1)], msg(1000), good)
    add("bad key after bad key: first wins", [in_curve_not_group, off_curve], msg(1000), good)
    add("bad key before bad signature", [off_curve], msg(1000), bytes(96))
    q = bo.iso3(bo.sswu((5, 7)))
    add("signature not in subgroup", [pk(i) for i in s], msg(1000), bo.g2_compress(q))
    add("signature compression bit clear", [pk(i) for i in s], msg(1000), bytes([good[0] & 0x7F]) + good[1:])
    add("signature x.c1 >= p", [pk(i) for i in s], msg(1000), bytes([0x9F]) + b"\xff" * 47 + good[48:])
    add("signature x.c0 >= p", [pk(i) for i in s], msg(1000), good[:48] + b"\xff" * 48)
    xs = 1
    while True:
        b = bytes([0x80]) + bytes(47) + xs.to_bytes(48, "big")
        if bo.g2_uncompress(b)[0] == bo.POINT_NOT_ON_CURVE:
            break
        xs += 1
    add("signature not on curve", [pk(i) for i in s], msg(1000), b)
    add("signature sign bit flipped", [pk(i) for i in s], msg(1000), bytes([good[0] ^ 0x20]) + good[1:])
    # non-32-byte messages through the single-call API
    for m in (b"", b"message", bytes(range(200))):
        add(f"valid, {len(m)}-byte message", [pk(0), pk(1)], m, agg_sig([0, 1], m))

    # aggregate_verify
    av = []

    def add_av(name, pks, msgs, sig):
        code = bo.aggregate_verify(pks, msgs, sig)
        av.append({"name": name, "pks": [p.hex() for p in pks], "msgs": [m.hex() for m in msgs], "sig": sig.hex(), "code": code})
        print(f"AV {name:32s} -> {code}", flush=True)

    def sig_of(i, m):
        return bo.sign(sk(i), m)

    ms = [msg(2000 + i) for i in range(4)]
    sigs = [sig_of(i, ms[i]) for i in range(4)]
    code, agg4 = bo.aggregate(sigs)
    assert code == 0
    add_av("4 distinct messages", [pk(i) for i in range(4)], ms, agg4)
    add_av("wrong order", [pk(i) for i in (1, 0, 2, 3)], ms, agg4)
    same = [ms[0]] * 3
    code, agg_same = bo.aggregate([sig_of(i, ms[0]) for i in range(3)])
    add_av("repeated message", [pk(i) for i in range(3)], same, agg_same)
    add_av("length mismatch", [pk(i) for i in range(4)], ms[:3], agg4)
    add_av("empty", [], [], agg4)
    add_av("bad key", [pk(0), off_curve, pk(2), pk(3)], ms, agg4)
    add_av("one message", [pk(0)], [ms[0]], sigs[0])

    # aggregate (signatures) and eth_aggregate_public_keys
    ag = []
    for name, lst in (("4 sigs", sigs), ("1 sig", sigs[:1]), ("with infinity", sigs[:2] + [bytes([0xC0]) + bytes(95)]),
                      ("sig + (-sig)", [sigs[0], bytes([sigs[0][0] ^ 0x20]) + sigs[0][1:]]),
                      ("not in subgroup", [sigs[0], bo.g2_compress(q)]), ("bad encoding second", [sigs[0], bytes(96)]),
                      ("not in group then bad encoding", [bo.g2_compress(q), bytes(96)]), ("40 sigs", [sigs[i % 4] for i in range(40)])):
        code, out = bo.aggregate(lst)
        ag.append({"name": name, "sigs": [x.hex() for x in lst], "code": code, "out": out.hex() if out else None})
        print(f"AGG {name:32s} -> {code}", flush=True)
    ak = []
    for name, lst in (("8 keys", [pk(i) for i in range(8)]), ("1 key", [pk(3)]), ("pk + (-pk)", [pk(0), neg0]),
                      ("with infinity", [pk(0), inf_pk]), ("not in group", [pk(0), in_curve_not_group]), ("70 keys", [pk(i) for i in range(70)])):
        code, out = bo.eth_aggregate_public_keys(lst)
        ak.append({"name": name, "pks": [x.hex() for x in lst], "code": code, "out": out.hex() if out else None})
        print(f"AGGPK {name:32s} -> {code}", flush=True)

    Path(__file__).with_name("bls_cases.json").write_text(json.dumps(
        {"seed": SEED, "fast_aggregate_verify": fav, "aggregate_verify": av, "aggregate": ag, "eth_aggregate_public_keys": ak}, indent=0) + "\n")


if __name__ == "__main__":
    main()
