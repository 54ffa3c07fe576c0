"""Generates tests/golden/ssz_roots.json with the pure-Python hashlib oracle (oracle/ssz_oracle.py).

Run from the repo root:  python tests/golden/make_ssz_golden.py [--full]
`--full` also computes the 2**20-validator mainnet state root of BASELINE.json config 3 (takes ~1 min of
hashlib time); that entry is produced by hashing the *serialized* state with the streaming helpers below so
that no 1M-element Python dicts are needed.
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ssz_oracle as so  # noqa: E402
from ethereum_consensus_b200 import state as S  # noqa: E402

CASES = [("minimal", 0), ("minimal", 1), ("minimal", 5), ("minimal", 300), ("mainnet", 77), ("mainnet", 1000),
         ("mainnet", 4097)]


def htr_state_from_arrays(st) -> bytes:
    """hash_tree_root of a SynthState using only hashlib + the oracle's merkleize helpers (streaming over arrays)."""
    P = so.PRESETS[st.preset]
    T = so.beacon_state_type(st.preset)
    ftypes = dict(T.fields)
    f = st.fixed
    vb = st.validators.tobytes()
    h = hashlib.sha256
    roots = bytearray()
    z24, z31 = bytes(24), bytes(31)
    for i in range(len(st.validators)):
        r = vb[121 * i:121 * i + 121]
        l0 = h(r[0:48] + bytes(16)).digest()
        a = h(l0 + r[48:80]).digest()
        b = h(r[80:88] + z24 + r[88:89] + z31).digest()
        c = h(r[89:97] + z24 + r[97:105] + z24).digest()
        d = h(r[105:113] + z24 + r[113:121] + z24).digest()
        roots += h(h(a + b).digest() + h(c + d).digest()).digest()
    lim = P["VALIDATOR_REGISTRY_LIMIT"]
    n = len(st.validators)
    small = S.to_oracle_value(_strip(st))
    fr = []
    for name, t in T.fields:
        if name == "validators":
            fr.append(so.mix_in_length(so.merkleize_bytes(bytes(roots), lim), n))
        elif name == "balances":
            fr.append(so.mix_in_length(so.merkleize_bytes(st.balances.tobytes(), lim // 4), n))
        elif name == "inactivity_scores":
            fr.append(so.mix_in_length(so.merkleize_bytes(st.inactivity_scores.tobytes(), lim // 4), n))
        elif name == "previous_epoch_participation":
            fr.append(so.mix_in_length(so.merkleize_bytes(st.previous_epoch_participation.tobytes(), lim // 32), n))
        elif name == "current_epoch_participation":
            fr.append(so.mix_in_length(so.merkleize_bytes(st.current_epoch_participation.tobytes(), lim // 32), n))
        else:
            fr.append(t.htr(small[name]))
    return so.merkleize_chunks(fr)


def _strip(st):
    import copy
    import numpy as np
    s2 = copy.copy(st)
    s2.validators = st.validators[:0]
    s2.balances = st.balances[:0]
    s2.inactivity_scores = st.inactivity_scores[:0]
    s2.previous_epoch_participation = st.previous_epoch_participation[:0]
    s2.current_epoch_participation = st.current_epoch_participation[:0]
    return s2


def main():
    out = {}
    T = {p: so.beacon_state_type(p) for p in ("mainnet", "minimal")}
    for preset, n in CASES:
        st = S.synth_state(n, preset, n_historical_summaries=3, n_historical_roots=2)
        root = T[preset].htr(S.to_oracle_value(st))
        assert root == htr_state_from_arrays(st)
        out[f"{preset}:{n}:hs3:hr2"] = root.hex()
        print(preset, n, root.hex())
    path = Path(__file__).with_name("ssz_roots.json")
    if path.exists():
        old = json.loads(path.read_text())
        for k, v in old.items():
            out.setdefault(k, v)
    if "--full" in sys.argv:
        st = S.synth_state(1 << 20, "mainnet")
        root = htr_state_from_arrays(st)
        out["mainnet:1048576:default"] = root.hex()
        print("full", root.hex())
    path.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
