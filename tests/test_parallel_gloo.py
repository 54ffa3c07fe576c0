"""CPU, world_size = 2 over gloo (127.0.0.1): the N>1 host logic — tuple sharding + verdict all_gather, and the
BeaconState slice math + root all_gather + combine — with oracle stand-ins for the two device calls."""
import ctypes
import json
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

from ethereum_consensus_b200 import parallel

ROOT = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent('''
    import ctypes, json, os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, sys.argv[1])
    from ethereum_consensus_b200 import parallel, state as S
    from oracle import ssz_oracle as so
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    out = {}

    # ---- SSZ: oracle stand-ins for b200_htr_beacon_state_deneb_{shard,combine}
    st = S.synth_state(1000, "mainnet", n_historical_summaries=3, n_historical_roots=2)
    ssz_bytes = S.serialize(st).tobytes()
    T = so.beacon_state_type("mainnet")
    val = S.to_oracle_value(st)
    lim = so.PRESETS["mainnet"]["VALIDATOR_REGISTRY_LIMIT"]
    def chunks_of(name):
        if name == "validators":
            return [so.Validator.htr(v) for v in val["validators"]]
        raw = {"balances": st.balances, "inactivity_scores": st.inactivity_scores,
               "previous_epoch_participation": st.previous_epoch_participation,
               "current_epoch_participation": st.current_epoch_participation}[name].tobytes()
        raw += bytes(-len(raw) % 32)
        return [raw[i:i + 32] for i in range(0, len(raw), 32)]
    NAMES = ["validators", "balances", "previous_epoch_participation", "current_epoch_participation", "inactivity_scores"]
    DEPTH = {"validators": 40, "balances": 38, "previous_epoch_participation": 35, "current_epoch_participation": 35, "inactivity_scores": 38}
    LENS = {"validators": 1000, "balances": 1000, "previous_epoch_participation": 1000, "current_epoch_participation": 1000, "inactivity_scores": 1000}
    def shard_fn(_ssz, _preset, r, w):
        outb = b""
        for nme in NAMES:
            ch = chunks_of(nme)
            lo, cnt, k = parallel.slice_of(len(ch), w, r)
            outb += so.merkleize_chunks(ch[lo:lo + cnt], 1 << k)
        return outb
    def combine_fn(_ssz, _preset, w, allr):
        roots = {}
        for q, nme in enumerate(NAMES):
            _, _, k = parallel.slice_of(len(chunks_of(nme)), w, 0)
            nodes = [allr[(r * 5 + q) * 32:(r * 5 + q) * 32 + 32] for r in range(w)]
            level = k
            while level < DEPTH[nme]:
                if len(nodes) & 1:
                    nodes.append(so.ZERO_HASHES[level])
                nodes = [so.hash_pair(nodes[i], nodes[i + 1]) for i in range(0, len(nodes), 2)]
                level += 1
            roots[nme] = so.mix_in_length(nodes[0], LENS[nme])
        fr = [roots[n] if n in roots else t.htr(val[n]) for n, t in T.fields]
        return so.merkleize_chunks(fr)
    out["ssz_root"] = parallel.sharded_beacon_state_root(ssz_bytes, "mainnet", rank, world, shard_fn, combine_fn).hex()
    out["ssz_want"] = T.htr(val).hex()

    # ---- BLS: the C oracle stands in for b200_fast_aggregate_verify_batch
    L = ctypes.CDLL(os.path.join(sys.argv[1], "oracle", "liboracle_bls.so"))
    L.orc_fast_aggregate_verify_batch.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int]
    cases = [c for c in json.load(open(os.path.join(sys.argv[1], "tests", "golden", "bls_cases.json")))["fast_aggregate_verify"]
             if len(c["msg"]) == 64 and len(c["pks"]) <= 8][:11]
    pks = np.frombuffer(b"".join(bytes.fromhex(p) for c in cases for p in c["pks"]), dtype=np.uint8)
    off = np.cumsum([0] + [len(c["pks"]) for c in cases]).astype(np.uint32)
    msgs = np.frombuffer(b"".join(bytes.fromhex(c["msg"]) for c in cases), dtype=np.uint8)
    sigs = np.frombuffer(b"".join(bytes.fromhex(c["sig"]) for c in cases), dtype=np.uint8)
    def verify_fn(p, o, m, s):
        n = len(o) - 1
        res = np.empty(max(n, 1), dtype=np.int32)
        p = np.ascontiguousarray(p); m = np.ascontiguousarray(m); s = np.ascontiguousarray(s)
        L.orc_fast_aggregate_verify_batch(p.ctypes.data, o.ctypes.data, m.ctypes.data, s.ctypes.data, n, res.ctypes.data, 1)
        return res[:n]
    got = parallel.sharded_fast_aggregate_verify(pks, off, msgs, sigs, rank, world, verify_fn)
    out["bls_codes"] = got.tolist()
    out["bls_want"] = [c["code"] for c in cases]
    out["shard"] = parallel.tuple_shard(len(cases), world, rank)
    print("RESULT" + json.dumps(out), flush=True)
    dist.destroy_process_group()
''')


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_slice_math():
    for n in (0, 1, 5, 1000, 4097, 1 << 20):
        for world in (1, 2, 4, 8):
            cover, k0 = 0, None
            for r in range(world):
                lo, cnt, k = parallel.slice_of(n, world, r)
                assert lo == min(n, r << k) and (k0 is None or k == k0)
                cover, k0 = cover + cnt, k
            assert cover == n and (world << k0) >= n
    for t in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            edges = [parallel.tuple_shard(t, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == t and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))


def test_two_ranks_over_gloo(tmp_path):
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(ROOT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        so_, se = p.communicate(timeout=300)
        assert p.returncode == 0, se[-2000:]
        outs.append(json.loads(next(l for l in so_.splitlines() if l.startswith("RESULT"))[6:]))
    for o in outs:
        assert o["ssz_root"] == o["ssz_want"]
        assert o["bls_codes"] == o["bls_want"]
    assert outs[0]["ssz_root"] == outs[1]["ssz_root"]
    assert outs[0]["shard"] != outs[1]["shard"]
