"""Per-phase timing of b200_state_root_incremental (B200_SSZ_TRACE=1): python tools/probe_incremental.py [validators]"""
import os, sys, time
os.environ.setdefault("B200_SSZ_TRACE", "1")
sys.path.insert(0, '.')
import numpy as np
from ethereum_consensus_b200 import _lib, ssz, state as S
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
_lib.init(0)
st = S.synth_state(N, "mainnet"); b = S.serialize(st); lay = S.layout(st)
dev = ssz.DeviceBeaconState(b, "mainnet")
rng = np.random.default_rng(1)
for it in range(4):
    att = np.unique(rng.choice(N, N // 32)).astype(np.uint64)
    t0 = time.perf_counter()
    dev.update_elements("current_epoch_participation", att, rng.integers(1, 8, len(att)).astype(np.uint8))
    bi = np.unique(rng.choice(N, 513)).astype(np.uint64)
    dev.update_elements("balances", bi, rng.integers(1, 2**40, len(bi)).astype("<u8"))
    t1 = time.perf_counter()
    if it >= 2:
        dev.update_bytes(lay["slot"][0], int(it).to_bytes(8, "little"))
        dev.update_bytes(lay["randao_mixes"][0] + 32 * it, bytes(32))
    t2 = time.perf_counter()
    r = dev.hash_tree_root_incremental()
    t3 = time.perf_counter()
    print(f"iter {it}: update_elements {1e3*(t1-t0):.3f} ms, update_bytes {1e3*(t2-t1):.3f} ms, root {1e3*(t3-t2):.3f} ms", file=sys.stderr)
print("full:", file=sys.stderr); dev.hash_tree_root()
