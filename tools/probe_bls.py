"""Quick perf probe (not the bench): T tuples x K keys replicated from a few oracle-signed combos."""
import sys, time, hashlib
sys.path.insert(0, '.')
import numpy as np
from oracle import bls_oracle as bo
from ethereum_consensus_b200 import crypto, _lib
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ND, NM = 64, 4
sks = [int.from_bytes(hashlib.sha256(b"p%d" % i).digest(), 'big') % bo.R for i in range(ND)]
pks = [bo.sk_to_pk(s) for s in sks]
combos = []
for m in range(NM):
    msg = hashlib.sha256(b"m%d" % m).digest()
    signers = [(m * 7 + j) % ND for j in range(K)]
    H = bo.pt_from_affine(bo.F2, bo.hash_to_g2(msg))
    sig = bo.g2_compress(bo.pt_to_affine(bo.F2, bo.pt_mul(bo.F2, H, sum(sks[i] for i in signers) % bo.R)))
    combos.append((signers, msg, sig))
pk_arr = np.frombuffer(b"".join(pks), dtype=np.uint8).reshape(ND, 48)
idx = np.concatenate([np.array(combos[t % NM][0], dtype=np.uint32) for t in range(T)])
off = (np.arange(T + 1) * K).astype(np.uint32)
flat = np.ascontiguousarray(pk_arr[idx]).reshape(-1)
msgs = np.frombuffer(b"".join(combos[t % NM][1] for t in range(T)), dtype=np.uint8)
sigs = np.frombuffer(b"".join(combos[t % NM][2] for t in range(T)), dtype=np.uint8).copy()
sigs[96 * 5 + 50] ^= 1   # one corrupted tuple
_lib.init(0)
for it in range(3):
    t0 = time.time(); codes = crypto.fast_aggregate_verify_batch(flat, off, msgs, sigs); dt = time.time() - t0
    print(f"strict  T={T} K={K}: wall {dt*1e3:.1f} ms, kernels {crypto.last_kernel_ms():.1f} ms, key_validate {crypto.last_dominant_kernel_ms():.1f} ms, "
          f"{T/dt:.0f} tuples/s, ok={int((codes==0).sum())} fail={int((codes==5).sum())} other={int(((codes!=0)&(codes!=5)).sum())}", flush=True)
reg = crypto.Registry(np.ascontiguousarray(pk_arr).reshape(-1))
for it in range(3):
    t0 = time.time(); codes = reg.verify_batch(idx, off, msgs, sigs); dt = time.time() - t0
    print(f"registry T={T} K={K}: wall {dt*1e3:.1f} ms, kernels {crypto.last_kernel_ms():.1f} ms, {T/dt:.0f} tuples/s, ok={int((codes==0).sum())}", flush=True)
