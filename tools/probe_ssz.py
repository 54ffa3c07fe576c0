import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from ethereum_consensus_b200 import _lib, ssz, state as S
_lib.init(0)
b = S.serialize(S.synth_state(1 << 20))
host = torch.from_numpy(b).pin_memory()
dev = torch.empty_like(host, device="cuda")
for _ in range(3): dev.copy_(host, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): dev.copy_(host, non_blocking=True); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"plain pinned H2D of {b.nbytes/1e6:.1f} MB: {dt*1e3:.2f} ms = {b.nbytes/dt/1e9:.1f} GB/s")
for name, buf in (("pinned", host), ("pageable", b)):
    for _ in range(3): ssz.hash_tree_root_beacon_state(buf)
    ts = []
    for _ in range(7):
        t0 = time.perf_counter(); r = ssz.hash_tree_root_beacon_state(buf); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"e2e from {name} host: min {min(ts):.2f} ms median {sorted(ts)[3]:.2f} ms, kernels-window {float(_lib.load().b200_last_kernel_ms()):.2f} ms")
d = ssz.DeviceBeaconState(host)
for _ in range(3): d.hash_tree_root()
print(f"resident: {float(_lib.load().b200_last_kernel_ms()):.3f} ms", r.hex()[:16])
