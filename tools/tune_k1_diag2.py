"""K1 (k_g1_validate_main) in-situ timing probe: does a foreign (torch) kernel having run in the process matter?
usage: tune_k1_diag2.py <startup> <per_iter>
  startup : none | empty (torch.empty only) | kernel (one torch fill kernel) | mm (one cuBLAS matmul)
  per_iter: none | rtmemset (cudaMemset 256 MiB via the runtime, no torch kernel) | zero (torch zero_ 256 MiB)
"""
import ctypes as C, sys, time
sys.path.insert(0, '.')
import numpy as np
from ethereum_consensus_b200 import _lib, crypto
L = C.CDLL('oracle/liboracle_bls.so')
L.orc_pk_sequence.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_void_p]
keys = np.empty((4096, 48), dtype=np.uint8)
L.orc_pk_sequence((12345).to_bytes(32, 'big'), (987654321).to_bytes(32, 'big'), 4096, keys.ctypes.data)
reg = np.tile(keys, ((1 << 21) // 4096, 1)).reshape(-1)
startup, per_iter = sys.argv[1], sys.argv[2]
if startup != "none":
    import torch
    torch.cuda.set_device(0)
    x = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rt = C.CDLL([m.split()[-1] for m in open("/proc/self/maps") if "libcudart" in m][0])
    if startup == "kernel":
        x[:1024].zero_(); torch.cuda.synchronize()
    elif startup == "mm":
        a = torch.empty(1024, 1024, device="cuda", dtype=torch.bfloat16); torch.mm(a, a); torch.cuda.synchronize()
_lib.init(0)
ms = []
for i in range(12):
    if per_iter == "rtmemset":
        rt.cudaMemset(C.c_void_p(x.data_ptr()), 0, C.c_size_t(256 << 20)); rt.cudaDeviceSynchronize()
    elif per_iter == "zero":
        x.zero_(); torch.cuda.synchronize()
    crypto.Registry(reg); ms.append(round(crypto.last_kernel_ms(), 1))
print(startup, per_iter, ms)
