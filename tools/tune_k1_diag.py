import ctypes as C, os, sys, time
sys.path.insert(0, '.')
import numpy as np
from ethereum_consensus_b200 import _lib, crypto
L = C.CDLL('oracle/liboracle_bls.so')
L.orc_pk_sequence.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_void_p]
keys = np.empty((4096, 48), dtype=np.uint8)
L.orc_pk_sequence((12345).to_bytes(32, 'big'), (987654321).to_bytes(32, 'big'), 4096, keys.ctypes.data)
reg = np.tile(keys, ((1 << 21) // 4096, 1)).reshape(-1)
mode = sys.argv[1]
if mode != "plain":
    import torch
    torch.cuda.set_device(0)
    x = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    y = torch.ones(256 << 20, dtype=torch.uint8, device="cuda")
_lib.init(0)
ms = []
for i in range(14):
    if mode == "torchflush":
        x.zero_(); torch.cuda.synchronize()
    elif mode == "readflush":      # read-only pass: leaves clean lines in L2
        y.view(torch.int64).sum(); torch.cuda.synchronize()
    elif mode == "writeread":      # write flush, then a read pass so the dirty lines are written back
        x.zero_(); y.view(torch.int64).sum(); torch.cuda.synchronize()
    elif mode == "writesleep":
        x.zero_(); torch.cuda.synchronize(); time.sleep(0.05)
    elif mode == "onesflush":      # non-zero fill (rules out zero-line compression effects)
        x.fill_(0x5a); torch.cuda.synchronize()
    elif mode == "small":          # 32 MiB write: smaller than L2
        x[:32 << 20].zero_(); torch.cuda.synchronize()
    crypto.Registry(reg); ms.append(round(crypto.last_kernel_ms(), 1))
print(mode, ms)
