"""Which property of a foreign CUDA module slows K1?  usage: tune_k1_diag3.py <kind> [after]
kind: -1 none, 0 plain, 1 assert, 2 printf, 3 malloc, 4 32KiB-lmem, 5 200KiB-smem, 6 plain on a new stream,
      10 torch fill kernel.  'after' = launch the foreign kernel after K1 has already run twice."""
import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np
from ethereum_consensus_b200 import _lib, crypto
L = C.CDLL('oracle/liboracle_bls.so')
L.orc_pk_sequence.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_void_p]
keys = np.empty((4096, 48), dtype=np.uint8)
L.orc_pk_sequence((12345).to_bytes(32, 'big'), (987654321).to_bytes(32, 'big'), 4096, keys.ctypes.data)
reg = np.tile(keys, ((1 << 21) // 4096, 1)).reshape(-1)
kind = int(sys.argv[1]); after = len(sys.argv) > 2
F = C.CDLL('tools/foreign/libforeign.so')
def foreign():
    if kind == 10:
        import torch
        x = torch.empty(1024, dtype=torch.uint8, device="cuda"); x.zero_(); torch.cuda.synchronize()
    elif kind >= 0:
        rc = F.foreign_run(kind); assert rc == 0, rc
_lib.init(0)
ms = []
if not after: foreign()
for i in range(8):
    if after and i == 3: foreign()
    crypto.Registry(reg); ms.append(round(crypto.last_kernel_ms(), 1))
print(kind, "after" if after else "before", ms)
