import ctypes as C, os, sys, json
sys.path.insert(0, '.')
import numpy as np
from ethereum_consensus_b200 import _lib, crypto
L = C.CDLL('oracle/liboracle_bls.so')
L.orc_pk_sequence.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_void_p]
keys = np.empty((4096, 48), dtype=np.uint8)
L.orc_pk_sequence((12345).to_bytes(32, 'big'), (987654321).to_bytes(32, 'big'), 4096, keys.ctypes.data)
reg = np.tile(keys, ((1 << 21) // 4096, 1)).reshape(-1)
_lib.init(0)
def k1():
    ms = []
    for _ in range(3):
        crypto.Registry(reg); ms.append(crypto.last_kernel_ms())
    return min(ms[1:])
print("K1 alone first:", round(k1(), 1))
c = json.load(open('tests/golden/bls_cases.json'))["fast_aggregate_verify"][0]
crypto.fast_aggregate_verify([bytes.fromhex(p) for p in c["pks"]], bytes.fromhex(c["msg"]), bytes.fromhex(c["sig"]))
print("K1 after the big-stack kernels ran once:", round(k1(), 1))
