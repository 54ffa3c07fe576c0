"""K1 (k_g1_validate) and SSZ variant timing: python tools/tune_k1.py   (reads B200_G1_VARIANT / B200_SSZ_MINB_* from env)"""
import ctypes as C, os, sys, hashlib
sys.path.insert(0, '.')
import numpy as np
from ethereum_consensus_b200 import _lib, crypto, ssz, state as S
L = C.CDLL('oracle/liboracle_bls.so')
L.orc_pk_sequence.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_void_p]
nd = int(os.environ.get('TUNE_ND', '4096'))
keys = np.empty((nd, 48), dtype=np.uint8)
L.orc_pk_sequence((12345).to_bytes(32, 'big'), (987654321).to_bytes(32, 'big'), nd, keys.ctypes.data)
reg = np.tile(keys, (1 << 21) // nd, 1) if False else np.tile(keys, ((1 << 21) // nd, 1))
_lib.init(0)
ms = []
for _ in range(4):
    r = crypto.Registry(reg.reshape(-1)); ms.append(crypto.last_kernel_ms())
assert (r.key_codes() == 0).all()
line = f"G1_VARIANT={os.environ.get('B200_G1_VARIANT','0')} k_g1_validate 2^21 keys: {min(ms[1:]):.1f} ms"
if os.environ.get("TUNE_SSZ"):
    b = S.serialize(S.synth_state(1 << 20))
    dev = ssz.DeviceBeaconState(b)
    t = []
    for _ in range(5):
        dev.hash_tree_root(); t.append(float(_lib.load().b200_last_kernel_ms()))
    line += f" | SSZ minb_val={os.environ.get('B200_SSZ_MINB_VALIDATORS','3')} minb_stage={os.environ.get('B200_SSZ_MINB_STAGE','3')}: {min(t):.3f} ms"
print(line, flush=True)
