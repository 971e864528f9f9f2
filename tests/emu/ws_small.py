import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import lewton_b200 as L
from lewton_b200 import _cabi as cabi
from helpers import make_setup
ctx = L.Context(0)
su = make_setup(ctx, 2, 8, 11)
S, P = 3, 5
pw = [L.PreviousWindowRight(su) for _ in range(S)]
spec = (np.random.default_rng(0).standard_normal((S, P, 2, 1024)) * 0.05).astype(np.float32)
stride = P * 1024
chains = [L.ChainSpec(pw[s], np.ones(P, np.uint8), coeff_offset=s*P*2048, out_offset=s*2*stride, out_stride=stride) for s in range(S)]
pcm = np.zeros((S, 2, stride), np.float32)
try:
    L.decode_chains(ctx, chains, cabi.ENTRY_SPECTRUM, cabi.MEM_HOST, spec, pcm, cabi.OUT_F32_PLANAR)
    print("ok", [c.n_samples for c in chains], float(np.abs(pcm).sum()))
except Exception as e:
    print("ERR", e)
