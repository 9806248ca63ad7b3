"""Per-phase cycle counts of the fused attention kernel (instrumented build generated from csrc/attn.cu by tools/exp/make_attn_prof.py, built into
tools/exp/libvar/libnhmc_attnprof.so): where a softmax warp's time goes.  NHMC_LIB_PATH=<that lib> python tools/exp/attn_prof.py"""
import ctypes as C, math, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nhmc_b200
from nhmc_b200 import kernels, _lib

B, T, nh, ch = 8, 1024, 14, 32
Cc = nh * ch
dev = torch.device("cuda")
qk = torch.randn(B, T, 2 * Cc, device=dev, dtype=torch.bfloat16)
v = torch.randn(B, Cc, T, device=dev, dtype=torch.bfloat16)
out = torch.empty(B, T, Cc, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    kernels.attn_fwd(qk, v, out, nh, 1 / math.sqrt(ch))
torch.cuda.synchronize()
lib = _lib.lib()
host = np.zeros((1024, 12), dtype=np.int64)
lib.nhmc_attn_prof_dump.argtypes = [C.c_void_p]
lib.nhmc_attn_prof_dump(host.ctypes.data)
names = ["total", "loop", "max_calc", "wait_sfull", "tmem_ld", "wait_pvdone", "rescale", "n_rescale", "wait_ofull", "smid", "start"]
h = host[:512]          # lower-half warpgroup of CTAs 0..511 (rows 512.. hold the upper-half warpgroup)
t0 = h[:, 10].min()
order = np.argsort(h[:, 10])
print("CTAs by start time: first 4, middle 4, last 4 (cycles)")
for idx in list(order[:4]) + list(order[254:258]) + list(order[-4:]):
    print({n: int(h[idx, i]) if n != "start" else int(h[idx, i] - t0) for i, n in enumerate(names)})
print("mean over CTAs, lower-half warpgroup:", {n: float(h[:, i].mean()) for i, n in enumerate(names[:9])})
print("mean over CTAs, upper-half warpgroup:", {n: float(host[512:1024, i].mean()) for i, n in enumerate(names[:9])})
