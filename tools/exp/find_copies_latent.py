"""Where do the strided aten copies of the latent U-Net forward come from?  python tools/exp/find_copies_latent.py"""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nhmc_b200
from nhmc_b200 import unet, kernels, latent
from torch.profiler import profile, ProfilerActivity

B = 8
ldm = latent.create_latent_model(dtype=torch.bfloat16)
unet.randomize_zero_tensors(ldm, seed=1234)
ldm.prepare("cuda")
z = torch.randn(B, 3, 64, 64, device="cuda")
t = torch.full((B,), 500, device="cuda", dtype=torch.long)
with kernels.gn_scope(kernels.GnWorkspace(z.device, B)), torch.no_grad():
    for _ in range(2):
        ldm.apply_model(z, t, None)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
        ldm.apply_model(z, t, None)
        torch.cuda.synchronize()
def parents(ev):
    out = []
    p_ = ev.cpu_parent
    while p_ is not None:
        out.append(p_.name)
        p_ = p_.cpu_parent
    return out
agg = collections.Counter(); tm = collections.Counter()
for ev in prof.events():
    if ev.device_type.name != "CPU" or ev.device_time_total <= 0:
        continue
    if ev.name in ("aten::copy_", "aten::cat", "aten::add", "aten::mul", "aten::add_", "aten::mul_", "aten::silu", "aten::cos", "aten::sin", "aten::exp"):
        st = [f.split("/")[-1] for f in (ev.stack or []) if "nhmc_b200" in f or "noise-space" in f][:3]
        key = (ev.name, str(ev.input_shapes)[:70], " < ".join(parents(ev)[:3]), " | ".join(st))
        agg[key] += 1; tm[key] += ev.device_time_total
for k, v in sorted(tm.items(), key=lambda kv: -kv[1])[:30]:
    print(round(v, 1), agg[k], k)
