"""Which tensors get re-laid-out (strided aten copies) in one gradient evaluation at B chains?  python tools/exp/find_copies.py 1"""
import collections, os, sys, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nhmc_b200
from nhmc_b200 import unet, kernels
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = unet.create_model(image_size=256, num_channels=128, num_res_blocks=1, channel_mult="", learn_sigma=True, attention_resolutions=16,
                      num_heads=4, num_head_channels=64, use_scale_shift_norm=True, resblock_updown=True, dtype=torch.bfloat16).prepare("cuda")
seen = collections.Counter()
orig_cl = unet._cl
def cl(x):
    if not unet._is_channels_last(x) and not (x.is_contiguous(memory_format=torch.channels_last)):
        fr = traceback.extract_stack(limit=4)[:-1]
        seen[(tuple(x.shape), tuple(x.stride()), " < ".join("%s:%d" % (f.name, f.lineno) for f in reversed(fr)))] += 1
    return orig_cl(x)
unet._cl = cl
x = torch.randn(B, 3, 256, 256, device="cuda", requires_grad=True)
t = torch.full((B,), 500.0, device="cuda")
with kernels.gn_scope(kernels.GnWorkspace(x.device, max(B, 1))):
    for it in range(2):
        out = m(x, t)
        (g,) = torch.autograd.grad(out.float().sum(), x)
    torch.cuda.synchronize()
    seen.clear()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
        out = m(x, t)
        (g,) = torch.autograd.grad(out.float().sum(), x)
        torch.cuda.synchronize()
for k, v in seen.most_common(30):
    print(v, k)
evs = [e for e in prof.events() if e.device_type.name == "CPU"]
def parents(ev):
    out = []
    p_ = ev.cpu_parent
    while p_ is not None:
        out.append(p_.name.replace("autograd::engine::evaluate_function: ", "EF:"))
        p_ = p_.cpu_parent
    return out
agg = collections.Counter(); tm = collections.Counter()
for ev in evs:
    if ev.name == "aten::copy_" and ev.device_time_total > 0:
        key = (str(ev.input_shapes)[:60], " < ".join(parents(ev)[:5]))
        agg[key] += 1; tm[key] += ev.device_time_total
for k, v in sorted(tm.items(), key=lambda kv: -kv[1])[:40]:
    print(round(v, 1), agg[k], k)
