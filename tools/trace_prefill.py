"""CUPTI kernel timeline of one context-encoding call (Llama-3.1-8B shapes, random weights, 8 layers): where does TTFT go?"""
import collections
import sys
import torch
sys.path.insert(0, ".")
from bench import LLAMA31_8B, build_app  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

L = 8
app = build_app(dict(LLAMA31_8B, num_hidden_layers=L), 1, 2, 272, 128, async_mode=False)
ids = torch.randint(0, 100, (2, 128))
mask = torch.ones_like(ids, dtype=torch.int32)
for _ in range(3):
    app.reset()
    app(ids, attention_mask=mask)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    app.reset()
    app(ids, attention_mask=mask)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
tot = 0.0
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
span = (evs[-1].time_range.end - evs[0].time_range.start)
for e in evs:
    d = e.time_range.end - e.time_range.start
    n = e.name[:70]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += d
    tot += d
print(f"layers={L} kernels={len(evs)} busy={tot:.0f} us span={span:.0f} us")
for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{d:9.1f} us  x{c:4d}  {d / c:8.1f} us/call  {n}")
