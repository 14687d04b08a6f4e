"""GPU timing of the event -> frame integration kernel (HBM-bound): python tools/time_events.py [B] [events_per_sample]"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from snnquantprune_b200 import input_pipeline as ip
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
T, wh = 20, 128
g = torch.Generator(device="cuda").manual_seed(0)
addrs = torch.stack([torch.randint(0, wh, (B * N,), device="cuda", generator=g, dtype=torch.int32),
                     torch.randint(0, wh, (B * N,), device="cuda", generator=g, dtype=torch.int32),
                     torch.randint(0, 2, (B * N,), device="cuda", generator=g, dtype=torch.int32)], 1).contiguous()
off = (torch.arange(B + 1, device="cuda", dtype=torch.int64) * N)
out = torch.empty((B, T, wh, wh, 2), device="cuda", dtype=torch.uint8)
fn = lambda: ip.events_to_frames(addrs, off, T, wh, 1, out=out, max_events_per_sample=N)
for _ in range(3): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
torch.cuda.synchronize(); e0.record()
for _ in range(n): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
bytes_ = addrs.numel() * 4 + out.numel()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = peaks.get("hbm_gbs", 6650.0)
print(json.dumps({"kernel": "k_events_to_frames", "B": B, "events_per_sample": N, "ms": ms, "samples_per_s": B / ms * 1e3,
                  "events_per_s": B * N / ms * 1e3, "algorithmic_bytes": bytes_, "GBps": bytes_ / ms / 1e6,
                  "hbm_peak_GBps": hbm, "frac": bytes_ / ms / 1e6 / hbm}))
x = (torch.rand((512 * 20, 64 * 64 * 128), device="cuda") < 0.1).to(torch.uint8)
fn2 = lambda: ip.density_stats(x, 512 * 20)
for _ in range(3): fn2()
torch.cuda.synchronize(); e0.record()
for _ in range(n): fn2()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"kernel": "k_slice_nonzeros", "bytes": x.numel(), "ms": ms, "GBps": x.numel() / ms / 1e6, "frac": x.numel() / ms / 1e6 / hbm}))
