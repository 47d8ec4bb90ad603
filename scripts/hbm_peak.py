#!/usr/bin/env python3
"""STREAM-style copy on this box's GPU, the way MEASURED_PEAKS.json was taken
(b.copy_(a) over 1 Gi bf16 elements, best of 10, CUDA events): prints GB/s so that kernel
experiments on different boxes of the pool can be normalised."""
import torch
a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
b = torch.empty_like(a)
a.zero_(); b.copy_(a); torch.cuda.synchronize()
best = 0.0
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * a.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
print("hbm_copy_gbs %.1f" % best)
