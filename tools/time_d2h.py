"""Where the end-to-end overhead of the C4 bench goes: device->pinned-host copy rate of one
307 MB ETC and the device-side re-layout (EnergyHistogram.dense) before it."""
import json
import time

import torch

n, t, pad = 19200, 2000, 64
dev = torch.device("cuda:0")
data = torch.rand((n, pad + 2048), dtype=torch.float64, device=dev)
rank = torch.randperm(n, device=dev)
host = torch.empty((n, 1, 1, t), dtype=torch.float64).pin_memory()
out = {}
for name, fn in (
        ("relayout_ms", lambda: data.view(1, n, 1, -1)[:, :, :, pad:pad + t].index_select(1, rank)
         .permute(1, 2, 0, 3).contiguous()),
        ("d2h_contiguous_ms", lambda: host.copy_(dense)),
        ("relayout_plus_d2h_ms", lambda: host.copy_(
            data.view(1, n, 1, -1)[:, :, :, pad:pad + t].index_select(1, rank).permute(1, 2, 0, 3)))):
    dense = data.view(1, n, 1, -1)[:, :, :, pad:pad + t].index_select(1, rank).permute(
        1, 2, 0, 3).contiguous()
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    out[name] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
out["bytes"] = host.numel() * 8
out["d2h_gb_s"] = round(out["bytes"] / out["d2h_contiguous_ms"] / 1e6, 2)
print(json.dumps(out))
