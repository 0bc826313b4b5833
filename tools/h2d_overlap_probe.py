"""Does a pinned host->device copy slow down when kernels run on another stream?  (platform probe, run under gpurun)"""
import json
import torch

n = 98 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
cs, ks = torch.cuda.Stream(), torch.cuda.Stream()


def copy_ms(busy: str):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if busy == "matmul":
        with torch.cuda.stream(ks):
            for _ in range(40):
                a @ a
    elif busy == "memset":
        with torch.cuda.stream(ks):
            for _ in range(200):
                d2.fill_(1)
    with torch.cuda.stream(cs):
        e0.record()
        d.copy_(h, non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


d2 = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for busy in ("idle", "matmul", "memset", "idle"):
    out.setdefault(busy, []).append(round(min(copy_ms(busy) for _ in range(3)), 3))
print(json.dumps({"h2d_98MB_ms": out, "GBps_idle": round(n / out["idle"][0] / 1e6, 1)}))
