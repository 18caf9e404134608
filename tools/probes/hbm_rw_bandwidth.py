import torch
x = torch.empty(8 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(8 << 30, dtype=torch.uint8, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_()); print(f"write-only (zero_ 8 GiB): {8.59/ms:.2f} TB/s")
ms = t(lambda: y.copy_(x)); print(f"copy 8 GiB: {2*8.59/ms:.2f} TB/s (read+write)")
ms = t(lambda: x.view(torch.float32).sum()); print(f"read-only (sum 8 GiB): {8.59/ms:.2f} TB/s")
