import torch, time
x = torch.empty((64, 1080, 1920, 3), dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d.copy_(x, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"H2D pinned {x.numel()/1e6:.0f} MB: {ms:.2f} ms  {x.numel()/ms/1e6:.1f} GB/s")
