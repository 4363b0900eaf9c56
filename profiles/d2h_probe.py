import torch, time
x = torch.empty(3840*2160*4, dtype=torch.uint8, device='cuda')
h = torch.empty(3840*2160*4, dtype=torch.uint8).pin_memory()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): h.copy_(x, non_blocking=True)
    s.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(20): h.copy_(x, non_blocking=True)
    b.record(s); s.synchronize()
ms = a.elapsed_time(b) / 20
print(f"D2H 33 MB pinned: {ms:.3f} ms = {x.numel()/ms/1e6:.1f} GB/s")
hp = torch.empty(20000*3, dtype=torch.float32).pin_memory(); d = torch.empty(20000*3, dtype=torch.float32, device='cuda')
