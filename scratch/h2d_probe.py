import torch, time
n = 2 * 1024**3
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(4): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("H2D pinned GB/s", 4*n/dt/1e9)
t=time.perf_counter(); x=torch.empty(4*1024**3, dtype=torch.uint8, device="cuda"); torch.cuda.synchronize(); print("alloc 4GB ms", (time.perf_counter()-t)*1e3)
t=time.perf_counter(); del x; torch.cuda.empty_cache(); torch.cuda.synchronize(); print("free 4GB ms", (time.perf_counter()-t)*1e3)
