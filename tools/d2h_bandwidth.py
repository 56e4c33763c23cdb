import torch, time
x = torch.empty((256, 1080, 1920), dtype=torch.int32, device="cuda")
h = torch.empty((256, 1080, 1920), dtype=torch.int32).pin_memory()
for n in (1, 16, 256):
    torch.cuda.synchronize()
    for _ in range(2):
        h[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        h[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 3
    print(n, "frames", n * 1080 * 1920 * 4 / dt / 1e9, "GB/s D2H")
t = time.perf_counter()
x.copy_(h, non_blocking=True); torch.cuda.synchronize()
print("H2D", 256 * 1080 * 1920 * 4 / (time.perf_counter() - t) / 1e9)
