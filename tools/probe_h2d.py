import torch, time, numpy as np
x = torch.empty(1 << 28, dtype=torch.float32).pin_memory()   # 1 GiB
d = torch.empty_like(x, device="cuda")
for n in (1 << 22, 1 << 25, 1 << 28):
    best = 0
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d[:n].copy_(x[:n], non_blocking=True); torch.cuda.synchronize()
        best = max(best, n * 4 / (time.perf_counter() - t0) / 1e9)
    print("H2D pinned %4d MB: %.1f GB/s" % (n * 4 >> 20, best))
a = np.empty(1 << 28, dtype=np.float32); a[:] = 1
rc = torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)
t = torch.from_numpy(a)
best = 0
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d.copy_(t, non_blocking=True); torch.cuda.synchronize()
    best = max(best, a.nbytes / (time.perf_counter() - t0) / 1e9)
print("H2D cudaHostRegister'ed numpy 1024 MB: %.1f GB/s (rc=%s, pinned=%s)" % (best, int(rc), t.is_pinned()))
