"""HBM copy bandwidth vs working-set size (torch copy_, CUDA events): is the large-footprint slowdown of the
scatter kernel a property of the memory system?"""
import torch
for gb in (1, 2, 4, 8, 16, 32, 64):
    n = gb * (1 << 30) // 4 // 2
    a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
    b = torch.empty_like(a)
    for _ in range(2):
        b.copy_(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("footprint %3d GB: %.3f ms  %.1f GB/s (read+write)" % (gb, best, 2 * n * 4 / best / 1e6))
    del a, b
    torch.cuda.empty_cache()
