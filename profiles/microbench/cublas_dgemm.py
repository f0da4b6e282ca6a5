"""cuBLAS DGEMM ceiling on this box (torch.matmul fp64) -- the practical FP64 roofline denominator."""
import json, torch, time
torch.backends.cuda.matmul.allow_tf32 = False
res = {}
for (m, n, k) in [(8192, 8192, 8192), (4096, 4096, 4096), (16384, 2048, 2048), (2048, 277, 100000), (65536, 277, 2052)]:
    a = torch.randn(m, k, device="cuda", dtype=torch.float64)
    b = torch.randn(k, n, device="cuda", dtype=torch.float64)
    for _ in range(2):
        c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # sustained: back to back ~2 s
    nrep = max(3, int(2000 / best))
    e0.record()
    for _ in range(nrep):
        c = a @ b
    e1.record(); torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / nrep
    res[f"{m}x{n}x{k}"] = {"burst_tflops": 2.0 * m * n * k / best / 1e9, "sustained_tflops": 2.0 * m * n * k / sus / 1e9}
    del a, b, c
print(json.dumps(res))
