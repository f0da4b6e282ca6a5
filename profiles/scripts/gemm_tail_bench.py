"""Time xtd_dgemm on shapes with short last tiles, with and without the tail warp layouts (XTD_GEMM_TAILS).
Run twice: XTD_GEMM_TAILS=0 python ... ; XTD_GEMM_TAILS=1 python ..."""
import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from xtddft_b200 import _lib
lib = _lib.load()
pad = lambda n: (n + 15) // 16 * 16
out = {}
for (m, n, k) in [(1780, 1553, 1553 * 24), (2740, 821, 821 * 48), (137, 821, 65536), (1792, 1536, 1553 * 24), (2770, 1777, 1777 * 16), (1780, 1553 - 17, 1553 * 24)]:
    a = torch.randn((m, pad(k)), dtype=torch.float64, device="cuda")
    b = torch.randn((n, pad(k)), dtype=torch.float64, device="cuda")
    c = torch.zeros((m, pad(n)), dtype=torch.float64, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: _lib.check(lib.xtd_dgemm(s, m, n, k, 1.0, C.c_void_p(a.data_ptr()), a.stride(0), 1, C.c_void_p(b.data_ptr()), b.stride(0), 1,
                                            C.c_void_p(c.data_ptr()), c.stride(0), 0), "gemm")
    for _ in range(2): call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[f"{m}x{n}x{k}"] = dict(ms=round(ms, 3), tflops=round(2.0 * m * n * k / ms / 1e9, 2))
    del a, b, c
print(json.dumps(dict(tails=os.environ.get("XTD_GEMM_TAILS", "1"), **out)))
