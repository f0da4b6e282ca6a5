"""Emulated-FP64 (INT8 tensor core, Ozaki splitting) contraction at the config-5 exchange shape against the DMMA GEMM:
C[2770, 1777] += sum_P U[P][2770, 1777] Lvv[P][1777, 1777]^T.  Prints one JSON line per slice count with the
FP64-equivalent TFLOP/s of the int8 kernel, the slicing times and the max relative error against the DMMA result."""
import ctypes as C
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from xtddft_b200 import _lib  # noqa: E402


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    m, n, k = 2770, 1777, 1777
    lib = _lib.load()
    ld = (k + 15) // 16 * 16
    g = torch.Generator(device="cuda").manual_seed(1)
    # DF-like magnitudes: a smooth decay over the aux index and a few orders of magnitude inside each row
    decay = torch.logspace(0, -3, nq, device="cuda", dtype=torch.float64)[:, None, None]
    a = torch.zeros((nq, m, ld), dtype=torch.float64, device="cuda")
    b = torch.zeros((nq, n, ld), dtype=torch.float64, device="cuda")
    a[:, :, :k] = torch.randn((nq, m, k), generator=g, device="cuda", dtype=torch.float64) * decay * \
        torch.pow(10.0, -3 * torch.rand((nq, m, k), generator=g, device="cuda", dtype=torch.float64))
    b[:, :, :k] = torch.randn((nq, n, k), generator=g, device="cuda", dtype=torch.float64) * decay * \
        torch.pow(10.0, -3 * torch.rand((nq, n, k), generator=g, device="cuda", dtype=torch.float64))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flops = 2.0 * m * n * k * nq
    # DMMA reference (the engine's GEMM) with timing
    ref = torch.zeros((m, 1792), dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(2):
        ref.zero_()
        e0.record()
        for q in range(nq):
            _lib.check(lib.xtd_dgemm(st, m, n, k, 1.0, C.c_void_p(a[q].data_ptr()), ld, 1, C.c_void_p(b[q].data_ptr()), ld, 1,
                                     C.c_void_p(ref.data_ptr()), 1792, 1), "dgemm")
        e1.record()
        torch.cuda.synchronize()
    dmma_ms = e0.elapsed_time(e1)
    print(json.dumps(dict(kind="dmma", nq=nq, ms=dmma_ms, tflops=flops / dmma_ms / 1e9)), flush=True)
    scale = ref[:, :n].abs().max().item()
    for slices in (4, 5, 6, 7, 8):
        for group in (0,):
            c = torch.zeros((m, 1792), dtype=torch.float64, device="cuda")
            ms = (C.c_double * 3)()
            for it in range(2):
                _lib.check(lib.xtd_ozaki_gemm(st, m, n, k, nq, slices, group, C.c_void_p(a.data_ptr()), ld, m * ld, C.c_void_p(b.data_ptr()), ld,
                                              n * ld, C.c_void_p(c.data_ptr()), 1792, 1.0, 0, ms), "ozaki")
            err = (c[:, :n] - ref[:, :n]).abs().max().item() / scale
            pairs = slices * (slices + 1) // 2
            print(json.dumps(dict(kind="ozaki", slices=slices, group=group, nq=nq, slice_a_ms=ms[0], slice_b_ms=ms[1], gemm_ms=ms[2],
                                  fp64_equiv_tflops=flops / ms[2] / 1e9, int8_tops=flops * pairs / ms[2] / 1e9, max_rel_err=err,
                                  speedup_vs_dmma=dmma_ms / ms[2])), flush=True)


if __name__ == "__main__":
    main()
