"""INT8 tensor-core peak of the box, the way SURVEY.md 8(d) asks for it: a dense M = N = K = 8192 INT8 x INT8 -> INT32 GEMM, burst (20
back-to-back calls) and sustained (>= 3 s of back-to-back calls), next to the bf16 figure of the same shape.  The GEMM is the library's
(torch._int_mm -> cuBLASLt; a plain library GEMM used as a yardstick, not part of the product path); the product's own instruction-level
ceiling is tools/microbench/mma_peak.cu.  Writes one JSON line (profiles/r02_int8_gemm_peak.json keeps the committed copy).
Usage: python tools/int8_gemm_peak.py"""
import json
import time

import torch


def bench(fn, seconds=None, iters=20):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    if seconds is None:
        for _ in range(iters):
            fn()
        n = iters
    else:
        t0 = time.time()
        while time.time() - t0 < seconds:
            for _ in range(20):
                fn()
            n += 20
            torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return n, e0.elapsed_time(e1) / 1e3


def main():
    torch.cuda.set_device(0)
    M = N = K = 8192
    a8 = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
    b8 = torch.randint(-127, 128, (N, K), dtype=torch.int8, device="cuda").t()          # column-major B, as cuBLASLt's int8 path wants it
    ab = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    bb = torch.randn(K, N, device="cuda", dtype=torch.bfloat16)
    ops = 2.0 * M * N * K
    out = {"shape": [M, N, K], "device": torch.cuda.get_device_name(0)}
    for name, fn in (("int8", lambda: torch._int_mm(a8, b8)), ("bf16", lambda: torch.mm(ab, bb))):
        n, s = bench(fn)
        out[name + "_burst_tops"] = round(ops * n / s / 1e12, 1)
        n, s = bench(fn, seconds=3.0)
        out[name + "_sustained_tops"] = round(ops * n / s / 1e12, 1)
        out[name + "_sustained_seconds"] = round(s, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
