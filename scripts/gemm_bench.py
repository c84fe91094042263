"""GEMM micro-benchmark on the shapes of the training step: our tcgen05 kernel (auto / forced tile widths) next to
torch.matmul (cuBLAS) on the same operands.  CUDA events, 30 iterations after 5 warm-ups, rotating 4 operand sets."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ekaid_b200 import lib
from ekaid_b200.functions import gemm

lib.require_device()
dev = torch.device("cuda:0")
SHAPES = [  # M, N, K, transA, transB, out (f32|bf16)
    (6656, 1024, 1024, 0, 0, "bf16"), (6656, 1024, 1024, 0, 1, "f32"), (6656, 4096, 1024, 0, 0, "bf16"),
    (6656, 1024, 4096, 0, 1, "f32"), (1024, 1024, 6656, 1, 1, "f32"), (4096, 1024, 6656, 1, 1, "f32"),
    (1024, 2048, 6656, 1, 1, "f32"), (6656, 2048, 2048, 0, 0, "f32"), (6656, 1024, 2048, 0, 0, "bf16"),
    (26624, 1024, 1024, 0, 0, "bf16"), (26624, 4096, 1024, 0, 0, "bf16"), (8192, 8192, 8192, 0, 0, "bf16"),
]

def timeit(fn, iters=30):
    for _ in range(5):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us

rows = []
for (M, N, K, ta, tb, out) in SHAPES:
    sets = []
    for s in range(4):
        A = (torch.randn((K, M) if ta else (M, K), device=dev) * 0.1).to(torch.bfloat16)
        B = (torch.randn((K, N) if tb else (N, K), device=dev) * 0.1).to(torch.bfloat16)
        C = torch.empty(M, N, device=dev, dtype=torch.float32 if out == "f32" else torch.bfloat16)
        sets.append((A, B, C))
    flops = 2.0 * M * N * K
    res = {"shape": (M, N, K, ta, tb, out)}
    for name, bn in (("auto", 0), ("bn64", 64), ("bn128", 128), ("bn256", 256), ("cl256", 1256), ("cg256", 2256), ("cg128", 2128)):
        def run(i, bn=bn):
            A, B, C = sets[i % 4]
            if out == "f32":
                gemm(A, B, M, N, K, ta, tb, C=C, force_bn=bn)
            else:
                gemm(A, B, M, N, K, ta, tb, Cb=C, force_bn=bn)
        try:
            us = timeit(run)
            res[name] = (round(us, 1), round(flops / us / 1e6))
        except Exception as e:
            res[name] = str(e)[:60]
    def ref(i):
        A, B, C = sets[i % 4]
        a = A.t() if ta else A
        b = B if tb else B.t()
        torch.matmul(a, b)
    us = timeit(ref)
    res["cublas"] = (round(us, 1), round(flops / us / 1e6))
    rows.append(res)
    print(res, flush=True)
json.dump(rows, open("gpurun_out/gemm_bench.json", "w"))
