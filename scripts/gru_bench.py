"""Times the one-launch GRU recurrence kernels (gru_seq.cu) against the per-step GEMM + cell path.  GPU only."""
import json
import sys

import torch

sys.path.insert(0, ".")
from ekaid_b200 import functions as F, lib  # noqa: E402
from ekaid_b200.lib import call  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    lib.require_device()
    dev = torch.device("cuda")
    out = []
    for B in (64, 256):
        H, L = 1024, 20
        gi = torch.randn(L * B, 3 * H, device=dev)
        Whh = (torch.randn(3 * H, H, device=dev) / 32).to(torch.bfloat16)
        bhh = torch.randn(3 * H, device=dev) * 0.1
        Hs = torch.empty(L * B, H, device=dev)
        HsT = torch.zeros((L + 1) * B, H, device=dev, dtype=torch.bfloat16)
        gates = torch.empty(L, B, 4 * H, device=dev)
        bar = torch.zeros(4, dtype=torch.int32, device=dev)
        dHs = torch.randn(L * B, H, device=dev)
        dgi = torch.empty(L * B, 3 * H, device=dev)
        dgh = torch.empty_like(dgi)
        dgiT = torch.empty(L * B, 3 * H, device=dev, dtype=torch.bfloat16)
        dghT = torch.empty_like(dgiT)

        def fwd():
            call("gru_seq_fwd", gi.data_ptr(), Whh.data_ptr(), bhh.data_ptr(), B, H, L, Hs.data_ptr(), HsT.data_ptr(),
                 gates.data_ptr(), bar.data_ptr(), 0, None)

        def bwd():
            call("gru_seq_bwd", dHs.data_ptr(), gates.data_ptr(), Hs.data_ptr(), Whh.data_ptr(), B, H, L, dgi.data_ptr(),
                 dgh.data_ptr(), dgiT.data_ptr(), dghT.data_ptr(), bar.data_ptr())

        gh = torch.empty(B, 3 * H, device=dev)

        def fwd_steps():
            call("copy_f32", bhh.data_ptr(), 0, gh.data_ptr(), 3 * H, B, 3 * H)
            for t in range(L):
                F.gemm(HsT[t * B:(t + 1) * B], Whh, B, 3 * H, H, addend=gh, C=gh)
                hp = Hs[(t - 1) * B:t * B] if t > 0 else None
                call("gru_cell_fwd", 1, gi[t * B:(t + 1) * B].data_ptr(), gh.data_ptr(), F.ptr(hp), B, H,
                     Hs[t * B:(t + 1) * B].data_ptr(), HsT[(t + 1) * B:(t + 2) * B].data_ptr(), gates[t].data_ptr(),
                     bhh.data_ptr())

        r = {"B": B, "seq_fwd_us": timeit(fwd), "seq_bwd_us": timeit(bwd)}
        # the same per-step loop inside a CUDA graph (how the step runs it)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fwd_steps()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                fwd_steps()
        r["steps_fwd_graph_us"] = timeit(g.replay)
        out.append(r)
        print(r, flush=True)
    json.dump(out, open("gpurun_out/gru_bench.json", "w"))


if __name__ == "__main__":
    main()
