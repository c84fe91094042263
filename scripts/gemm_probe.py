"""Where does the time of a tcgen05 GEMM launch go?  (measurement tooling; B200 only)

1. mixed 16-bit operand formats (fp16 x bf16 in one MMA) and the 16-bit output routing: correctness vs fp32 matmul
2. role ablations through ekaid_gemm_debug (no epilogue / no TMA / no MMA ...): CUDA events over back-to-back launches
3. per-CTA globaltimer stamps of one launch: set-up, first data, per-tile MMA issue / accumulator ready / epilogue done

    python scripts/gemm_probe.py [out.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ekaid_b200 import lib  # noqa: E402
from ekaid_b200.functions import gemm  # noqa: E402

lib.require_device()
dev = torch.device("cuda:0")
so = lib.load()
OUT = {"mixed": [], "ablate": [], "timeline": []}


def check_mixed():
    torch.manual_seed(0)
    M, N, K = 384, 512, 320
    a32 = torch.randn(M, K, device=dev)
    b32 = torch.randn(N, K, device=dev)
    # (a_format != b_format in one MMA was tried: the B200 raises an illegal-instruction error, so both operands of a
    # GEMM share one 16-bit format)
    for ta in (torch.bfloat16, torch.float16):
        for tb in (ta,):
            A, B = a32.to(ta), b32.to(tb)
            ref = A.float() @ B.float().t()
            C = torch.empty(M, N, device=dev)
            Cb = torch.empty(M, N, device=dev, dtype=torch.float16)
            Cb2 = torch.empty(M, N - 256, device=dev, dtype=torch.bfloat16)
            gemm(A, B, M, N, K, C=C, Cb=Cb, Cb2=Cb2, cb_n1=256, cb2_n0=256)
            torch.cuda.synchronize()
            e = float((C - ref).abs().max() / ref.abs().max())
            e16 = float((Cb[:, :256].float() - ref[:, :256]).abs().max() / ref.abs().max())
            e2 = float((Cb2.float() - ref[:, 256:]).abs().max() / ref.abs().max())
            OUT["mixed"].append({"A": str(ta), "B": str(tb), "err_f32": e, "err_cb_fp16": e16, "err_cb2_bf16": e2})
            print("mixed", ta, tb, "%.2e %.2e %.2e" % (e, e16, e2), flush=True)
    # transposed operands with mixed formats (wgrad: dY^T bf16, X fp16)
    A = torch.randn(K, M, device=dev).to(torch.bfloat16)
    A = A.to(torch.float16)
    B = torch.randn(K, N, device=dev).to(torch.float16)
    ref = A.float().t() @ B.float()
    C = torch.empty(M, N, device=dev)
    gemm(A, B, M, N, K, 1, 1, C=C, splits=1)
    torch.cuda.synchronize()
    e = float((C - ref).abs().max() / ref.abs().max())
    OUT["mixed"].append({"case": "wgrad fp16^T x fp16", "err_f32": e})
    print("mixed wgrad", "%.2e" % e, flush=True)
    # saturation: fp16 output clamps instead of overflowing
    A = torch.full((128, 64), 300.0, device=dev).to(torch.bfloat16)
    B = torch.full((128, 64), 300.0, device=dev).to(torch.bfloat16)
    Cb = torch.empty(128, 128, device=dev, dtype=torch.float16)
    gemm(A, B, 128, 128, 64, Cb=Cb)
    torch.cuda.synchronize()
    OUT["mixed"].append({"case": "saturation", "max": float(Cb.float().max()), "finite": bool(torch.isfinite(Cb.float()).all())})
    print("saturation", float(Cb.float().max()), flush=True)


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
    return e0.elapsed_time(e1) / iters * 1e3


SHAPES = [  # M, N, K, tA, tB, out, extras
    (6656, 1024, 1024, 0, 0, "bf16", ""),
    (6656, 1024, 1024, 0, 0, "bf16", "bias"),
    (6656, 1024, 1024, 0, 0, "bf16+f32", "bias+rowb"),
    (6656, 1024, 1024, 0, 1, "f32", "addend"),
    (6656, 1024, 1024, 0, 1, "f32", "addend+drop"),
    (6656, 4096, 1024, 0, 0, "bf16", "bias"),
    (6656, 1024, 4096, 0, 1, "bf16", ""),
    (1024, 1024, 6656, 1, 1, "f32", ""),
    (4096, 1024, 6656, 1, 1, "f32", ""),
    (6656, 2048, 2048, 0, 0, "f32", "bias"),
    (26624, 4096, 1024, 0, 0, "bf16", "bias"),
]
ABL = [("normal", 0), ("nostore", 16), ("noepi", 256), ("notma", 32), ("nomma", 64), ("mma_only", 32 | 256),
       ("tma_only", 64 | 256), ("no_gst", 512), ("no_stage", 1024), ("no_gst_no_stage", 1536), ("epi_only", 32 | 64),
       ("epi_only_no_gst", 32 | 64 | 512)]
VARIANTS = [("auto", 0), ("bn128", 128), ("bn256", 256), ("cg256", 2256)]


def make(M, N, K, ta, tb, out, extras):
    sets = []
    for s in range(3):
        A = (torch.randn((K, M) if ta else (M, K), device=dev) * 0.1).to(torch.bfloat16)
        B = (torch.randn((K, N) if tb else (N, K), device=dev) * 0.1).to(torch.bfloat16)
        kw = {}
        if "f32" in out:
            kw["C"] = torch.empty(M, N, device=dev)
        if "bf16" in out:
            kw["Cb"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        if "bias" in extras:
            kw["bias"] = torch.randn(N, device=dev)
        if "rowb" in extras:
            kw["rowb"] = torch.randn(64, N, device=dev)
            kw["rowb_div"], kw["rowb_mod"] = 52, 64
            kw["rowflag"] = torch.zeros(M, dtype=torch.uint8, device=dev)
            kw["rowb_alt"] = torch.randn(N, device=dev)
        if "addend" in extras:
            kw["addend"] = torch.randn(M, N, device=dev)
        if "drop" in extras:
            from ekaid_b200.functions import rng_state
            kw["drop"] = (rng_state(dev).data_ptr(), 7, 0.2)
        sets.append((A, B, kw))
    return sets


def main():
    check_mixed()
    ts = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
    for shp in SHAPES:
        M, N, K, ta, tb, out, extras = shp
        sets = make(*shp)
        flops = 2.0 * M * N * K

        def run(i, bn=0):
            A, B, kw = sets[i % 3]
            gemm(A, B, M, N, K, ta, tb, force_bn=bn, **kw)

        def ref(i):
            A, B, kw = sets[i % 3]
            torch.matmul(A.t() if ta else A, B if tb else B.t())

        row = {"shape": list(shp)}
        us = timeit(ref)
        row["cublas"] = [round(us, 2), round(flops / us / 1e6)]
        for vname, bn in VARIANTS:
            try:
                so.ekaid_gemm_debug(0, None)
                us = timeit(lambda i: run(i, bn))
                row[vname] = [round(us, 2), round(flops / us / 1e6)]
            except Exception as e:       # noqa: BLE001
                row[vname] = str(e)[:80]
        for vname, bn in (("bn256", 256),):
            for aname, fl in ABL[1:]:
                try:
                    so.ekaid_gemm_debug(fl, None)
                    us = timeit(lambda i: run(i, bn))
                    row[vname + "/" + aname] = round(us, 2)
                except Exception as e:   # noqa: BLE001
                    row[vname + "/" + aname] = str(e)[:80]
            so.ekaid_gemm_debug(0, None)
        OUT["ablate"].append(row)
        print(json.dumps(row), flush=True)
        # timeline of one launch (after warm-up, back to back with a preceding launch of the same kernel)
        for vname, bn in (("bn256", 256),):
            try:
                for i in range(3):
                    run(i, bn)
                ts.zero_()
                torch.cuda.synchronize()
                run(0, bn)
                so.ekaid_gemm_debug(0, ts.data_ptr())
                run(1, bn)
                so.ekaid_gemm_debug(0, None)
                run(2, bn)
                torch.cuda.synchronize()
                t = ts.view(148, 32).cpu().double()
                used = t[:, 0] > 0
                t = t[used]
                t0 = float(t[:, 0].min())

                def stat(col, base=None):
                    v = t[:, col]
                    ok = v > 0
                    if ok.sum() == 0:
                        return None
                    v = v[ok] - (t0 if base is None else t[ok][:, base])
                    v = v / 1e3
                    return [round(float(v.min()), 2), round(float(v.median()), 2), round(float(v.max()), 2), int(ok.sum())]

                tl = {"shape": list(shp), "variant": vname, "ctas": int(used.sum()),
                      "entry": stat(0), "after_pdl_wait": stat(1), "first_tma_issue": stat(2), "first_data": stat(3),
                      "first_data_minus_wait": stat(3, 1),
                      "mma_issued_unit": [stat(4 + 2 * i) for i in range(4)],
                      "acc_ready_unit": [stat(16 + i) for i in range(4)],
                      "epi_done_unit": [stat(5 + 2 * i) for i in range(4)],
                      "epi_minus_acc_unit0": None, "end": stat(31),
                      "chunk0": [stat(22 + k, 16) for k in range(4)], "chunk2": [stat(26 + k, 16) for k in range(4)]}
                a, e = t[:, 16], t[:, 5]
                ok = (a > 0) & (e > 0)
                if ok.sum():
                    d = (e[ok] - a[ok]) / 1e3
                    tl["epi_minus_acc_unit0"] = [round(float(d.min()), 2), round(float(d.median()), 2), round(float(d.max()), 2)]
                OUT["timeline"].append(tl)
                print(json.dumps(tl), flush=True)
            except Exception as e:       # noqa: BLE001
                print("timeline failed", vname, e, flush=True)
                so.ekaid_gemm_debug(0, None)
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/gemm_probe.json"
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    json.dump(OUT, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
