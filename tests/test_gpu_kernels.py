"""Kernel-level checks on a B200, through the C ABI: the tcgen05 GEMM (all operand layouts, tails, split-K,
fused epilogues) and the SIMT fp32 GEMM against torch matmul on the same (bf16-rounded) operands, and the
small kernels against closed-form torch expressions."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    from ekaid_b200 import lib
    lib.require_device()
    return torch.device("cuda:0")


def _mk(shape, dev, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * 0.5).to(dev).to(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("transA,transB", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (304, 200, 136), (1040, 1024, 1024), (64, 3072, 1024),
                                   (208, 6144, 1024), (1024, 600, 416)])
def test_gemm_layouts(dtype, transA, transB, M, N, K):
    from ekaid_b200.functions import gemm
    dev = _dev()
    A = _mk((K, M) if transA else (M, K), dev, 1, dtype)
    B = _mk((K, N) if transB else (N, K), dev, 2, dtype)
    C = torch.full((M, N), float("nan"), device=dev)
    gemm(A, B, M, N, K, transA, transB, C=C, splits=1)
    Af = A.float().t() if transA else A.float()
    Bf = B.float() if transB else B.float().t()
    ref = Af.double() @ Bf.double()
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    assert err < (2e-5 if dtype == torch.float32 else 1e-5), err      # operands identical -> only fp32 accumulation order


@pytest.mark.parametrize("bn", [64, 128, 256, 1128, 1256, 2128, 2256])     # 1000 + width: CTA-pair multicast; 2000 + width: cta_group::2
def test_gemm_tc_tile_widths_and_splitk(bn):
    from ekaid_b200.functions import gemm
    dev = _dev()
    M, N, K = 1024, 1024, 3328
    A = _mk((K, M), dev, 3, torch.bfloat16)
    B = _mk((K, N), dev, 4, torch.bfloat16)
    ref = A.float().t().double() @ B.float().double()
    for splits in (1, 0, 5):
        C = torch.full((M, N), float("nan"), device=dev)
        gemm(A, B, M, N, K, 1, 1, C=C, splits=splits, force_bn=bn)
        err = float((C.double() - ref).abs().max() / ref.abs().max())
        assert err < 1e-5, (bn, splits, err)
    # K-major operands, odd number of M tiles (the second CTA of the last pair has no rows), fused epilogue
    M2, N2, K2 = 128 * 5 + 40, 512, 1024
    A2 = _mk((M2, K2), dev, 5, torch.bfloat16)
    B2 = _mk((N2, K2), dev, 6, torch.bfloat16)
    bias = _mk((N2,), dev, 7, torch.float32)
    ref2 = A2.float().double() @ B2.float().double().t() + bias.double()
    C2 = torch.full((M2, N2), float("nan"), device=dev)
    gemm(A2, B2, M2, N2, K2, 0, 0, bias=bias, C=C2, force_bn=bn)
    assert float((C2.double() - ref2).abs().max() / ref2.abs().max()) < 1e-5, bn


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_epilogue(dtype):
    from ekaid_b200.functions import gemm, ACT_TANH, ACT_RELU
    dev = _dev()
    Bsz, Nn = 3, 52
    M, N, K = 2 * Bsz * Nn, 1024, 512
    A = _mk((M, K), dev, 5, dtype)
    W = _mk((N, 2 * K), dev, 6, dtype)          # pitch 2K: use the left half as a strided view
    bias = _mk((N,), dev, 7, torch.float32)
    add = _mk((M, N), dev, 8, torch.float32)
    rowb = _mk((Bsz, N), dev, 9, torch.float32)
    alt = _mk((N,), dev, 10, torch.float32)
    flags = (torch.arange(M, device=dev) % 7 == 0).to(torch.uint8)
    ref = A.float() @ W[:, :K].float().t() + bias + add
    rows = (torch.arange(M, device=dev) // Nn) % Bsz
    ref = ref + torch.where(flags.bool().unsqueeze(1), alt.unsqueeze(0), rowb[rows])
    C = torch.empty(M, N, device=dev)
    Cb = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if dtype == torch.bfloat16 else None
    gemm(A, W[:, :K], M, N, K, bias=bias, addend=add, rowb=rowb, rowb_div=Nn, rowb_mod=Bsz, rowflag=flags,
         rowb_alt=alt, act=ACT_TANH, C=C, Cb=Cb)
    # fp32 operands take the split-precision tensor-core product: operand planes carry 2^-17 relative, the pre-activation
    # spans +-25, so near tanh's linear range that is ~1e-4 absolute (4e-6 of the pre-activation scale)
    assert float((C - torch.tanh(ref)).abs().max()) < (2e-5 if dtype == torch.bfloat16 else 2e-4)
    assert float((torch.atanh(C.double().clamp(-0.999, 0.999)) - torch.atanh(torch.tanh(ref.double()).clamp(-0.999, 0.999))).abs().max()
                 / ref.abs().max()) < 2e-5
    if Cb is not None:
        assert float((Cb.float() - torch.tanh(ref)).abs().max()) < 1e-2
    # fp32 output split by column (C2 from column 512 on) with the addend on the first 256 columns only, plus dropout-free
    # bias: what the relation backward uses to get both halves of d[v | q] from one product
    Ca = torch.full((M, 512), float("nan"), device=dev)
    Cc = torch.full((M, N - 512), float("nan"), device=dev)
    gemm(A, W[:, :K], M, N, K, bias=bias, addend=add, add_n1=256, C=Ca, C2=Cc, c_n1=512)
    full = A.float() @ W[:, :K].float().t() + bias
    full[:, :256] += add[:, :256]
    tol2 = 1e-5 * float(full.abs().max()) + 2e-5
    assert float((Ca - full[:, :512]).abs().max()) < tol2 and float((Cc - full[:, 512:]).abs().max()) < tol2
    # in-place accumulate (addend aliases C), narrow N with scalar tail path
    C2 = add[:, :100].clone()
    gemm(A, W[:100, :K], M, 100, K, addend=C2, act=ACT_RELU, C=C2)
    ref2 = torch.relu(A.float() @ W[:100, :K].float().t() + add[:, :100])
    assert float((C2 - ref2).abs().max()) < 2e-5


def test_colsum_rowflags_onehot():
    from ekaid_b200.functions import colsum, onehot_adj
    from ekaid_b200.lib import call
    from oracle import ekaid_oracle as O
    from ekaid_b200.synthetic import synthetic_batch
    dev = _dev()
    x = _mk((777, 130), dev, 11, torch.float32)
    sc = _mk((777,), dev, 12, torch.float32)
    assert float((colsum(x, 777, 130) - x.sum(0)).abs().max()) < 1e-3
    assert float((colsum(x, 777, 130, rowscale=sc) - (x * sc[:, None]).sum(0)).abs().max()) < 1e-3
    xb = x.to(torch.bfloat16)
    assert float((colsum(xb, 777, 130) - xb.float().sum(0)).abs().max()) < 1e-3
    # several sums of different widths / operand types in one launch
    from ekaid_b200.functions import colsum_many
    a = torch.randn(700, 1024, device=dev)
    b = torch.randn(700, 2048, device=dev).to(torch.bfloat16)
    c = torch.randn(700, 130, device=dev)
    outs = [torch.empty(1024, device=dev), torch.empty(1024, device=dev), torch.empty(1024, device=dev), torch.empty(130, device=dev)]
    for _ in range(2):          # twice: the ticket counters must be back at zero
        colsum_many([(a, outs[0]), (b[:, :1024], outs[1]), (b[:, 1024:], outs[2]), (c, outs[3])], 700)
        assert float((outs[0] - a.sum(0)).abs().max()) < 1e-3
        assert float((outs[1] - b[:, :1024].float().sum(0)).abs().max()) < 1e-3
        assert float((outs[2] - b[:, 1024:].float().sum(0)).abs().max()) < 1e-3
        assert float((outs[3] - c.sum(0)).abs().max()) < 1e-3
    # calls with different widths share one workspace (ticket counters must survive the partial sums of other widths)
    for n in (1024, 3072, 1, 6144, 33, 3072):
        y = torch.randn(60, n, device=dev)
        assert float((colsum(y, 60, n) - y.sum(0)).abs().max()) < 1e-4, n
    X = _mk((100, 256), dev, 13, torch.float32)
    X[3] = 0
    X[17] = 0
    flags = torch.empty(100, dtype=torch.uint8, device=dev)
    call("row_zero_flags", X.data_ptr(), 100, 256, flags.data_ptr())
    assert flags.nonzero().flatten().tolist() == [3, 17]
    b = synthetic_batch(3, 52, seed=3)
    for idx, L in ((6, 11), (8, 3)):
        got = onehot_adj(b[idx].to(dev), 52, L)
        assert torch.equal(got.cpu(), O.process_matrix(b[idx], 52, L))


def test_spatial_labels_from_boxes_golden_and_oracle():
    """ekaid_spatial_labels == the reference's get_adj_matrix (golden fixture, decision-boundary boxes included) and ==
    the oracle on seeded loader boxes; padding rows / columns are zero; feeds onehot_adj directly."""
    import os
    import numpy as np
    from helpers import GOLDEN
    from oracle import ekaid_oracle as O
    from ekaid_b200.functions import onehot_adj, spatial_labels
    from ekaid_b200.synthetic import synthetic_batch
    dev = _dev()
    z = np.load(os.path.join(GOLDEN, "spatial_labels.npz"))
    for k in range(5):
        bb = torch.from_numpy(z["boxes%d" % k])
        want = torch.from_numpy(z["labels%d" % k].astype(np.int64))
        got = spatial_labels(bb.to(dev))
        assert got.dtype == torch.float64 and got.shape == want.shape
        assert torch.equal(got.cpu().long(), want), k
    b = synthetic_batch(2, 60, seed=21)
    got = spatial_labels(b[10].to(dev), size=100)
    assert torch.equal(got.cpu().long(), O.spatial_adj_matrix(b[10]))
    assert torch.equal(got.cpu(), b[6])                      # the loader's own labels for these boxes
    assert torch.equal(onehot_adj(got, 60, 11).cpu(), O.process_matrix(b[6], 60, 11))
    assert spatial_labels(torch.zeros(0, 52, 4, device=dev)).shape == (0, 100, 100)
    with pytest.raises(ValueError):
        spatial_labels(torch.zeros(2, 52, 3, device=dev))


def test_spatial_labels_full_size_properties():
    """Inference-config size (512 pairs) and the stress graph (126 nodes): the kernel against the loader's vectorised
    rule, plus the size-independent properties of get_adj_matrix: self edges are label 3 (IoU = 1), entry (j,i) is
    reverse_type of (i,j), containment is antisymmetric (1 <-> 2), padding is zero."""
    from ekaid_b200.functions import spatial_labels
    from ekaid_b200.synthetic import _REVERSE, spatial_labels_from_boxes, synthetic_batch
    dev = _dev()
    for B, N in ((512, 52), (64, 126)):
        bb = synthetic_batch(B, N, seed=77)[10]
        got = spatial_labels(bb.to(dev)).cpu().long()
        S = max(100, N)
        assert got.shape == (B, S, S)
        lab = got[:, :N, :N]
        assert torch.equal(lab, spatial_labels_from_boxes(bb))
        assert int(got[:, N:].abs().sum()) == 0 and int(got[:, :, N:].abs().sum()) == 0
        assert bool((lab.diagonal(dim1=1, dim2=2) == 3).all())
        upper = torch.triu(torch.ones(N, N, dtype=torch.bool), 1)
        assert torch.equal(lab.transpose(1, 2)[:, upper], _REVERSE[lab[:, upper]])
        assert int(lab.min()) >= 0 and int(lab.max()) <= 11


def test_adam_matches_torch():
    from ekaid_b200.lib import call
    dev = _dev()
    p = _mk((10000,), dev, 14, torch.float32)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    pw = torch.ones(2, device=dev)
    for step in range(3):
        g = _mk((10000,), dev, 20 + step, torch.float32)
        ref.grad = g.clone()
        opt.step()
        call("adam_advance", pw.data_ptr(), 0.9, 0.999)
        call("adam_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 1e-3, 0.9, 0.999, 1e-8,
             0.0, pw.data_ptr(), 0 if step else 7)      # step 0 with a capped grid (background mode)
    assert float((p - ref.detach()).abs().max()) < 1e-6


@pytest.mark.parametrize("transA,transB", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(6656, 1024, 192), (6656, 1000, 200), (4992, 1024, 128), (6656, 768, 64)])
def test_gemm_tail_halving(transA, transB, M, N, K):
    """Shapes whose last wave of 256-wide tiles is at most half full: those tiles run as half-width units
    (gemm_tc.cu decode()).  Bias + bf16/fp32 outputs so the whole epilogue is exercised on both unit kinds."""
    from ekaid_b200.functions import gemm
    dev = _dev()
    A = _mk((K, M) if transA else (M, K), dev, 31, torch.bfloat16)
    B = _mk((K, N) if transB else (N, K), dev, 32, torch.bfloat16)
    bias = _mk((N,), dev, 33, torch.float32)
    C = torch.full((M, N), float("nan"), device=dev)
    Cb = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    gemm(A, B, M, N, K, transA, transB, bias=bias, C=C, Cb=Cb, force_bn=256)
    Af = A.float().t() if transA else A.float()
    Bf = B.float() if transB else B.float().t()
    ref = Af.double() @ Bf.double() + bias.double()
    assert float((C.double() - ref).abs().max() / ref.abs().max()) < 1e-5
    assert float((Cb.double() - ref).abs().max() / ref.abs().max()) < 1e-2


def test_small_linear_weighted_sums_wn_many():
    """The small fused entry points against their torch expressions."""
    import ctypes
    from ekaid_b200.functions import SmallLinearFn, WeightedSumsFn, WNormFn, WNormManyFn
    dev = _dev()
    x = _mk((64, 1024), dev, 41, torch.float32).requires_grad_(True)
    W = _mk((6, 1024), dev, 42, torch.float32).requires_grad_(True)
    b = _mk((6,), dev, 43, torch.float32).requires_grad_(True)
    y = SmallLinearFn.apply(x, W, b)
    ref = torch.nn.functional.linear(x.detach(), W.detach(), b.detach())
    assert float((y - ref).abs().max()) < 1e-4
    y.sum().backward()
    assert float((W.grad - x.detach().sum(0).expand(6, -1)).abs().max()) < 1e-3 and float((b.grad - 64).abs().max()) < 1e-4

    ts = [_mk((64, 1024), dev, 50 + k, torch.float32).requires_grad_(True) for k in range(3)]
    ts += [_mk((3328,), dev, 60 + k, torch.float32).requires_grad_(True) for k in range(2)]
    ws = [_mk((64, 1024), dev, 70 + k, torch.float32) for k in range(3)] + [None, None]
    coefs = (1.0, 1.0, 0.5, 2.5e-3, -1.25e-3)
    loss = WeightedSumsFn.apply(coefs, ws, *ts)
    ref = sum(c * ((t.detach().double() * w.double()).sum() if w is not None else t.detach().double().sum()) for c, w, t in zip(coefs, ws, ts))
    assert abs(float(loss) - float(ref)) < 1e-3 * (1 + abs(float(ref)))
    loss.backward()
    assert float((ts[2].grad - 0.5 * ws[2]).abs().max()) < 1e-6 and float((ts[4].grad + 1.25e-3).abs().max()) < 1e-8
    l2 = WeightedSumsFn.apply(coefs, ws, *ts)          # deterministic: bitwise equal on a second call
    assert float(l2) == float(loss)

    vs = [_mk(s, dev, 80 + i, torch.float32).requires_grad_(True) for i, s in enumerate([(1024, 2048), (1024, 1024), (4, 64), (1, 11)])]
    gs = [torch.tensor(1.5 + i, device=dev).requires_grad_(True) for i in range(4)]
    outs = WNormManyFn.apply(None, *[t for v, g in zip(vs, gs) for t in (v, g)])
    cot = [_mk(v.shape, dev, 90 + i, torch.float32) for i, v in enumerate(vs)]
    sum((o * c).sum() for o, c in zip(outs, cot)).backward()
    for v, g, o, c in zip(vs, gs, outs, cot):
        v2, g2 = v.detach().clone().requires_grad_(True), g.detach().clone().requires_grad_(True)
        o2 = WNormFn.apply(v2, g2)
        (o2 * c).sum().backward()
        refw = v.detach() * (g.detach() / v.detach().norm())
        assert float((o - refw).abs().max()) < 1e-5 * float(refw.abs().max()) + 1e-7
        assert float((o - o2).abs().max()) == 0.0
        assert float((v.grad - v2.grad).abs().max()) <= 1e-6 * float(v2.grad.abs().max()) + 1e-9
        assert abs(float(g.grad) - float(g2.grad)) <= 1e-5 * abs(float(g2.grad)) + 1e-7


@pytest.mark.parametrize("transA,transB,M,N,K", [(0, 0, 6656, 1024, 1024), (0, 1, 6656, 1024, 4096), (1, 1, 1024, 2048, 6656),
                                                 (0, 0, 1280, 3072, 600), (1, 0, 520, 1024, 1032)])
def test_gemm_fp32_split3_on_tensor_cores(transA, transB, M, N, K, monkeypatch):
    """fp32 operands -> three bf16-plane products in one tcgen05 GEMM (split3_bf16 + K' = 3K): close to the SIMT fp32
    kernel (its test oracle) and to the fp64 product, with the fused epilogue (bias + addend) unchanged."""
    from ekaid_b200 import functions, lib
    from ekaid_b200.functions import gemm
    dev = _dev()
    A = _mk((K, M) if transA else (M, K), dev, 11, torch.float32)
    B = _mk((K, N) if transB else (N, K), dev, 12, torch.float32)
    bias = _mk((N,), dev, 13, torch.float32)
    add = _mk((M, N), dev, 14, torch.float32)
    ref = (A.double().t() if transA else A.double()) @ (B.double() if transB else B.double().t()) + bias.double() + add.double()
    out = {}
    for mode in ("split", "simt"):
        monkeypatch.setattr(functions, "FP32_GEMM", mode)
        C = torch.full((M, N), float("nan"), device=dev)
        before = lib.LAUNCHES
        gemm(A, B, M, N, K, transA, transB, bias=bias, addend=add, C=C)
        assert lib.LAUNCHES - before == (3 if mode == "split" else 1)      # two operand splits + one tcgen05 GEMM
        out[mode] = C
    e_split = float((out["split"].double() - ref).abs().max() / ref.abs().max())
    e_simt = float((out["simt"].double() - ref).abs().max() / ref.abs().max())
    print("split3 err %.2e, simt err %.2e" % (e_split, e_simt))
    # (the tensor core accumulates with truncation: the error grows with the contraction length, 6656 / 4096 here)
    assert e_split < 2e-5 and e_simt < 1e-5
    # the planes reproduce the operand to 2^-16: lo + hi of a value, and the two plane orders (A: lo, hi, hi; B: hi, lo, hi)
    X = _mk((64, 256), dev, 15, torch.float32)
    P = functions._split3(X, 64, 256, 0, 0).float()
    assert torch.equal(P[:, 256:512], P[:, 512:]) and float((P[:, :256] + P[:, 256:512] - X).abs().max() / X.abs().max()) < 2 ** -16
    Q = functions._split3(X, 64, 256, 1, 1).float()
    assert torch.equal(Q[:64], Q[128:]) and torch.equal(Q[:64], P[:, 256:512]) and torch.equal(Q[64:128], P[:, :256])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("transB", [0, 1])
@pytest.mark.parametrize("M,N,K", [(64, 6656, 512), (64, 2048, 1536), (3, 1024, 2048), (64, 512, 6656), (1, 148, 512),
                                   (64, 1536, 2048), (17, 304, 72)])
def test_gemm_skinny_rows(dtype, transB, M, N, K):
    """Products with at most 64 rows (the answer decoder's per-step GEMMs) take the warp-MMA kernel (gemm_skinny.cu): plain
    output (split-K), fused epilogue (bias + addend + activation, fp32 and 16-bit outputs), both weight layouts, both 16-bit
    formats; same results as the tcgen05 kernel and as the fp64 product."""
    from ekaid_b200 import lib
    from ekaid_b200.functions import gemm, ACT_RELU, ACT_SIGMOID
    dev = _dev()
    so = lib.load()
    if transB and N % 8:
        pytest.skip("N-contiguous weights need 16-byte rows")
    A = _mk((M, K), dev, 21, dtype)
    B = _mk((K, N) if transB else (N, K), dev, 22, dtype)
    bias = _mk((N,), dev, 23, torch.float32)
    add = _mk((M, N), dev, 24, torch.float32)
    prod = A.double() @ (B.double() if transB else B.double().t())
    scale = float(prod.abs().max())
    n0 = so.ekaid_gemm_skinny_count()
    C = torch.full((M, N), float("nan"), device=dev)
    gemm(A, B, M, N, K, 0, transB, C=C)                                    # plain: may split K
    assert so.ekaid_gemm_skinny_count() == n0 + 1
    assert float((C.double() - prod).abs().max()) / scale < 1e-5
    Ct = torch.full((M, N), float("nan"), device=dev)
    gemm(A, B, M, N, K, 0, transB, C=Ct, force_bn=64)                      # the tcgen05 kernel on the same operands
    assert so.ekaid_gemm_skinny_count() == n0 + 1
    assert float((C - Ct).abs().max()) / scale < 1e-5
    C2 = torch.full((M, N), float("nan"), device=dev)
    Cb = torch.full((M, N), float("nan"), device=dev, dtype=dtype)
    gemm(A, B, M, N, K, 0, transB, bias=bias, addend=add, act=ACT_RELU, C=C2, Cb=Cb)
    ref = torch.relu(prod + bias.double() + add.double())
    assert so.ekaid_gemm_skinny_count() == n0 + 2
    assert float((C2.double() - ref).abs().max()) / scale < 1e-5
    assert float((Cb.double() - ref).abs().max()) / scale < (1e-2 if dtype == torch.bfloat16 else 2e-3)
    C3 = add.clone()
    gemm(A, B, M, N, K, 0, transB, addend=C3, act=ACT_SIGMOID, C=C3)       # in place: C = act(C + A B)
    # sigmoid' <= 1/4: the 1e-5 * scale bound on the pre-activation carries over with that factor
    assert float((C3.double() - torch.sigmoid(prod + add.double())).abs().max()) < 0.25e-5 * scale + 2e-6


def test_semantic_labels_from_classes_golden():
    """semantic_labels kernel (get_semantic_adj, "feature extraction/combine_dicts.py":106-151): bit-identical to labels made
    by the reference's own function; background detections, classes present in both name lists, asymmetric co-occurrence
    values are all in the fixture; the result feeds the relation encoders as it is (int8 label matrix)."""
    import json
    import os
    import numpy as np
    from helpers import GOLDEN
    from ekaid_b200.functions import onehot_adj, semantic_labels, semantic_tables
    from oracle import ekaid_oracle as O
    dev = _dev()
    z = np.load(os.path.join(GOLDEN, "semantic_labels.npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    tables = semantic_tables(meta["ana"], meta["di"], meta["kg"], meta["name2idx"], z["small_adj"], dev)
    got = semantic_labels(torch.from_numpy(z["classes"]).to(dev), tables)
    assert got.dtype == torch.int8 and got.shape == (6, 100, 100)
    assert torch.equal(got.cpu(), torch.from_numpy(z["labels"]))
    assert torch.equal(onehot_adj(got, 52, 3).cpu(), O.process_matrix(torch.from_numpy(z["labels"]).double(), 52, 3))
