"""Train-mode (dropout) path on a B200.  Bitwise RNG parity with ATen is impossible, so:
  * the dropout RNG is checked statistically and for determinism / site independence;
  * the train-mode code path (K = 2048 self_weights GEMM on the dropped concat, separate query/key/Z GEMMs, masked
    backward) is run with every probability forced to 0 and must reproduce the eval-mode oracle parity exactly;
  * with the real probabilities, the analytic gradients are checked against central finite differences of the SAME
    masked function (the seed is frozen), which validates that forward and backward regenerate identical masks."""
import contextlib
import io

import pytest
import torch

from helpers import OUT_NAMES, case_inputs, load_case, loss_weights, oracle_forward, rel_err
from test_gpu_parity import build_model, to_dev, grad_err, param_err, GTOL, TOL

pytestmark = pytest.mark.gpu


def _dev():
    from ekaid_b200 import lib
    lib.require_device()
    return torch.device("cuda:0")


def test_dropout_rng_statistics_and_determinism():
    from ekaid_b200.functions import drop_combine, rng_advance, rng_state
    dev = _dev()
    seed = rng_state(dev).data_ptr()
    x = torch.ones(2048, 1024, device=dev)
    outs = {}
    for site, p in ((1, 0.2), (2, 0.2), (3, 0.5)):
        y = torch.empty_like(x)
        drop_combine([x], [(seed, site, p)], 2048, 1024, outf=y)
        zero = float((y == 0).float().mean())
        assert abs(zero - p) < 4e-3, (site, p, zero)
        kept = y[y != 0]
        assert float((kept - 1.0 / (1.0 - p)).abs().max()) < 1e-6
        assert abs(float(y.mean()) - 1.0) < 1e-2
        outs[site] = y
    assert float(((outs[1] == 0) == (outs[2] == 0)).float().mean()) < 0.75      # independent sites (0.68 expected)
    again = torch.empty_like(x)
    drop_combine([x], [(seed, 1, 0.2)], 2048, 1024, outf=again)
    assert torch.equal(again, outs[1])                                           # same (seed, site, idx) -> same mask
    xb = x.to(torch.bfloat16)
    yb = torch.empty(2048, 1024, device=dev, dtype=torch.bfloat16)
    drop_combine([xb], [(seed, 1, 0.2)], 2048, 1024, outT=yb)
    assert torch.equal(yb == 0, outs[1] == 0)                                    # mask independent of storage type
    rng_advance(dev)
    drop_combine([x], [(seed, 1, 0.2)], 2048, 1024, outf=again)
    assert not torch.equal(again, outs[1])
    # row/col structure: no correlation between neighbouring elements
    m = (outs[1] != 0).float()
    assert abs(float((m[:, 1:] * m[:, :-1]).mean()) - 0.64) < 5e-3
    assert abs(float((m[1:] * m[:-1]).mean()) - 0.64) < 5e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_train_path_with_p0_matches_oracle(precision):
    from ekaid_b200 import functions
    from oracle import ekaid_oracle as O
    dev = _dev()
    z, meta = load_case("c1_b3_n52_all_grads")
    sd, inp, _ = case_inputs(meta)
    m = build_model(meta, sd, precision, dev)
    m.train()
    m.dropout_override = 0.0
    functions.DEBUG_SINK = []
    try:
        outs = m(*to_dev(inp, dev), setting="mode2", graph=meta["graph"])
        masks = functions.DEBUG_SINK
    finally:
        functions.DEBUG_SINK = None
    with torch.no_grad():
        ref = oracle_forward(sd, inp, meta)
    for k, o, r in zip(OUT_NAMES[1:5], outs[1:5], ref[1:5]):
        assert rel_err(o, r) < TOL[precision], (k, rel_err(o, r))
    ws = loss_weights(outs)
    sum((o * w.to(dev)).sum() for o, w in zip(outs[1:], ws)).backward()
    sdg = {k: v.clone().requires_grad_(v.is_floating_point() and k != "w_emb.emb_.weight") for k, v in sd.items()}
    cd = m.cfg.model.change_detector
    ro = O.change_detector_forward(sdg, *inp, graph=meta["graph"], num_heads=cd.att_head, nongt_dim=meta["nongt_dim"],
                                   relu_masks=masks)
    sum((o * w).sum() for o, w in zip(ro[1:], ws)).backward()
    bad = []
    for k, p in m.named_parameters():
        g_ref = sdg[k].grad
        if g_ref is None or float(g_ref.abs().max()) < 1e-4:
            continue
        e = param_err(precision, k, p.grad, {n: t.grad for n, t in sdg.items() if t.grad is not None}, sd)
        if e > GTOL[precision]:
            bad.append((k, e))
    assert not bad, bad


def test_train_mode_gradients_match_finite_differences():
    """fp32 path, real dropout probabilities, frozen seed: <analytic grad, d> == (L(w + e d) - L(w - e d)) / 2e."""
    dev = _dev()
    z, meta = load_case("c0_b2_n52_all")
    sd, inp, _ = case_inputs(meta)
    m = build_model(meta, sd, "fp32", dev)
    m.train()
    torch.manual_seed(1238)                       # the device seed is re-derived from torch's: same masks in every run
    m.freeze_dropout_seed = True
    dinp = to_dev(inp, dev)
    gw = torch.Generator().manual_seed(7)
    cot = None

    def loss():
        nonlocal cot
        outs = m(*dinp, setting="mode2", graph="all")
        if cot is None:
            cot = [torch.randn(o.shape, generator=gw).to(dev) for o in outs[1:]]
        return sum((o.double() * c.double()).sum() for o, c in zip(outs[1:], cot))

    l0 = loss()
    l0.backward()
    with torch.no_grad():
        l0b = loss()
    assert float((l0 - l0b).abs()) < 1e-6 * max(1.0, float(l0.abs()))         # frozen seed -> same masks
    eval_out = m.eval()(*dinp)[3]
    m.train()
    assert rel_err(m(*dinp)[3], eval_out) > 1e-2                               # dropout is really active
    names = ["img.weight", "semantic_relation.explicit_relation.self_weights.main.1.weight_v",
             "spatial_relation.explicit_relation.neighbor_net.1.query.main.1.weight_v",
             "spatial_relation.explicit_relation.neighbor_net.1.key.main.1.weight_v",
             "imp_relation.implicit_relation.neighbor_net.1.pair_pos_fc1.main.1.weight_v",
             "imp_relation.implicit_relation.neighbor_net.1.linear_out_2.weight",
             "spatial_relation.explicit_relation.bias.main.0.weight_v",
             "context1.weight", "gate2.weight", "embed.0.weight", "att.weight",
             "q_emb.rnn.weight_hh_l0", "w_emb.emb.weight", "q_att.W1_self_att_q.main.1.weight_v"]
    params = dict(m.named_parameters())
    report = []
    for k in names:
        p = params[k]
        g = p.grad.detach().clone()
        d = g / (g.norm() + 1e-30)
        analytic = float((g.double() * d.double()).sum())
        eps = max(2e-4 * float(p.detach().norm()), 5e-4)     # fp32 loss rounding vs curvature (ReLU / clamp kinks)
        with torch.no_grad():
            p.add_(eps * d)
            lp = float(loss())
            p.sub_(2 * eps * d)
            lm = float(loss())
            p.add_(eps * d)
        fd = (lp - lm) / (2 * eps)
        report.append((k, analytic, fd))
        # tiny, kink-rich parameters (clamp / ReLU / mask boundaries inside the difference quotient): a coarse band; their
        # exact values are pinned by test_train_mode_forward_and_gradients_match_oracle_on_the_same_masks
        tol = 0.12 if ("pair_pos" in k or ".bias.main" in k) else 3e-2
        assert abs(fd - analytic) < tol * abs(analytic) + 1e-3, (k, analytic, fd, eps)
    print("finite-difference check:", [(k.split(".")[-3:], round(a, 4), round(f, 4)) for k, a, f in report])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_every_train_forward_draws_fresh_masks(precision):
    """The reference's nn.Dropout modules draw new masks on every forward; so must the drop-in when it is used in the
    reference's own loop (no GraphFusionStep): two consecutive train-mode forwards differ, a frozen seed repeats, and
    torch.manual_seed re-derives the device seed."""
    from ekaid_b200.functions import rng_state
    dev = _dev()
    z, meta = load_case("c0_b2_n52_all")
    sd, inp, _ = case_inputs(meta)
    m = build_model(meta, sd, precision, dev)
    m.train()
    dinp = to_dev(inp, dev)
    with torch.no_grad():
        a = m(*dinp)[3].clone()
        b = m(*dinp)[3].clone()
        assert rel_err(a, b) > 1e-3, "two train-mode forwards reused the same dropout masks"
        m.freeze_dropout_seed = True
        c = m(*dinp)[3].clone()
        d = m(*dinp)[3].clone()
        assert torch.equal(c, d)
        m.freeze_dropout_seed = False
        # stand-alone encoder entry point advances too
        enc = m.spatial_relation.train()
        v = torch.randn(2, 52, 1024, device=dev)
        q = torch.randn(2, 1024, device=dev)
        o1 = enc(v.clone(), dinp[2], q)[0].clone()
        o2 = enc(v.clone(), dinp[2], q)[0].clone()
        assert rel_err(o1, o2) > 1e-3
        s0 = int(rng_state(dev).item())
        torch.manual_seed(torch.initial_seed() + 1)
        s1 = int(rng_state(dev).item())
        assert s0 != s1, "torch.manual_seed did not re-derive the device-side dropout seed"


def _site_mask(dev, site, p, shape):
    """The product's own dropout multipliers of one site for the current seed (ekaid_drop_mask), on the CPU."""
    from ekaid_b200 import lib
    from ekaid_b200.functions import rng_state
    n = 1
    for s_ in shape:
        n *= s_
    out = torch.empty(n, dtype=torch.float32, device=dev)
    lib.call("drop_mask", rng_state(dev).data_ptr(), site, float(p), n, out.data_ptr())
    return out.cpu().view(*shape)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_train_mode_forward_and_gradients_match_oracle_on_the_same_masks(precision):
    """Train mode with the REAL dropout probabilities: the masks every kernel derives from (seed, site, element) are read
    back through ekaid_drop_mask and handed to the oracle, whose dropout sits where the reference's modules put it --
    in front of every FCNet Linear (fc.py:25-32), on the doubled GAT output before the ReLU (graph_att.py:103-104), after
    tanh / sigmoid of the fusion (modules.py:279-287), between embed's Linear and ReLU (modules.py:105-111), in front of
    W1 and on the pooled question vector (language_model.py:141,155).  A mask applied at the wrong place, to the wrong
    tensor or with the wrong scale shows up as an O(1) output error."""
    from ekaid_b200 import functions
    from oracle import ekaid_oracle as O
    dev = _dev()
    z, meta = load_case("c1_b3_n52_all_grads")
    sd, inp, _ = case_inputs(meta)
    m = build_model(meta, sd, precision, dev)
    m.train()
    torch.manual_seed(20261017)                   # the device seed is re-derived from torch's: same masks in every run
    m.freeze_dropout_seed = True                  # the forward below and ekaid_drop_mask see the same seed
    functions.DEBUG_SINK = []
    try:
        outs = m(*to_dev(inp, dev), setting="mode2", graph=meta["graph"])
        relu_masks = functions.DEBUG_SINK
    finally:
        functions.DEBUG_SINK = None
    B, N, D, L = meta["B"], meta["N"], 1024, inp[8].shape[1]
    K, BN, M, dim = N, B * N, 2 * B * N, m.dim
    halves = lambda t, *shape: (t[:t.shape[0] // 2].reshape(B, *shape), t[t.shape[0] // 2:].reshape(B, *shape))   # noqa: E731
    drop = {"question": {"w1": _site_mask(dev, 10, m.q_att.W1_self_att_q.main[0].p, (L, B, D)).transpose(0, 1),
                         "qv": _site_mask(dev, 11, m.q_att.drop.p, (B, D))}}
    gats = {"sem": (m.semantic_relation.explicit_relation, 100), "spa": (m.spatial_relation.explicit_relation, 200),
            "imp": (m.imp_relation.implicit_relation, 300)}
    for key, (g, s0) in gats.items():
        p_fc, p_gat = g.self_weights.main[0].p, g.dropout.p
        assert 0.0 < p_fc < 1.0 and 0.0 < p_gat < 1.0
        vq = halves(_site_mask(dev, s0 + 1, p_fc, (M, 2 * D)), N, 2 * D)
        q = halves(_site_mask(dev, s0 + 2, p_fc, (M, D)), N, D)
        k = halves(_site_mask(dev, s0 + 3, p_fc, (M, D)), N, D)
        out = halves(_site_mask(dev, s0 + 5, p_gat, (M, D)), N, D)
        pos = halves(_site_mask(dev, s0 + 4, p_fc, (2 * B, N * K, 64)), N * K, 64) if key == "imp" else (None, None)
        drop[key] = tuple({"vq": vq[i], "q": q[i], "k": k[i], "out": out[i], "pos": pos[i]} for i in range(2))
    drop["ctx"] = halves(_site_mask(dev, 20, m.dropout.p, (M, D)), N, D)
    drop["gate"] = halves(_site_mask(dev, 21, m.dropout.p, (M, D)), N, D)
    drop["embed"] = halves(_site_mask(dev, 22, m.embed[1].p, (M, dim)), N, dim)
    frac = float((drop["ctx"][0] == 0).float().mean())
    assert 0.45 < frac < 0.55 and abs(float(drop["ctx"][0].max()) - 2.0) < 1e-6        # p = 0.5, scale 1 / (1 - p)
    cd = m.cfg.model.change_detector
    sdg = {k_: v.clone().requires_grad_(v.is_floating_point() and k_ != "w_emb.emb_.weight") for k_, v in sd.items()}
    ro = O.change_detector_forward(sdg, *inp, graph=meta["graph"], num_heads=cd.att_head, nongt_dim=meta["nongt_dim"],
                                   relu_masks=relu_masks, drop=drop)
    with torch.no_grad():
        ev = oracle_forward(sd, inp, meta)
    errs = {k_: rel_err(o, r) for k_, o, r in zip(OUT_NAMES[1:], outs[1:], ro[1:])}
    print(precision, "train mode, same masks:", {k_: "%.1e" % v for k_, v in errs.items()},
          "| eval vs train (oracle): %.2f" % rel_err(ro[3], ev[3]))
    assert rel_err(ro[3], ev[3]) > 5e-2                       # the masks really change the result
    for k_, e in errs.items():
        assert e < TOL[precision], (precision, k_, e)
    ws = loss_weights(outs)
    sum((o * w.to(dev)).sum() for o, w in zip(outs[1:], ws)).backward()
    sum((o * w).sum() for o, w in zip(ro[1:], ws)).backward()
    bad = []
    n_checked = 0
    for k_, p in m.named_parameters():
        g_ref = sdg[k_].grad
        if g_ref is None or float(g_ref.abs().max()) < 1e-4 or k_ == "q_att.W2_self_att_q.main.0.bias":
            continue
        e = param_err(precision, k_, p.grad, {n_: t.grad for n_, t in sdg.items() if t.grad is not None}, sd)
        n_checked += 1
        if e > GTOL[precision]:
            bad.append((k_, e))
    assert n_checked > 40 and not bad, bad
