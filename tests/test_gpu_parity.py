"""Parity of the CUDA path (through the drop-in modules -> C ABI) with the CPU oracle and with the golden vectors
the reference produced.  Tolerances are the north-star's: 1e-4 relative (fp32 path), 2e-2 (bf16 tensor-core path),
measured as max|a-b| / max|b| per output tensor; gradients likewise per parameter."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from helpers import (CASES, OUT_NAMES, case_inputs, check_fixture_inputs, load_case, loss_weights, oracle_forward,
                     rel_err, spec_for, speaker_spec)

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}
# Gradients are compared with the oracle evaluated on the SAME ReLU active sets as the kernels chose
# (functions.DEBUG_SINK -> oracle relu_masks): a pre-activation within rounding distance of zero may legitimately
# fall on either side, which changes individual bias/gain gradients by ~1e-3 without any kernel being wrong.
GTOL = {"fp32": 5e-4, "bf16": 5e-2}


# analytically zero: a constant added to every logit of the batch-axis softmax (Q4) changes nothing; what either side
# computes for it is cancellation noise
ZERO_GRAD = ("q_att.W2_self_att_q.main.0.bias",)


def grad_err(g, ref, precision):
    """fp32 path: max-abs error / max-abs.  16-bit path: relative L2 error (activation gradients are stored in bf16,
    so single entries of cancelling sums carry ~2^-9 of the TERM size; the L2 norm is the meaningful scale)."""
    g = g.detach().double().cpu()
    ref = ref.detach().double().cpu()
    if precision == "fp32":
        return float((g - ref).abs().max() / (ref.abs().max() + 1e-30))
    return float((g - ref).norm() / (ref.norm() + 1e-30))


def gain_err(g, k, ref_grads, sd):
    """Scalar gain of a legacy weight_norm layer (w = g v / ||v||): its gradient dg = <dW, v> / ||v|| is the projection
    of the effective-weight gradient dW onto one direction -- a cancelling inner product of ~1e6 terms for the big
    layers.  The 16-bit path bounds the error of dW in relative L2 (GTOL); the consistent bound for its projection is
    the same fraction of ||dW||_F, which follows from the two reference gradients:
    ||dW||^2 = dg^2 + (||v|| / g)^2 ||dv||^2."""
    kv = k[:-len("weight_g")] + "weight_v"
    dg_ref = float(ref_grads[k].double())
    v, gval = sd[kv].double(), float(sd[k].double())
    dv_ref = ref_grads[kv].double()
    dW = (dg_ref ** 2 + (float(v.norm()) / gval) ** 2 * float(dv_ref.norm()) ** 2) ** 0.5
    return abs(float(g.detach().double().cpu()) - dg_ref) / (dW + 1e-30)


def param_err(precision, k, g, ref_grads, values):
    """Error of one parameter gradient in the metric of its kind (k = full reference name)."""
    if k.endswith("weight_g") and (k[:-len("weight_g")] + "weight_v") in ref_grads:
        if precision == "bf16":
            return gain_err(g, k, ref_grads, values)
        # fp32 path: the direct relative error, or -- where the projection cancels so strongly that the 5e-6 of the
        # split-precision tensor-core products shows -- a 10x tighter bound than GTOL on the same ||dW||_F scale
        return min(grad_err(g, ref_grads[k], precision), 10.0 * gain_err(g, k, ref_grads, values))
    return grad_err(g, ref_grads[k], precision)


def _dev():
    from ekaid_b200 import lib
    lib.require_device()
    return torch.device("cuda:0")


def build_model(meta, sd, precision, dev):
    from ekaid_b200.config import WORD_TO_IDX, default_cfg
    from ekaid_b200.modules import ChangeDetector
    cfg = default_cfg(meta["graph"], nongt_dim=meta["nongt_dim"])
    cfg.data.train.empty_image = meta["empty_image"]
    with contextlib.redirect_stdout(io.StringIO()):
        m = ChangeDetector(cfg, WORD_TO_IDX)
    m.load_state_dict(sd)
    m.to(dev).eval().set_precision(precision)
    return m


def to_dev(inp, dev):
    # boxes (6, 7) stay on the CPU as float64, like the reference step (train_mimic.py:206-218)
    return tuple(t if i in (6, 7) else t.to(dev) for i, t in enumerate(inp))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_forward_matches_golden_and_oracle(name, precision):
    dev = _dev()
    z, meta = load_case(name)
    sd, inp, _ = case_inputs(meta)
    check_fixture_inputs(z, sd, inp)
    m = build_model(meta, sd, precision, dev)
    with torch.no_grad():
        outs = m(*to_dev(inp, dev), setting="mode2", graph=meta["graph"])
        ref = oracle_forward(sd, inp, meta)
    errs = {}
    for k, o, r in zip(OUT_NAMES, outs, ref):
        assert tuple(o.shape) == z[k].shape, k
        errs[k] = (rel_err(o, z[k]), rel_err(o, r))
    print(name, precision, {k: "%.1e/%.1e" % v for k, v in errs.items()})
    # every output -- including the 40x cancelling difference `input_attended` and the fc1 head `pred` -- at the
    # north-star tolerance relative to ITS OWN maximum, against the reference's golden vectors and against the oracle
    for k, (eg, eo) in errs.items():
        tol = TOL[precision]
        if k in ("pred", "input_attended") and float(np.abs(z[k]).max()) == 0.0:
            continue        # empty_image: both images identical, the difference is exactly zero on both sides
        assert eg < tol, (name, precision, k, "vs golden", eg)
        assert eo < tol, (name, precision, k, "vs oracle", eo)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["c1_b3_n52_all_grads", "c7_b2_n52_zero_bias", "c4_b2_n60_k52_all",
                                  "c2_b2_n52_ips"])
def test_gradients_match_oracle(name, precision):
    z, meta = load_case(name)
    check_gradients(meta, precision, name)


def check_gradients(meta, precision, name=""):
    """Forward + gradient of every live parameter against the oracle's autograd on the inputs / weights `meta` seeds."""
    from ekaid_b200 import functions
    dev = _dev()
    sd, inp, _ = case_inputs(meta)
    m = build_model(meta, sd, precision, dev)
    functions.DEBUG_SINK = []
    try:
        outs = m(*to_dev(inp, dev), setting="mode2", graph=meta["graph"])
        masks = functions.DEBUG_SINK
    finally:
        functions.DEBUG_SINK = None
    ws = loss_weights(outs)
    loss = sum((o * w.to(dev)).sum() for o, w in zip(outs[1:], ws))
    loss.backward()
    sdg = {k: v.clone().requires_grad_(v.is_floating_point() and k != "w_emb.emb_.weight") for k, v in sd.items()}
    cd = m.cfg.model.change_detector
    from oracle import ekaid_oracle as O
    ro = O.change_detector_forward(sdg, *inp, graph=meta["graph"], num_heads=cd.att_head, nongt_dim=meta["nongt_dim"],
                                   pos_emb_dim=cd.pos_emb_dim, coef_sem=cd.coef_sem, coef_spa=cd.coef_spa,
                                   relu_masks=masks)
    for k, o, r in zip(OUT_NAMES[1:5], outs[1:5], ro[1:5]):
        assert rel_err(o, r) < TOL[precision], ("masked oracle forward", k, rel_err(o, r))
    rl = sum((o * w).sum() for o, w in zip(ro[1:], ws))
    rl.backward()
    bad = []
    table = []
    for k, p in m.named_parameters():
        g_ref = sdg[k].grad
        if g_ref is None or float(g_ref.abs().max()) < 1e-4 or k in ZERO_GRAD:
            # dead parameters (Q2, Q3, Q11) and analytically-zero gradients (softmax shift invariance)
            if p.grad is not None:
                assert float(p.grad.abs().max()) < (1e-3 if precision == "fp32" else 0.5), (k, float(p.grad.abs().max()))
            continue
        assert p.grad is not None, k
        e = param_err(precision, k, p.grad, {n: t.grad for n, t in sdg.items() if t.grad is not None}, sd)
        table.append((e, k))
        if e > GTOL[precision]:
            bad.append((k, e))
    table.sort(reverse=True)
    print(name, precision, "worst grads:", [("%.1e" % e, k) for e, k in table[:6]])
    assert len(table) > 40
    assert not bad, bad


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["c0_b2_n52_all", "c1_b3_n52_all_grads"])
def test_argmax_answer_tokens(name, precision):
    """Arg-max answer tokens through the answer decoder (oracle restatement of DynamicSpeaker, pinned to the
    reference's tokens in tests/test_oracle_golden.py).

    With random-init decoder weights the 148-way logits are nearly tied at many of the 90 steps, and
    greedy decoding feeds every flip back into the recurrence.  The check is therefore made step-wise along the
    REFERENCE token path (teacher forcing on the golden tokens): at every step where the reference's top-2 logit
    margin exceeds `tau`, the arg-max computed from the CUDA path's (bef, aft, diff) must be the reference token.
    tau: 1e-3 nat for the fp32 path, 0.01 nat for the 16-bit path (twice its largest log-probability deviation); the free-running greedy decode must additionally
    agree on the whole prefix before the first sub-margin step."""
    from ekaid_b200.synthetic import synthetic_state_dict
    from oracle import ekaid_oracle as O
    dev = _dev()
    z, meta = load_case(name)
    sd, inp, _ = case_inputs(meta)
    m = build_model(meta, sd, precision, dev)
    tau = 1e-3 if precision == "fp32" else 0.01      # 16-bit path: measured max |dlogp| 4-5e-3 nat
    with torch.no_grad():
        outs = m(*to_dev(inp, dev), setting="mode2", graph=meta["graph"])
        ssd = synthetic_state_dict(speaker_spec(), 4321)
        gold = torch.from_numpy(z["tokens"])
        B = gold.shape[0]
        tf_in = torch.cat([torch.full((B, 1), 2, dtype=torch.long), gold], 1)     # <bos>=2 then the reference tokens
        ref_feats = [torch.from_numpy(z[k]) for k in ("attended_1", "attended_2", "input_attended")]
        my_feats = [outs[3].cpu(), outs[4].cpu(), outs[5].cpu()]

        def run(feats):
            state = (torch.zeros(2, B, 512), torch.zeros(2, B, 512))
            lps = []
            for t in range(90):
                lp, state, _ = O.speaker_logprobs(ssd, tf_in[:, t], *feats, state)
                if t == 0:
                    lp = lp.clone()
                    lp[:, 0] = float("-inf")
                lps.append(lp)
            return torch.stack(lps, 1)                                            # [B, 90, V]

        lr, lm = run(ref_feats), run(my_feats)
        assert torch.equal(lr.argmax(2), gold), "teacher-forced reference decode must reproduce the golden tokens"
        top2 = lr.topk(2, dim=2).values
        margin = top2[..., 0] - top2[..., 1]
        solid = margin > tau
        agree = lm.argmax(2) == gold
        frac = float(solid.float().mean())
        print(name, precision, "steps with margin > %g: %.0f%%; agreement on them: %.4f; overall agreement %.4f"
              % (tau, 100 * frac, float(agree[solid].float().mean()), float(agree.float().mean())))
        print("   margin quantiles (5/25/50/75%%): %s; max |dlogp| %.3g" % (
            [round(float(q), 4) for q in torch.quantile(margin.flatten(), torch.tensor([.05, .25, .5, .75]))],
            float((lm - lr)[:, 1:].abs().max())))
        assert frac > (0.5 if precision == "fp32" else 0.3)
        assert float((lm - lr)[:, 1:].abs().max()) < (1e-3 if precision == "fp32" else 1e-2)
        assert bool(agree[solid].all()), "arg-max token differs at a step with a solid reference margin"
        seq = O.speaker_greedy(ssd, *my_feats, 90, 512)
        for b in range(B):
            weak = (~solid[b]).nonzero().flatten()
            upto = int(weak[0]) if len(weak) else 90
            assert torch.equal(seq[b, :upto], gold[b, :upto]), (b, upto)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_relation_encoders_standalone(precision):
    """ExplicitRelationEncoder / ImplicitRelationEncoder as individually swappable modules (single image batch)."""
    from ekaid_b200.modules import ExplicitRelationEncoder, ImplicitRelationEncoder
    from ekaid_b200.synthetic import synthetic_batch, synthetic_state_dict
    from oracle import ekaid_oracle as O
    dev = _dev()
    B, N, D = 3, 52, 1024
    full = synthetic_state_dict(spec_for("all"), 1238)
    batch = synthetic_batch(B, N, seed=21)
    g = torch.Generator().manual_seed(5)
    v = torch.randn(B, N, D, generator=g)
    v[1, 7] = 0                                   # a zero row: quirk Q10
    q = torch.randn(B, D, generator=g)
    adj = O.process_matrix(batch[6], N, 11)
    for kind in ("explicit", "implicit"):
        with contextlib.redirect_stdout(io.StringIO()):
            if kind == "explicit":
                enc = ExplicitRelationEncoder(D, D, D, 2, 11, num_heads=4, nongt_dim=52, label_bias=False)
                prefix, R = "spatial_relation.", O.REL_SPA
            else:
                enc = ImplicitRelationEncoder(D, D, D, 2, 64, 52, num_heads=4, label_bias=False)
                prefix, R = "imp_relation.", O.REL_IMP
        sub = {k[len(prefix):]: t for k, t in full.items() if k.startswith(prefix)}
        enc.load_state_dict(sub)
        enc.to(dev).eval()
        enc.precision = precision
        vd = v.clone().to(dev).requires_grad_(True)
        qd = q.clone().to(dev).requires_grad_(True)
        geo = adj.to(dev) if kind == "explicit" else batch[10]
        from ekaid_b200 import functions
        functions.DEBUG_SINK = []
        try:
            out, aff = enc(vd, geo, qd)
            mask = functions.DEBUG_SINK[0]
        finally:
            functions.DEBUG_SINK = None
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(6))
        (out * w.to(dev)).sum().backward()
        sdg = {k: t.clone().requires_grad_(True) for k, t in full.items() if k.startswith(prefix)}
        vr = v.clone().requires_grad_(True)
        qr = q.clone().requires_grad_(True)
        pe = None if kind == "explicit" else O.position_embedding(O.position_matrix(batch[10], 52), 64)
        ref, aux = O.gat_relation(sdg, R, vr, qr, adj if kind == "explicit" else None, pe, 4, 52, return_aux=True,
                                  relu_mask=mask.view(B, N, D))
        (ref * w).sum().backward()
        assert rel_err(out, ref) < TOL[precision], (kind, rel_err(out, ref))
        assert rel_err(aff[1], aux["P"]) < TOL[precision] * 5, (kind, "P", rel_err(aff[1], aux["P"]))
        assert grad_err(vd.grad, vr.grad, precision) < GTOL[precision], (kind, "dv", grad_err(vd.grad, vr.grad, precision))
        assert grad_err(qd.grad, qr.grad, precision) < GTOL[precision], (kind, "dq", grad_err(qd.grad, qr.grad, precision))
        for k, p in enc.named_parameters():
            gr = sdg[prefix + k].grad
            if gr is None or float(gr.abs().max()) < 1e-4:
                continue
            e = param_err(precision, prefix + k, p.grad, {n: t.grad for n, t in sdg.items() if t.grad is not None}, full)
            assert e < GTOL[precision], (kind, k, e)
        # outside autograd the encoder mutates and returns its first argument (quirk Q1)
        with torch.no_grad():
            v2 = v.clone().to(dev)
            o2, _ = enc(v2, geo, q.to(dev))
            assert o2 is v2 and rel_err(v2, ref) < TOL[precision]


def test_question_path_matches_oracle():
    from ekaid_b200.functions import PC
    from oracle import ekaid_oracle as O
    dev = _dev()
    z, meta = load_case("c1_b3_n52_all_grads")
    sd, inp, _ = case_inputs(meta)
    for precision in ("fp32", "bf16"):
        m = build_model(meta, sd, precision, dev)
        qv = m.question_vector(PC(precision), inp[8].to(dev))
        w = torch.randn(qv.shape, generator=torch.Generator().manual_seed(3))
        (qv * w.to(dev)).sum().backward()
        sdg = {k: v.clone().requires_grad_(v.is_floating_point() and k != "w_emb.emb_.weight") for k, v in sd.items()
               if k.startswith(("w_emb", "q_emb", "q_att"))}
        ref = O.question_vector(sdg, inp[8])
        (ref * w).sum().backward()
        assert rel_err(qv, ref) < TOL[precision], rel_err(qv, ref)
        for k, p in m.named_parameters():
            if k in sdg and sdg[k].grad is not None and float(sdg[k].grad.abs().max()) > 1e-4:
                e = param_err(precision, k, p.grad, {n: t.grad for n, t in sdg.items() if t.grad is not None}, sd)
                assert e < GTOL[precision], (precision, k, e)


def test_batch_coupling_q4_and_local_batch_contract():
    """Quirk Q4: outputs of one sample depend on the others in the batch; a shard processed alone must equal the
    oracle on that shard (the data-parallel contract of SURVEY.md section 8(e))."""
    dev = _dev()
    z, meta = load_case("c1_b3_n52_all_grads")
    sd, inp, _ = case_inputs(meta)
    m = build_model(meta, sd, "fp32", dev)
    shard = tuple(t[:2] for t in inp)
    with torch.no_grad():
        full = m(*to_dev(inp, dev))
        part = m(*to_dev(shard, dev))
        ref_part = oracle_forward(sd, shard, dict(meta, B=2))
    assert rel_err(part[3], ref_part[3]) < 1e-4
    assert rel_err(part[3], full[3][:2]) > 1e-3        # genuinely different function of the batch


def test_error_behaviour():
    from ekaid_b200.config import WORD_TO_IDX, default_cfg
    from ekaid_b200.modules import ChangeDetector
    dev = _dev()
    z, meta = load_case("c0_b2_n52_all")
    sd, inp, _ = case_inputs(meta)
    m = build_model(meta, sd, "fp32", dev)
    with pytest.raises(NotImplementedError):
        m(*to_dev(inp, dev), setting="mode1")
    with pytest.raises(ValueError):
        m(*to_dev(inp, dev), graph="bogus")
    bad = list(to_dev(inp, dev))
    bad[2] = bad[2][:, :40]                     # adjacency with the wrong node count
    with pytest.raises(ValueError):
        m(*bad)


def test_full_size_batch64_forward_and_consistency():
    """BASELINE.json configs[1] size (batch 64 x 52 nodes x 1024-d): both precision paths against the CPU oracle
    (forward, a few seconds of CPU), run-to-run determinism of the fp32 path, and the size-independent property that
    the relation encoders treat images independently given the question vector (rows of a sub-batch are unchanged)."""
    from ekaid_b200.synthetic import synthetic_batch, synthetic_state_dict
    from oracle import ekaid_oracle as O
    dev = _dev()
    B, N = 64, 52
    meta = dict(graph="all", nongt_dim=52, empty_image=False, B=B, N=N)
    sd = synthetic_state_dict(spec_for("all"), 1238)
    b = synthetic_batch(B, N, seed=4321)
    inp = (b[0], b[1], O.process_matrix(b[6], N, 11), O.process_matrix(b[7], N, 11), O.process_matrix(b[8], N, 3),
           O.process_matrix(b[9], N, 3), b[10], b[11], b[12])
    with torch.no_grad():
        ref = oracle_forward(sd, inp, meta)
    outs = {}
    for precision in ("fp32", "bf16"):
        m = build_model(meta, sd, precision, dev)
        with torch.no_grad():
            outs[precision] = m(*to_dev(inp, dev))
            again = m(*to_dev(inp, dev))
        for k, o, r in zip(OUT_NAMES[1:5], outs[precision][1:5], ref[1:5]):
            assert rel_err(o, r) < TOL[precision], (precision, k, rel_err(o, r))
        if precision == "fp32":
            assert all(torch.equal(x, y) for x, y in zip(outs[precision], again))       # deterministic
    assert rel_err(outs["fp32"][5], ref[5]) < 1e-4
    scale = max(float(ref[3].abs().max()), float(ref[4].abs().max()))
    assert float((outs["bf16"][5].cpu() - ref[5]).abs().max()) / scale < 2e-2
    # image independence of one relation step given q: first 8 samples alone == first 8 rows of the full batch
    from ekaid_b200.modules import ExplicitRelationEncoder
    with contextlib.redirect_stdout(io.StringIO()):
        enc = ExplicitRelationEncoder(1024, 1024, 1024, 2, 11, num_heads=4, nongt_dim=52, label_bias=False)
    enc.load_state_dict({k[len("spatial_relation."):]: t for k, t in sd.items() if k.startswith("spatial_relation.")})
    enc.to(dev).eval()
    enc.precision = "fp32"
    g = torch.Generator().manual_seed(1)
    v = torch.randn(B, N, 1024, generator=g).to(dev)
    q = torch.randn(B, 1024, generator=g).to(dev)
    adj = inp[2].to(dev)
    with torch.no_grad():
        full, _ = enc(v.clone(), adj, q)
        part, _ = enc(v[:8].clone(), adj[:8], q[:8])
    assert torch.equal(full[:8], part)


@pytest.mark.parametrize("B", [1, 3, 64, 130])
def test_gru_one_launch_recurrence_matches_step_kernels(B):
    """gru_seq.cu (all 20 steps in one persistent launch, grid barrier between steps) against the per-step GEMM + cell
    kernels it replaces on the bf16 path, forward and BPTT, including ragged and multi-pass batch sizes."""
    from ekaid_b200 import functions as F
    from ekaid_b200.functions import PC
    dev = _dev()
    z, meta = load_case("c1_b3_n52_all_grads")
    sd, inp, _ = case_inputs(meta)
    g = torch.Generator().manual_seed(11 + B)
    question = torch.randint(0, 100, (B, 20), generator=g).to(dev)
    w = torch.randn(B, 1024, generator=g).to(dev)
    res = {}
    old = (F.GRU_SEQ, F.GRU_SEQ_MAX_BATCH)
    F.GRU_SEQ_MAX_BATCH = 1 << 20          # exercise the multi-pass code too, whatever the dispatch heuristic says
    try:
        for mode in (True, False):
            F.GRU_SEQ = mode
            m = build_model(meta, sd, "bf16", dev)
            qv = m.question_vector(PC("bf16"), question)
            (qv * w).sum().backward()
            res[mode] = (qv.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    finally:
        F.GRU_SEQ, F.GRU_SEQ_MAX_BATCH = old
    assert rel_err(res[True][0], res[False][0]) < 2e-3, rel_err(res[True][0], res[False][0])
    assert set(res[True][1]) == set(res[False][1])
    for k, gseq in res[True][1].items():
        gstep = res[False][1][k]
        # (the attention bias b2 has an analytically zero gradient -- softmax shift invariance -- hence the atol)
        e = float((gseq - gstep).norm() / (gstep.norm() + 1e-3 * gstep.numel() ** 0.5))
        assert e < 3e-2, (k, e, float(gstep.abs().max()))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gattnet_and_attention_layer_standalone(precision):
    """GAttNet.forward(v_feat = [v | q], adj, pos_emb) and, through it, GraphSelfAttentionLayer.forward -- the inner
    signatures of SURVEY.md section 8(b) -- against the oracle: output = relu(2 out) = X_new - X, aff[1] = P."""
    from ekaid_b200.modules import ExplicitRelationEncoder, ImplicitRelationEncoder, q_expand_v_cat
    from ekaid_b200.synthetic import synthetic_batch, synthetic_state_dict
    from oracle import ekaid_oracle as O
    dev = _dev()
    B, N, D = 2, 52, 1024
    full = synthetic_state_dict(spec_for("all"), 1238)
    batch = synthetic_batch(B, N, seed=23)
    g = torch.Generator().manual_seed(9)
    v = torch.randn(B, N, D, generator=g)
    v[1, 5] = 0
    q = torch.randn(B, D, generator=g)
    adj = O.process_matrix(batch[6], N, 11)
    tol = TOL[precision]
    for kind in ("explicit", "implicit"):
        with contextlib.redirect_stdout(io.StringIO()):
            if kind == "explicit":
                enc = ExplicitRelationEncoder(D, D, D, 2, 11, num_heads=4, nongt_dim=52, label_bias=False)
                prefix, R, gat = "spatial_relation.", O.REL_SPA, None
            else:
                enc = ImplicitRelationEncoder(D, D, D, 2, 64, 52, num_heads=4, label_bias=False)
                prefix, R = "imp_relation.", O.REL_IMP
        enc.load_state_dict({k[len(prefix):]: t for k, t in full.items() if k.startswith(prefix)})
        enc.to(dev).eval()
        for m in enc.modules():
            if hasattr(m, "precision"):
                m.precision = precision
        gat = enc.explicit_relation if kind == "explicit" else enc.implicit_relation
        vq = q_expand_v_cat(q, v).to(dev).requires_grad_(True)
        if kind == "explicit":
            out, aff = gat(vq, adj.to(dev))
            pos = None
        else:
            pos = O.position_embedding(O.position_matrix(batch[10], 52), 64)          # [B,N,K,64] fp64, as the reference
            out, aff = gat(vq, torch.ones(B, N, N, 1, device=dev), pos.to(dev))
        sdg = {k: t.clone().requires_grad_(True) for k, t in full.items() if k.startswith(prefix)}
        vr = v.clone().requires_grad_(True)
        Xn, aux = O.gat_relation(sdg, R, vr, q, adj if kind == "explicit" else None, pos, 4, 52, return_aux=True)
        ref = (Xn - vr).detach()
        assert rel_err(out, ref) < tol, (kind, precision, rel_err(out, ref))
        assert len(aff) == 2 and tuple(aff[1].shape) == (B, N, 4, 52)
        assert float((aff[1].cpu() - aux["P"].detach()).abs().max()) < (1e-4 if precision == "fp32" else 2e-2)
        # gradients through the stand-alone path: d/d(v half of the input) and a weight
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(10))
        (out * w.to(dev)).sum().backward()
        ((Xn - vr) * w).sum().backward()
        gv = vq.grad[..., :D].cpu()
        # the oracle's vr.grad also contains the residual-free path only (Xn - vr removes the identity)
        e = float((gv - vr.grad).norm() / vr.grad.norm())
        assert e < (1e-2 if precision == "fp32" else 0.1), (kind, precision, "d v", e)
        pname = "neighbor_net.1.linear_out_2.weight"
        pg = dict(gat.named_parameters())[pname].grad.cpu()
        rg = sdg[prefix + ("explicit_relation." if kind == "explicit" else "implicit_relation.") + pname].grad
        e = float((pg - rg).norm() / rg.norm())
        assert e < (1e-2 if precision == "fp32" else 0.1), (kind, precision, pname, e)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_question_modules_standalone(precision):
    """WordEmbedding -> QuestionEmbedding.forward_all -> QuestionSelfAttention.forward called one by one (the inner
    signatures of SURVEY.md section 8(b)) must reproduce the oracle's question vector and its gradients."""
    from oracle import ekaid_oracle as O
    dev = _dev()
    z, meta = load_case("c1_b3_n52_all_grads")
    sd, inp, _ = case_inputs(meta)
    m = build_model(meta, sd, precision, dev)
    question = inp[8].to(dev)
    w_emb = m.w_emb(question)                                       # [B,L,600]
    hs = m.q_emb.forward_all(w_emb)                                 # [B,L,1024]
    qv = m.q_att(hs)                                                # [B,1024]
    assert tuple(hs.shape) == (question.shape[0], question.shape[1], 1024)
    assert float((m.q_emb(w_emb).detach() - hs[:, -1].detach()).abs().max()) < (1e-5 if precision == "fp32" else 2e-2)
    sdg = {k: v.clone().requires_grad_(v.is_floating_point() and k != "w_emb.emb_.weight") for k, v in sd.items()
           if k.startswith(("w_emb", "q_emb", "q_att"))}
    ref = O.question_vector(sdg, inp[8])
    assert rel_err(qv, ref) < TOL[precision], rel_err(qv, ref)
    w = torch.randn(qv.shape, generator=torch.Generator().manual_seed(3))
    (qv * w.to(dev)).sum().backward()
    (ref * w).sum().backward()
    for k in ("q_emb.rnn.weight_hh_l0", "q_emb.rnn.weight_ih_l0", "q_att.W1_self_att_q.main.1.weight_v", "w_emb.emb.weight"):
        g = dict(m.named_parameters())[k].grad
        assert g is not None, k
        e = grad_err(g, sdg[k].grad, precision)
        assert e < GTOL[precision], (precision, k, e)
