"""The step object (ekaid_b200.step) on a B200: gradients written straight into the flat slots must equal the plain
autograd gradients, the flat Adam must equal torch.optim.Adam, and a CUDA-graph replay must equal the eager step."""
import contextlib
import io

import pytest
import torch

from helpers import case_inputs, load_case, rel_err
from test_gpu_parity import build_model, to_dev

pytestmark = pytest.mark.gpu


def _dev():
    from ekaid_b200 import lib
    lib.require_device()
    return torch.device("cuda:0")


def _loss(outs):
    return outs[3].sum() * 0.01 + outs[4].sum() * 0.02 - outs[5].sum() * 0.03 + 2.5e-3 * (outs[1].sum() + outs[2].sum())


def test_slot_gradients_and_flat_adam_match_autograd_and_torch_adam():
    from ekaid_b200 import functions
    from ekaid_b200.step import FlatAdam
    dev = _dev()
    z, meta = load_case("c0_b2_n52_all")
    sd, inp, _ = case_inputs(meta)
    dinp = to_dev(inp, dev)
    ref = build_model(meta, sd, "fp32", dev)
    _loss(ref(*dinp)).backward()
    ref_grads = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
    opt_ref = torch.optim.Adam([p for p in ref.parameters() if p.grad is not None], lr=1e-3)
    opt_ref.step()

    m = build_model(meta, sd, "fp32", dev)
    assert not functions.GRAD_SLOTS
    try:
        opt = FlatAdam(m.live_parameters(), lr=1e-3)
        assert functions.GRAD_SLOTS
        opt.zero_grad()
        _loss(m(*dinp)).backward()
        opt.check_slots()
        names = {id(p): k for k, p in m.named_parameters()}
        n_checked = 0
        for p, slot in zip(opt.params, opt.slots):
            k = names[id(p)]
            if k in ref_grads:
                assert torch.equal(slot, ref_grads[k]), k          # same kernels, same order -> bit-identical
                n_checked += 1
        assert n_checked > 60
        assert set(ref_grads) <= {names[id(p)] for p in opt.params}
        opt.step()
        torch.cuda.synchronize()
        for k, p in m.named_parameters():
            q = dict(ref.named_parameters())[k]
            assert float((p.detach() - q.detach()).abs().max()) < 2e-6, k
    finally:
        functions.GRAD_SLOTS.clear()


def test_graph_replay_matches_eager_step():
    from ekaid_b200 import functions
    from ekaid_b200.step import GraphFusionStep, expand_adjacency, select_fields
    from ekaid_b200.synthetic import synthetic_batch
    dev = _dev()
    z, meta = load_case("c0_b2_n52_all")
    sd, _, _ = case_inputs(meta)
    batches = [tuple(t.to(dev) for t in select_fields(synthetic_batch(2, 52, seed=s))) for s in (1, 2, 3)]
    losses = {}
    params = {}
    try:
        for mode in ("eager", "graph"):
            functions.GRAD_SLOTS.clear()
            m = build_model(meta, sd, "fp32", dev)
            step = GraphFusionStep(m, m.cfg, lr=1e-3)
            if mode == "graph":
                # capture runs warm-up steps that update the weights: restore them afterwards
                step.capture(batches[0], train=True, warmup=2)
                with torch.no_grad():
                    m.load_state_dict(sd)
                step.opt.m.zero_()
                step.opt.v.zero_()
                step.opt.pow_state.fill_(1.0)
            out = []
            for b in batches:
                if mode == "graph":
                    out.append(float(step.replay(b)))
                else:
                    out.append(float(step.train_step(expand_adjacency(b, m.cfg), b[9], b[10].float())))
            losses[mode] = out
            params[mode] = {k: p.detach().clone() for k, p in m.named_parameters()}
    finally:
        functions.GRAD_SLOTS.clear()
    assert losses["eager"] == pytest.approx(losses["graph"], rel=1e-6)
    assert losses["eager"][0] != losses["eager"][1]
    for k in params["eager"]:
        assert float((params["eager"][k] - params["graph"][k]).abs().max()) < 1e-6, k


def test_step_pipeline_matches_plain_replay():
    """StepPipeline (prefetch of the next host batch on a copy stream, results read one step late) must give the same
    sequence of losses and the same parameters as replaying the same host batches one after the other."""
    from ekaid_b200 import functions
    from ekaid_b200.step import GraphFusionStep, select_fields
    from ekaid_b200.synthetic import synthetic_batch
    dev = _dev()
    z, meta = load_case("c0_b2_n52_all")
    sd, _, _ = case_inputs(meta)
    host = [tuple(t.contiguous().pin_memory() for t in select_fields(synthetic_batch(2, 52, seed=s))) for s in (1, 2, 3, 4, 5)]
    losses, params = {}, {}
    try:
        for mode in ("plain", "pipeline"):
            functions.GRAD_SLOTS.clear()
            m = build_model(meta, sd, "fp32", dev)            # eval mode: no dropout, so both runs are comparable
            step = GraphFusionStep(m, m.cfg, lr=1e-3)
            step.capture(tuple(t.to(dev) for t in host[0]), train=True, warmup=2)
            with torch.no_grad():
                m.load_state_dict(sd)
            step.opt.m.zero_()
            step.opt.v.zero_()
            step.opt.pow_state.fill_(1.0)
            out = []
            if mode == "plain":
                for b in host:
                    out.append(float(step.replay(b)))
            else:
                pipe = step.pipeline()
                pipe.prefetch(host[0])
                pending = None
                for i in range(len(host)):
                    if i + 1 < len(host):
                        pipe.prefetch(host[i + 1])
                    k = pipe.run()
                    if pending is not None:
                        out.append(pipe.result(pending))
                    pending = k
                out.append(pipe.result(pending))
            torch.cuda.synchronize()
            losses[mode] = out
            params[mode] = {k: p.detach().clone() for k, p in m.named_parameters()}
    finally:
        functions.GRAD_SLOTS.clear()
    assert losses["plain"] == pytest.approx(losses["pipeline"], rel=1e-6)
    assert len(set(losses["plain"])) == len(host)
    for k in params["plain"]:
        assert float((params["plain"][k] - params["pipeline"][k]).abs().max()) < 1e-6, k


def test_second_gradient_contribution_accumulates_instead_of_overwriting():
    """A parameter that receives two gradient contributions before its .grad is reset -- a second backward without
    zero_grad (gradient accumulation), or one module applied twice in a graph -- must end up with the SUM: only the
    first contribution may be written into the flat slot, later ones are added by autograd."""
    from ekaid_b200 import functions
    from ekaid_b200.step import FlatAdam
    dev = _dev()
    z, meta = load_case("c0_b2_n52_all")
    sd, inp, _ = case_inputs(meta)
    dinp = to_dev(inp, dev)
    ref = build_model(meta, sd, "fp32", dev)
    _loss(ref(*dinp)).backward()
    g1 = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
    m = build_model(meta, sd, "fp32", dev)
    try:
        opt = FlatAdam(m.live_parameters(), lr=1e-3)
        opt.zero_grad()
        _loss(m(*dinp)).backward()
        _loss(m(*dinp)).backward()            # no zero_grad in between
        torch.cuda.synchronize()
        n = 0
        for k, p in m.named_parameters():
            if k in g1 and p.grad is not None:
                err = float((p.grad - 2 * g1[k]).abs().max() / (2 * g1[k].abs().max() + 1e-30))
                assert err < 1e-5, (k, err)
                n += 1
        assert n > 60
        # the encoder applied twice in ONE graph (reference style: forward(bef) then forward(aft))
        opt.zero_grad()
        enc = m.spatial_relation
        v1 = torch.randn(2, 52, 1024, device=dev)
        v2 = torch.randn(2, 52, 1024, device=dev)
        q = torch.randn(2, 1024, device=dev)
        o1, _ = enc(v1.clone().requires_grad_(True), dinp[2], q)
        o2, _ = enc(v2.clone().requires_grad_(True), dinp[3], q)
        (o1.sum() + o2.sum()).backward()
        w = enc.explicit_relation.neighbor_net[1].linear_out_2.weight
        both = w.grad.clone()
        opt.zero_grad()
        o1, _ = enc(v1.clone().requires_grad_(True), dinp[2], q)
        o1.sum().backward()
        first = w.grad.clone()
        opt.zero_grad()
        o2, _ = enc(v2.clone().requires_grad_(True), dinp[3], q)
        o2.sum().backward()
        second = w.grad.clone()
        err = float((both - (first + second)).abs().max() / (both.abs().max() + 1e-30))
        assert err < 1e-5, err
        opt.close()
        assert not functions.GRAD_SLOTS
    finally:
        functions.GRAD_SLOTS.clear()
        functions.slots_reset()


@pytest.mark.parametrize("case", ["c1_b3_n52_all_grads", "c4_b2_n60_k52_all"])
def test_int8_label_matrices_equal_the_onehot_path_bit_for_bit(case):
    """ChangeDetector fed the loader's label matrices (int8 [B,S,S]) must give exactly what it
    gives on process_matrix's one-hot tensors (utils/mimic_utils.py:119-149) -- every output and every parameter
    gradient, including the label-bias tables that the labels index -- and onehot_adj(int8) == process_matrix."""
    from ekaid_b200.functions import onehot_adj
    from ekaid_b200.step import select_fields
    from oracle import ekaid_oracle as O
    dev = _dev()
    z, meta = load_case(case)
    sd, inp, batch = case_inputs(meta)
    dinp = to_dev(inp, dev)
    raw = tuple(t.to(dev) for t in select_fields(batch))
    assert raw[2].dtype == torch.int8 and raw[2].dim() == 3
    N = meta["N"]
    assert torch.equal(onehot_adj(raw[2], N, 11).cpu(), O.process_matrix(batch[6], N, 11))
    assert torch.equal(onehot_adj(raw[4], N, 3).cpu(), O.process_matrix(batch[8], N, 3))
    for precision in ("fp32", "bf16"):
        res = []
        for inputs in (dinp, raw[:9]):
            m = build_model(meta, sd, precision, dev)
            outs = m(*inputs)
            _loss(outs).backward()
            res.append(([o.detach().clone() for o in outs],
                        {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
        for a, b in zip(res[0][0], res[1][0]):
            assert torch.equal(a, b)
        assert set(res[0][1]) == set(res[1][1])
        for k in res[0][1]:
            # fp32 path: same kernels in the same order except the table gradient (per-label select vs fma with a one-hot).
            # 16-bit path: split-K partial sums land through fp32 atomics (order varies run to run) and a last-bit difference
            # of a wgrad feeds bf16-rounded gradients downstream; scalar gains are cancelling projections of those
            a, b = res[1][1][k].double(), res[0][1][k].double()
            if precision == "fp32":
                assert rel_err(a, b) < 1e-5, (precision, k)
            elif float(b.abs().max()) >= 1e-3:      # (below: analytically zero, e.g. the key bias -- noise on both sides)
                # run-to-run noise of the 16-bit path, in the gradient metric of the parity tests (relative L2)
                # (scalar gains of weight-normalised layers are cancelling projections: the widest band)
                assert float((a - b).norm() / b.norm()) < (8e-2 if a.numel() == 1 else 3e-2), (precision, k)


def test_eval_pass_and_second_model_between_training_steps_do_not_interfere():
    """The reference's loop runs a validation pass inside training (train_mimic.py:292-345) and nothing stops a process
    from holding two models: an eval forward of the SAME model, and a forward + backward of a SECOND ChangeDetector, between
    two training steps must leave the training trajectory (losses, parameters) exactly as it is without them."""
    from ekaid_b200 import functions
    from ekaid_b200.step import GraphFusionStep, expand_adjacency, select_fields
    from ekaid_b200.synthetic import synthetic_batch
    dev = _dev()
    z, meta = load_case("c0_b2_n52_all")
    sd, inp, _ = case_inputs(meta)
    batches = [tuple(t.to(dev) for t in select_fields(synthetic_batch(2, 52, seed=s))) for s in (11, 12, 13)]
    runs = {}
    try:
        for mode in ("plain", "interleaved"):
            functions.GRAD_SLOTS.clear()
            m = build_model(meta, sd, "fp32", dev)
            m.train()
            m.dropout_override = 0.0            # same function in both runs regardless of how often the seed advances
            step = GraphFusionStep(m, m.cfg, lr=1e-3)
            other = build_model(meta, sd, "fp32", dev) if mode == "interleaved" else None
            losses = []
            for i, b in enumerate(batches):
                losses.append(float(step.train_step(expand_adjacency(b, m.cfg), b[9], b[10].float())))
                if mode == "interleaved":
                    m.eval()
                    with torch.no_grad():
                        ev = m(*expand_adjacency(batches[(i + 1) % 3], m.cfg)[:9])
                    assert torch.isfinite(ev[3]).all()
                    m.train()
                    o = other(*expand_adjacency(b, m.cfg)[:9])          # a second model, with its own backward
                    _loss(o).backward()
                    assert other.img.weight.grad is not None
                    other.zero_grad()
            runs[mode] = (losses, {k: p.detach().clone() for k, p in m.named_parameters()})
    finally:
        functions.GRAD_SLOTS.clear()
    assert runs["plain"][0] == runs["interleaved"][0]
    for k in runs["plain"][1]:
        assert torch.equal(runs["plain"][1][k], runs["interleaved"][1][k]), k


def test_loader_batches_drive_the_captured_step():
    """ekaid_b200.loader (the reference's dataset schema -> rcc_collate 13-tuples, int8 label matrices, pinned) feeding
    StepPipeline: same losses as replaying the same batches from device memory."""
    from ekaid_b200 import functions
    from ekaid_b200.loader import RCCArrays, RCCDataset, batches
    from ekaid_b200.step import GraphFusionStep, select_fields
    dev = _dev()
    z, meta = load_case("c0_b2_n52_all")
    sd, inp, _ = case_inputs(meta)
    ds = RCCDataset(RCCArrays.synthetic(8, 52, seed=31))
    host = [select_fields(b) for b in batches(ds, 2, compact=True, pin=True)]
    assert len(host) == 4 and host[0][2].dtype == torch.int8 and host[0][0].is_pinned()
    res = {}
    try:
        for mode in ("device", "pipeline"):
            functions.GRAD_SLOTS.clear()
            m = build_model(meta, sd, "fp32", dev)
            step = GraphFusionStep(m, m.cfg, lr=1e-3)
            step.capture(tuple(t.to(dev) for t in host[0]), train=True, warmup=2)
            with torch.no_grad():
                m.load_state_dict(sd)
            step.opt.m.zero_()
            step.opt.v.zero_()
            step.opt.pow_state.fill_(1.0)
            if mode == "device":
                res[mode] = [float(step.replay(tuple(t.to(dev) for t in hb))) for hb in host]
            else:
                pipe = step.pipeline()
                pipe.prefetch(host[0])
                out, pending = [], None
                for i in range(len(host)):
                    if i + 1 < len(host):
                        pipe.prefetch(host[i + 1])
                    k = pipe.run()
                    if pending is not None:
                        out.append(pipe.result(pending))
                    pending = k
                out.append(pipe.result(pending))
                res[mode] = out
    finally:
        functions.GRAD_SLOTS.clear()
    assert res["device"] == pytest.approx(res["pipeline"], rel=1e-6)
    assert len(set(round(x, 6) for x in res["device"])) == 4
