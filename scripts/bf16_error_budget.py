"""CPU emulation of where the tensor-core path rounds, to budget the error of `input_attended` (analysis tooling).

The product's bf16 path keeps the residual stream, softmax and all reductions in fp32 and rounds GEMM operands /
stored activations to 16 bits at the points named below.  This script re-runs the oracle's closed form in fp64 with a
rounding hook at every such point, so each point can be switched between {none, bf16, fp16} and the effect on every
output read off -- in particular on `input_attended = attended_2 - attended_1`, which cancels ~40x.

    python scripts/bf16_error_budget.py            # table: one point at a time, then chosen combinations
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import case_inputs, load_case  # noqa: E402
from oracle import ekaid_oracle as O  # noqa: E402

POINTS = ["x_in", "w_img", "X_op", "qv_op", "w_sw", "sf", "w_qk", "w_z", "q", "k", "z", "p", "cat_x", "cat_diff", "cat_xs",
          "w_cg", "cx_gt", "w_e", "question"]


def rnd(t, how):
    if how == "none":
        return t
    if how == "bf16":
        return t.float().bfloat16().double()
    if how == "fp16":
        return t.float().half().double()
    if how == "bf16x2":          # hi + lo planes
        hi = t.float().bfloat16()
        lo = (t.float() - hi.float()).bfloat16()
        return hi.double() + lo.double()
    raise ValueError(how)


def forward(sd, inp, cfg):
    """fp64 closed form with rounding hooks; cfg: {point: how}."""
    r = lambda name, t: rnd(t, cfg.get(name, "none"))     # noqa: E731
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    i1, i2, d_adj, q_adj, d_sem, q_sem, d_bb, q_bb, question = inp
    B, N, D = i1.shape[0], i1.shape[1], 1024
    H, dh, K = 4, 256, min(52, N)
    Wimg = r("w_img", sd["img.weight"])
    X = torch.cat([r("x_in", i1.double()), r("x_in", i2.double())], 0) @ Wimg.t() + sd["img.bias"]       # [2B,N,D]
    qv = question_path(sd, question, cfg.get("question", "none"))
    qv2 = torch.cat([qv, qv], 0)
    adjs = {"sem": torch.cat([d_sem, q_sem], 0).double(), "spa": torch.cat([d_adj, q_adj], 0).double()}
    pe = torch.cat([O.position_embedding(O.position_matrix(d_bb, 52), 64), O.position_embedding(O.position_matrix(q_bb, 52), 64)], 0)
    for R, key in ((O.REL_SEM, "sem"), (O.REL_SPA, "spa"), (O.REL_IMP, "imp")):
        Wsw = r("w_sw", O.wn(sd, R + ".self_weights.main.1"))
        flag = (X.sum(-1, keepdim=True) != 0).double()
        qpart = r("qv_op", qv2) @ Wsw[:, D:].t()
        sf = r("X_op", X) @ Wsw[:, :D].t() + flag * qpart[:, None, :] + sd[R + ".self_weights.main.1.bias"]
        sf = r("sf", sf)
        nn_ = R + ".neighbor_net.1"
        q = r("q", sf @ r("w_qk", O.wn(sd, nn_ + ".query.main.1")).t() + sd[nn_ + ".query.main.1.bias"])
        k = r("k", sf[:, :K] @ r("w_qk", O.wn(sd, nn_ + ".key.main.1")).t() + sd[nn_ + ".key.main.1.bias"])
        q = q.view(2 * B, N, H, dh).transpose(1, 2)
        k = k.view(2 * B, K, H, dh).transpose(1, 2)
        aff = (q @ k.transpose(2, 3)) / math.sqrt(dh)
        aff = aff.transpose(1, 2)
        if key == "imp":
            pf = F.relu(pe.double().reshape(2 * B, -1, 64) @ O.wn(sd, nn_ + ".pair_pos_fc1.main.1").t() + sd[nn_ + ".pair_pos_fc1.main.1.bias"])
            aff = aff + torch.log(torch.clamp(pf.view(2 * B, -1, K, H).transpose(2, 3), min=1e-6))
            adj_t = X.new_ones(2 * B, N, K, 1)
        else:
            adj_t = adjs[key].transpose(1, 2)[:, :, :K, :]
        cond = adj_t.sum(-1)
        bias = (adj_t @ O.wn(sd, R + ".bias.main.0").t()).squeeze(-1)
        affT = aff.transpose(2, 3)
        masked = torch.where(cond.unsqueeze(3) > 0, affT, torch.full_like(affT, -9e15)) + bias.unsqueeze(3)
        P = r("p", F.softmax(masked.transpose(2, 3), 3))                       # [2B,N,H,K]
        Wo = r("w_z", sd[nn_ + ".linear_out_2.weight"])                       # [D, H*D]
        Z = r("z", sf[:, :K] @ Wo.t().reshape(D, H * D).contiguous().view(D, H * D)) if False else None
        # Z_h = sf W_out[:, hD:(h+1)D]^T
        out = 0
        for h in range(H):
            Zh = r("z", sf[:, :K] @ Wo[:, h * D:(h + 1) * D].t())              # [2B,K,D]
            out = out + P[:, :, h, :] @ Zh
        out = out + sd[nn_ + ".linear_out_2.bias"]
        X = X + F.relu(out + out)
    c = 0.333
    X = c * X + c * X + (1 - c - c) * X
    Xb, Xa = X[:B], X[B:]
    diff = Xa - Xb
    Wc1, Wc2 = r("w_cg", sd["context1.weight"]), r("w_cg", sd["context2.weight"])
    Wg1, Wg2 = r("w_cg", sd["gate1.weight"]), r("w_cg", sd["gate2.weight"])
    We = r("w_e", sd["embed.0.weight"])
    dq = r("cat_diff", diff)
    outs = []
    for Xi in (Xb, Xa):
        xo = r("cat_x", Xi)
        ctx = r("cx_gt", torch.tanh(dq @ Wc1.t() + xo @ Wc2.t() + sd["context2.bias"]))
        gate = r("cx_gt", torch.sigmoid(dq @ Wg1.t() + xo @ Wg2.t() + sd["gate2.bias"]))
        xs = r("cat_xs", gate * ctx)
        e = F.relu(torch.cat([xo, dq, xs], -1) @ We.t() + sd["embed.0.bias"])
        att = torch.sigmoid(e @ sd["att.weight"].t() + sd["att.bias"])
        outs.append((att.transpose(1, 2), (Xi * att).sum(1)))
    ia = outs[1][1] - outs[0][1]
    pred = ia @ sd["fc1.weight"].t() + sd["fc1.bias"]
    return {"att_bef": outs[0][0], "att_aft": outs[1][0], "attended_1": outs[0][1], "attended_2": outs[1][1],
            "input_attended": ia, "pred": pred}


def question_path(sd, question, how):
    """language_model.py:48-156 with the product's rounding points: E, W_ih, W_hh, h_{t-1} (GEMM operands), Hd, W1."""
    rq = lambda t: rnd(t, how)     # noqa: E731
    x = rq(O.word_embedding(sd, question))
    w_ih, w_hh = rq(sd["q_emb.rnn.weight_ih_l0"]), rq(sd["q_emb.rnn.weight_hh_l0"])
    b_ih, b_hh = sd["q_emb.rnn.bias_ih_l0"], sd["q_emb.rnn.bias_hh_l0"]
    B, L, _ = x.shape
    Hd = w_hh.shape[1]
    h = x.new_zeros(B, Hd)
    outs = []
    gi_all = x @ w_ih.t() + b_ih
    for t in range(L):
        gi = gi_all[:, t]
        gh = rq(h) @ w_hh.t() + b_hh
        r_ = torch.sigmoid(gi[:, :Hd] + gh[:, :Hd])
        z_ = torch.sigmoid(gi[:, Hd:2 * Hd] + gh[:, Hd:2 * Hd])
        n_ = torch.tanh(gi[:, 2 * Hd:] + r_ * gh[:, 2 * Hd:])
        h = (1 - z_) * n_ + z_ * h
        outs.append(h)
    hs = torch.stack(outs, 1)
    w1 = rq(O.wn(sd, "q_att.W1_self_att_q.main.1"))
    a1 = rq(torch.tanh(rq(hs).reshape(-1, Hd) @ w1.t() + sd["q_att.W1_self_att_q.main.1.bias"]))
    a = (a1 @ O.wn(sd, "q_att.W2_self_att_q.main.0").t() + sd["q_att.W2_self_att_q.main.0.bias"]).view(B, L)
    weight = F.softmax(a.t(), dim=1).contiguous().view(-1, 1, L)
    return torch.bmm(weight, hs).view(-1, Hd)


def err(a, b):
    return float((a - b).abs().max() / b.abs().max())


def main():
    cases = sys.argv[1:] or ["c0_b2_n52_all", "c1_b3_n52_all_grads"]
    for name in cases:
        z, meta = load_case(name)
        sd, inp, _ = case_inputs(meta)
        ref = forward(sd, inp, {})
        print("== %s   |attended| %.1f  |input_attended| %.2f  |pred| %.2f" % (
            name, float(ref["attended_2"].abs().max()), float(ref["input_attended"].abs().max()), float(ref["pred"].abs().max())))
        allbf = {p: "bf16" for p in POINTS}
        allbf["p"] = "bf16x2"

        def show(tag, cfg):
            o = forward(sd, inp, cfg)
            print("  %-46s ia %.2e  pred %.2e  att %.1e  attended %.1e" % (
                tag, err(o["input_attended"], ref["input_attended"]), err(o["pred"], ref["pred"]),
                err(o["att_aft"], ref["att_aft"]), err(o["attended_2"], ref["attended_2"])))

        show("all points bf16 (today's path)", allbf)
        for p in POINTS:
            show("only %s bf16" % p, {p: "bf16"})
        for tag, over in (
            ("z fp16", {"z": "fp16"}),
            ("z fp16, p fp16", {"z": "fp16", "p": "fp16"}),
            ("z,sf fp16", {"z": "fp16", "sf": "fp16"}),
            ("z,sf,w_z fp16", {"z": "fp16", "sf": "fp16", "w_z": "fp16"}),
            ("z,sf,w_z,w_qk fp16", {"z": "fp16", "sf": "fp16", "w_z": "fp16", "w_qk": "fp16"}),
            ("z,sf,w_z,w_qk,X_op,w_sw fp16", {"z": "fp16", "sf": "fp16", "w_z": "fp16", "w_qk": "fp16", "X_op": "fp16", "w_sw": "fp16"}),
            ("z,sf,w_z,w_qk,X_op,w_sw,q,k fp16", {"z": "fp16", "sf": "fp16", "w_z": "fp16", "w_qk": "fp16", "X_op": "fp16", "w_sw": "fp16", "q": "fp16", "k": "fp16"}),
            ("+ x_in, w_img fp16", {"z": "fp16", "sf": "fp16", "w_z": "fp16", "w_qk": "fp16", "X_op": "fp16", "w_sw": "fp16", "q": "fp16", "k": "fp16", "x_in": "fp16", "w_img": "fp16"}),
            ("everything fp16", {p: "fp16" for p in POINTS}),
            ("PLAN: weights,question,qv,sf,z,p fp16; X_op,x_in,cat,q,k bf16", {"w_img": "fp16", "w_sw": "fp16", "w_qk": "fp16", "w_z": "fp16", "w_cg": "fp16", "w_e": "fp16", "question": "fp16", "qv_op": "fp16", "sf": "fp16", "z": "fp16", "p": "fp16"}),
            ("PLAN without z", {"w_img": "fp16", "w_sw": "fp16", "w_qk": "fp16", "w_z": "fp16", "w_cg": "fp16", "w_e": "fp16", "question": "fp16", "qv_op": "fp16", "sf": "fp16"}),
            ("PLAN + X_op fp16", {"w_img": "fp16", "w_sw": "fp16", "w_qk": "fp16", "w_z": "fp16", "w_cg": "fp16", "w_e": "fp16", "question": "fp16", "qv_op": "fp16", "sf": "fp16", "z": "fp16", "p": "fp16", "X_op": "fp16"}),
            ("weights + question + qv fp16 only", {"w_img": "fp16", "w_sw": "fp16", "w_qk": "fp16", "w_z": "fp16", "w_cg": "fp16", "w_e": "fp16", "question": "fp16", "qv_op": "fp16"}),
            ("z bf16x2", {"z": "bf16x2"}),
            ("z,sf bf16x2", {"z": "bf16x2", "sf": "bf16x2"}),
        ):
            cfg = dict(allbf)
            cfg.update(over)
            show(tag, cfg)


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
