"""CPU oracle for the EKAID graph+fusion hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a closed-form restatement (stock torch CPU ops, fp32 or fp64) of what the reference
`ChangeDetector.forward` computes for setting='mode2'.  It exists so the CUDA product path can be
checked; it is never imported by `ekaid_b200/` (only by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`).

Pinning: the reference ships no tests or golden vectors for this path, so the oracle is
pinned against outputs of the reference itself, produced in the build container by
`tests/golden/make_golden.py` (which imports /root/reference/model) and committed under
`tests/golden/*.npz`; `tests/test_oracle_golden.py` replays them.

Every function cites the reference lines it restates (paths relative to /root/reference/model).
Quirk numbers (Q1..Q13) refer to SURVEY.md section 3.4.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

REL_SEM = "semantic_relation.explicit_relation"
REL_SPA = "spatial_relation.explicit_relation"
REL_IMP = "imp_relation.implicit_relation"


def wn(sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """Legacy weight_norm(dim=None): w = g * v / ||v||_F  (models/fc.py:33-34)."""
    v = sd[prefix + ".weight_v"]
    g = sd[prefix + ".weight_g"]
    return v * (g / v.norm())


# ----------------------------------------------------------------------------------------------
# question path
# ----------------------------------------------------------------------------------------------
def word_embedding(sd, question: Tensor) -> Tensor:
    """models/language_model.py:48-53  (two tables, concatenated; dropout p=0).  Both tables are
    nn.Embedding(ntoken + 1, 300, padding_idx=ntoken) (:26-29): row ntoken never receives a gradient."""
    w, w_ = sd["w_emb.emb.weight"], sd["w_emb.emb_.weight"]
    return torch.cat((F.embedding(question, w, padding_idx=w.shape[0] - 1),
                      F.embedding(question, w_, padding_idx=w_.shape[0] - 1)), 2)


def gru_all(sd, x: Tensor) -> Tensor:
    """models/language_model.py:106-115  1-layer unidirectional GRU, h0 = 0, all steps returned.

    Gate order in weight_ih/hh is (r, z, n) -- torch.nn.GRU semantics."""
    w_ih, w_hh = sd["q_emb.rnn.weight_ih_l0"], sd["q_emb.rnn.weight_hh_l0"]
    b_ih, b_hh = sd["q_emb.rnn.bias_ih_l0"], sd["q_emb.rnn.bias_hh_l0"]
    B, L, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    outs = []
    gi_all = x @ w_ih.t() + b_ih
    for t in range(L):
        gi = gi_all[:, t]
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        outs.append(h)
    return torch.stack(outs, 1)


def question_self_attention(sd, h: Tensor, drop: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """models/language_model.py:127-156 incl. quirk Q4 (softmax over the BATCH axis, then the
    contiguous [L,B] result is re-viewed as [B,1,L]).
    drop (train mode with GIVEN masks; multipliers 0 or 1/(1-p)): 'w1' [B, L, Hd] = the Dropout in front of W1's Linear
    (fc.py:25-32), 'qv' [B, Hd] = self.drop on the pooled vector (:155)."""
    B, L, Hd = h.shape
    w1 = wn(sd, "q_att.W1_self_att_q.main.1")
    b1 = sd["q_att.W1_self_att_q.main.1.bias"]
    w2 = wn(sd, "q_att.W2_self_att_q.main.0")
    b2 = sd["q_att.W2_self_att_q.main.0.bias"]
    hin = h if drop is None else h * drop["w1"]
    a1 = torch.tanh(hin.reshape(-1, Hd) @ w1.t() + b1)
    a = (a1 @ w2.t() + b2).view(B, L)
    weight = F.softmax(a.t(), dim=1).contiguous().view(-1, 1, L)      # Q4
    out = torch.bmm(weight, h).view(-1, Hd)
    return out if drop is None else out * drop["qv"]


def question_vector(sd, question: Tensor, drop: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """models/modules.py:200-206."""
    return question_self_attention(sd, gru_all(sd, word_embedding(sd, question)), drop)


# ----------------------------------------------------------------------------------------------
# geometry (implicit relation)
# ----------------------------------------------------------------------------------------------
def position_matrix(bbox: Tensor, nongt_dim: int) -> Tensor:
    """utils/mimic_utils.py:152-190.  NOTE (Q13): the slice is on the ROW axis, so the result is
    [B, min(nongt,N), N, 4], not the documented [B, N, nongt, 4]."""
    xmin, ymin, xmax, ymax = torch.split(bbox, 1, dim=-1)
    w = xmax - xmin + 1.0
    h = ymax - ymin + 1.0
    cx = 0.5 * (xmin + xmax)
    cy = 0.5 * (ymin + ymax)
    dx = torch.log(torch.clamp(torch.abs((cx - cx.transpose(1, 2)) / w), min=1e-3))
    dy = torch.log(torch.clamp(torch.abs((cy - cy.transpose(1, 2)) / h), min=1e-3))
    dw = torch.log(w / w.transpose(1, 2))
    dh = torch.log(h / h.transpose(1, 2))
    return torch.stack([m[:, :nongt_dim] for m in (dx, dy, dw, dh)], 3)


def position_embedding(pos_mat: Tensor, feat_dim: int = 64, wave_length: float = 1000.0) -> Tensor:
    """utils/mimic_utils.py:192-208.  dim_mat is computed in fp32 (torch.ones default dtype) and
    promoted to the dtype of pos_mat (fp64 when the boxes arrive as .double())."""
    feat_range = torch.arange(0, feat_dim / 8)
    dim_mat = torch.pow(torch.ones((1,)) * wave_length, (8.0 / feat_dim) * feat_range).view(1, 1, 1, -1)
    dim_mat = dim_mat.to(pos_mat.device)      # (bench.py's GPU-eager comparator runs this restatement on cuda:0)
    div = (100.0 * pos_mat).unsqueeze(4) / dim_mat
    emb = torch.cat([torch.sin(div), torch.cos(div)], -1)
    return emb.view(emb.shape[0], emb.shape[1], emb.shape[2], feat_dim)


# ----------------------------------------------------------------------------------------------
# one relation encoder step (GAT)
# ----------------------------------------------------------------------------------------------
def gat_relation(sd, R: str, X: Tensor, qv: Tensor, adj: Optional[Tensor], pos_emb: Optional[Tensor],
                 num_heads: int, nongt_dim: int, return_aux: bool = False, relu_mask: Optional[Tensor] = None,
                 drop: Optional[Dict[str, Tensor]] = None):
    """X <- X + GAT(cat(X, q), adj)   (models/relation_encoder.py:57-84 / :112-132,
    models/graph_att.py:53-106, models/graph_att_layer.py:60-178), eval mode.

    adj: [B,N,N,label] float (explicit) or None (implicit: all-ones, bias is a constant shift, Q5).
    Only direction 1 (transposed adjacency) is live and its output is doubled (Q2).
    drop (train mode with GIVEN masks; multipliers 0 or 1/(1-p)), every FCNet applies Dropout BEFORE its Linear
    (fc.py:25-32): 'vq' [B,N,D+Dq] self_weights' input, 'q' / 'k' [B,N,D] the query / key inputs, 'pos' [B,N*K,64]
    pair_pos_fc1's input, 'out' [B,N,D] GAttNet.dropout on the doubled output before the ReLU (graph_att.py:103-104)."""
    B, N, D = X.shape
    K = min(nongt_dim, N)
    Hn = num_heads
    dh = D // Hn
    # q_expand_v_cat with mask=True (Q10)  relation_encoder.py:19-29
    qe = qv.view(B, 1, -1).expand(B, N, qv.shape[1]).clone()
    qe = qe * (X.sum(-1, keepdim=True) != 0).to(X.dtype)
    vq = torch.cat((X, qe), -1)
    if drop is not None:
        vq = vq * drop["vq"]
    sf = vq @ wn(sd, R + ".self_weights.main.1").t() + sd[R + ".self_weights.main.1.bias"]
    nn_ = R + ".neighbor_net.1"                                       # Q2: direction 1 only
    sq, sk = (sf, sf) if drop is None else (sf * drop["q"], sf * drop["k"])
    q = (sq @ wn(sd, nn_ + ".query.main.1").t() + sd[nn_ + ".query.main.1.bias"]).view(B, N, Hn, dh).transpose(1, 2)
    k = (sk[:, :K] @ wn(sd, nn_ + ".key.main.1").t() + sd[nn_ + ".key.main.1.bias"]).view(B, K, Hn, dh).transpose(1, 2)
    aff = (1.0 / math.sqrt(float(dh))) * (q @ k.transpose(2, 3))      # [B,H,N,K]
    aff = aff.transpose(1, 2)                                         # [B,N,H,K]
    if pos_emb is not None:
        # Q7 / Q13  graph_att_layer.py:113-135
        pe = pos_emb.to(X.dtype).reshape(B, -1, pos_emb.shape[-1])
        if drop is not None:
            pe = pe * drop["pos"]
        pf = F.relu(pe @ wn(sd, nn_ + ".pair_pos_fc1.main.1").t() + sd[nn_ + ".pair_pos_fc1.main.1.bias"])
        aw = pf.view(B, -1, K, Hn).transpose(2, 3)
        aff = aff + torch.log(torch.clamp(aw, min=1e-6))
    if adj is None:
        adj_t = X.new_ones(B, N, K, 1)
    else:
        adj_t = adj.to(X.dtype).transpose(1, 2)[:, :, :K, :]          # graph_att.py:76,88
    cond = adj_t.sum(-1)                                              # graph_att.py:89
    bias = (adj_t @ wn(sd, R + ".bias.main.0").t()).squeeze(-1)       # Q5  graph_att.py:92
    affT = aff.transpose(2, 3)                                        # [B,N,K,H]
    masked = torch.where(cond.unsqueeze(3) > 0, affT, torch.full_like(affT, -9e15))   # Q6
    masked = masked + bias.unsqueeze(3)
    P = F.softmax(masked.transpose(2, 3), 3)                          # [B,N,H,K]
    out_t = P.reshape(B, N * Hn, K) @ sf[:, :K]                       # Q3
    out = out_t.reshape(B * N, Hn * D) @ sd[nn_ + ".linear_out_2.weight"].t() + sd[nn_ + ".linear_out_2.bias"]
    out = out.view(B, N, D)
    out2 = (out + out) if drop is None else (out + out) * drop["out"]
    if relu_mask is None:
        Xn = X + F.relu(out2)                                         # Q2 doubling; graph_att.py:102-104
    else:
        # gradient checks only: use the active set chosen by the implementation under test, so that ReLU kinks
        # (pre-activations within rounding distance of 0) do not show up as gradient differences
        Xn = X + out2 * relu_mask.to(out.dtype).view_as(out)
    if return_aux:
        return Xn, {"self_feat": sf, "P": P, "out": out}
    return Xn


# ----------------------------------------------------------------------------------------------
# whole ChangeDetector forward, setting='mode2'
# ----------------------------------------------------------------------------------------------
def change_detector_forward(sd: Dict[str, Tensor], input_1: Tensor, input_2: Tensor,
                            d_adj: Tensor, q_adj: Tensor, d_sem_adj: Tensor, q_sem_adj: Tensor,
                            d_bb: Tensor, q_bb: Tensor, question: Tensor, *, graph: str = "all",
                            num_heads: int = 4, nongt_dim: int = 52, pos_emb_dim: int = 64,
                            coef_sem: float = 0.333, coef_spa: float = 0.333, return_aux: bool = False,
                            relu_masks: Optional[Sequence[Tensor]] = None, drop: Optional[Dict] = None):
    """models/modules.py:169-313 (eval mode, empty_image False, feature_mode != 'mode0').
    drop: train mode with GIVEN dropout multipliers (the reference draws them from torch's generator; a test that wants
    to compare a train-mode forward hands both sides the same ones): {'question': {...}, 'sem' / 'spa' / 'imp':
    ({...bef}, {...aft}) as gat_relation takes them, 'ctx' / 'gate': (bef, aft) [B,N,D] (modules.py:279-287),
    'embed': (bef, aft) [B,N,dim] (Linear -> Dropout -> ReLU, modules.py:105-111)}."""
    dt = input_1.dtype
    aux = {}
    Xb = input_1 @ sd["img.weight"].t() + sd["img.bias"]              # modules.py:195-196
    Xa = input_2 @ sd["img.weight"].t() + sd["img.bias"]
    qv = question_vector(sd, question, None if drop is None else drop["question"])        # modules.py:200-206
    dr = (lambda key, i: None) if drop is None else (lambda key, i: drop[key][i])
    aux["qv"] = qv
    kw = dict(num_heads=num_heads, nongt_dim=nongt_dim)
    # relu_masks (gradient checks only): one [2*B*N, D] mask per relation in execution order (main rows, then
    # reference rows), then one [2*B*N, dim] mask for the embed ReLU
    masks = list(relu_masks) if relu_masks is not None else None
    BN = Xb.shape[0] * Xb.shape[1]

    def nxt():
        if masks is None:
            return None, None
        mk = masks.pop(0)
        return mk[:BN], mk[BN:]

    if graph in ("semantic", "all"):                                  # modules.py:216-218
        mb, ma = nxt()
        Xb = gat_relation(sd, REL_SEM, Xb, qv, d_sem_adj, None, relu_mask=mb, drop=dr("sem", 0), **kw)
        Xa = gat_relation(sd, REL_SEM, Xa, qv, q_sem_adj, None, relu_mask=ma, drop=dr("sem", 1), **kw)
    if graph in ("spatial", "all", "i+s"):                            # modules.py:221-223
        mb, ma = nxt()
        Xb = gat_relation(sd, REL_SPA, Xb, qv, d_adj, None, relu_mask=mb, drop=dr("spa", 0), **kw)
        Xa = gat_relation(sd, REL_SPA, Xa, qv, q_adj, None, relu_mask=ma, drop=dr("spa", 1), **kw)
    if graph in ("implicit", "all", "i+s"):                           # modules.py:226-230
        pe_b = position_embedding(position_matrix(d_bb, nongt_dim), pos_emb_dim)
        pe_a = position_embedding(position_matrix(q_bb, nongt_dim), pos_emb_dim)
        mb, ma = nxt()
        Xb = gat_relation(sd, REL_IMP, Xb, qv, None, pe_b, relu_mask=mb, drop=dr("imp", 0), **kw)
        Xa = gat_relation(sd, REL_IMP, Xa, qv, None, pe_a, relu_mask=ma, drop=dr("imp", 1), **kw)
    emb, ema = nxt()
    # Q1: input_bef1/2/3 alias ONE tensor  (modules.py:233-247)
    if graph == "all":
        Xb = coef_sem * Xb + coef_spa * Xb + (1 - coef_sem - coef_spa) * Xb
        Xa = coef_sem * Xa + coef_spa * Xa + (1 - coef_sem - coef_spa) * Xa
    elif graph == "i+s":
        Xb = (Xb + Xb) / 2
        Xa = (Xa + Xa) / 2
    diff = Xa - Xb                                                    # modules.py:250
    aux["X_bef"], aux["X_aft"], aux["diff"] = Xb, Xa, diff
    c1, g1 = sd["context1.weight"], sd["gate1.weight"]
    c2, g2 = sd["context2.weight"], sd["gate2.weight"]

    def fuse(X, i):                                                   # modules.py:278-288
        ctx = torch.tanh(diff @ c1.t() + X @ c2.t() + sd["context2.bias"])
        gate = torch.sigmoid(diff @ g1.t() + X @ g2.t() + sd["gate2.bias"])
        if drop is not None:
            ctx, gate = ctx * drop["ctx"][i], gate * drop["gate"][i]
        return gate * ctx

    def pool(X, Xs, emask, i):                                        # modules.py:290-308
        pre = torch.cat([X, diff, Xs], -1) @ sd["embed.0.weight"].t() + sd["embed.0.bias"]
        if drop is not None:
            pre = pre * drop["embed"][i]
        e = F.relu(pre) if emask is None else pre * emask.to(pre.dtype).view_as(pre)
        att = torch.sigmoid(e @ sd["att.weight"].t() + sd["att.bias"])          # [B,N,1]
        return att.transpose(1, 2), (X * att).sum(1)

    att_b, attended_1 = pool(Xb, fuse(Xb, 0), emb, 0)
    att_a, attended_2 = pool(Xa, fuse(Xa, 1), ema, 1)
    input_attended = attended_2 - attended_1                          # modules.py:309
    pred = input_attended @ sd["fc1.weight"].t() + sd["fc1.bias"]     # modules.py:310
    outs = (pred.to(dt), att_b, att_a, attended_1, attended_2, input_attended)
    return (outs, aux) if return_aux else outs


# ----------------------------------------------------------------------------------------------
# step glue: integer label matrix -> one-hot planes
# ----------------------------------------------------------------------------------------------
def process_matrix(adj: Tensor, num_objects: int, label_num: int) -> Tensor:
    """utils/mimic_utils.py:119-149: plane c is (adj == c+1) as float32; label 0 = no edge."""
    a = adj[:, :num_objects, :num_objects]
    return torch.stack([(a == i).to(torch.float32) for i in range(1, label_num + 1)], 3)


def spatial_adj_matrix(boxes, size: int = 100, lx: float = 1024.0, ly: float = 1024.0) -> Tensor:
    """get_adj_matrix with bbox_relation_type / reverse_type ("feature extraction/ana_bbox_generator.py":213-259,
    266-302,320-335), scalar double arithmetic pair by pair like the reference: boxes [B,N,4] -> int64 [B,S,S],
    S = max(size, N).  Pinned by tests/golden/spatial_labels.npz (made by the reference's own functions)."""
    import math
    bb = torch.as_tensor(boxes, dtype=torch.float64).tolist()
    n = len(bb[0]) if bb else 0
    S = max(size, n)
    out = torch.zeros(len(bb), S, S, dtype=torch.int64)
    rev = (0, 2, 1, 3, 8, 9, 10, 11, 4, 5, 6, 7)

    def rel(a, b):
        if a[0] < b[0] and a[1] < b[1] and a[2] > b[2] and a[3] > b[3]:
            return 1
        if a[0] > b[0] and a[1] > b[1] and a[2] < b[2] and a[3] < b[3]:
            return 2
        iw = max(min(a[2], b[2]) - max(a[0], b[0]) + 1.0, 0.0)
        ih = max(min(a[3], b[3]) - max(a[1], b[1]) + 1.0, 0.0)
        inter = iw * ih
        uni = (a[2] - a[0] + 1.0) * (a[3] - a[1] + 1.0) + (b[2] - b[0] + 1.0) * (b[3] - b[1] + 1.0) - inter
        if inter / uni >= 0.5:
            return 3
        dx = (b[2] + b[0]) / 2 - (a[2] + a[0]) / 2
        dy = (b[3] + b[1]) / 2 - (a[3] + a[1]) / 2
        if math.sqrt(dx * dx + dy * dy) >= (lx + ly) / 3:
            return 0
        ang = math.atan2(dy, dx) / math.pi * 180
        if ang < 0:
            ang += 360
        return math.ceil(ang / 45) + 3

    for k, boxes_k in enumerate(bb):
        for i in range(n):
            for j in range(i, n):
                t = rel(boxes_k[i], boxes_k[j])
                out[k, i, j] = t
                out[k, j, i] = rev[t]
    return out


def semantic_adj_matrix(pred_classes, ana_classes, di_classes, kg_ana, small_adj, small_name2index, size: int = 100):
    """get_semantic_adj ("feature extraction/combine_dicts.py":106-151) pair by pair like the reference: pred_classes
    [B, T] ints (anatomy ids, then disease ids already offset by len(ana_classes); len(ana) + len(di) = background)
    -> int64 [B, S, S].  Pinned by tests/golden/semantic_labels.npz (made by the reference's own function)."""
    import numpy as np
    thing = list(ana_classes) + list(di_classes)
    ana_set, di_set = set(ana_classes), set(di_classes)
    pred = np.asarray(pred_classes)
    B, T = pred.shape
    S = max(size, T)
    out = np.zeros([B, S, S], dtype=np.int64)
    for b in range(B):
        pc = pred[b]
        for i in range(T):
            for j in range(i, T):
                if pc[i] == len(thing) or pc[j] == len(thing):
                    continue
                ni, nj = thing[pc[i]], thing[pc[j]]
                if kg_ana[ni] == kg_ana[nj] and ((ni in ana_set and nj in di_set) or (nj in ana_set and ni in di_set)):
                    out[b, i, j] = out[b, j, i] = 1
                if ni.lower() in small_name2index and nj.lower() in small_name2index:
                    v = max(small_adj[small_name2index[ni.lower()], small_name2index[nj.lower()]], out[b, i, j])
                    out[b, i, j] = out[b, j, i] = v
    return torch.from_numpy(out)


# ----------------------------------------------------------------------------------------------
# answer decoder (boundary consumer) -- greedy decode used for the arg-max token parity check
# ----------------------------------------------------------------------------------------------
def speaker_core(sp, xt, bef, aft, diff, state):
    """models/dynamic_speaker_change_pos.py:94-131 (eval mode)."""
    h, c = state
    prev_h = h[1]
    emb = F.relu(torch.cat([bef, diff, aft], 1) @ sp["core.embed.0.weight"].t() + sp["core.embed.0.bias"])

    def lstm(prefix, x, hc):
        gates = x @ sp[prefix + ".weight_ih"].t() + sp[prefix + ".bias_ih"] + hc[0] @ sp[prefix + ".weight_hh"].t() + sp[prefix + ".bias_hh"]
        i, f, g, o = gates.chunk(4, 1)
        c_new = torch.sigmoid(f) * hc[1] + torch.sigmoid(i) * torch.tanh(g)
        return torch.sigmoid(o) * torch.tanh(c_new), c_new

    h_mod, c_mod = lstm("core.module_att_lstm", torch.cat([emb, prev_h], 1), (h[0], c[0]))
    mw = F.softmax(h_mod @ sp["core.weight_fc.0.weight"].t() + sp["core.weight_fc.0.bias"], dim=1)
    vpos = F.relu(prev_h @ sp["core.pos1.0.weight"].t() + sp["core.pos1.0.bias"])
    dpos = vpos @ sp["core.weight_pos.weight"].t() + sp["core.weight_pos.bias"]
    ppos = F.softmax(dpos, dim=1) @ sp["core.pos2.weight"].t() + sp["core.pos2.bias"]
    feats = torch.stack([bef, diff, aft], 1)
    att_feat = (feats * mw.unsqueeze(2)).sum(1)
    gi = torch.cat([prev_h, ppos, att_feat], -1)
    g1 = F.relu(gi @ sp["core.gate1x.0.weight"].t() + sp["core.gate1x.0.bias"])
    gate = torch.sigmoid(g1 @ sp["core.gate2x.weight"].t() + sp["core.gate2x.bias"])
    h_lang, c_lang = lstm("core.lang_lstm", torch.cat([xt, gate * att_feat], 1), (h[1], c[1]))
    return h_lang, (torch.stack([h_mod, h_lang]), torch.stack([c_mod, c_lang])), dpos


def speaker_logprobs(sp, it, bef, aft, diff, state):
    """dynamic_speaker_change_pos.py:225-240."""
    xt = F.relu(sp["embed.0.weight"][it])
    out, state, dpos = speaker_core(sp, xt, bef, aft, diff, state)
    return F.log_softmax(out @ sp["logit.weight"].t() + sp["logit.bias"], dim=1), state, dpos


def speaker_greedy(sp, bef, aft, diff, seq_length: int, rnn_size: int):
    """dynamic_speaker_change_pos.py:287-357 with sample_max=1, beam_size=1."""
    B = bef.shape[0]
    state = (bef.new_zeros(2, B, rnn_size), bef.new_zeros(2, B, rnn_size))
    seq = torch.zeros(B, seq_length, dtype=torch.long)
    it = torch.full((B,), 2, dtype=torch.long)
    unfinished = None
    for t in range(seq_length + 1):
        logp, state, _ = speaker_logprobs(sp, it, bef, aft, diff, state)
        if t == 0:
            logp = logp.clone()
            logp[:, 0] = float("-inf")
        if t == seq_length:
            break
        it = logp.argmax(1)
        unfinished = (it > 0) if t == 0 else unfinished & (it > 0)
        it = it * unfinished.to(it.dtype)
        seq[:, t] = it
        if unfinished.sum() == 0:
            break
    return seq


def speaker_teacher_forced(sp, bef, aft, diff, seq: Tensor, seq_length: int, rnn_size: int):
    """dynamic_speaker_change_pos.py:182-222 (eval mode, ss_prob = 0): log-probs [B, seq_length, V]."""
    B = bef.shape[0]
    state = (bef.new_zeros(2, B, rnn_size), bef.new_zeros(2, B, rnn_size))
    V = sp["logit.weight"].shape[0]
    outs = [bef.new_zeros(B, V) for _ in range(seq_length)]
    for i in range(seq_length):
        if i >= 1 and int(seq[:, i].sum()) == 0:
            break
        logp, state, _ = speaker_logprobs(sp, seq[:, i], bef, aft, diff, state)
        outs[i] = logp
    return torch.stack(outs, 1)


def lm_criterion(logp: Tensor, target: Tensor, mask: Tensor) -> Tensor:
    """utils/utils.py:204-216."""
    target = target[:, :logp.size(1)]
    mask = mask[:, :logp.size(1)]
    out = -logp.gather(2, target.unsqueeze(2)).squeeze(2) * mask
    return out.sum() / mask.sum()
