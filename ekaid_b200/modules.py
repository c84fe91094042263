"""Drop-in nn.Modules for the EKAID graph+fusion hot path.

Same class names, constructor signatures, forward signatures, error behaviour and state_dict keys as the
reference modules (paths relative to /root/reference/model):

    ChangeDetector                      models/modules.py:81-313
    ExplicitRelationEncoder,
    ImplicitRelationEncoder,
    q_expand_v_cat                      models/relation_encoder.py:19-132
    GAttNet                             models/graph_att.py:17-106
    GraphSelfAttentionLayer             models/graph_att_layer.py:19-178
    FCNet                               models/fc.py:15-49
    WordEmbedding, QuestionEmbedding,
    QuestionSelfAttention               models/language_model.py:17-156

The module tree only owns parameters; the arithmetic runs in hand-written sm_100a kernels reached through
the C ABI (ekaid_b200.functions).  There is no PyTorch/CPU fallback: calling forward without a B200 raises.

`precision`: 'bf16' (tcgen05 tensor cores, 2e-2 parity) or 'fp32' (SIMT, 1e-4 parity).  Default from the
environment variable EKAID_B200_PRECISION, else 'bf16'; settable per module tree via `set_precision`.
"""
from __future__ import annotations

import math
import os
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    from torch.nn.utils import weight_norm as _weight_norm

from . import lib as _lib
from .functions import (PC, Drop, EdgeAttentionFn, Fork, FusionFn, GRUFn, LinearFn, QuestionFn, RelationFn, SmallLinearFn, WNormFn,
                        WNormManyFn, rng_advance, wn_many_compute)


def _default_precision() -> str:
    return os.environ.get("EKAID_B200_PRECISION", "bf16")


def _wn(module: nn.Module) -> nn.Module:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return _weight_norm(module, dim=None)


def wn_weight(lin: nn.Module) -> torch.Tensor:
    """Effective weight of a legacy weight_norm(dim=None) layer: g * v / ||v||_F (fc.py:33-34)."""
    v, g = lin.weight_v, lin.weight_g
    if v.is_cuda:
        return WNormFn.apply(v, g)
    return v * (g / v.norm())


class FCNet(nn.Module):
    """[Dropout] -> weight_norm(Linear, dim=None) -> [activation], stacked (models/fc.py:15-49)."""

    def __init__(self, dims, act='ReLU', dropout=0, bias=True):
        super().__init__()
        layers = []
        for i in range(len(dims) - 2):
            if 0 < dropout:
                layers.append(nn.Dropout(dropout))
            layers.append(_wn(nn.Linear(dims[i], dims[i + 1], bias=bias)))
            if '' != act and act is not None:
                layers.append(getattr(nn, act)())
        if 0 < dropout:
            layers.append(nn.Dropout(dropout))
        layers.append(_wn(nn.Linear(dims[-2], dims[-1], bias=bias)))
        if '' != act and act is not None:
            layers.append(getattr(nn, act)())
        self.main = nn.Sequential(*layers)
        self.precision = _default_precision()

    def linear(self, idx: int = -1) -> nn.Module:
        lins = [m for m in self.main if isinstance(m, nn.Linear)]
        return lins[idx]

    def forward(self, x):
        pc = PC(self.precision)
        out = x
        for m in self.main:
            if isinstance(m, nn.Linear):
                shp = out.shape
                if m.in_features % 8 or m.out_features % 8:
                    # a handful of inputs / outputs (label bias 3|11 -> 1, pair_pos_fc1 64 -> 4): one warp per output
                    y = SmallLinearFn.apply(out.reshape(-1, shp[-1]).float(), wn_weight(m), m.bias)
                else:
                    y, _ = LinearFn.apply(pc, out.reshape(-1, shp[-1]), None, wn_weight(m), m.bias)
                out = y.view(*shp[:-1], y.shape[-1])
            else:
                out = m(out)
        return out


def q_expand_v_cat(q, v, mask=True):
    """models/relation_encoder.py:19-29 (kept for API compatibility; the CUDA path never materialises it)."""
    q = q.view(q.size(0), 1, q.size(1))
    q_expand = q.expand(-1, v.shape[1], -1).clone()
    if mask:
        q_expand = q_expand * (v.sum(-1, keepdim=True) != 0).to(q_expand.dtype)
    return torch.cat((v, q_expand), dim=-1)


class GraphSelfAttentionLayer(nn.Module):
    """Parameter container with the reference layout (models/graph_att_layer.py:20-57).  Inside a relation
    encoder its arithmetic is executed by RelationFn."""

    def __init__(self, feat_dim, nongt_dim=20, pos_emb_dim=-1, num_heads=16, dropout=[0.2, 0.5]):
        super().__init__()
        self.fc_dim = num_heads
        self.feat_dim = feat_dim
        self.dim = (feat_dim, feat_dim, feat_dim)
        self.dim_group = (int(self.dim[0] / num_heads), int(self.dim[1] / num_heads), int(self.dim[2] / num_heads))
        self.num_heads = num_heads
        self.pos_emb_dim = pos_emb_dim
        if self.pos_emb_dim > 0:
            self.pair_pos_fc1 = FCNet([pos_emb_dim, self.fc_dim], None, dropout[0])
        self.query = FCNet([feat_dim, self.dim[0]], None, dropout[0])
        self.nongt_dim = nongt_dim
        self.key = FCNet([feat_dim, self.dim[1]], None, dropout[0])
        # never called in the reference either (quirk Q3) but part of the state_dict
        self.linear_out_ = _wn(nn.Conv2d(in_channels=self.fc_dim * feat_dim, out_channels=self.dim[2],
                                         kernel_size=(1, 1), groups=self.fc_dim))
        self.linear_out_2 = nn.Linear(self.fc_dim * feat_dim, self.dim[2])
        self.precision = _default_precision()

    def forward(self, roi_feat, adj_matrix, position_embedding, label_biases_att):
        """Stand-alone form of models/graph_att_layer.py:60-178: (output [B,N,D], aff_softmax [B,N,H,K]).  Projections
        go through the GEMM kernels (FCNet / LinearFn), scores + mask + biases + softmax and the aggregation through the
        edge kernels (EdgeAttentionFn); the few reshapes in between are torch ops.  The relation encoders and
        ChangeDetector do not come through here: they run the fused RelationFn."""
        pc = PC(getattr(self, "precision", _default_precision()))
        B, N, D = roi_feat.shape
        H = self.num_heads
        Kn = self.nongt_dim if self.nongt_dim < N else N
        nongt = roi_feat[:, :Kn, :]
        q = self.query(roi_feat)                                        # [B,N,D]
        k = self.key(nongt)                                             # [B,Kn,D]
        # Z_h = v_data W_out2[:, hD:(h+1)D]^T: the out-projection applied before the aggregation (Q3 re-association)
        Wz = self.linear_out_2.weight.view(D, H, D).permute(1, 0, 2).reshape(H * D, D)
        z, _ = LinearFn.apply(pc, nongt.reshape(B * Kn, D), None, Wz, None)
        z = z.view(B, Kn, H * D)
        if Kn < N:
            k = F.pad(k, (0, 0, 0, N - Kn))
            z = F.pad(z, (0, 0, 0, N - Kn))
        qkz = torch.cat((q, k, z), dim=-1).reshape(B * N, (2 + H) * D)
        gbias = None
        if position_embedding is not None and self.pos_emb_dim > 0:
            pe = position_embedding.float().reshape(B, -1, self.pos_emb_dim)
            feat = F.relu(self.pair_pos_fc1(pe)).view(B, -1, Kn, self.fc_dim)          # :113-127
            if feat.shape[1] != N:
                raise ValueError("position_embedding must hold N x K pairs per image, got %s" % (tuple(position_embedding.shape),))
            gbias = torch.log(torch.clamp(feat, min=1e-6))                               # :131-135
        cond = lbias = None
        if adj_matrix is not None:
            cond = adj_matrix.float().reshape(B, N, Kn)
            lbias = label_biases_att.float().reshape(B, N, Kn)
        out, P = EdgeAttentionFn.apply(pc, (B, N, Kn, D, H), qkz, cond, lbias, gbias, self.linear_out_2.bias)
        return out.view(B, N, D), P


class GAttNet(nn.Module):
    """models/graph_att.py:17-106 (parameters) + the fused relation step."""

    def __init__(self, dir_num, label_num, in_feat_dim, out_feat_dim, nongt_dim=20, dropout=0.2, label_bias=True,
                 num_heads=16, pos_emb_dim=-1):
        super().__init__()
        assert dir_num <= 2, "Got more than two directions in a graph."
        self.dir_num = dir_num
        self.label_num = label_num
        self.in_feat_dim = in_feat_dim
        self.out_feat_dim = out_feat_dim
        self.dropout = nn.Dropout(dropout)
        self.self_weights = FCNet([in_feat_dim, out_feat_dim], '', dropout)
        self.bias = FCNet([label_num, 1], '', 0, label_bias)
        self.nongt_dim = nongt_dim
        self.pos_emb_dim = pos_emb_dim
        self.num_heads = num_heads
        self.neighbor_net = nn.ModuleList([
            GraphSelfAttentionLayer(pos_emb_dim=pos_emb_dim, num_heads=num_heads, feat_dim=out_feat_dim,
                                    nongt_dim=nongt_dim) for _ in range(dir_num)])

    def live_layer(self) -> GraphSelfAttentionLayer:
        # quirk Q2: the output of every direction but the last is overwritten
        return self.neighbor_net[self.dir_num - 1]

    def make_drop(self, dev, override=None) -> Drop:
        """Train-mode dropout of this GAT: 0.2 on the inputs of self_weights / query / key / pair_pos_fc1 (fc.py:25-32)
        and 0.2 on the doubled output before the ReLU (graph_att.py:103)."""
        p_fc = self.self_weights.main[0].p if isinstance(self.self_weights.main[0], nn.Dropout) else 0.0
        p_gat = self.dropout.p
        if override is not None:
            p_fc = p_gat = override
        return Drop(dev, self.training, p_fc=p_fc, p_gat=p_gat)

    def effective_weights(self):
        """Weight-normalised matrices of the live branch (independent of the inputs: ChangeDetector computes them while
        the question GRU runs on its own stream)."""
        if self.dir_num != 2:
            raise NotImplementedError("only dir_num == 2 (the reference configuration) is implemented")
        return dict(zip(("sw", "q", "k", "p0"), (wn_weight(lin) for lin in self.wn_linears())))

    def wn_linears(self):
        """The four weight-normalised layers of the live branch, in the order (sw, q, k, p0)."""
        if self.dir_num != 2:
            raise NotImplementedError("only dir_num == 2 (the reference configuration) is implemented")
        layer = self.live_layer()
        p0 = layer.pair_pos_fc1.linear() if self.pos_emb_dim > 0 else self.bias.linear()
        return [self.self_weights.linear(), layer.query.linear(), layer.key.linear(), p0]

    def _step_args(self, N, weights):
        layer = self.live_layer()
        w = weights if weights is not None else self.effective_weights()
        if self.pos_emb_dim > 0:
            kind, p1 = "implicit", layer.pair_pos_fc1.linear().bias
        else:
            kind, p1 = "explicit", None
        return layer, w, kind, p1, layer.num_heads, min(self.nongt_dim, N)

    def prepare_step(self, pc, geo0, geo1, g_split, G, B, N, drop=None, site0=100, weights=None):
        """The activation-independent part of relation_step (functions.relation_prepare); hand the result back through
        relation_step(prep=...)."""
        from .functions import relation_prepare
        layer, w, kind, p1, H, Kn = self._step_args(N, weights)
        dev = w["sw"].device
        if drop is None:
            drop = self.make_drop(dev)
        dims = (G, B, N, Kn, self.out_feat_dim, H)
        return relation_prepare(pc, drop, site0, kind, dims, w["sw"], w["q"], layer.query.linear().bias, w["k"],
                                layer.key.linear().bias, layer.linear_out_2.weight, w["p0"], p1, geo0, geo1, g_split,
                                need_bwd=torch.is_grad_enabled())

    def relation_step(self, pc, X, XT, q, geo0, geo1, g_split, G, B, N, drop=None, site0=100, weights=None, prep=None):
        """X [G*N, D] -> X + relu(2 * attention output).  geo*: adjacency (explicit) or fp64 boxes (implicit)."""
        if drop is None:
            drop = self.make_drop(X.device)
        D = self.out_feat_dim
        layer, w, kind, p1, H, Kn = self._step_args(N, weights)
        sw = self.self_weights.linear()
        ql, kl = layer.query.linear(), layer.key.linear()
        dims = (G, B, N, Kn, D, H)
        return RelationFn.apply(pc, drop, site0, kind, dims, X, XT, q, w["sw"], sw.bias, w["q"], ql.bias, w["k"], kl.bias,
                                layer.linear_out_2.weight, layer.linear_out_2.bias, w["p0"], p1, geo0, geo1, g_split,
                                prep)

    def forward(self, v_feat, adj_matrix, pos_emb=None):
        if self.pos_emb_dim > 0 and pos_emb is None:
            raise ValueError(f"position embedding is set to None with pos_emb_dim {self.pos_emb_dim}")
        elif self.pos_emb_dim < 0 and pos_emb is not None:
            raise ValueError("position embedding is NOT None with pos_emb_dim < 0")
        # Stand-alone form of models/graph_att.py:71-106 on a materialised [v | q] tensor (the relation encoders and
        # ChangeDetector never build it: they run the fused RelationFn).  Returns (output, [aff_d0, aff_d1]).
        B, N, _ = v_feat.shape
        nongt = self.nongt_dim
        adj = adj_matrix.float()
        adj_list = [adj, adj.transpose(1, 2)]
        self_feat = self.self_weights(v_feat)                           # [B,N,out]
        output = self_feat
        aff = []
        for d in range(self.dir_num):
            a = adj_list[d][:, :, :nongt, :]
            cond = a.sum(-1)                                            # :88
            lbias = self.bias(a).squeeze(-1)                            # :92
            layer = self.neighbor_net[d]
            if d < self.dir_num - 1:
                # quirk Q2: this direction's output is overwritten below; only its attention map is returned
                with torch.no_grad():
                    _, p = layer(self_feat, cond, pos_emb, lbias)
                aff.append(p)
                continue
            out, p = layer(self_feat, cond, pos_emb, lbias)
            aff.append(p)
            output = out + out                                          # :99-101 (neighbor_emb[d] is `output` itself)
        output = F.relu(self.dropout(output))
        return output, aff


def _maybe_inplace(v: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """Reference encoders mutate and return their first argument (quirk Q1).  Outside autograd we do the same."""
    if not torch.is_grad_enabled() or (not v.requires_grad and not out.requires_grad):
        v.copy_(out.view_as(v))
        return v
    return out.view_as(v)


class ImplicitRelationEncoder(nn.Module):
    """models/relation_encoder.py:33-84.  `position_embedding` may be the reference's [B,N,K,64] tensor -- not
    supported on the CUDA path -- or, preferred, the raw boxes [B,N,4] (fp64): the geometry bias is computed in
    the edge kernel straight from the boxes."""

    def __init__(self, v_dim, q_dim, out_dim, dir_num, pos_emb_dim, nongt_dim, num_heads=16, num_steps=1,
                 residual_connection=True, label_bias=True):
        super().__init__()
        self.v_dim, self.q_dim, self.out_dim = v_dim, q_dim, out_dim
        self.residual_connection = residual_connection
        self.num_steps = num_steps
        print("In ImplicitRelationEncoder, num of graph propogate steps:",
              "%d, residual_connection: %s" % (self.num_steps, self.residual_connection))
        self.v_transform = FCNet([v_dim, out_dim]) if self.v_dim != self.out_dim else None
        self.implicit_relation = GAttNet(dir_num, 1, out_dim + q_dim, out_dim, nongt_dim=nongt_dim,
                                         label_bias=label_bias, num_heads=num_heads, pos_emb_dim=pos_emb_dim)
        self.precision = _default_precision()

    def forward(self, v, position_embedding, q):
        if self.v_transform is not None or not self.residual_connection or self.num_steps != 1:
            raise NotImplementedError("only v_dim == out_dim, residual_connection=True, num_steps=1 (reference config)")
        if position_embedding.dim() == 4:
            # the reference's materialised [B,N,K,pos_emb_dim] embedding: the stand-alone GAttNet path
            # (relation_encoder.py:68-84); the fused path below takes the boxes and never builds it
            B, N, _ = v.shape
            ones = torch.ones(B, N, N, 1, device=v.device)
            rel, affs = self.implicit_relation(q_expand_v_cat(q, v, mask=True), ones, position_embedding)
            return _maybe_inplace(v, v + rel), affs
        if position_embedding.dim() != 3 or position_embedding.shape[-1] != 4:
            raise ValueError("the implicit encoder takes the boxes [B, N, 4] (geometry computed inside the kernel) or "
                             "the reference's [B, N, K, %d] position embedding (got shape %s)"
                             % (self.implicit_relation.pos_emb_dim, tuple(position_embedding.shape)))
        pc = PC(self.precision)
        B, N, D = v.shape
        if self.training:
            rng_advance(v.device)        # fresh dropout masks per forward, like nn.Dropout
        out, _, P = self.implicit_relation.relation_step(pc, v.reshape(B * N, D), None, q, position_embedding, None,
                                                         B, B, B, N)
        return _maybe_inplace(v, out), [None, P]


class ExplicitRelationEncoder(nn.Module):
    """models/relation_encoder.py:88-132."""

    def __init__(self, v_dim, q_dim, out_dim, dir_num, label_num, nongt_dim=20, num_heads=16, num_steps=1,
                 residual_connection=True, label_bias=True):
        super().__init__()
        self.v_dim, self.q_dim, self.out_dim = v_dim, q_dim, out_dim
        self.num_steps = num_steps
        self.residual_connection = residual_connection
        print("In ExplicitRelationEncoder, num of graph propogation steps:",
              "%d, residual_connection: %s" % (self.num_steps, self.residual_connection))
        self.v_transform = FCNet([v_dim, out_dim]) if self.v_dim != self.out_dim else None
        self.explicit_relation = GAttNet(dir_num, label_num, out_dim + q_dim, out_dim, nongt_dim=nongt_dim,
                                         num_heads=num_heads, label_bias=label_bias, pos_emb_dim=-1)
        self.precision = _default_precision()

    def forward(self, v, exp_adj_matrix, q):
        if self.v_transform is not None or not self.residual_connection or self.num_steps != 1:
            raise NotImplementedError("only v_dim == out_dim, residual_connection=True, num_steps=1 (reference config)")
        pc = PC(self.precision)
        B, N, D = v.shape
        if self.training:
            rng_advance(v.device)        # fresh dropout masks per forward, like nn.Dropout
        out, _, P = self.explicit_relation.relation_step(pc, v.reshape(B * N, D), None, q, exp_adj_matrix, None,
                                                         B, B, B, N)
        return _maybe_inplace(v, out), [None, P]


class WordEmbedding(nn.Module):
    """models/language_model.py:17-53 (parameters; gather runs in QuestionFn)."""

    def __init__(self, ntoken, emb_dim, dropout, op=''):
        super().__init__()
        self.op = op
        self.emb = nn.Embedding(ntoken + 1, emb_dim, padding_idx=ntoken)
        if 'c' in op:
            self.emb_ = nn.Embedding(ntoken + 1, emb_dim, padding_idx=ntoken)
            self.emb_.weight.requires_grad = False
        self.dropout = nn.Dropout(dropout)
        self.ntoken = ntoken
        self.emb_dim = emb_dim

    def forward(self, x):
        # tiny gather; kept as a torch op for standalone use (the fused path gathers inside QuestionFn)
        emb = self.emb(x)
        if 'c' in self.op:
            emb = torch.cat((emb, self.emb_(x)), 2)
        return self.dropout(emb)


class QuestionEmbedding(nn.Module):
    """models/language_model.py:56-115 (parameters; recurrence runs in QuestionFn)."""

    def __init__(self, in_dim, num_hid, nlayers, bidirect, dropout, rnn_type='GRU'):
        super().__init__()
        assert rnn_type == 'LSTM' or rnn_type == 'GRU'
        rnn_cls = nn.LSTM if rnn_type == 'LSTM' else nn.GRU
        self.rnn = rnn_cls(in_dim, num_hid, nlayers, bidirectional=bidirect, dropout=dropout, batch_first=True)
        self.in_dim = in_dim
        self.num_hid = num_hid
        self.nlayers = nlayers
        self.rnn_type = rnn_type
        self.ndirections = 1 + int(bidirect)
        self.precision = _default_precision()

    def forward_all(self, x):
        """All GRU outputs [B,L,H] for x [B,L,in] (language_model.py:106-115), stand-alone; inside ChangeDetector the
        recurrence is part of QuestionFn."""
        if self.rnn_type != 'GRU' or self.nlayers != 1 or self.ndirections != 1:
            raise NotImplementedError("only the reference configuration (1-layer unidirectional GRU) is implemented")
        rnn = self.rnn
        return GRUFn.apply(PC(self.precision), x, rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0)

    def forward(self, x):
        """Last GRU output [B,H] (language_model.py:87-98)."""
        return self.forward_all(x)[:, -1]


class QuestionSelfAttention(nn.Module):
    """models/language_model.py:118-156 (parameters; pooling incl. quirk Q4 runs in QuestionFn)."""

    def __init__(self, num_hid, dropout):
        super().__init__()
        self.num_hid = num_hid
        self.drop = nn.Dropout(dropout)
        self.W1_self_att_q = FCNet(dims=[num_hid, num_hid], dropout=dropout, act=None)
        self.W2_self_att_q = FCNet(dims=[num_hid, 1], act=None)

    def forward(self, ques_feat):
        """[B,L,H] -> [B,H] (language_model.py:127-156), stand-alone: the two projections run through the GEMM kernels,
        the batch-axis softmax and its reinterpretation (quirk Q4) are written with the same tensor ops as the
        reference so that the quirk is reproduced exactly."""
        B, L = ques_feat.shape[0], ques_feat.shape[1]
        flat = ques_feat.contiguous().view(-1, self.num_hid)
        atten = self.W2_self_att_q(torch.tanh(self.W1_self_att_q(flat))).view(B, L)
        weight = F.softmax(atten.t(), dim=1).view(-1, 1, L)             # Q4: softmax over the batch axis, then reinterpreted
        out = torch.bmm(weight, ques_feat).view(-1, self.num_hid)
        return self.drop(out)


class SelfAttention(nn.Module):
    """models/modules.py:17-77.  Dead in setting='mode2' (SSRE is constructed but never called); kept so that
    reference checkpoints load (quirk Q11)."""

    def __init__(self, cfg):
        super().__init__()
        cd = cfg.model.change_detector
        if cd.att_dim % cd.att_head != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention "
                             "heads (%d)" % (cd.att_dim, cd.att_head))
        self.num_attention_heads = cd.att_head
        self.attention_head_size = int(cd.att_dim / cd.att_head)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(cd.att_dim * 2, self.all_head_size)
        self.key = nn.Linear(cd.att_dim * 2, self.all_head_size)
        self.value = nn.Linear(cd.att_dim * 2, self.all_head_size)
        self.dropout = nn.Dropout(0.1)
        self.layer_norm = nn.LayerNorm(cd.att_dim, eps=1e-6)


class ChangeDetector(nn.Module):
    """Drop-in for models/modules.py:81-313 (setting='mode2')."""

    def __init__(self, cfg, word_to_idx):
        super().__init__()
        cd = cfg.model.change_detector
        self.input_dim = cd.input_dim
        self.dim = cd.dim
        self.feat_dim = cd.feat_dim - 2
        self.att_head = cd.att_head
        self.att_dim = cd.att_dim
        self.nongt_dim = cd.nongt_dim
        self.pos_emb_dim = cd.pos_emb_dim
        self.img = nn.Linear(self.feat_dim, self.att_dim)
        self.SSRE = SelfAttention(cfg)
        self.context1 = nn.Linear(self.att_dim, self.att_dim, bias=False)
        self.context2 = nn.Linear(self.att_dim, self.att_dim)
        self.gate1 = nn.Linear(self.att_dim, self.att_dim, bias=False)
        self.gate2 = nn.Linear(self.att_dim, self.att_dim)
        self.dropout = nn.Dropout(0.5)
        self.embed = nn.Sequential(nn.Linear(self.att_dim * 3, self.dim), nn.Dropout(0.5), nn.ReLU())
        self.att = nn.Linear(self.dim, 1)
        self.fc1 = nn.Linear(self.att_dim, 6)
        self.coef_sem = cd.coef_sem
        self.coef_spa = cd.coef_spa
        assert self.coef_sem + self.coef_spa <= 1
        q_dim = cfg.model.speaker.embed_dim
        if cfg.train.setting == 'mode2':
            g = cfg.train.graph
            if g == 'all' or g == 'semantic':
                self.semantic_relation = ExplicitRelationEncoder(
                    cd.att_dim, q_dim, cd.att_dim, cd.dir_num, cd.sem_label_num, num_heads=cd.att_head, num_steps=1,
                    nongt_dim=cd.nongt_dim, residual_connection=True, label_bias=False)
            if g == 'all' or g == 'spatial' or g == 'i+s':
                self.spatial_relation = ExplicitRelationEncoder(
                    cd.att_dim, q_dim, cd.att_dim, cd.dir_num, cd.spa_label_num, num_heads=cd.att_head, num_steps=1,
                    nongt_dim=cd.nongt_dim, residual_connection=True, label_bias=False)
            if g == 'all' or g == 'implicit' or g == 'i+s':
                self.imp_relation = ImplicitRelationEncoder(
                    cd.att_dim, q_dim, cd.att_dim, cd.dir_num, 64, cd.nongt_dim, num_heads=cd.att_head, num_steps=1,
                    residual_connection=True, label_bias=False)
        self.w_emb = WordEmbedding(len(word_to_idx), 300, .0, 'c')
        self.q_emb = QuestionEmbedding(600, q_dim, 1, False, .0)
        self.q_att = QuestionSelfAttention(q_dim, .2)
        self.cfg = cfg
        if cfg.data.feature_mode == 'mode0':
            raise NotImplementedError("feature_mode 'mode0' (ResNet-101 on raw images) is outside the hot path")
        self.precision = _default_precision()
        # tests only: force every dropout probability (0.0 runs the train-mode code path without masks)
        self.dropout_override = None
        # tests only: keep the dropout seed where it is (finite-difference checks need the SAME masks in every forward)
        self.freeze_dropout_seed = False
        self._side = None
        self._qv_keepalive = None

    def live_parameters(self):
        """Parameters that can receive a gradient in setting='mode2'.  The rest exist only so reference checkpoints
        load (quirk Q11): SSRE.*, direction-0 attention layers (Q2), the never-called grouped conv linear_out_ (Q3)
        and the frozen embedding table.  torch.optim.Adam skips them too (their .grad stays None)."""
        return [p for _, p in self.live_named_parameters()]

    def live_named_parameters(self):
        """(name, parameter) pairs of live_parameters()."""
        out = []
        for name, p in self.named_parameters():
            if not p.requires_grad or name.startswith("SSRE."):
                continue
            if ".linear_out_." in name:
                continue
            parts = name.split(".")
            if "neighbor_net" in parts:
                owner = self.get_submodule(".".join(parts[:parts.index("neighbor_net")]))
                if int(parts[parts.index("neighbor_net") + 1]) != owner.dir_num - 1:
                    continue
            out.append((name, p))
        return out

    def set_precision(self, precision: str) -> "ChangeDetector":
        PC(precision)
        for m in self.modules():
            if hasattr(m, "precision"):
                m.precision = precision
        return self

    # -- question path --------------------------------------------------------------------------
    def question_vector(self, pc: PC, question: torch.Tensor) -> torch.Tensor:
        rnn = self.q_emb.rnn
        w1 = self.q_att.W1_self_att_q.linear()
        w2 = self.q_att.W2_self_att_q.linear()
        ov = self.dropout_override
        drop = Drop(question.device, self.training,
                    p_fc=self.q_att.W1_self_att_q.main[0].p if ov is None else ov,
                    p_qv=self.q_att.drop.p if ov is None else ov)
        return QuestionFn.apply(pc, drop, question, self.w_emb.emb.weight, self.w_emb.emb_.weight, rnn.weight_ih_l0,
                                rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0, wn_weight(w1), w1.bias,
                                wn_weight(w2), w2.bias,
                                self.w_emb.emb.padding_idx if self.w_emb.emb.padding_idx is not None else -1)

    def position_emb(self, bb):
        """The reference materialises a [B,N,K,64] fp64 embedding here (modules.py:162-166); the CUDA path
        consumes the boxes directly, so this returns them unchanged."""
        return bb

    def forward(self, input_1, input_2, d_adj_matrix, q_adj_matrix, d_sem_adj_matrix, q_sem_adj_matrix, d_bb, q_bb,
                question, setting='mode2', graph='all'):
        if self.cfg.data.train.empty_image == True:  # noqa: E712  (modules.py:170-178)
            input_1, input_2 = torch.ones_like(input_1), torch.ones_like(input_2)
            if d_adj_matrix.dim() == 3:        # label matrices: the reference sees all-ones ONE-HOT tensors here
                n_ = input_1.shape[1]
                ones = lambda L: torch.ones(input_1.shape[0], n_, n_, L, device=input_1.device)  # noqa: E731
                d_adj_matrix, q_adj_matrix = ones(self.cfg.model.change_detector.spa_label_num), ones(self.cfg.model.change_detector.spa_label_num)
                d_sem_adj_matrix, q_sem_adj_matrix = ones(self.cfg.model.change_detector.sem_label_num), ones(self.cfg.model.change_detector.sem_label_num)
            d_adj_matrix, q_adj_matrix = torch.ones_like(d_adj_matrix), torch.ones_like(q_adj_matrix)
            d_sem_adj_matrix, q_sem_adj_matrix = torch.ones_like(d_sem_adj_matrix), torch.ones_like(q_sem_adj_matrix)
            d_bb, q_bb = torch.ones_like(d_bb), torch.ones_like(q_bb)
        if setting != 'mode2':
            raise NotImplementedError("only setting='mode2' is live in the reference (modes 1/3/4 use "
                                      "self.graph_relation, which does not exist; mode0 is the SSRE ablation)")
        if graph not in ('all', 'semantic', 'spatial', 'implicit', 'i+s'):
            raise ValueError("unknown graph mode %r" % (graph,))
        pc = PC(self.precision)
        ov = self.dropout_override
        B, N, C = input_1.size()
        D = self.att_dim
        G = 2 * B
        _lib.require_device()            # no CPU / PyTorch fallback: fail loudly without an sm_100 device
        dev = input_1.device
        if self.training and not self.freeze_dropout_seed:
            # every train-mode forward draws fresh dropout masks (the reference's nn.Dropout modules do): the device-side
            # seed is advanced by a kernel, so this also holds for replays of a captured CUDA graph
            rng_advance(dev)
        # The question path is a chain of 20 small, latency-bound GRU steps: it runs on its own stream while this
        # stream does the work that does not depend on it (ROI projection, weight normalisation).  Autograd replays
        # the same split in backward (BPTT next to the weight-norm / img gradients); a captured CUDA graph keeps it.
        cur = torch.cuda.current_stream(dev)
        if self._side is None:
            self._side = torch.cuda.Stream(dev, priority=-1)    # its small kernels go first when SM slots free up
        side = self._side
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            qv = self.question_vector(pc, question)
        self._qv_keepalive = qv          # allocated on the side stream, consumed on this one
        X, XT = LinearFn.apply(pc, input_1, input_2, self.img.weight, self.img.bias)      # [2BN, D]
        gats = {}
        if graph in ('semantic', 'all'):
            gats['sem'] = self.semantic_relation.explicit_relation
        if graph in ('spatial', 'all', 'i+s'):
            gats['spa'] = self.spatial_relation.explicit_relation
        if graph in ('implicit', 'all', 'i+s'):
            gats['imp'] = self.imp_relation.implicit_relation
        # the four weight-normalised matrices of each encoder in two launches (and two more in backward).  One call per
        # encoder, not one for all: its backward then runs as soon as THAT encoder's weight gradients exist, so the
        # gradients of the encoders that finish first can be exchanged between ranks long before backward ends
        wn_args = {k: [t for lin in g.wn_linears() for t in (lin.weight_v, lin.weight_g)] for k, g in gats.items()}
        with torch.no_grad():
            wn_pre = {k: wn_many_compute(wn_args[k]) for k in gats}
        pre_w = {k: dict(zip(("sw", "q", "k", "p0"), wn_pre[k][2])) for k in gats}
        # ... and everything else that needs neither the activations nor the question vector: operand-type weight
        # copies, adjacency condition / label bias, geometry bias.  All of it overlaps the question path.
        geos = {'sem': (d_sem_adj_matrix, q_sem_adj_matrix, 100), 'spa': (d_adj_matrix, q_adj_matrix, 200),
                'imp': (d_bb, q_bb, 300)}
        drops = {k: g.make_drop(dev, ov) for k, g in gats.items()}
        # one stream per encoder: the preparation of the second / third encoder (its weight casts, adjacency bias or
        # geometry bias) must not delay the first encoder's GEMMs -- each is joined right before the step that needs it
        fkp = Fork(dev, len(gats), pool="prep")
        preps = {}
        for i, (k, g) in enumerate(gats.items()):
            with fkp.branch(i):
                preps[k] = g.prepare_step(pc, geos[k][0], geos[k][1], B, G, B, N, drop=drops[k], site0=geos[k][2],
                                          weights=pre_w[k])
        cur.wait_stream(side)
        for i, k in enumerate(gats):
            if fkp.on:
                cur.wait_stream(fkp.streams[i])
            # the autograd node of this encoder's weight normalisation is created HERE (see WNormManyFn)
            eff = dict(zip(("sw", "q", "k", "p0"), WNormManyFn.apply(wn_pre[k], *wn_args[k])))
            X, XT, _ = gats[k].relation_step(pc, X, XT, qv, geos[k][0], geos[k][1], B, G, B, N, drop=drops[k],
                                             site0=geos[k][2], weights=eff, prep=preps[k])
        mode = 1 if graph == 'all' else (2 if graph == 'i+s' else 0)
        coefs = (float(self.coef_sem), float(self.coef_spa), float(1 - self.coef_sem - self.coef_spa))
        fdrop = Drop(dev, self.training, p_fuse=self.dropout.p if ov is None else ov,
                     p_embed=self.embed[1].p if ov is None else ov)
        att_weight_before, att_weight_after, attended_1, attended_2, input_attended = FusionFn.apply(
            pc, fdrop, (B, N, D, self.dim), mode, coefs, X, self.context1.weight, self.context2.weight, self.context2.bias,
            self.gate1.weight, self.gate2.weight, self.gate2.bias, self.embed[0].weight, self.embed[0].bias,
            self.att.weight, self.att.bias)
        pred = SmallLinearFn.apply(input_attended, self.fc1.weight, self.fc1.bias)      # [B,6], unused by the loss (Q11)
        return pred, att_weight_before, att_weight_after, attended_1, attended_2, input_attended
