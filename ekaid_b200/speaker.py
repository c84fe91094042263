"""Answer decoder: drop-in `DynamicSpeaker` / `DynamicCore` (reference models/dynamic_speaker_change_pos.py) and the
masked language-model criterion (utils/utils.py:204-216) on the library's kernels.

Same constructor, attribute tree and state_dict keys as the reference, same entry points (`_forward`, `_sample`,
`get_logprobs_state`, `init_hidden`, `get_module_weights`).  What differs is how a sequence is executed
(`SpeakerSeqFn`):

* everything that does not depend on the recurrent state leaves the token loop: `core.embed`'s Linear + ReLU
  (:98-99, step-invariant), its product with the module LSTM's input weights (eval mode), and the word-embedding part
  of the language LSTM's input product for ALL steps at once (teacher forcing: the tokens are known up front; sampling:
  a [vocab, 4R] table indexed by the sampled token);
* the four products whose input is the previous language-LSTM state (module LSTM input half, pos1, gate1x's prev_h
  columns, the language LSTM's recurrent half) are ONE GEMM per step (N = 6656, K = 512);
* per step that leaves 5 GEMMs (tcgen05, weights L2-resident) + 4 small kernels; the logit layer, log-softmax and the
  masked NLL run once over all steps; in backward all weight gradients are batched over the steps as well
  (K = T * B), only the data-gradient chain is walked step by step;
* greedy sampling keeps `unfinished` / the stop condition on the device (`dec_token`): no host synchronisation per
  step (:354); the loop is cut short by a flag read every `check_every` steps.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.nn as nn

from . import lib
from .functions import (ACT_RELU, PC, Drop, _dst, _f32c, cast_many, colsum, gemm, ptr, rng_advance)
from .lib import call

# dropout sites of the decoder (functions.Drop site numbering: relation encoders 100-399, question path 10-11, fusion 20-21)
SITE_WORD, SITE_EMBED, SITE_POS1, SITE_DPOS, SITE_GATE1, SITE_OUT, SITE_SAMPLE = 400, 401, 402, 403, 404, 405, 406


def _default_precision() -> str:
    from .modules import _default_precision as d
    return d()


class DynamicCore(nn.Module):
    """dynamic_speaker_change_pos.py:43-92 (parameters; the step itself runs in SpeakerSeqFn / DynamicSpeaker)."""

    def __init__(self, cfg):
        super().__init__()
        sp = cfg.model.speaker
        self.rnn_num_layers = 2
        self.drop_prob_lm = sp.drop_prob_lm
        self.input_dim = sp.input_dim
        self.embed_input_dim = sp.embed_input_dim
        self.embed_dim = sp.embed_dim
        self.embed = nn.Sequential(nn.Linear(self.embed_input_dim, self.embed_dim), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.module_att_lstm = nn.LSTMCell(self.embed_dim + sp.rnn_size, sp.rnn_size)
        self.weight_fc = nn.Sequential(nn.Linear(sp.rnn_size, 3), nn.Softmax(dim=1))
        self.pos1 = nn.Sequential(nn.Linear(512, 512), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.softmax = nn.Softmax(dim=1)
        self.dropout = nn.Dropout(0.5)
        self.weight_pos = nn.Linear(sp.rnn_size, 16)
        self.pos2 = nn.Linear(16, sp.rnn_size)
        g = sp.rnn_size * 2 + sp.input_dim
        self.gate1x = nn.Sequential(nn.Linear(g, g), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.gate2x = nn.Linear(g, sp.input_dim)
        self.lang_lstm = nn.LSTMCell(sp.word_embed_size + sp.input_dim, sp.rnn_size)
        self.module_weights = None

    def get_module_weights(self):
        return self.module_weights

    def forward(self, xt, loc_feat_bef, loc_feat_aft, feat_diff, state):
        raise NotImplementedError("DynamicCore runs inside DynamicSpeaker (get_logprobs_state / _forward / _sample): its "
                                  "step is fused with the word embedding and the logit layer")


def _op(pc: PC, t: torch.Tensor) -> torch.Tensor:
    """fp32 2-D view -> GEMM operand (bf16 copy on the tensor-core path, the fp32 view itself on the parity path)."""
    t = t.detach()
    if not pc.bf16:
        return t if t.stride(-1) == 1 else t.contiguous()
    out = torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
    cast_many(pc, [(t, out)])
    return out


class _W:
    """Operand-type views / copies of the decoder's weights for one call (weights change every optimizer step)."""

    def __init__(self, pc: PC, sp: "DynamicSpeaker"):
        c = sp.core
        R, E, D, We = sp.rnn_size, c.embed_dim, c.input_dim, sp.word_embed_size
        dev = sp.logit.weight.device
        self.R, self.E, self.D, self.We, self.G, self.V = R, E, D, We, 2 * R + D, sp.vocab_size
        self.Wex = (We + 7) // 8 * 8                         # word-embedding width padded to a 16-byte operand pitch
        T = pc.T
        Wm, Wl, Wg1 = c.module_att_lstm.weight_ih.detach(), c.lang_lstm.weight_ih.detach(), c.gate1x[0].weight.detach()
        NH = 4 * R + 512 + self.G + 4 * R
        self.NH = NH
        self.o_p1, self.o_g1, self.o_lh = 4 * R, 4 * R + 512, 4 * R + 512 + self.G
        # rows: [module LSTM input weights, prev_h columns | pos1 | gate1x, prev_h columns | language LSTM recurrent]
        self.Wcat_h = torch.empty(NH, R, dtype=T, device=dev)
        self.W_ih_x = torch.zeros(4 * R, self.Wex, dtype=T, device=dev)
        jobs = [(Wm[:, E:], self.Wcat_h[:4 * R]), (c.pos1[0].weight.detach(), self.Wcat_h[self.o_p1:self.o_g1]),
                (Wg1[:, :R], self.Wcat_h[self.o_g1:self.o_lh]), (c.lang_lstm.weight_hh.detach(), self.Wcat_h[self.o_lh:]),
                (Wl[:, :We], self.W_ih_x[:, :We])]
        if pc.bf16:
            mk = lambda src: torch.empty(src.shape, dtype=T, device=dev)       # noqa: E731
            self.W_hh_m, self.W_ih_me, self.W_g1r = mk(c.module_att_lstm.weight_hh), mk(Wm[:, :E]), mk(Wg1[:, R:])
            self.W_g2, self.W_ih_lg, self.W_lo = mk(c.gate2x.weight), mk(Wl[:, We:]), mk(sp.logit.weight)
            self.W_e = mk(c.embed[0].weight)
            jobs += [(c.module_att_lstm.weight_hh.detach(), self.W_hh_m), (Wm[:, :E], self.W_ih_me), (Wg1[:, R:], self.W_g1r),
                     (c.gate2x.weight.detach(), self.W_g2), (Wl[:, We:], self.W_ih_lg), (sp.logit.weight.detach(), self.W_lo),
                     (c.embed[0].weight.detach(), self.W_e)]
        else:
            self.W_hh_m, self.W_ih_me, self.W_g1r = c.module_att_lstm.weight_hh.detach(), Wm[:, :E], Wg1[:, R:]
            self.W_g2, self.W_ih_lg, self.W_lo = c.gate2x.weight.detach(), Wl[:, We:], sp.logit.weight.detach()
            self.W_e = c.embed[0].weight.detach()
        cast_many(pc, jobs)
        f = lambda t: _f32c(t)                                                   # noqa: E731
        self.b_e, self.b_g1, self.b_g2, self.b_lo = f(c.embed[0].bias), f(c.gate1x[0].bias), f(c.gate2x.bias), f(sp.logit.bias)
        self.b_mi, self.b_mh = f(c.module_att_lstm.bias_ih), f(c.module_att_lstm.bias_hh)
        self.b_li, self.b_lh = f(c.lang_lstm.bias_ih), f(c.lang_lstm.bias_hh)
        self.emb = f(sp.embed[0].weight)
        self.att_w = [f(c.weight_fc[0].weight), f(c.weight_fc[0].bias), f(c.pos1[0].bias), f(c.weight_pos.weight),
                      f(c.weight_pos.bias), f(c.pos2.weight), f(c.pos2.bias)]
        self.att_ptrs = (ctypes.c_void_p * 7)(*[t.data_ptr() for t in self.att_w])


def _opf(pc: PC) -> int:
    return 1 if pc.bf16 else 0


def _gemm_op(pc, A, Wt, M, N, K, out_op, **kw):
    """GEMM whose result is a GEMM operand again (bf16 output on the tensor-core path, fp32 on the parity path)."""
    if pc.bf16:
        gemm(A, Wt, M, N, K, Cb=out_op, **kw)
    else:
        gemm(A, Wt, M, N, K, C=out_op, **kw)


class SpeakerSeqFn(torch.autograd.Function):
    """The teacher-forced decoder over T steps (DynamicSpeaker._forward, :182-222, with get_logprobs_state :225-240 and
    DynamicCore.forward :94-131 inside) -> (log-probs [B, seq_length, V], output_pos [B, seq_length, 16], module weights
    [B, T, 3]) or, with `fused_nll`, the masked NLL of utils/utils.py:204-216 directly (train_mimic.py:242)."""

    @staticmethod
    def forward(ctx, pc: PC, drop, sp: "DynamicSpeaker", T: int, fused_nll: bool, seq, masks, bef, aft, diff, *params):
        lib.require_device()
        dev = bef.device
        w = _W(pc, sp)
        R, E, D, G, V, NH, Wex = w.R, w.E, w.D, w.G, w.V, w.NH, w.Wex
        B = bef.shape[0]
        TB = T * B
        opf, OT = _opf(pc), pc.T
        don = drop is not None and drop.on
        p_lm = float(sp.drop_prob_lm)
        d = (lambda site, p: drop.a(site, p)) if don else (lambda site, p: (None, 0, 0.0))
        bef, aft, diff = _f32c(bef), _f32c(aft), _f32c(diff)
        seq = seq.detach().contiguous()
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)      # noqa: E731
        opt = lambda *s: torch.empty(*s, dtype=OT, device=dev)                 # noqa: E731
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        # ---- step-invariant part -------------------------------------------------------------------------------
        EI = opt(B, 3 * D)
        if pc.bf16:
            cast_many(pc, [(bef, EI[:, :D]), (diff, EI[:, D:2 * D]), (aft, EI[:, 2 * D:])])
        else:
            cast_many(pc, [(bef, EI[:, :D]), (diff, EI[:, D:2 * D]), (aft, EI[:, 2 * D:])])
        emb0 = f32(B, E)
        gemm(EI, w.W_e, B, E, 3 * D, bias=w.b_e, act=ACT_RELU, C=emb0)
        # module-LSTM input product of the embedded features: one GEMM over all steps (each step has its own Dropout
        # mask in train mode; in eval mode every step sees the same rows)
        if don:
            EMBD = opt(TB, E)
            call("dec_drop_op", emb0.data_ptr(), E, TB, E, B, *d(SITE_EMBED, p_lm), 0, EMBD.data_ptr(), E, opf)
            S3 = f32(TB, 4 * R)
            gemm(EMBD, w.W_ih_me, TB, 4 * R, E, C=S3)
        else:
            EMBD = _op(pc, emb0)
            S3 = f32(B, 4 * R)
            gemm(EMBD, w.W_ih_me, B, 4 * R, E, C=S3)
        # word-embedding part of the language LSTM's input product, all steps at once
        XT = opt(TB, Wex)
        call("dec_embed", seq.data_ptr(), seq.stride(0), seq.stride(1), 0, B, TB, w.emb.data_ptr(), V, w.We, XT.data_ptr(),
             Wex, opf, *d(SITE_WORD, p_lm), err.data_ptr())
        XP = f32(TB, 4 * R)
        gemm(XT, w.W_ih_x, TB, 4 * R, Wex, C=XP)
        # ---- slabs ----------------------------------------------------------------------------------------------
        HL = torch.zeros((T + 1) * B, R, dtype=OT, device=dev)        # h_lang as operand, block t = state before step t
        HM = torch.zeros((T + 1) * B, R, dtype=OT, device=dev)
        HMf = torch.zeros((T + 1) * B, R, dtype=torch.float32, device=dev)
        HLf = torch.zeros((T + 1) * B, R, dtype=torch.float32, device=dev)
        Cm = torch.zeros((T + 1) * B, R, dtype=torch.float32, device=dev)
        Cl = torch.zeros((T + 1) * B, R, dtype=torch.float32, device=dev)
        Gm, Gl = f32(TB, 4 * R), f32(TB, 4 * R)
        OUT = opt(TB, R) if don else None
        MW, PW, DPOS, VPOS, ATT = f32(TB, 4), f32(TB, 16), f32(TB, 16), f32(TB, 512), f32(TB, D)
        GI2, G1, GATE, GATED = opt(TB, R + D), opt(TB, G), f32(TB, D), opt(TB, D)
        S1, S2, S6, g2pre = f32(B, NH), f32(B, 4 * R), f32(B, 4 * R), f32(B, D)
        g1tmp = f32(B, G) if don else None
        for t in range(T):
            r0, r1, r2 = t * B, (t + 1) * B, (t + 2) * B
            gemm(HL[r0:r1], w.Wcat_h, B, NH, R, C=S1)
            gemm(HM[r0:r1], w.W_hh_m, B, 4 * R, R, C=S2)
            s3 = S3[r0:r1] if don else S3
            call("dec_lstm_fwd", S1.data_ptr(), NH, S2.data_ptr(), 4 * R, s3.data_ptr(), 4 * R, None, None,
                 w.b_mi.data_ptr(), w.b_mh.data_ptr(), Cm[r0:r1].data_ptr(), B, R, Gm[r0:r1].data_ptr(),
                 Cm[r1:r2].data_ptr(), HMf[r1:r2].data_ptr(), HM[r1:r2].data_ptr(), R, None, 0, opf, None, 0, 0.0, 0)
            call("dec_att_fwd", HMf[r1:r2].data_ptr(), S1[:, w.o_p1:].data_ptr(), NH, ctypes.addressof(w.att_ptrs),
                 bef.data_ptr(), diff.data_ptr(), aft.data_ptr(), B, R, 512, D,
                 drop.seed if don else None, SITE_POS1, p_lm if don else 0.0, SITE_DPOS, 0.5 if don else 0.0, r0,
                 MW[r0:r1].data_ptr(), PW[r0:r1].data_ptr(), DPOS[r0:r1].data_ptr(), VPOS[r0:r1].data_ptr(),
                 ATT[r0:r1].data_ptr(), GI2[r0:r1].data_ptr(), R + D, opf)
            kw = dict(bias=w.b_g1, addend=S1[:, w.o_g1:w.o_lh], act=ACT_RELU)
            if don:
                gemm(GI2[r0:r1], w.W_g1r, B, G, R + D, C=g1tmp, **kw)
                call("dec_drop_op", g1tmp.data_ptr(), G, B, G, 0, *d(SITE_GATE1, p_lm), r0 * G, G1[r0:r1].data_ptr(), G, opf)
            else:
                _gemm_op(pc, GI2[r0:r1], w.W_g1r, B, G, R + D, G1[r0:r1], **kw)
            gemm(G1[r0:r1], w.W_g2, B, D, G, bias=w.b_g2, C=g2pre)
            call("dec_gate_fwd", g2pre.data_ptr(), ATT[r0:r1].data_ptr(), B * D, GATE[r0:r1].data_ptr(),
                 GATED[r0:r1].data_ptr(), opf)
            gemm(GATED[r0:r1], w.W_ih_lg, B, 4 * R, D, C=S6)
            call("dec_lstm_fwd", S6.data_ptr(), 4 * R, S1[:, w.o_lh:].data_ptr(), NH, XP[r0:r1].data_ptr(), 4 * R, None, None,
                 w.b_li.data_ptr(), w.b_lh.data_ptr(), Cl[r0:r1].data_ptr(), B, R, Gl[r0:r1].data_ptr(),
                 Cl[r1:r2].data_ptr(), HLf[r1:r2].data_ptr(), HL[r1:r2].data_ptr(), R,
                 OUT[r0:r1].data_ptr() if don else None, R, opf, *d(SITE_OUT, p_lm), r0 * R)
        # ---- logit layer + log-softmax (+ masked NLL) over all steps ---------------------------------------------
        logits = f32(TB, V)
        Xo = OUT if don else HL[B:]
        gemm(Xo, w.W_lo, TB, V, R, bias=w.b_lo, C=logits)
        L = sp.seq_length
        ctx.pc, ctx.drop, ctx.sp, ctx.T, ctx.fused, ctx.w, ctx.B = pc, drop, sp, T, fused_nll, w, B
        ctx.saved = (seq, masks, bef, aft, diff, emb0, EMBD, XT, HL, HM, HMf, Cm, Cl, Gm, Gl, OUT, MW, PW, VPOS, ATT, GI2, G1,
                     GATE, GATED, logits, EI)
        ctx.params = params
        sp._last_state = (torch.stack([HMf[TB:], HLf[TB:]]), torch.stack([Cm[TB:], Cl[TB:]]))
        sp._tok_err = err
        if fused_nll:
            masks = _f32c(masks)
            row_loss, res = f32(TB), f32(2)
            call("dec_nll", logits.data_ptr(), V, TB, B, V, seq.data_ptr(), seq.stride(0), masks.data_ptr(), masks.stride(0),
                 2, None, 0, row_loss.data_ptr(), None, None, None, 0, opf, err.data_ptr())
            call("dec_nll_reduce", row_loss.data_ptr(), TB, masks.data_ptr(), masks.stride(0), B,
                 min(L, masks.shape[1] - 1), res.data_ptr())
            ctx.res = res
            ctx.masks = masks
            return res[0].clone()
        outputs = torch.zeros(B, L, V, dtype=torch.float32, device=dev)
        ones = torch.ones(B, T + 1, dtype=torch.float32, device=dev)
        dummy_lab = torch.zeros(B, T + 1, dtype=torch.int64, device=dev)
        call("dec_nll", logits.data_ptr(), V, TB, B, V, dummy_lab.data_ptr(), T + 1, ones.data_ptr(), T + 1, 1,
             outputs.data_ptr(), L, None, None, None, None, 0, opf, None)
        outputs_pos = torch.zeros(B, L, 16, dtype=torch.float32, device=dev)
        outputs_pos[:, :T] = torch.log_softmax(DPOS.view(T, B, 16).transpose(0, 1), dim=2)        # :239 (not in the loss)
        mws = MW.view(T, B, 4)[:, :, :3].transpose(0, 1).contiguous()
        ctx.outputs = outputs
        ctx.mark_non_differentiable(outputs_pos, mws)
        return outputs, outputs_pos, mws

    @staticmethod
    def backward(ctx, *grads):
        pc, drop, sp, T, w, B = ctx.pc, ctx.drop, ctx.sp, ctx.T, ctx.w, ctx.B
        (seq, masks, bef, aft, diff, emb0, EMBD, XT, HL, HM, HMf, Cm, Cl, Gm, Gl, OUT, MW, PW, VPOS, ATT, GI2, G1, GATE, GATED,
         logits, EI) = ctx.saved
        R, E, D, G, V, NH, Wex = w.R, w.E, w.D, w.G, w.V, w.NH, w.Wex
        dev = bef.device
        TB = T * B
        opf, OT = _opf(pc), pc.T
        don = drop is not None and drop.on
        p_lm = float(sp.drop_prob_lm)
        d = (lambda site, p: drop.a(site, p)) if don else (lambda site, p: (None, 0, 0.0))
        keep = 1.0 / (1.0 - p_lm) if don else 1.0
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)      # noqa: E731
        opt = lambda *s: torch.empty(*s, dtype=OT, device=dev)                 # noqa: E731
        # ---- gradient of the logits ------------------------------------------------------------------------------
        Vp = (V + 7) // 8 * 8
        dLG = opt(TB, Vp)
        if ctx.fused:
            g = _f32c(grads[0]).reshape(1)
            call("dec_nll", logits.data_ptr(), V, TB, B, V, seq.data_ptr(), seq.stride(0), ctx.masks.data_ptr(),
                 ctx.masks.stride(0), 4, None, 0, None, g.data_ptr(), ctx.res[1:].data_ptr(), dLG.data_ptr(), Vp, opf, None)
        else:
            dout = _f32c(grads[0])
            call("dec_lsm_bwd", dout.data_ptr(), ctx.outputs.data_ptr(), TB, B, V, sp.seq_length, dLG.data_ptr(), Vp, opf)
        Xo = OUT if don else HL[B:]
        # parameter gradients are written where the optimizer wants them (its flat slot on first use after a reset)
        P = dict(zip(sp._param_names, ctx.params))
        slot = lambda name, *shape: _dst(P[name].data_ptr(), tuple(shape), dev)     # noqa: E731
        dW_lo, db_lo = slot("logit.weight", V, R), slot("logit.bias", V)
        gemm(dLG, Xo, V, R, TB, transA=1, transB=1, C=dW_lo)
        colsum(dLG, TB, V, out=db_lo)
        dOUT = f32(TB, R)
        gemm(dLG, w.W_lo, TB, R, V, transB=1, C=dOUT)
        # ---- the data-gradient chain, step by step ------------------------------------------------------------------
        dS1 = opt(TB, NH)                  # [dgm | d pos1-pre | d gate1x-pre | dgl] per step: operand of the dgrad AND the wgrad
        dGm_f, dGl_f = f32(TB, 4 * R), f32(TB, 4 * R)
        dG1_f = f32(TB, G)
        DG2 = opt(TB, D)
        DGI2 = f32(TB, R + D)
        DFC, DDPOS = f32(TB, 4), f32(TB, 16)
        dbef, ddiff, daft = (torch.zeros(B, D, dtype=torch.float32, device=dev) for _ in range(3))
        dhl, dhm, dhm_fc = f32(B, R), f32(B, R), f32(B, R)
        dcl, dcm = f32(B, R), f32(B, R)
        dgated, dg1, datt_g = f32(B, D), f32(B, G), f32(B, D)
        for t in range(T - 1, -1, -1):
            r0, r1, r2 = t * B, (t + 1) * B, (t + 2) * B
            last = t == T - 1
            s1 = dS1[r0:r1]
            call("dec_lstm_bwd", dOUT[r0:r1].data_ptr(), R, *d(SITE_OUT, p_lm), r0 * R, None if last else dhl.data_ptr(), R,
                 None, 0, None if last else dcl.data_ptr(), Gl[r0:r1].data_ptr(), Cl[r1:r2].data_ptr(), Cl[r0:r1].data_ptr(),
                 B, R, s1[:, w.o_lh:].data_ptr(), NH, opf, dGl_f[r0:r1].data_ptr(), dcl.data_ptr())
            gemm(s1[:, w.o_lh:], w.W_ih_lg, B, D, 4 * R, transB=1, C=dgated)
            call("dec_gate_bwd", dgated.data_ptr(), GATE[r0:r1].data_ptr(), ATT[r0:r1].data_ptr(), B * D, datt_g.data_ptr(),
                 DG2[r0:r1].data_ptr(), opf)
            gemm(DG2[r0:r1], w.W_g2, B, G, D, transB=1, C=dg1)
            call("dec_relu_drop_bwd", dg1.data_ptr(), G, G1[r0:r1].data_ptr(), G, opf, B, G, keep,
                 s1[:, w.o_g1:].data_ptr(), NH, opf, dG1_f[r0:r1].data_ptr(), G)
            gemm(s1[:, w.o_g1:w.o_lh], w.W_g1r, B, R + D, G, transB=1, C=DGI2[r0:r1])
            call("dec_att_bwd", DGI2[r0:r1].data_ptr(), R + D, datt_g.data_ptr(), ctypes.addressof(w.att_ptrs), bef.data_ptr(),
                 diff.data_ptr(), aft.data_ptr(), MW[r0:r1].data_ptr(), PW[r0:r1].data_ptr(), VPOS[r0:r1].data_ptr(), B, R,
                 512, D, drop.seed if don else None, SITE_POS1, p_lm if don else 0.0, SITE_DPOS, 0.5 if don else 0.0, r0,
                 dbef.data_ptr(), ddiff.data_ptr(), daft.data_ptr(), DFC[r0:r1].data_ptr(), DDPOS[r0:r1].data_ptr(),
                 dhm_fc.data_ptr(), s1[:, w.o_p1:].data_ptr(), NH, opf)
            call("dec_lstm_bwd", dhm_fc.data_ptr(), R, None, 0, 0.0, 0, None if last else dhm.data_ptr(), R, None, 0,
                 None if last else dcm.data_ptr(), Gm[r0:r1].data_ptr(), Cm[r1:r2].data_ptr(), Cm[r0:r1].data_ptr(), B, R,
                 s1.data_ptr(), NH, opf, dGm_f[r0:r1].data_ptr(), dcm.data_ptr())
            if t > 0:
                gemm(s1, w.Wcat_h, B, R, NH, transB=1, C=dhl)                    # all four uses of prev_h = h_lang(t-1)
                gemm(s1[:, :4 * R], w.W_hh_m, B, R, 4 * R, transB=1, C=dhm)
        # ---- weight gradients, batched over the steps (K = T * B) -------------------------------------------------
        c = sp.core
        dWcat = f32(NH, R)
        gemm(dS1, HL[:TB], NH, R, TB, transA=1, transB=1, C=dWcat)
        dW_hh_m = slot("core.module_att_lstm.weight_hh", 4 * R, R)
        gemm(dS1[:, :4 * R], HM[:TB], 4 * R, R, TB, transA=1, transB=1, C=dW_hh_m)
        dW_ih_m = slot("core.module_att_lstm.weight_ih", 4 * R, E + R)
        dEMB = f32(TB, E)
        gemm(dS1[:, :4 * R], w.W_ih_me, TB, E, 4 * R, transB=1, C=dEMB)
        demb0 = f32(B, E)
        if don:
            gemm(dS1[:, :4 * R], EMBD, 4 * R, E, TB, transA=1, transB=1, C=dW_ih_m[:, :E])
            call("dec_masked_sum_t", dEMB.data_ptr(), E, T, B, E, *d(SITE_EMBED, p_lm), 0, demb0.data_ptr())
        else:
            # every step saw the same embedded rows: sum the gate gradients over the steps first
            dGsum = dGm_f.view(T, B, 4 * R).sum(0)
            gemm(_op(pc, dGsum), EMBD, 4 * R, E, B, transA=1, transB=1, C=dW_ih_m[:, :E], splits=1)
            demb0 = dEMB.view(T, B, E).sum(0)
        dW_ih_m[:, E:] = dWcat[:4 * R]
        dW_p1 = slot("core.pos1.0.weight", 512, R)
        dW_p1.copy_(dWcat[w.o_p1:w.o_g1])
        dW_g1 = slot("core.gate1x.0.weight", G, G)
        dW_g1[:, :R] = dWcat[w.o_g1:w.o_lh]
        gemm(dS1[:, w.o_g1:w.o_lh], GI2, G, R + D, TB, transA=1, transB=1, C=dW_g1[:, R:])
        dW_g2 = slot("core.gate2x.weight", D, G)
        gemm(DG2, G1, D, G, TB, transA=1, transB=1, C=dW_g2)
        dW_ih_l = slot("core.lang_lstm.weight_ih", 4 * R, w.We + D)
        gemm(dS1[:, w.o_lh:], GATED, 4 * R, D, TB, transA=1, transB=1, C=dW_ih_l[:, w.We:])
        dWx = f32(4 * R, Wex)
        gemm(dS1[:, w.o_lh:], XT, 4 * R, Wex, TB, transA=1, transB=1, C=dWx)
        dW_ih_l[:, :w.We] = dWx[:, :w.We]
        dW_hh_l = slot("core.lang_lstm.weight_hh", 4 * R, R)
        dW_hh_l.copy_(dWcat[w.o_lh:])
        # biases: column sums of the pre-activation gradients (fp32 copies)
        db_m = colsum(dGm_f, TB, 4 * R, out=slot("core.module_att_lstm.bias_ih", 4 * R))
        db_m2 = slot("core.module_att_lstm.bias_hh", 4 * R)
        db_m2.copy_(db_m)
        db_l = colsum(dGl_f, TB, 4 * R, out=slot("core.lang_lstm.bias_ih", 4 * R))
        db_l2 = slot("core.lang_lstm.bias_hh", 4 * R)
        db_l2.copy_(db_l)
        db_g1 = colsum(dG1_f, TB, G, out=slot("core.gate1x.0.bias", G))
        db_g2 = colsum(DG2, TB, D, out=slot("core.gate2x.bias", D))
        db_p1 = colsum(dS1[:, w.o_p1:w.o_g1], TB, 512, out=slot("core.pos1.0.bias", 512))
        # word embedding: dXT = dgl W_ih[:, :We] -> ReLU / Dropout mask (XT > 0) -> scatter-add over the tokens
        dXT = f32(TB, Wex)
        gemm(dS1[:, w.o_lh:], w.W_ih_x, TB, Wex, 4 * R, transB=1, C=dXT)
        dXTm = f32(TB, Wex)
        call("dec_relu_drop_bwd", dXT.data_ptr(), Wex, XT.data_ptr(), Wex, opf, TB, Wex, keep, None, 0, 0, dXTm.data_ptr(), Wex)
        tok = seq[:, :T].t().reshape(-1)
        demb_w = slot("embed.0.weight", V, w.We)
        demb_w.zero_().index_add_(0, tok, dXTm[:, :w.We])
        # core.embed: ReLU mask, then its own weight / input gradients
        demb0p = f32(B, E)
        call("dec_relu_drop_bwd", demb0.data_ptr(), E, emb0.data_ptr(), E, 0, B, E, 1.0, None, 0, 0, demb0p.data_ptr(), E)
        dE_op = _op(pc, demb0p)
        dW_e = slot("core.embed.0.weight", E, 3 * D)
        gemm(dE_op, EI, E, 3 * D, B, transA=1, transB=1, C=dW_e, splits=1)
        db_e = colsum(demb0p, B, E, out=slot("core.embed.0.bias", E))
        dEI = f32(B, 3 * D)
        gemm(dE_op, w.W_e, B, 3 * D, E, transB=1, C=dEI)
        dbef += dEI[:, :D]
        ddiff += dEI[:, D:2 * D]
        daft += dEI[:, 2 * D:]
        # the three tiny layers
        dW_fc, dW_wp = slot("core.weight_fc.0.weight", 3, R), slot("core.weight_pos.weight", 16, 512)
        dW_p2 = slot("core.pos2.weight", R, 16)
        nck = max(1, min(64, TB // 32))            # row chunks: partial sums, then the deterministic column-sum kernel

        def outer_small(a, lda, m, bmat, ldb, n, out, transpose):
            part = f32(nck, m * n)
            call("dec_outer_small", a.data_ptr(), lda, m, bmat.data_ptr(), ldb, n, TB, part.data_ptr(), nck, transpose)
            colsum(part, nck, m * n, out=out.view(-1))

        outer_small(DFC, 4, 3, HMf[B:], R, R, dW_fc, 0)
        outer_small(DDPOS, 16, 16, VPOS, 512, 512, dW_wp, 0)
        outer_small(PW, 16, 16, DGI2, R + D, R, dW_p2, 1)
        db_fc = slot("core.weight_fc.0.bias", 3)
        db_fc.copy_(colsum(DFC, TB, 4)[:3])
        db_wp = colsum(DDPOS, TB, 16, out=slot("core.weight_pos.bias", 16))
        db_p2 = colsum(DGI2, TB, R, out=slot("core.pos2.bias", R))
        by_name = {
            "embed.0.weight": demb_w, "core.embed.0.weight": dW_e, "core.embed.0.bias": db_e,
            "core.module_att_lstm.weight_ih": dW_ih_m, "core.module_att_lstm.weight_hh": dW_hh_m,
            "core.module_att_lstm.bias_ih": db_m, "core.module_att_lstm.bias_hh": db_m2,
            "core.weight_fc.0.weight": dW_fc, "core.weight_fc.0.bias": db_fc,
            "core.pos1.0.weight": dW_p1, "core.pos1.0.bias": db_p1,
            "core.weight_pos.weight": dW_wp, "core.weight_pos.bias": db_wp,
            "core.pos2.weight": dW_p2, "core.pos2.bias": db_p2,
            "core.gate1x.0.weight": dW_g1, "core.gate1x.0.bias": db_g1,
            "core.gate2x.weight": dW_g2, "core.gate2x.bias": db_g2,
            "core.lang_lstm.weight_ih": dW_ih_l, "core.lang_lstm.weight_hh": dW_hh_l,
            "core.lang_lstm.bias_ih": db_l, "core.lang_lstm.bias_hh": db_l2,
            "logit.weight": dW_lo, "logit.bias": db_lo,
        }
        pg = tuple(by_name[n] for n in sp._param_names)
        return (None, None, None, None, None, None, None, dbef, daft, ddiff) + pg


class DynamicSpeaker(nn.Module):
    """dynamic_speaker_change_pos.py:139-357 (beam search excluded: the reference's test script decodes greedily,
    test_mimic.py:119-122)."""

    def __init__(self, cfg, vocab_size=0):
        super().__init__()
        sp = cfg.model.speaker
        self.vocab_size = sp.vocab_size if vocab_size == 0 else vocab_size
        self.word_embed_size = sp.word_embed_size
        self.rnn_size = sp.rnn_size
        self.drop_prob_lm = sp.drop_prob_lm
        self.seq_length = sp.seq_length
        self.ss_prob = 0.0
        self.embed = nn.Sequential(nn.Embedding(self.vocab_size, self.word_embed_size), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.core = DynamicCore(cfg)
        self.rnn_num_layers = self.core.rnn_num_layers
        self.logit_layers = getattr(sp, 'logit_layers', 1)
        if self.logit_layers != 1:
            raise NotImplementedError("logit_layers > 1 (the reference's config uses the single Linear logit layer)")
        self.logit = nn.Linear(self.rnn_size, self.vocab_size)
        if self.rnn_size != 512:
            raise ValueError("DynamicCore.pos1 is Linear(512, 512) on the previous hidden state: rnn_size must be 512")
        self.module_weights = []
        self.precision = _default_precision()
        self._param_names = [n for n, _ in self.named_parameters()]
        self._last_state = None
        self._tok_err = None
        self._runners = {}

    def set_precision(self, precision: str):
        PC(precision)
        self.precision = precision
        return self

    def init_hidden(self, batch_size):
        weight = next(self.parameters())
        return (weight.new_zeros(self.rnn_num_layers, batch_size, self.rnn_size),
                weight.new_zeros(self.rnn_num_layers, batch_size, self.rnn_size))

    def _steps(self, seq) -> int:
        """Number of steps the reference's loop executes (:210-214): it stops at the first column i >= 1 that is all zero."""
        L = min(self.seq_length, seq.shape[1])
        nz = (seq[:, :L] != 0).any(0)
        nz[0] = True
        idx = (~nz).nonzero()
        return int(idx[0]) if idx.numel() else L

    def _run(self, feat_bef, feat_aft, feat_diff, seq, masks, fused, steps):
        if self.training and self.ss_prob > 0.0:
            raise NotImplementedError("scheduled sampling (ss_prob > 0) is not implemented; the reference's config keeps "
                                      "it off (scheduled_sampling_start = -1)")
        pc = PC(self.precision)
        dev = feat_bef.device
        if seq.dim() != 2 or seq.shape[0] != feat_bef.shape[0]:
            raise ValueError("seq must be [batch, tokens] with the batch of the features, got %s" % (tuple(seq.shape),))
        seq = seq.to(device=dev, dtype=torch.int64)
        T = int(steps) if steps is not None else self._steps(seq)
        if T < 1 or T > min(self.seq_length, seq.shape[1]):
            raise ValueError("steps=%d outside 1..%d" % (T, min(self.seq_length, seq.shape[1])))
        if fused:
            if masks is None or tuple(masks.shape) != tuple(seq.shape):
                raise ValueError("masked_nll needs masks of the labels' shape %s" % (tuple(seq.shape),))
            if seq.shape[1] < T + 1:
                raise ValueError("labels need %d columns (step t predicts column t + 1), got %d" % (T + 1, seq.shape[1]))
            masks = masks.to(device=dev, dtype=torch.float32)
        if self.training:
            rng_advance(dev)
        drop = Drop(dev, self.training)
        params = [p for _, p in self.named_parameters()]
        return SpeakerSeqFn.apply(pc, drop, self, T, fused, seq, masks, feat_bef, feat_aft, feat_diff, *params)

    def _forward(self, feat_bef, feat_aft, feat_diff, seq, steps: Optional[int] = None):
        """-> (log-probs [B, seq_length, V], log position probs [B, seq_length, 16]); rows of steps the loop did not reach
        are zero, like the reference's."""
        outputs, outputs_pos, mws = self._run(feat_bef, feat_aft, feat_diff, seq, None, False, steps)
        self.module_weights = list(mws.unbind(1))
        return outputs, outputs_pos

    def masked_nll(self, feat_bef, feat_aft, feat_diff, labels, masks, steps: Optional[int] = None):
        """lang_criterion(_forward(...)[0], labels[:, 1:], masks[:, 1:]) (train_mimic.py:236-242, utils/utils.py:204-216)
        with log-softmax, gather, mask and normalisation fused into one pass over the logits; differentiable.
        steps: run a fixed number of steps (CUDA-graph capture); steps beyond the last non-empty column carry mask 0 and
        change neither the loss nor any gradient."""
        return self._run(feat_bef, feat_aft, feat_diff, labels, masks, True, steps)

    def get_module_weights(self):
        if len(self.module_weights) == 0:
            print('no module weights accumulated')
            return None
        return torch.stack(self.module_weights, dim=1)

    # ---- inference -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def get_logprobs_state(self, it, feat_bef, feat_aft, feat_diff, state):
        """:225-240, one eval-mode step: (log_probs [B,V], state, log_pos_probs [B,16])."""
        if self.training:
            raise NotImplementedError("get_logprobs_state is the inference step; training goes through _forward / masked_nll")
        r = _StepRunner(self, feat_bef.shape[0], feat_bef.device)
        r.load(feat_bef, feat_aft, feat_diff, state)
        r.prepare()
        r.tok.copy_(it.view(-1))
        r.step(0, sample=False)
        self.core.module_weights = r.mw[:, :3].clone()
        self.module_weights.append(self.core.module_weights)
        return torch.log_softmax(r.logits, dim=1), r.state(), torch.log_softmax(r.dpos, dim=1)

    @torch.no_grad()
    def _sample(self, feat_bef, feat_aft, feat_diff, seq, cfg={}, sample_max=0, check_every=16, use_graph=True):
        """:287-357 with beam_size = 1: greedy (sample_max = 1, the reference's test path, test_mimic.py:119-122) or
        multinomial sampling from exp(logprobs / temperature) (sample_max = 0, :341-349; the uniforms come from the
        library's counter RNG, so the draws are not torch.multinomial's -- their distribution is)
        -> (seq [B, seq_length] int64, seq_logprobs [B, seq_length]).
        The token loop runs in blocks of `check_every` steps; the host reads the device-side stop flag once per block
        (0 = never, always seq_length steps).  use_graph: every block is a captured CUDA graph, kept per batch size."""
        sp = cfg.model.speaker if hasattr(cfg, "model") else {}
        if (sp.get('beam_size', 1) if hasattr(sp, "get") else 1) > 1:
            raise NotImplementedError("beam search is not implemented: the reference's test script uses beam_size 1, and its "
                                      "_sample_beam cannot run (it unpacks get_logprobs_state's three results into two "
                                      "names, dynamic_speaker_change_pos.py:273)")
        temperature = float(sp.get('temperature', 1.0) if hasattr(sp, "get") else 1.0)
        if not sample_max and temperature <= 0.0:
            raise ValueError("temperature must be positive, got %r" % (temperature,))
        if self.training:
            raise NotImplementedError("_sample is an inference entry point: call eval() first")
        self.module_weights = []
        B, dev = feat_bef.shape[0], feat_bef.device
        T = self.seq_length
        ce = check_every if check_every and check_every > 0 else T + 1
        key = (B, str(dev), self.precision, ce, bool(use_graph), T, bool(sample_max), temperature,
               self.logit.weight.data_ptr())
        r = self._runners.get(key)
        if r is None:
            if len(self._runners) > 4:
                self._runners.clear()
            r = _StepRunner(self, B, dev)
            r.multinomial, r.temperature = (0 if sample_max else 1), temperature
            self._runners[key] = r
        if not sample_max:
            rng_advance(dev)             # a fresh stream of uniforms per call (a kernel: valid inside a captured graph too)
        r.load(feat_bef, feat_aft, feat_diff, None)
        blocks = [(t0, min(t0 + ce, T + 1)) for t0 in range(0, T + 1, ce)]
        for bi, (t0, t1) in enumerate(blocks):
            if use_graph:
                g = r.graphs.get((t0, t1))
                if g is None:
                    if not r.graphs:
                        # first use: one eager pass warms every kernel up outside the capture (and is thrown away)
                        r.run_block(0, min(2, T + 1), first=True)
                        r.load(feat_bef, feat_aft, feat_diff, None)
                        torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    before = lib.LAUNCHES
                    with lib.graph_capture(g):
                        r.run_block(t0, t1, first=(bi == 0))
                    r.graphs[(t0, t1)] = g
                    r.graph_launches[(t0, t1)] = lib.LAUNCHES - before
                g.replay()
                lib.LAUNCHES += r.graph_launches[(t0, t1)]
            else:
                r.run_block(t0, t1, first=(bi == 0))
            if bi + 1 < len(blocks):
                r.flag_host.copy_(r.running)
                if int(r.flag_host[0]) == 0:
                    break
        return r.seq.clone(), r.seq_lp.clone()


class _StepRunner:
    """Buffers and launches of the eval-mode decode step for one batch size (static addresses: its blocks of steps can be
    captured into CUDA graphs and replayed on new inputs)."""

    def __init__(self, sp: DynamicSpeaker, B: int, dev):
        lib.require_device()
        self.sp, self.B, self.dev = sp, B, dev
        self.pc = pc = PC(sp.precision)
        c = sp.core
        R, E, D, V = sp.rnn_size, c.embed_dim, c.input_dim, sp.vocab_size
        G, T = 2 * R + D, sp.seq_length
        self.NH = 4 * R + 512 + G + 4 * R
        OT = pc.T
        f32 = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)      # noqa: E731
        opt = lambda *s: torch.zeros(*s, dtype=OT, device=dev)                 # noqa: E731
        self.bef, self.aft, self.diff = f32(B, D), f32(B, D), f32(B, D)
        self.h0, self.c0 = f32(2, B, R), f32(2, B, R)
        self.hm, self.hl = opt(B, R), opt(B, R)
        self.hmf, self.hlf, self.cm, self.cl = f32(B, R), f32(B, R), f32(B, R), f32(B, R)
        self.gm, self.gl = f32(B, 4 * R), f32(B, 4 * R)
        self.S1, self.S2, self.S3, self.S6, self.g2pre = f32(B, self.NH), f32(B, 4 * R), f32(B, 4 * R), f32(B, 4 * R), f32(B, D)
        self.mw, self.pw, self.dpos, self.vpos, self.att = f32(B, 4), f32(B, 16), f32(B, 16), f32(B, 512), f32(B, D)
        self.gi2, self.g1, self.gate, self.gated = opt(B, R + D), opt(B, G), f32(B, D), opt(B, D)
        self.logits = f32(B, V)
        self.seq = torch.zeros(B, T, dtype=torch.int64, device=dev)
        self.seq_lp = f32(B, T)
        self.unfinished = torch.ones(B, dtype=torch.uint8, device=dev)
        self.running = torch.ones(1, dtype=torch.int32, device=dev)
        self.tok = torch.full((B,), 2, dtype=torch.int64, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.flag_host = torch.ones(1, dtype=torch.int32).pin_memory()
        self.vocab = torch.arange(V, device=dev, dtype=torch.int64).view(V, 1)
        self.graphs, self.graph_launches = {}, {}
        self.w = None
        self.multinomial, self.temperature = 0, 1.0

    def load(self, bef, aft, diff, state):
        self.bef.copy_(bef)
        self.aft.copy_(aft)
        self.diff.copy_(diff)
        if state is None:
            self.h0.zero_()
            self.c0.zero_()
        else:
            self.h0.copy_(state[0])
            self.c0.copy_(state[1])

    def state(self):
        return torch.stack([self.hmf, self.hlf]), torch.stack([self.cm, self.cl])

    def prepare(self):
        """Per sequence: operand copies of the weights, the step-invariant embedding products, the token table, state."""
        pc, sp = self.pc, self.sp
        self.w = w = _W(pc, sp)
        B, dev = self.B, self.dev
        R, E, D, V, Wex = w.R, w.E, w.D, w.V, w.Wex
        OT, opf = pc.T, _opf(pc)
        EI = torch.empty(B, 3 * D, dtype=OT, device=dev)
        cast_many(pc, [(self.bef, EI[:, :D]), (self.diff, EI[:, D:2 * D]), (self.aft, EI[:, 2 * D:])])
        emb0 = torch.empty(B, E, dtype=OT, device=dev)
        _gemm_op(pc, EI, w.W_e, B, E, 3 * D, emb0, bias=w.b_e, act=ACT_RELU)
        gemm(emb0, w.W_ih_me, B, 4 * R, E, C=self.S3)
        # word-embedding half of the language LSTM's input product as a [V, 4R] table: one row per possible token
        ET = torch.empty(V, Wex, dtype=OT, device=dev)
        call("dec_embed", self.vocab.data_ptr(), 1, 0, 0, V, V, w.emb.data_ptr(), V, w.We, ET.data_ptr(), Wex, opf, None, 0, 0.0,
             self.err.data_ptr())
        self.TBL = torch.empty(V, 4 * R, dtype=torch.float32, device=dev)
        gemm(ET, w.W_ih_x, V, 4 * R, Wex, C=self.TBL)
        self.hmf.copy_(self.h0[0])
        self.hlf.copy_(self.h0[1])
        self.cm.copy_(self.c0[0])
        self.cl.copy_(self.c0[1])
        cast_many(pc, [(self.hmf, self.hm), (self.hlf, self.hl)])
        self.seq.zero_()
        self.seq_lp.zero_()
        self.unfinished.fill_(1)
        self.running.fill_(1)
        self.tok.fill_(2)                                  # <bos> (:303)

    def step(self, t: int, sample: bool = True):
        pc, w, B = self.pc, self.w, self.B
        R, D, G, V, NH = w.R, w.D, w.G, w.V, w.NH
        opf = _opf(pc)
        gemm(self.hl, w.Wcat_h, B, NH, R, C=self.S1)
        gemm(self.hm, w.W_hh_m, B, 4 * R, R, C=self.S2)
        call("dec_lstm_fwd", self.S1.data_ptr(), NH, self.S2.data_ptr(), 4 * R, self.S3.data_ptr(), 4 * R, None, None,
             w.b_mi.data_ptr(), w.b_mh.data_ptr(), self.cm.data_ptr(), B, R, self.gm.data_ptr(), self.cm.data_ptr(),
             self.hmf.data_ptr(), self.hm.data_ptr(), R, None, 0, opf, None, 0, 0.0, 0)
        call("dec_att_fwd", self.hmf.data_ptr(), self.S1[:, w.o_p1:].data_ptr(), NH, ctypes.addressof(w.att_ptrs),
             self.bef.data_ptr(), self.diff.data_ptr(), self.aft.data_ptr(), B, R, 512, D, None, 0, 0.0, 0, 0.0, 0,
             self.mw.data_ptr(), self.pw.data_ptr(), self.dpos.data_ptr(), self.vpos.data_ptr(), self.att.data_ptr(),
             self.gi2.data_ptr(), R + D, opf)
        _gemm_op(pc, self.gi2, w.W_g1r, B, G, R + D, self.g1, bias=w.b_g1, addend=self.S1[:, w.o_g1:w.o_lh], act=ACT_RELU)
        gemm(self.g1, w.W_g2, B, D, G, bias=w.b_g2, C=self.g2pre)
        call("dec_gate_fwd", self.g2pre.data_ptr(), self.att.data_ptr(), B * D, self.gate.data_ptr(), self.gated.data_ptr(), opf)
        gemm(self.gated, w.W_ih_lg, B, 4 * R, D, C=self.S6)
        call("dec_lstm_fwd", self.S6.data_ptr(), 4 * R, self.S1[:, w.o_lh:].data_ptr(), NH, None, 0, self.TBL.data_ptr(),
             self.tok.data_ptr(), w.b_li.data_ptr(), w.b_lh.data_ptr(), self.cl.data_ptr(), B, R, self.gl.data_ptr(),
             self.cl.data_ptr(), self.hlf.data_ptr(), self.hl.data_ptr(), R, None, 0, opf, None, 0, 0.0, 0)
        gemm(self.hl, w.W_lo, B, V, R, bias=w.b_lo, C=self.logits)
        if sample:
            T = self.sp.seq_length
            from .functions import rng_state
            call("dec_token", self.logits.data_ptr(), V, B, V, t, T, self.seq.data_ptr(), self.seq_lp.data_ptr(),
                 self.unfinished.data_ptr(), self.running.data_ptr(), self.tok.data_ptr(), None, self.multinomial,
                 float(self.temperature), rng_state(self.dev).data_ptr(), SITE_SAMPLE)

    def run_block(self, t0: int, t1: int, first: bool):
        """Steps t0 .. t1-1 of the reference's `for t in range(seq_length + 1)` loop (:299); its last iteration only advances
        the state (:334-335), so nothing is launched for it."""
        if first:
            self.prepare()
        for t in range(t0, min(t1, self.sp.seq_length)):
            self.step(t)


class LanguageModelCriterion(nn.Module):
    """utils/utils.py:204-216 on log-probabilities (plain tensor ops: the fused form is DynamicSpeaker.masked_nll)."""

    def forward(self, input, target, mask, emphasize_last=False):
        target = target[:, :input.size(1)]
        mask = mask[:, :input.size(1)]
        output = -input.gather(2, target.unsqueeze(2)).squeeze(2) * mask
        return torch.sum(output) / torch.sum(mask)
