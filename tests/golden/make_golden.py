"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference/model).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz + state_dict_spec.json.  Inputs and weights are NOT stored: they are
regenerated from seeds by ekaid_b200.synthetic (a checksum of each is stored so drift is detected).

Test-side shims (reference files untouched):
  1. config merge through yaml.safe_load (configs/config.py:168 calls yaml.load without Loader);
  2. torch.Tensor.cuda -> identity on CPU (models/graph_att_layer.py:131);
  3. models.modules.torch_extract_position_embedding rebound with device=cpu (utils/mimic_utils.py:192-193).
"""
import functools
import io
import json
import contextlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
REF = "/root/reference/model"
sys.path.insert(0, REF)

from ekaid_b200.synthetic import synthetic_batch, synthetic_state_dict   # noqa: E402
from ekaid_b200.config import WORD_TO_IDX                                # noqa: E402
from oracle import ekaid_oracle as O                                     # noqa: E402

import yaml                                                              # noqa: E402
from configs import config as refcfg                                    # noqa: E402
from utils.attr_dict import AttrDict as RefAttrDict                     # noqa: E402


def to_ref_attr(d):
    if isinstance(d, dict):
        a = RefAttrDict()
        for k, v in d.items():
            a[k] = to_ref_attr(v)
        return a
    return d


def ref_cfg(graph="all", nongt_dim=52, empty_image=False):
    with open(os.path.join(REF, "configs/dynamic/dynamic_change_pos_mimic.yaml")) as f:
        y = to_ref_attr(yaml.safe_load(f))
    refcfg.merge_cfg_from_cfg(y)                      # shim 1
    cfg = refcfg.cfg
    cfg.data.feature_mode = "both"
    cfg.train.graph = graph
    cfg.train.setting = "mode2"
    cfg.model.change_detector.nongt_dim = nongt_dim
    cfg.data.train.empty_image = empty_image
    return cfg


torch.Tensor.cuda = lambda self, *a, **k: self        # shim 2
import models.modules as refmod                       # noqa: E402
from utils import mimic_utils as refutils             # noqa: E402
refmod.torch_extract_position_embedding = functools.partial(
    refutils.torch_extract_position_embedding, device=torch.device("cpu"))   # shim 3
from models.dynamic_speaker_change_pos import DynamicSpeaker            # noqa: E402


def build_ref(graph="all", nongt_dim=52, empty_image=False, seed=1238, zero_img_bias=False):
    cfg = ref_cfg(graph, nongt_dim, empty_image)
    with contextlib.redirect_stdout(io.StringIO()):
        m = refmod.ChangeDetector(cfg, WORD_TO_IDX)
    spec = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = synthetic_state_dict(spec, seed)
    if zero_img_bias:
        sd["img.bias"].zero_()
    m.load_state_dict(sd)
    m.eval()
    return m, cfg, spec, sd


def checksum(t):
    t = t.detach().double().flatten()
    idx = torch.arange(t.numel(), dtype=torch.float64)
    return np.array([float(t.sum()), float((t * torch.cos(idx * 0.37)).sum())])


def prep_inputs(batch, cfg, N):
    (d_feats, q_feats, labels, _, masks, _, d_adj, q_adj, d_sem, q_sem, d_bb, q_bb, question) = batch
    dev = torch.device("cpu")
    pm = refutils.process_matrix
    return (d_feats, q_feats, pm(d_adj, cfg, N, dev, "spatial"), pm(q_adj, cfg, N, dev, "spatial"),
            pm(d_sem, cfg, N, dev, "semantic"), pm(q_sem, cfg, N, dev, "semantic"), d_bb, q_bb, question)


GRAD_PROBE = 16


def grad_summary(p):
    g = p.grad.detach().flatten().double()
    n = g.numel()
    idx = (torch.arange(GRAD_PROBE, dtype=torch.int64) * 7919 + 13) % n
    return np.concatenate([[float(g.norm()), float(g.sum())], g[:GRAD_PROBE if n >= GRAD_PROBE else n].numpy(),
                           g[idx].numpy()])


def run_case(name, B, N, graph="all", nongt_dim=52, empty_image=False, seed=1234, zero_img_bias=False,
             with_grads=False, with_tokens=False):
    m, cfg, spec, sd = build_ref(graph, nongt_dim, empty_image, zero_img_bias=zero_img_bias)
    batch = synthetic_batch(B, N, seed=seed)
    if zero_img_bias:
        # make some nodes' projected rows exactly zero so quirk Q10 fires
        pass
    inp = prep_inputs(batch, cfg, N)
    rec = {"meta": json.dumps(dict(B=B, N=N, graph=graph, nongt_dim=nongt_dim, empty_image=empty_image,
                                   seed=seed, zero_img_bias=zero_img_bias, weight_seed=1238))}
    rec["in_checksum"] = np.concatenate([checksum(t) for t in inp])
    rec["w_checksum"] = np.concatenate([checksum(sd[k]) for k in sorted(sd)][:40])
    with torch.no_grad():
        outs = m(*[t.clone() if torch.is_tensor(t) else t for t in inp], setting="mode2", graph=graph)
    for k, o in zip(("pred", "att_bef", "att_aft", "attended_1", "attended_2", "input_attended"), outs):
        rec[k] = o.numpy()
    # oracle must agree with the reference before the fixture is accepted
    with torch.no_grad():
        cd = cfg.model.change_detector
        oo = O.change_detector_forward(
            sd, *[t.clone() if torch.is_tensor(t) else t for t in inp], graph=graph,
            num_heads=cd.att_head, nongt_dim=nongt_dim, pos_emb_dim=cd.pos_emb_dim,
            coef_sem=cd.coef_sem, coef_spa=cd.coef_spa) if not empty_image else None
    if oo is not None:
        for a, b in zip(outs, oo):
            err = float((a - b).abs().max() / (a.abs().max() + 1e-12))
            assert err < 2e-5, (name, err)
            print(f"  {name}: oracle-vs-reference max-rel {err:.2e}")
    if with_grads:
        m2, _, _, _ = build_ref(graph, nongt_dim, empty_image, zero_img_bias=zero_img_bias)
        gw = torch.Generator().manual_seed(99)
        outs2 = m2(*[t.clone() if torch.is_tensor(t) else t for t in inp], setting="mode2", graph=graph)
        loss = 0
        for o in outs2[1:]:
            loss = loss + (o * torch.randn(o.shape, generator=gw)).sum()
        loss.backward()
        rec["loss"] = np.array(float(loss))
        names = []
        rows = []
        for k, p in m2.named_parameters():
            if p.grad is None:
                continue
            names.append(k)
            rows.append(grad_summary(p))
        rec["grad_names"] = np.array(names)
        width = max(len(r) for r in rows)
        rec["grad_rows"] = np.stack([np.pad(r, (0, width - len(r))) for r in rows])
    if with_tokens:
        with contextlib.redirect_stdout(io.StringIO()):
            sp = DynamicSpeaker(cfg, 148)
        sspec = {k: tuple(v.shape) for k, v in sp.state_dict().items()}
        ssd = synthetic_state_dict(sspec, 4321)
        sp.load_state_dict(ssd)
        sp.eval()
        with torch.no_grad():
            seq, _ = sp._sample(outs[3], outs[4], outs[5], None, cfg, sample_max=1)
            labels = batch[2].squeeze(1)
            logp, _ = sp._forward(outs[3], outs[4], outs[5], labels)
        rec["tokens"] = seq.numpy()
        rec["tf_logp_checksum"] = checksum(logp)
        o_seq = O.speaker_greedy(ssd, outs[3], outs[4], outs[5], cfg.model.speaker.seq_length,
                                 cfg.model.speaker.rnn_size)
        assert torch.equal(o_seq, seq), "oracle speaker tokens differ from reference"
        o_logp = O.speaker_teacher_forced(ssd, outs[3], outs[4], outs[5], labels,
                                          cfg.model.speaker.seq_length, cfg.model.speaker.rnn_size)
        assert float((o_logp - logp).abs().max()) < 1e-4
        json.dump({k: list(v) for k, v in sspec.items()}, open(os.path.join(HERE, "speaker_spec.json"), "w"), indent=0)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    return spec


def main():
    torch.manual_seed(0)
    spec = run_case("c0_b2_n52_all", 2, 52, with_tokens=True)
    json.dump({k: list(v) for k, v in spec.items()}, open(os.path.join(HERE, "state_dict_spec.json"), "w"), indent=0)
    run_case("c1_b3_n52_all_grads", 3, 52, seed=77, with_grads=True, with_tokens=True)
    for g in ("semantic", "spatial", "implicit", "i+s"):
        spec_g = run_case("c2_b2_n52_" + g.replace("+", "p"), 2, 52, graph=g, seed=5)
        json.dump({k: list(v) for k, v in spec_g.items()},
                  open(os.path.join(HERE, "state_dict_spec_%s.json" % g.replace("+", "p")), "w"), indent=0)
    run_case("c3_b2_n26_all", 2, 26, seed=6)
    run_case("c4_b2_n60_k52_all", 2, 60, seed=7)                       # Q9 + Q13
    run_case("c5_b1_n52_all", 1, 52, seed=8)
    run_case("c6_b2_n52_empty_image", 2, 52, seed=9, empty_image=True)  # multi-hot adjacency
    run_case("c7_b2_n52_zero_bias", 2, 52, seed=10, zero_img_bias=True, with_grads=True)   # Q10
    run_case("c8_b2_n126_k126_all", 2, 126, nongt_dim=126, seed=11)    # stress shape
    # process_matrix fixture
    b = synthetic_batch(2, 52, seed=3)
    cfg = ref_cfg()
    pm = refutils.process_matrix(b[6], cfg, 52, torch.device("cpu"), "spatial")
    assert torch.equal(pm, O.process_matrix(b[6], 52, 11))
    ps = refutils.process_matrix(b[8], cfg, 52, torch.device("cpu"), "semantic")
    assert torch.equal(ps, O.process_matrix(b[8], 52, 3))
    np.savez_compressed(os.path.join(HERE, "process_matrix.npz"), spa_sum=pm.sum((0, 3)).numpy(),
                        sem_sum=ps.sum((0, 3)).numpy(), spa_chk=checksum(pm), sem_chk=checksum(ps))
    print("golden written")


if __name__ == "__main__":
    main()
