"""Shared test helpers: golden fixtures, seeded inputs/weights, error metrics."""
import json
import os

import numpy as np
import torch

from ekaid_b200.config import default_cfg
from ekaid_b200.synthetic import synthetic_batch, synthetic_state_dict
from oracle import ekaid_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OUT_NAMES = ("pred", "att_bef", "att_aft", "attended_1", "attended_2", "input_attended")
GRAD_PROBE = 16

CASES = ["c0_b2_n52_all", "c1_b3_n52_all_grads", "c2_b2_n52_semantic", "c2_b2_n52_spatial", "c2_b2_n52_implicit",
         "c2_b2_n52_ips", "c3_b2_n26_all", "c4_b2_n60_k52_all", "c5_b1_n52_all", "c6_b2_n52_empty_image",
         "c7_b2_n52_zero_bias", "c8_b2_n126_k126_all"]


def spec_for(graph):
    name = {"all": "state_dict_spec.json", "i+s": "state_dict_spec_ips.json"}.get(graph, "state_dict_spec_%s.json" % graph)
    return {k: tuple(v) for k, v in json.load(open(os.path.join(GOLDEN, name))).items()}


def speaker_spec():
    return {k: tuple(v) for k, v in json.load(open(os.path.join(GOLDEN, "speaker_spec.json"))).items()}


def checksum(t):
    t = t.detach().double().flatten().cpu()
    idx = torch.arange(t.numel(), dtype=torch.float64)
    return np.array([float(t.sum()), float((t * torch.cos(idx * 0.37)).sum())])


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return z, meta


def case_inputs(meta):
    """(state_dict, inputs 9-tuple as ChangeDetector.forward takes them, raw batch) regenerated from seeds (CPU)."""
    sd = synthetic_state_dict(spec_for(meta["graph"]), meta["weight_seed"])
    if meta["zero_img_bias"]:
        sd["img.bias"].zero_()
    batch = synthetic_batch(meta["B"], meta["N"], seed=meta["seed"])
    N = meta["N"]
    inp = (batch[0], batch[1], O.process_matrix(batch[6], N, 11), O.process_matrix(batch[7], N, 11),
           O.process_matrix(batch[8], N, 3), O.process_matrix(batch[9], N, 3), batch[10], batch[11], batch[12])
    return sd, inp, batch


def check_fixture_inputs(z, sd, inp):
    got = np.concatenate([checksum(t) for t in inp])
    np.testing.assert_allclose(got, z["in_checksum"], rtol=1e-9, atol=1e-6,
                               err_msg="seeded inputs differ from the ones the golden file was made with")
    gw = np.concatenate([checksum(sd[k]) for k in sorted(sd)][:40])
    np.testing.assert_allclose(gw, z["w_checksum"], rtol=1e-9, atol=1e-6,
                               err_msg="seeded weights differ from the ones the golden file was made with")


def oracle_forward(sd, inp, meta, dtype=torch.float32, return_aux=False):
    if meta["empty_image"]:
        inp = tuple(torch.ones_like(t) if i < 8 else t for i, t in enumerate(inp))
    inp = tuple(t.to(dtype) if (t.is_floating_point() and i < 6) else t for i, t in enumerate(inp))
    sdd = {k: v.to(dtype) for k, v in sd.items()}
    cd = default_cfg().model.change_detector
    return O.change_detector_forward(sdd, *inp, graph=meta["graph"], num_heads=cd.att_head,
                                     nongt_dim=meta["nongt_dim"], pos_emb_dim=cd.pos_emb_dim,
                                     coef_sem=cd.coef_sem, coef_spa=cd.coef_spa, return_aux=return_aux)


def rel_err(a, b):
    """max |a-b| / max |b|  -- the tolerance metric of the parity bar."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def grad_summary(g):
    g = g.detach().flatten().double().cpu()
    n = g.numel()
    idx = (torch.arange(GRAD_PROBE, dtype=torch.int64) * 7919 + 13) % n
    return np.concatenate([[float(g.norm()), float(g.sum())], g[:GRAD_PROBE if n >= GRAD_PROBE else n].numpy(),
                           g[idx].numpy()])


def loss_weights(outs):
    gw = torch.Generator().manual_seed(99)
    return [torch.randn(o.shape, generator=gw) for o in outs[1:]]
