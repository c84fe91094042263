"""Synthetic feature loader and synthetic weights.

Replaces the reference's HDF5 loader (reference model/datasets/rcc_dataset_pos_mimic.py:171-313):
``synthetic_batch`` emits the same 13-tuple ``rcc_collate`` returns, with the distributions of
SURVEY.md section 8(d).  Detectron2 feature extraction is out of scope, so ROI features are
``relu(N(0,1))`` (Faster-RCNN fc features are post-ReLU, reference
"feature extraction/ana_bbox_generator.py":476-487) and missing detections are all-zero nodes
(:590-594).  Spatial labels follow the reference's 12-way geometry bucketing (:266-302, :320-335).

Everything is generated with a CPU ``torch.Generator`` so the same seed gives the same batch on
every box; tensors are then moved to the requested device.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

_REVERSE = torch.tensor([0, 2, 1, 3, 8, 9, 10, 11, 4, 5, 6, 7])   # ana_bbox_generator.py:278-302


def spatial_labels_from_boxes(bb: torch.Tensor, lx: float = 1024.0, ly: float = 1024.0) -> torch.Tensor:
    """[B,N,4] boxes (xmin,ymin,xmax,ymax) -> [B,N,N] int64 labels 0..11.

    Restates bbox_relation_type / get_adj_matrix ("feature extraction/ana_bbox_generator.py":266-335):
    1 = i strictly contains j, 2 = j strictly contains i, 3 = IoU >= 0.5, 0 = centres further apart
    than (lx+ly)/3, else ceil(angle/45)+3; entry [j,i] (j>i) is reverse_type of [i,j]."""
    bb = bb.double()
    b1 = bb.unsqueeze(2)      # i
    b2 = bb.unsqueeze(1)      # j
    inside = (b1[..., 0] < b2[..., 0]) & (b1[..., 1] < b2[..., 1]) & (b1[..., 2] > b2[..., 2]) & (b1[..., 3] > b2[..., 3])
    cover = (b1[..., 0] > b2[..., 0]) & (b1[..., 1] > b2[..., 1]) & (b1[..., 2] < b2[..., 2]) & (b1[..., 3] < b2[..., 3])
    iw = (torch.minimum(b1[..., 2], b2[..., 2]) - torch.maximum(b1[..., 0], b2[..., 0]) + 1.0).clamp(min=0)
    ih = (torch.minimum(b1[..., 3], b2[..., 3]) - torch.maximum(b1[..., 1], b2[..., 1]) + 1.0).clamp(min=0)
    inter = iw * ih
    a1 = (b1[..., 2] - b1[..., 0] + 1.0) * (b1[..., 3] - b1[..., 1] + 1.0)
    a2 = (b2[..., 2] - b2[..., 0] + 1.0) * (b2[..., 3] - b2[..., 1] + 1.0)
    iou = inter / (a1 + a2 - inter)
    c1x, c1y = (b1[..., 0] + b1[..., 2]) / 2, (b1[..., 1] + b1[..., 3]) / 2
    c2x, c2y = (b2[..., 0] + b2[..., 2]) / 2, (b2[..., 1] + b2[..., 3]) / 2
    dist = torch.sqrt((c2x - c1x) ** 2 + (c2y - c1y) ** 2)
    ang = torch.atan2(c2y - c1y, c2x - c1x) / math.pi * 180.0
    ang = torch.where(ang < 0, ang + 360.0, ang)
    t = torch.ceil(ang / 45.0).long() + 3
    t = torch.where(dist >= (lx + ly) / 3.0, torch.zeros_like(t), t)
    t = torch.where(iou >= 0.5, torch.full_like(t, 3), t)
    t = torch.where(cover, torch.full_like(t, 2), t)
    t = torch.where(inside, torch.full_like(t, 1), t)
    n = bb.shape[1]
    upper = torch.triu(torch.ones(n, n, dtype=torch.bool))
    return torch.where(upper, t, _REVERSE[t.transpose(1, 2)])


def synthetic_batch(batch_size: int, num_nodes: int = 52, seed: int = 1234, device="cpu",
                    feat_dim: int = 1024, q_len: int = 20, seq_len: int = 90, ntoken: int = 147,
                    p_missing: float = 0.1, adj_size: int = 100) -> Tuple[torch.Tensor, ...]:
    """The 13-tuple of rcc_collate (rcc_dataset_pos_mimic.py:273,311-313):
    (d_feats, sc_feats, labels, sc_pos_labels, masks, pair_index, d_adj, q_adj, d_sem_adj,
     q_sem_adj, d_bb, q_bb, question).

    Adjacency matrices are integer labels stored as float64 [B,S,S] with S = max(adj_size, N)
    (the loader emits `.double()`, :178-181); boxes are float64 and, as in the reference step
    (train_mimic.py:206-218), are NOT moved to the device."""
    g = torch.Generator().manual_seed(seed)
    B, N = batch_size, num_nodes
    S = max(adj_size, N)

    def image():
        feats = torch.relu(torch.randn(B, N, feat_dim, generator=g))
        xy = torch.rand(B, N, 2, generator=g) * 800.0
        wh = torch.rand(B, N, 2, generator=g) * 200.0 + 8.0
        bb = torch.cat([xy, (xy + wh).clamp(max=1024.0)], -1).double()
        keep = (torch.rand(B, N, generator=g) >= p_missing)
        feats = feats * keep.unsqueeze(-1)
        bb = bb * keep.unsqueeze(-1)
        spa = torch.zeros(B, S, S, dtype=torch.float64)
        spa[:, :N, :N] = spatial_labels_from_boxes(bb).double()
        # semantic: {0: .90, 1: .06, 2: .04}, symmetric, whole rows empty with p ~ .3
        u = torch.rand(B, N, N, generator=g)
        lab = (u > 0.90).long() + (u > 0.96).long()
        lab = torch.triu(lab) + torch.triu(lab, 1).transpose(1, 2)
        alive = (torch.rand(B, N, generator=g) >= 0.3)
        lab = lab * alive.unsqueeze(2) * alive.unsqueeze(1)
        sem = torch.zeros(B, S, S, dtype=torch.float64)
        sem[:, :N, :N] = lab.double()
        return feats, bb, spa, sem

    d_feats, d_bb, d_adj, d_sem = image()
    q_feats, q_bb, q_adj, q_sem = image()
    qlen = torch.randint(4, 13, (B,), generator=g)
    question = torch.randint(1, ntoken + 1, (B, q_len), generator=g)
    question = question * (torch.arange(q_len).unsqueeze(0) < qlen.unsqueeze(1))
    alen = torch.randint(2, 31, (B,), generator=g)
    labels = torch.randint(2, ntoken + 1, (B, seq_len + 1), generator=g)
    pos = torch.arange(seq_len + 1).unsqueeze(0)
    labels = labels * ((pos >= 1) & (pos <= alen.unsqueeze(1)))
    labels[:, 0] = 1
    nz = (labels != 0).sum(1) + 1
    masks = (pos < nz.unsqueeze(1)).long()
    dev = torch.device(device)
    return (d_feats.to(dev), q_feats.to(dev), labels.unsqueeze(1).to(dev),
            torch.zeros(B, 1, seq_len + 1, dtype=torch.long, device=dev), masks.unsqueeze(1).to(dev),
            torch.arange(B), d_adj.to(dev), q_adj.to(dev), d_sem.to(dev), q_sem.to(dev),
            d_bb, q_bb, question.to(dev))


def synthetic_state_dict(spec: Dict[str, Tuple[int, ...]], seed: int = 1238,
                         dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic random weights for a {key: shape} spec, independent of module construction
    order (one generator per key), so the reference model, the oracle and the CUDA modules can all
    be loaded with bit-identical parameters from the same seed."""
    out = {}
    for i, key in enumerate(sorted(spec)):
        shape = tuple(spec[key])
        g = torch.Generator().manual_seed(seed * 1000003 + i)
        if key.endswith("weight_g"):
            t = None                                  # filled below from weight_v
        elif len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if "emb" in key and "rnn" not in key and "embed" not in key:
                t = torch.randn(shape, generator=g) * 0.5
            else:
                t = (torch.rand(shape, generator=g) * 2 - 1) * (1.0 / math.sqrt(fan_in))
        elif "layer_norm.weight" in key:
            t = torch.ones(shape)
        else:
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        out[key] = t
    for key in out:
        if key.endswith("weight_g"):
            v = out[key[:-1] + "v"]
            g = torch.Generator().manual_seed(seed * 7919 + len(key))
            out[key] = (v.norm() * (0.8 + 0.4 * torch.rand((), generator=g))).reshape(spec[key])
    return {k: v.to(dtype) for k, v in out.items()}
