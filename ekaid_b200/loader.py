"""Feature loader with the reference's on-disk schema.

The reference reads two HDF5 files ("feature extraction/combine_dicts.py":162-187 writes the feature file,
dataset_preparation.py:27-54 the label file) through `RCCDataset` (model/datasets/rcc_dataset_pos_mimic.py:24-273) and
batches with `rcc_collate` (:284-309).  This module keeps that contract -- dataset names, shapes, the 13-tuple, the mask
rule -- on plain arrays:

    feature file: image_features [n_img, 52, 1024] f32 | image_bb [n_img, 52, 4] f32 | image_adj_matrix [n_img, 100, 100] i64
                  | semantic_adj_matrix [n_img, 100, 100] i64 | bbox_label [n_img, 52] i64
    label file:   questions [n, 20] | answers [n, 90] (`labels`) | pos [n, 90] | feature_idx [n, 2]
                  | label_start_idx [n] | label_end_idx [n]

`RCCArrays.from_hdf5` opens the reference's files when h5py is installed (it is not in this image: the call then raises
ImportError, nothing is emulated); `RCCArrays.synthetic` builds arrays of the same schema from ekaid_b200.synthetic, and
`from_npz` / `save_npz` move them through numpy archives.  `RCCDataset.__getitem__` and `rcc_collate` restate the
reference's sample assembly; `collate(..., compact=True)` additionally narrows the four label matrices to int8 (what the
relation encoders consume directly, see step.select_fields).
"""
from __future__ import annotations

import random
from typing import Dict, Iterator, Optional, Sequence, Tuple

import numpy as np
import torch

FEATURE_KEYS = ("image_features", "image_bb", "image_adj_matrix", "semantic_adj_matrix", "bbox_label")
LABEL_KEYS = ("questions", "answers", "pos", "feature_idx", "label_start_idx", "label_end_idx")


class RCCArrays:
    """The two files' datasets as arrays (numpy, or h5py datasets: both index the same way)."""

    def __init__(self, arrays: Dict[str, object]):
        missing = [k for k in FEATURE_KEYS[:4] + LABEL_KEYS if k not in arrays]
        if missing:
            raise KeyError("missing datasets: %s" % ", ".join(missing))
        self.a = arrays
        f = arrays["image_features"]
        if len(f.shape) != 3 or arrays["image_bb"].shape[:2] != f.shape[:2] or arrays["image_bb"].shape[2] != 4:
            raise ValueError("image_features must be [n_img, nodes, dim] and image_bb [n_img, nodes, 4]")
        s = arrays["image_adj_matrix"].shape
        if len(s) != 3 or s[1] != s[2] or s[1] < f.shape[1] or arrays["semantic_adj_matrix"].shape != s:
            raise ValueError("adjacency datasets must be [n_img, S, S] with S >= nodes")

    def __getitem__(self, k):
        return self.a[k]

    @property
    def num_pairs(self) -> int:
        return int(self.a["feature_idx"].shape[0])

    @classmethod
    def from_hdf5(cls, feature_path: str, label_path: str) -> "RCCArrays":
        """The reference's own files (rcc_dataset_pos_mimic.py:60-110)."""
        try:
            import h5py
        except ImportError as e:      # pragma: no cover - h5py is not part of this image
            raise ImportError("reading the reference's HDF5 files needs h5py; convert them with save_npz on a machine that "
                              "has it, or install h5py") from e
        hf, hl = h5py.File(feature_path, "r"), h5py.File(label_path, "r")
        arrays = {k: hf[k] for k in FEATURE_KEYS if k in hf}
        arrays.update({k: hl[k] for k in LABEL_KEYS if k in hl})
        if "labels" in hl and "answers" not in arrays:
            arrays["answers"] = hl["labels"]
        return cls(arrays)

    @classmethod
    def from_npz(cls, path: str) -> "RCCArrays":
        z = np.load(path, allow_pickle=False)
        return cls({k: z[k] for k in z.files})

    def save_npz(self, path: str) -> None:
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in self.a.items()})

    @classmethod
    def synthetic(cls, n_pairs: int, nodes: int = 52, seed: int = 1234, adj_size: int = 100) -> "RCCArrays":
        """Arrays with the reference's schema filled from ekaid_b200.synthetic (two images per pair)."""
        from .synthetic import synthetic_batch
        b = synthetic_batch(n_pairs, nodes, seed=seed, adj_size=adj_size)
        cat = lambda x, y: torch.cat([x, y], 0).numpy()        # noqa: E731  image i = main of pair i, n + i = its reference
        labels = b[2].squeeze(1).numpy()
        n = n_pairs
        return cls({
            "image_features": cat(b[0], b[1]).astype(np.float32),
            "image_bb": cat(b[10], b[11]).astype(np.float32),
            "image_adj_matrix": cat(b[6], b[7]).astype(np.int64),
            "semantic_adj_matrix": cat(b[8], b[9]).astype(np.int64),
            "bbox_label": np.zeros((2 * n, nodes), dtype=np.int64),
            "questions": b[12].numpy().astype(np.int64),
            "answers": labels[:, :90].astype(np.int64),
            "pos": np.zeros((n, 90), dtype=np.int64),
            "feature_idx": np.stack([np.arange(n), n + np.arange(n)], 1).astype(np.int64),
            "label_start_idx": np.arange(n, dtype=np.int64),
            "label_end_idx": np.arange(n, dtype=np.int64),
        })


class RCCDataset:
    """rcc_dataset_pos_mimic.py:24-273 for feature_mode 'both' (the configuration the hot path runs): one sample =
    (d_feature, q_feature, seq [1, L+1], pos [1, L+1], mask [1, L+1], pair index, d_adj, q_adj, d_sem_adj, q_sem_adj,
    d_bb, q_bb, question)."""

    def __init__(self, arrays: RCCArrays, split_idxs: Optional[Sequence[int]] = None, seq_per_img: int = 1,
                 max_seq_length: int = 90, feature_mode: str = "both"):
        if feature_mode != "both":
            raise NotImplementedError("feature_mode %r: the graph + fusion path runs on 'both' (anatomy + disease nodes)"
                                      % (feature_mode,))
        self.arr = arrays
        self.split_idxs = list(range(arrays.num_pairs)) if split_idxs is None else list(split_idxs)
        self.seq_per_img = seq_per_img
        self.max_seq_length = max_seq_length

    def __len__(self) -> int:
        return len(self.split_idxs)

    def __getitem__(self, index: int):
        random.seed(1111)                                       # :172 (the reference re-seeds on every access)
        a = self.arr
        img_idx = self.split_idxs[index]
        i0, i1 = int(a["feature_idx"][img_idx, 0]), int(a["feature_idx"][img_idx, 1])
        d_feature, q_feature = np.asarray(a["image_features"][i0]), np.asarray(a["image_features"][i1])
        dbl = lambda x: torch.from_numpy(np.asarray(x)).double()      # noqa: E731   (:179-184)
        d_adj, q_adj = dbl(a["image_adj_matrix"][i0]), dbl(a["image_adj_matrix"][i1])
        d_sem, q_sem = dbl(a["semantic_adj_matrix"][i0]), dbl(a["semantic_adj_matrix"][i1])
        d_bb, q_bb = dbl(a["image_bb"][i0]), dbl(a["image_bb"][i1])
        ix1, ix2 = int(a["label_start_idx"][img_idx]), int(a["label_end_idx"][img_idx])
        n_cap = ix2 - ix1 + 1
        L = self.max_seq_length
        seq = np.zeros([self.seq_per_img, L + 1], dtype=int)
        pos = np.zeros([self.seq_per_img, L + 1], dtype=int)
        if n_cap < self.seq_per_img:
            for q in range(self.seq_per_img):
                ixl = random.randint(ix1, ix2)
                seq[q, :L] = a["answers"][ixl, :L]
                pos[q, :L] = a["pos"][ixl, :L]
        else:
            ixl = random.randint(ix1, ix2 - self.seq_per_img + 1)
            seq[:, :L] = a["answers"][ixl: ixl + self.seq_per_img, :L]
            pos[:, :L] = a["pos"][ixl: ixl + self.seq_per_img, :L]
        mask = np.zeros_like(seq)                               # :258-263: ones over the tokens and one position more
        for ix, row in enumerate(mask):
            row[:int((seq[ix] != 0).sum()) + 1] = 1
        question = np.asarray(a["questions"][ixl])
        return (d_feature, q_feature, seq, pos, mask, img_idx, d_adj, q_adj, d_sem, q_sem, d_bb, q_bb, question)


def rcc_collate(batch, compact: bool = False, pin: bool = False) -> Tuple[torch.Tensor, ...]:
    """rcc_dataset_pos_mimic.py:284-309: stack every field.  compact: label matrices as int8 (lossless: labels 0..11)."""
    cols = list(zip(*batch))
    st = lambda xs: torch.stack([torch.as_tensor(np.asarray(x)) if not torch.is_tensor(x) else x for x in xs], 0)   # noqa: E731
    out = [st(cols[0]).float(), st(cols[1]).float(), st(cols[2]).long(), st(cols[3]).long(), st(cols[4]).long(),
           torch.as_tensor(cols[5]), st(cols[6]), st(cols[7]), st(cols[8]), st(cols[9]), st(cols[10]), st(cols[11]),
           st(cols[12]).long()]
    if compact:
        for i in (6, 7, 8, 9):
            out[i] = out[i].to(torch.int8)
    if pin and torch.cuda.is_available():
        out = [t.pin_memory() for t in out]
    return tuple(out)


def batches(dataset: RCCDataset, batch_size: int, shuffle: bool = False, seed: int = 0, compact: bool = False,
            pin: bool = False, drop_last: bool = True) -> Iterator[Tuple[torch.Tensor, ...]]:
    """Minimal RCCDataLoader (:312-316): 13-tuples of `batch_size` samples."""
    order = list(range(len(dataset)))
    if shuffle:
        random.Random(seed).shuffle(order)
    for lo in range(0, len(order), batch_size):
        idx = order[lo:lo + batch_size]
        if len(idx) < batch_size and drop_last:
            return
        yield rcc_collate([dataset[i] for i in idx], compact=compact, pin=pin)
