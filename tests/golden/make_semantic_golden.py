"""Golden vectors for the semantic adjacency labels: runs the REFERENCE's own get_semantic_adj
("/root/reference/feature extraction/combine_dicts.py":106-151) on seeded detections and synthetic knowledge tables and
stores classes, tables and labels in tests/golden/semantic_labels.npz.  The module itself cannot be imported here (its
imports need the authors' data files), so the function definition is taken out of its syntax tree and executed
unmodified; nothing is copied into this repo.

    python tests/golden/make_semantic_golden.py      (in the build container, where /root/reference exists)
"""
import ast
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/feature extraction/combine_dicts.py"
T = 52                                      # test_topk_per_image: 26 anatomy + 26 disease detections


def reference_function():
    tree = ast.parse(open(SRC).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_semantic_adj"]
    assert len(body) == 1
    ns = {"np": np, "test_topk_per_image": T}
    exec(compile(ast.Module(body=body, type_ignores=[]), SRC, "exec"), ns)
    return ns["get_semantic_adj"]


def tables(seed=3):
    g = np.random.default_rng(seed)
    small = ["atelectasis", "cardiomegaly", "consolidation", "edema", "enlarged cardiomediastinum", "fracture",
             "lung lesion", "lung opacity", "no finding", "pleural effusion", "pleural other", "pneumonia",
             "pneumothorax", "support devices"]
    ana = ["anatomy %02d" % i for i in range(34)] + ["Cardiomegaly", "Lung Lesion"]      # two names live in both lists
    di = [s.title() for s in small[:11]] + ["disease %02d" % i for i in range(12)] + ["Cardiomegaly", "Lung Lesion"]
    organs = ["lung", "heart", "mediastinum", "pleura", "bone", "abdomen", "spine", "hilum"]
    kg = {n: organs[int(g.integers(0, len(organs)))] for n in set(ana + di)}
    small_adj = g.integers(0, 3, size=(14, 14))
    return ana, di, kg, small_adj, {n: i for i, n in enumerate(small)}


def main():
    fn = reference_function()
    ana, di, kg, small_adj, name2idx = tables()
    g = np.random.default_rng(17)
    n = 6
    cls_ana = g.integers(0, len(ana), size=(n, T // 2))
    cls_loc = g.integers(0, len(di) + 1, size=(n, T // 2))             # len(di) = background for the disease head
    labels, classes = [], []
    for k in range(n):
        loc = cls_loc[k].copy()
        lab = fn(cls_ana[k].copy(), loc, list(ana), list(di), kg, small_adj, name2idx)     # (offsets `loc` in place)
        labels.append(lab)
        classes.append(np.hstack((cls_ana[k], cls_loc[k] + len(ana))))
    out = {"classes": np.stack(classes).astype(np.int32), "labels": np.stack(labels).astype(np.int8),
           "small_adj": small_adj.astype(np.int32),
           "meta": json.dumps({"ana": ana, "di": di, "kg": kg, "name2idx": name2idx})}
    np.savez_compressed(os.path.join(HERE, "semantic_labels.npz"), **out)
    print(out["classes"].shape, out["labels"].shape, np.bincount(out["labels"].ravel().astype(np.int64), minlength=3))


if __name__ == "__main__":
    main()
