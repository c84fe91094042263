"""Golden vectors for the spatial adjacency labels: runs the REFERENCE's own functions (get_iou, get_center,
get_distance, get_angle, cal_angle, bbox_relation_type, reverse_type, get_adj_matrix of
"/root/reference/feature extraction/ana_bbox_generator.py":213-335) on seeded boxes and stores boxes + labels in
tests/golden/spatial_labels.npz.  The module itself cannot be imported here (detectron2, cv2 missing), so the eight
function definitions are taken out of its syntax tree and executed unmodified; nothing is copied into this repo.

    python tests/golden/make_spatial_golden.py      (in the build container, where /root/reference exists)
"""
import ast
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
SRC = "/root/reference/feature extraction/ana_bbox_generator.py"
WANTED = ("get_iou", "get_center", "get_distance", "get_angle", "cal_angle", "bbox_relation_type", "reverse_type",
          "get_adj_matrix")


def reference_functions():
    tree = ast.parse(open(SRC).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    assert sorted(n.name for n in body) == sorted(WANTED)
    ns = {"np": np, "math": math, "tqdm": lambda x: x}
    exec(compile(ast.Module(body=body, type_ignores=[]), SRC, "exec"), ns)
    return ns


def main():
    from ekaid_b200.synthetic import synthetic_batch
    ref = reference_functions()
    b = synthetic_batch(3, 52, seed=5)
    boxes = [b[10].numpy(), b[11].numpy()]                      # loader boxes, with zeroed (missing) nodes
    g = np.random.default_rng(11)
    # hand-made cases on the decision boundaries: containment, identical boxes, axis-aligned and diagonal
    # neighbours (angles 0, 45, 90, ... exactly), far apart, touching IoU = 0.5
    edge = np.array([[0, 0, 100, 100], [10, 10, 50, 50], [0, 0, 100, 100], [200, 0, 300, 100], [0, 200, 100, 300],
                     [200, 200, 300, 300], [-200, 0, -100, 100], [0, -200, 100, -100], [-200, -200, -100, -100],
                     [200, -200, 300, -100], [-200, 200, -100, 300], [900, 900, 1000, 1000], [0, 0, 100, 49.5],
                     [0, 0, 0, 0], [0, 0, 0, 0], [0, 682, 100, 782.7], [0, 0, 1023, 1023], [50, 0, 150, 100]],
                    dtype=np.float64)[None]
    ints = np.sort(g.integers(0, 1024, size=(2, 30, 2, 2)), axis=2).transpose(0, 1, 3, 2).reshape(2, 30, 4)
    ints = np.stack([ints[..., 0], ints[..., 2], ints[..., 1], ints[..., 3]], -1).astype(np.float64)
    big = synthetic_batch(1, 126, seed=9)[10].numpy()
    out = {}
    for k, bb in enumerate(boxes + [edge, ints, big]):
        lab = ref["get_adj_matrix"]([list(map(list, x)) for x in bb.tolist()],
                                    np.zeros([bb.shape[0], max(100, bb.shape[1]), max(100, bb.shape[1])], int))
        out["boxes%d" % k] = bb
        out["labels%d" % k] = lab.astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "spatial_labels.npz"), **out)
    for k in range(5):
        print(k, out["boxes%d" % k].shape, np.bincount(out["labels%d" % k].ravel(), minlength=12))


if __name__ == "__main__":
    main()
