"""The CPU oracle (oracle/ekaid_oracle.py) against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import ekaid_oracle as O
from ekaid_b200.synthetic import synthetic_batch, synthetic_state_dict
from helpers import (CASES, GOLDEN, OUT_NAMES, case_inputs, check_fixture_inputs, grad_summary, load_case, loss_weights,
                     oracle_forward, rel_err, speaker_spec, checksum)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_outputs(name):
    z, meta = load_case(name)
    sd, inp, _ = case_inputs(meta)
    check_fixture_inputs(z, sd, inp)
    with torch.no_grad():
        outs = oracle_forward(sd, inp, meta)
    for k, o in zip(OUT_NAMES, outs):
        assert tuple(o.shape) == z[k].shape
        assert rel_err(o, z[k]) < 2e-5, (name, k, rel_err(o, z[k]))


@pytest.mark.parametrize("name", ["c1_b3_n52_all_grads", "c7_b2_n52_zero_bias"])
def test_oracle_matches_reference_gradients(name):
    z, meta = load_case(name)
    sd, inp, _ = case_inputs(meta)
    # w_emb.emb_ is frozen in the reference (language_model.py:28-29)
    sdg = {k: v.clone().requires_grad_(v.is_floating_point() and k != "w_emb.emb_.weight") for k, v in sd.items()}
    outs = oracle_forward(sdg, inp, meta)
    loss = sum((o * w).sum() for o, w in zip(outs[1:], loss_weights(outs)))
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-3 * max(1.0, abs(float(z["loss"])))
    loss.backward()
    names = [str(n) for n in z["grad_names"]]
    rows = z["grad_rows"]
    assert len(names) > 40
    for n, row in zip(names, rows):
        g = sdg[n].grad
        assert g is not None, n
        s = grad_summary(g)
        ref = row[:len(s)]
        scale = max(abs(ref[0]), 1e-12)          # gradient norm
        # atol: parameters whose gradient is analytically zero (key bias, implicit label bias, W2 bias: softmax
        # shift invariance) only carry rounding noise ~1e-5
        assert abs(s[0] - ref[0]) < 2e-4 * scale + 5e-5, (n, s[0], ref[0])
        assert np.abs(s[2:] - ref[2:]).max() < 5e-4 * scale + 5e-5, (n, np.abs(s[2:] - ref[2:]).max(), scale)
    # parameters the reference leaves without gradient (quirks Q2, Q3, Q11) get none from the oracle either
    dead = [k for k in sdg if sdg[k].requires_grad and k not in names]
    for k in dead:
        g = sdg[k].grad
        assert g is None or float(g.abs().max()) == 0.0 or k.startswith("fc1."), k


@pytest.mark.parametrize("name", ["c0_b2_n52_all", "c1_b3_n52_all_grads"])
def test_oracle_speaker_tokens(name):
    z, meta = load_case(name)
    ssd = synthetic_state_dict(speaker_spec(), 4321)
    bef, aft, diff = (torch.from_numpy(z[k]) for k in ("attended_1", "attended_2", "input_attended"))
    with torch.no_grad():
        seq = O.speaker_greedy(ssd, bef, aft, diff, 90, 512)
    assert np.array_equal(seq.numpy(), z["tokens"])
    batch = synthetic_batch(meta["B"], meta["N"], seed=meta["seed"])
    with torch.no_grad():
        logp = O.speaker_teacher_forced(ssd, bef, aft, diff, batch[2].squeeze(1), 90, 512)
    np.testing.assert_allclose(checksum(logp), z["tf_logp_checksum"], rtol=1e-4)


def test_process_matrix_fixture():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "process_matrix.npz"))
    b = synthetic_batch(2, 52, seed=3)
    pm = O.process_matrix(b[6], 52, 11)
    ps = O.process_matrix(b[8], 52, 3)
    assert np.array_equal(pm.sum((0, 3)).numpy(), z["spa_sum"])
    assert np.array_equal(ps.sum((0, 3)).numpy(), z["sem_sum"])
    np.testing.assert_allclose(checksum(pm), z["spa_chk"])
    np.testing.assert_allclose(checksum(ps), z["sem_chk"])


def test_spatial_labels_fixture():
    """Spatial adjacency labels from boxes: the oracle's scalar restatement and the loader's vectorised rule against
    labels produced by the reference's own get_adj_matrix (tests/golden/make_spatial_golden.py)."""
    from ekaid_b200.synthetic import spatial_labels_from_boxes
    z = np.load(os.path.join(GOLDEN, "spatial_labels.npz"))
    for k in range(5):
        bb = torch.from_numpy(z["boxes%d" % k])
        want = torch.from_numpy(z["labels%d" % k].astype(np.int64))
        n = bb.shape[1]
        assert want.shape[1] == max(100, n) and int(want[:, n:].abs().sum()) == 0 and int(want[:, :, n:].abs().sum()) == 0
        assert torch.equal(O.spatial_adj_matrix(bb), want), k
        assert torch.equal(spatial_labels_from_boxes(bb), want[:, :n, :n]), k
    assert O.spatial_adj_matrix(torch.zeros(0, 5, 4)).shape == (0, 100, 100)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_spatial_labels_scalar_vs_vectorised_on_integer_boxes(seed):
    """Integer box coordinates put many pairs exactly on the decision boundaries (shared edges, equal centres, exact
    45-degree diagonals): the scalar restatement of get_adj_matrix and the loader's vectorised rule must agree there
    too, and the reverse_type / self-edge properties hold."""
    from ekaid_b200.synthetic import _REVERSE, spatial_labels_from_boxes
    g = torch.Generator().manual_seed(seed)
    lo = torch.randint(0, 40, (2, 24, 2), generator=g) * 16
    wh = torch.randint(0, 12, (2, 24, 2), generator=g) * 16
    bb = torch.cat([lo, lo + wh], 2).double()
    bb[0, 3] = bb[0, 2]                                   # identical boxes
    bb[1, 5] = 0                                          # a missing detection (all-zero box)
    want = O.spatial_adj_matrix(bb, size=30)
    assert want.shape == (2, 30, 30)
    lab = spatial_labels_from_boxes(bb)
    assert torch.equal(lab, want[:, :24, :24])
    assert bool((lab.diagonal(dim1=1, dim2=2) == 3).all())
    upper = torch.triu(torch.ones(24, 24, dtype=torch.bool), 1)
    assert torch.equal(lab.transpose(1, 2)[:, upper], _REVERSE[lab[:, upper]])
    assert len(torch.unique(lab)) >= 8                    # the draw exercises most label values


def test_semantic_adjacency_oracle_matches_reference_labels():
    """oracle.semantic_adj_matrix (get_semantic_adj, "feature extraction/combine_dicts.py":106-151) against labels made by
    the reference's own function (tests/golden/make_semantic_golden.py)."""
    import json
    import os
    import numpy as np
    import torch
    from helpers import GOLDEN
    from oracle import ekaid_oracle as O
    z = np.load(os.path.join(GOLDEN, "semantic_labels.npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    got = O.semantic_adj_matrix(z["classes"], meta["ana"], meta["di"], meta["kg"], z["small_adj"], meta["name2idx"])
    assert got.shape == (6, 100, 100)
    assert torch.equal(got, torch.from_numpy(z["labels"].astype(np.int64)))
    assert int(got.max()) == 2 and torch.equal(got, got.transpose(1, 2))
