"""Parity at the sizes BASELINE.json quotes (the golden cases are B <= 3): gradients of every live parameter at the
training batch (64 pairs x 52 nodes) and on the scaled graph (126 nodes, K = 126); forward at the inference batches
(256 and 512 pairs).  Same bars as tests/test_gpu_parity.py: 1e-4 (fp32 path) / 2e-2 (16-bit path) per output relative to
its own maximum, gradients 5e-4 / 5e-2.  The oracle (CPU, stock torch ops) takes seconds at these sizes."""
import pytest
import torch

from helpers import OUT_NAMES, case_inputs, oracle_forward, rel_err
from test_gpu_parity import TOL, build_model, check_gradients, to_dev, _dev

pytestmark = pytest.mark.gpu


def _meta(B, N, nongt, seed):
    return {"graph": "all", "nongt_dim": nongt, "empty_image": False, "B": B, "N": N, "seed": seed,
            "weight_seed": 1238, "zero_img_bias": False}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,N,nongt", [(64, 52, 52), (2, 126, 126)])
def test_gradients_match_oracle_at_size(B, N, nongt, precision):
    check_gradients(_meta(B, N, nongt, 7000 + B + N), precision, "b%d_n%d" % (B, N))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B", [256, 512])
def test_forward_matches_oracle_at_inference_batch(B, precision):
    """BASELINE config 2 (test_mimic path, batch 512 per GPU) and config 3's batch (256): the large-batch GRU path and
    the M = 2*B*52 GEMMs."""
    dev = _dev()
    meta = _meta(B, 52, 52, 8000 + B)
    sd, inp, _ = case_inputs(meta)
    m = build_model(meta, sd, precision, dev)
    with torch.no_grad():
        outs = m(*to_dev(inp, dev), setting="mode2", graph="all")
        ref = oracle_forward(sd, inp, meta)
    errs = {k: rel_err(o, r) for k, o, r in zip(OUT_NAMES, outs, ref)}
    print("B=%d" % B, precision, {k: "%.1e" % v for k, v in errs.items()})
    for k, e in errs.items():
        assert e < TOL[precision], (B, precision, k, e)
