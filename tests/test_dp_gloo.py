"""world_size-2 data-parallel host logic on CPU (gloo): batch sharding, the flat live-parameter gradient buffer and
its single all-reduce.  Contract: rank r == reference on its LOCAL batch, the
exchanged gradient == mean over ranks of the per-rank oracle gradients."""
import contextlib
import io
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import spec_for


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_grads(rank, world, sd):
    """oracle gradients of the live parameters on this rank's shard of a global batch of 2."""
    from ekaid_b200.step import shard_batch
    from ekaid_b200.synthetic import synthetic_batch
    from oracle import ekaid_oracle as O
    batch = shard_batch(synthetic_batch(2, 52, seed=55), rank, world)
    inp = (batch[0], batch[1], O.process_matrix(batch[6], 52, 11), O.process_matrix(batch[7], 52, 11),
           O.process_matrix(batch[8], 52, 3), O.process_matrix(batch[9], 52, 3), batch[10], batch[11], batch[12])
    sdg = {k: v.clone().requires_grad_(v.is_floating_point() and k != "w_emb.emb_.weight") for k, v in sd.items()}
    outs = O.change_detector_forward(sdg, *inp)
    (outs[3].sum() + outs[4].sum() + outs[1].sum()).backward()
    return {k: v.grad for k, v in sdg.items() if v.grad is not None}


def _worker(rank, world, port, outfile):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ekaid_b200.config import WORD_TO_IDX, default_cfg
    from ekaid_b200.modules import ChangeDetector
    from ekaid_b200.step import FlatAdam, allreduce_mean_
    from ekaid_b200.synthetic import synthetic_state_dict
    with contextlib.redirect_stdout(io.StringIO()):
        cd = ChangeDetector(default_cfg(), WORD_TO_IDX)
    sd = synthetic_state_dict(spec_for("all"), 1238)
    cd.load_state_dict(sd)
    live = cd.live_parameters()
    opt = FlatAdam(live)
    names = {id(p): k for k, p in cd.named_parameters()}
    grads = _rank_grads(rank, world, sd)
    opt.zero_grad()
    from ekaid_b200 import functions
    assert len(functions.GRAD_SLOTS) == len(live)
    opt.grad.zero_()
    for p, slot in zip(opt.params, opt.slots):
        k = names[id(p)]
        assert slot.data_ptr() >= opt.grad.data_ptr() and p.data_ptr() % 256 == opt.flat.data_ptr() % 256
        assert functions.GRAD_SLOTS[p.data_ptr()][0] is slot and slot.shape == p.shape
        if k in grads:
            slot.copy_(grads[k])
    allreduce_mean_(opt.grad)
    out = {names[id(p)]: slot.clone() for p, slot in zip(opt.params, opt.slots) if names[id(p)] in grads}
    stats = {"n_live": len(live), "numel": int(sum(p.numel() for p in live)),
             "dead_with_grad": [k for k in grads if k not in out]}
    if rank == 0:
        torch.save((out, stats), outfile)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gradient_mean_over_local_batches(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    outfile = str(tmp_path / "rank0.pt")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, outfile)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=500)
        assert p.exitcode == 0
    got, stats = torch.load(outfile, weights_only=False)
    from ekaid_b200.synthetic import synthetic_state_dict
    sd = synthetic_state_dict(spec_for("all"), 1238)
    g0, g1 = _rank_grads(0, world, sd), _rank_grads(1, world, sd)
    assert stats["dead_with_grad"] == []            # every parameter the oracle gives a gradient to is "live"
    assert 30e6 < stats["numel"] < 40e6             # 36.9 M live of 68.4 M parameters (Q2, Q3, Q11)
    for k, g in got.items():
        ref = (g0[k] + g1[k]) / 2
        # (worker processes run the oracle with a different thread count -> fp32 summation order differs)
        assert float((g - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-4, k   # atol: analytically-zero gradients are rounding noise


def test_shard_batch_is_contiguous_partition():
    from ekaid_b200.step import shard_batch
    from ekaid_b200.synthetic import synthetic_batch
    b = synthetic_batch(4, 52, seed=2)
    parts = [shard_batch(b, r, 2) for r in range(2)]
    for i, t in enumerate(b):
        assert torch.equal(torch.cat([parts[0][i], parts[1][i]]), t)
