"""Kernel timeline of one CUDA-graph replay of the training step (torch.profiler / CUPTI), for analysis only:
writes gpurun_out/timeline.json = [[name, stream, start_us, dur_us], ...] sorted by start.  Never a bench number."""
import contextlib
import io
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from ekaid_b200 import lib  # noqa: E402
from ekaid_b200.config import WORD_TO_IDX, default_cfg  # noqa: E402
from ekaid_b200.modules import ChangeDetector  # noqa: E402
from ekaid_b200.step import GraphFusionStep, select_fields  # noqa: E402
from ekaid_b200.synthetic import synthetic_batch, synthetic_state_dict  # noqa: E402


def main():
    B, N = int(os.environ.get("TL_BATCH", 64)), 52
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib.require_device()
    cfg = default_cfg("all", nongt_dim=52)
    with contextlib.redirect_stdout(io.StringIO()):
        cd = ChangeDetector(cfg, WORD_TO_IDX)
    spec = {k: tuple(v.shape) for k, v in cd.state_dict().items()}
    cd.load_state_dict(synthetic_state_dict(spec, 1238))
    cd.to(dev).set_precision("bf16")
    cd.train()
    step = GraphFusionStep(cd, cfg, graph="all")
    res = [tuple(t.to(dev) for t in select_fields(synthetic_batch(B, N, seed=1234 + i))) for i in range(2)]
    step.capture(res[0], train=True)
    for i in range(5):
        step.replay(res[i % 2])
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(3):
            step.replay(res[i % 2])
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            tr = e.time_range
            evs.append([e.name[:90], int(getattr(e, "device_resource_id", getattr(e, "thread", 0)) or 0), tr.start, tr.end - tr.start])
    evs.sort(key=lambda x: x[2])
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(evs, open("gpurun_out/timeline.json", "w"))
    print("events", len(evs))


if __name__ == "__main__":
    main()
