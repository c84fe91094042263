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
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    lib.require_device()
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    cfg = default_cfg("all", nongt_dim=52)
    with contextlib.redirect_stdout(io.StringIO()):
        cd = ChangeDetector(cfg, WORD_TO_IDX)
    spec = {k: tuple(v.shape) for k, v in cd.state_dict().items()}
    cd.load_state_dict(synthetic_state_dict(spec, 1238))
    cd.to(dev).set_precision("bf16")
    cd.train()
    step = GraphFusionStep(cd, cfg, graph="all", process_group=pg)
    res = [tuple(t.to(dev) for t in select_fields(synthetic_batch(B, N, seed=1234 + 17 * rank + i))) for i in range(2)]
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
    if rank == 0:
        tag = os.environ.get("TL_TAG", "%dgpu" % world)
        json.dump(evs, open("gpurun_out/timeline_%s.json" % tag, "w"))
        # summary of the LAST replay: span, busy time (union of kernel intervals), per-kernel totals, NCCL intervals
        starts = [e[2] for e in evs]
        gaps = sorted(((starts[i + 1] - (evs[i][2] + evs[i][3]), i) for i in range(len(evs) - 1)), reverse=True)[:2]
        cut = max(i for _, i in gaps) + 1          # the two largest gaps separate the three replays
        last = evs[cut:]
        t0, t1 = last[0][2], max(e[2] + e[3] for e in last)
        busy, cur_s, cur_e = 0.0, None, None
        for e in last:
            s_, e_ = e[2], e[2] + e[3]
            if cur_s is None:
                cur_s, cur_e = s_, e_
            elif s_ <= cur_e:
                cur_e = max(cur_e, e_)
            else:
                busy += cur_e - cur_s
                cur_s, cur_e = s_, e_
        busy += cur_e - cur_s
        tot = {}
        for e in last:
            k = e[0].split("<")[0].split("(")[0][:60]
            d = tot.setdefault(k, [0.0, 0])
            d[0] += e[3]
            d[1] += 1
        lines = ["replay span %.1f us, some kernel running %.1f us (%.1f%%), %d kernels" % (t1 - t0, busy, 100 * busy / (t1 - t0), len(last))]
        for k, (d, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:28]:
            lines.append("%9.1f us %4d x  %s" % (d, n, k))
        for e in last:
            if "nccl" in e[0].lower():
                lines.append("NCCL kernel at +%.1f us for %.1f us: %s" % (e[2] - t0, e[3], e[0][:60]))
        open("gpurun_out/timeline_%s_summary.txt" % tag, "w").write("\n".join(lines) + "\n")
        print("\n".join(lines))
    if world > 1:
        step._graph = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
