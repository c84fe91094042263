"""Per-kernel GPU time of the answer decoder (teacher-forced forward + backward, then greedy decoding), eager launches
bracketed by CUDA events (lib.PROFILE).  Analysis tooling.   python scripts/decoder_profile.py [batch]"""
import collections
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ekaid_b200 import lib  # noqa: E402
from ekaid_b200.config import default_cfg  # noqa: E402
from ekaid_b200.speaker import DynamicSpeaker  # noqa: E402
from ekaid_b200.synthetic import synthetic_batch, synthetic_state_dict  # noqa: E402

lib.require_device()
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = default_cfg("all")
with contextlib.redirect_stdout(io.StringIO()):
    sp = DynamicSpeaker(cfg, vocab_size=148)
sp.load_state_dict(synthetic_state_dict({k: tuple(v.shape) for k, v in sp.state_dict().items()}, 4321))
sp.to(dev).set_precision("bf16").train()
b = synthetic_batch(B, 52, seed=3)
labels, masks = b[2].squeeze(1).to(dev), b[4].squeeze(1).float().to(dev)
feats = [torch.randn(B, 1024, device=dev).requires_grad_(True) for _ in range(3)]


def run():
    loss = sp.masked_nll(feats[0], feats[1], feats[2], labels, masks)
    loss.backward()


def report(prof, title):
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, a, b_, info in prof:
        key = name
        if name == "gemm_tc" and info:
            key = "gemm %s" % (info["shape"],)
        agg[key][0] += 1
        agg[key][1] += a.elapsed_time(b_) * 1e3
    tot = sum(v[1] for v in agg.values())
    print("== %s: %d calls, %.0f us of bracketed GPU time" % (title, sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        print("  %8.1f us  %4d x %6.1f us   %s" % (v[1], v[0], v[1] / v[0], k))


for _ in range(2):
    run()
torch.cuda.synchronize()
lib.PROFILE = []
run()
torch.cuda.synchronize()
prof, lib.PROFILE = lib.PROFILE, None
report(prof, "teacher-forced fwd + bwd, batch %d, %d steps" % (B, sp._steps(labels)))
sp.eval()
with torch.no_grad():
    for _ in range(2):
        sp._sample(feats[0], feats[1], feats[2], None, cfg, sample_max=1, check_every=0, use_graph=False)
    torch.cuda.synchronize()
    lib.PROFILE = []
    sp._sample(feats[0], feats[1], feats[2], None, cfg, sample_max=1, check_every=0, use_graph=False)
    torch.cuda.synchronize()
    prof, lib.PROFILE = lib.PROFILE, None
report(prof, "greedy decode, 90 steps")
print("skinny launches so far:", lib.load().ekaid_gemm_skinny_count())
