"""Two eager training steps at a small batch, for `compute-sanitizer --tool memcheck python scripts/sanitizer_step.py 4`.
The slot check of FlatAdam is skipped and reported instead: under the sanitizer autograd does not adopt the gradient
views (it clones them); GraphFusionStep.train_step copies such gradients into their slots, this script reports them."""
import contextlib, io, sys, torch
sys.path.insert(0, ".")
from ekaid_b200 import lib
from ekaid_b200.config import WORD_TO_IDX, default_cfg
from ekaid_b200.modules import ChangeDetector
from ekaid_b200.step import GraphFusionStep, select_fields, expand_adjacency
from ekaid_b200.synthetic import synthetic_batch, synthetic_state_dict
for B in [int(a) for a in sys.argv[1:]] or [4, 64]:
    dev = torch.device("cuda", 0)
    cfg = default_cfg("all", nongt_dim=52)
    with contextlib.redirect_stdout(io.StringIO()):
        cd = ChangeDetector(cfg, WORD_TO_IDX)
    spec = {k: tuple(v.shape) for k, v in cd.state_dict().items()}
    cd.load_state_dict(synthetic_state_dict(spec, 1238))
    cd.to(dev).set_precision("bf16"); cd.train()
    step = GraphFusionStep(cd, cfg, graph="all")
    names = {p.data_ptr(): n for n, p in cd.named_parameters()}
    raw = tuple(t.to(dev) for t in select_fields(synthetic_batch(B, 52, seed=1234)))
    step.opt._checked = 99
    for it in range(2):
        step.train_step(expand_adjacency(raw, cfg), raw[9], raw[10].float())
        torch.cuda.synchronize()
        bad = [names.get(p.data_ptr()) for p, s in zip(step.opt.params, step.opt.slots) if p.grad is not None and p.grad.data_ptr() != s.data_ptr()]
        print("B", B, "iter", it, "not in slot:", bad)
