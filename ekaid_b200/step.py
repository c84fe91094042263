"""Training / inference step glue for the graph+fusion path (reference model/train_mimic.py:203-269 and
model/test_mimic.py:92-132), as a reusable object instead of script-level code.

    unpack the 13-tuple -> H2D -> process_matrix x4 (ONE kernel each instead of 14 host-syncing label loops)
    -> ChangeDetector -> loss terms that touch the path -> backward -> (NCCL gradient all-reduce) -> Adam

The answer decoder (DynamicSpeaker) is the boundary consumer (SURVEY.md section 8(f) "next"): a caller-supplied
`decoder_loss(bef, aft, diff, labels, masks)` closes the loop exactly like `speaker._forward` + LanguageModelCriterion
do in the reference; without one, a fixed cotangent stands in for the decoder gradient (bench / tests).
"""
from __future__ import annotations

import os
import weakref
from typing import Callable, Optional, Sequence

import torch

from . import functions, lib
from .functions import onehot_adj
from .modules import ChangeDetector


class FlatAdam:
    """torch.optim.Adam semantics (reference utils/utils.py:96-99: lr 1e-4, betas (0.9, 0.999), eps 1e-8, wd 0) on
    ONE flat fp32 buffer: parameters and gradients of the module are re-pointed into two contiguous buffers, so the
    optimizer is a single kernel launch and the data-parallel gradient exchange a single NCCL all-reduce."""

    def __init__(self, params: Sequence[torch.nn.Parameter], lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [p for p in params]
        dev = params[0].device
        al = 64                                   # every parameter starts on a 256-byte boundary (vector loads, TMA)
        n = sum((p.numel() + al - 1) // al * al for p in params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.pow_state = torch.ones(2, dtype=torch.float32, device=dev)
        off = 0
        self.slots = []
        self.offsets = [0]            # start of every parameter in the flat buffers (+ the total at the end)
        for p in params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            slot = self.grad[off:off + k].view(p.shape)
            self.slots.append(slot)
            p.grad = None
            # the backward kernels write this parameter's gradient straight into its slot (functions.GRAD_SLOTS)
            functions.GRAD_SLOTS[p.data_ptr()] = (slot, weakref.ref(p))
            off += (k + al - 1) // al * al
            self.offsets.append(off)
        self.params = params
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self._checked = 0

    def zero_grad(self):
        """set_to_none semantics: every live slot is fully overwritten by the next backward, so no memset is needed and
        autograd adopts the slot tensors instead of launching one `grad += new` kernel per parameter."""
        for p in self.params:
            p.grad = None
        functions.slots_reset()

    def close(self):
        """Unregister the gradient slots (the parameters keep their storage in the flat buffer)."""
        for p in self.params:
            functions.GRAD_SLOTS.pop(p.data_ptr(), None)

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001  (interpreter shutdown)
            pass

    def sync_slots(self, first: int = 0, last: Optional[int] = None, dry_run: bool = False) -> int:
        """Gradients that autograd did NOT adopt as their slot view (it clones a gradient whenever somebody else still
        references it -- seen e.g. under compute-sanitizer) are copied into the flat buffer so that the update and the
        all-reduce see them.  Returns how many needed that (dry_run: only counts)."""
        n = 0
        for p, slot in list(zip(self.params, self.slots))[first:last]:
            if p.grad is not None and p.grad.data_ptr() != slot.data_ptr():
                n += 1
                if not dry_run:
                    slot.copy_(p.grad)
        return n

    def check_slots(self):
        """After a backward: every gradient autograd holds must BE its slot (otherwise Adam would miss it)."""
        for p, slot in zip(self.params, self.slots):
            if p.grad is not None and p.grad.data_ptr() != slot.data_ptr():
                raise RuntimeError("gradient of a parameter of shape %s did not land in its flat slot" % (tuple(p.shape),))

    def advance(self):
        """beta^t bookkeeping of the step about to be applied (once per step, before the first update_range)."""
        lib.call("adam_advance", self.pow_state.data_ptr(), self.betas[0], self.betas[1])

    def update_range(self, lo: int, hi: int, max_ctas: int = 0):
        """Adam update of flat elements [lo, hi) (a whole number of parameters)."""
        if hi <= lo:
            return
        f, g, m, v = (t[lo:hi] for t in (self.flat, self.grad, self.m, self.v))
        lib.call("adam_step", f.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), hi - lo, self.lr, self.betas[0],
                 self.betas[1], self.eps, self.wd, self.pow_state.data_ptr(), max_ctas)

    def step(self):
        self.advance()
        self.update_range(0, self.flat.numel())


def allreduce_mean_(flat: torch.Tensor, group=None, wire_bf16: bool = False) -> torch.Tensor:
    """Data-parallel gradient exchange: ONE all-reduce over a slice of the flat gradient buffer, mean over ranks
    (NCCL: ReduceOp.AVG over NVLink; gloo, used by the CPU tests, has no AVG -> SUM then scale).
    wire_bf16: the slice crosses NVLink as bf16 (half the bytes; own cast kernels on both sides) -- the 16-bit path's
    gradients carry bf16 operand noise anyway; the fp32 path always exchanges fp32."""
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        n = flat.numel()
        world = dist.get_world_size(group)
        if wire_bf16 and flat.is_cuda and n % 64 == 0 and flat.data_ptr() % 16 == 0:
            # SUM, not AVG: NCCL implements AVG as a pre-multiplied sum, which rules out the in-switch (NVLS) reduction;
            # the 1 / world factor is applied by the cast back to fp32
            wire = torch.empty(n, dtype=torch.bfloat16, device=flat.device)
            lib.call("cast_f32_bf16", flat.data_ptr(), 64, wire.data_ptr(), 64, n // 64, 64)
            dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=group)
            lib.call("cast_bf16_f32_scaled", wire.data_ptr(), 64, flat.data_ptr(), 64, n // 64, 64, 1.0 / world)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(dist.get_world_size(group))
    return flat


def shard_batch(batch, rank: int, world: int):
    """Contiguous shard of the loader's 13-tuple for one rank: no data-path collective."""
    b = batch[0].shape[0]
    per = b // world
    return tuple(t[rank * per:(rank + 1) * per] for t in batch)


def select_fields(batch, compact: bool = True):
    """The 11 tensors of the loader's 13-tuple that the step consumes (train_mimic.py:206-218), in step order:
    (d_feats, sc_feats, d_adj, q_adj, d_sem, q_sem, d_bb, q_bb, question, labels, masks); adjacency still as integer
    label matrices [B,S,S].  compact: the four label matrices (values 0..11, stored as float64 by the reference's
    collate) are narrowed to int8 on the host -- the collate-side half of SURVEY.md section 8f row 2: 20 MB -> 2.5 MB
    of the 48 MB a batch of 64 moves over PCIe -- and the relation encoders read them as they are."""
    (d_feats, sc_feats, labels, sc_pos_labels, masks, pair_index, d_adj, q_adj, d_sem, q_sem, d_bb, q_bb,
     question) = batch
    if compact:
        d_adj, q_adj, d_sem, q_sem = (t if t.dtype == torch.int8 else t.to(torch.int8)
                                      for t in (d_adj, q_adj, d_sem, q_sem))
    return (d_feats, sc_feats, d_adj, q_adj, d_sem, q_sem, d_bb, q_bb, question, labels.squeeze(1), masks.squeeze(1))


def to_device(batch, device, compact: bool = True):
    """Host 13-tuple -> device tensors (pinned host memory makes the copies asynchronous)."""
    return tuple(t.to(device, non_blocking=True) for t in select_fields(batch, compact))


def expand_adjacency(raw, cfg, onehot: bool = None):
    """train_mimic.py:223-227.  -> the 9 ChangeDetector inputs.  int8 label matrices (select_fields(compact=True)) are
    handed to ChangeDetector as they are: its relation encoders take the label of an edge straight from the matrix
    (`adj_labels_fwd/bwd`), so the fp32 one-hot tensors of process_matrix never exist.  onehot=True (default for any
    other dtype) runs process_matrix x4, one kernel per matrix, and yields the reference's [B,N,N,L] tensors."""
    cd = cfg.model.change_detector
    n = raw[0].shape[1]
    if onehot is None:
        onehot = raw[2].dtype != torch.int8
    if not onehot:
        return tuple(raw[:9])
    return (raw[0], raw[1], onehot_adj(raw[2], n, cd.spa_label_num), onehot_adj(raw[3], n, cd.spa_label_num),
            onehot_adj(raw[4], n, cd.sem_label_num), onehot_adj(raw[5], n, cd.sem_label_num), raw[6], raw[7], raw[8])


def process_batch(batch, cfg, device, compact: bool = True):
    """to_device + expand_adjacency.  Returns (inputs, labels, masks)."""
    raw = to_device(batch, device, compact)
    return expand_adjacency(raw, cfg), raw[9], raw[10].float()


class GraphFusionStep:
    """One data-parallel rank of the reference's train/test step for the graph+fusion path."""

    def __init__(self, change_detector: ChangeDetector, cfg, graph: str = "all", lr: float = 1e-4,
                 decoder_loss: Optional[Callable] = None, process_group=None, speaker=None,
                 decoder_steps: Optional[int] = None):
        """speaker: an ekaid_b200.speaker.DynamicSpeaker closes the loop like train_mimic.py:230-248 -- its fused masked
        NLL is the decoder loss and its parameters join the optimizer (the reference's Adam covers both modules,
        train_mimic.py:163-165).  decoder_steps: fixed number of decoder steps (needed to capture the step in a CUDA
        graph; None = stop at the first all-empty label column like the reference's loop, one host read per step)."""
        self.cd = change_detector
        self.cfg = cfg
        self.graph = graph
        self.speaker = speaker
        self.decoder_steps = decoder_steps
        if speaker is not None and decoder_loss is None:
            def decoder_loss(bef, aft, diff, labels, masks):
                return speaker.masked_nll(bef, aft, diff, labels, masks, steps=self.decoder_steps)
        self.decoder_loss = decoder_loss
        self.pg = process_group
        # Flat parameter order = the order in which backward FINISHES the gradients, in contiguous segments:
        #   0 fusion stage | 1 implicit encoder | 2 spatial encoder | 3 semantic encoder   (every parameter of the stage,
        #     incl. its weight-normalised v / g pairs: each encoder's weight-norm backward runs right after the encoder's)
        #   4 the ROI projection `img` (its gradient is the last one of the image path)
        #   5 question path (embedding, GRU, question attention): BPTT is the serial tail of the step
        # As soon as a segment is complete (autograd hooks count its gradients) its all-reduce (N > 1) and Adam update
        # go out on the optimizer stream and overlap the rest of backward; segment 4 goes out when BPTT is launched
        # and runs next to the recurrence; only segment 5 is left for the end.
        named = change_detector.live_named_parameters()
        if speaker is not None:
            # the decoder's backward is the first thing backward does: its gradients complete with segment 0
            named = named + [("speaker." + n, p) for n, p in speaker.named_parameters()]

        def segment_of(name: str) -> int:
            if name.startswith("speaker."):
                return 0
            if name.startswith(("w_emb.", "q_emb.", "q_att.")):
                return 5
            if name.startswith(("context1.", "context2.", "gate1.", "gate2.", "embed.", "att.", "fc1.")):
                return 0
            if name.startswith("imp_relation."):
                return 1
            if name.startswith("spatial_relation."):
                return 2
            if name.startswith("semantic_relation."):
                return 3
            return 4                      # img: its gradient is the last one of the image path

        seg = [segment_of(n) for n, _ in named]
        order = sorted(range(len(named)), key=lambda i: (seg[i], i))
        params = [named[i][1] for i in order]
        self.opt = FlatAdam(params, lr=lr)
        self._seg_of_param = [seg[i] for i in order]
        self._nseg = 6
        # [first, last) parameter index and [lo, hi) flat offsets of every segment
        self._seg_pidx, self._seg_range = [], []
        for k in range(self._nseg):
            idx = [j for j, sk in enumerate(self._seg_of_param) if sk == k]
            first, last = (idx[0], idx[-1] + 1) if idx else (0, 0)
            self._seg_pidx.append((first, last))
            self._seg_range.append((self.opt.offsets[first], self.opt.offsets[last]) if idx else (0, 0))
        self._armed = False
        self._fired = [0] * self._nseg
        self._expected = None      # per segment: how many parameters receive a gradient (learned on the first step:
        self._done = [False] * self._nseg      # e.g. fc1 never does, its output is not in the reference's loss -- Q11)
        self._advanced = False
        self._opt_stream = None
        self._hooked = bool(params and params[0].is_cuda and os.environ.get("EKAID_B200_EARLY_ADAM", "1") != "0")
        # EKAID_B200_EARLY_SEGMENTS=1: segments 0-3 go out right when they complete instead of at the BPTT launch.
        # Measured (B200, 1 and 2 GPUs over NVLink): no gain -- 4.19 vs 4.15 ms and 4.52 vs 4.48 ms; the all-reduce is
        # cheap and the extra Adam launches compete with the GEMMs -- so it is off by default.
        self._early_segments = os.environ.get("EKAID_B200_EARLY_SEGMENTS", "0") == "1"
        # Data parallel: the gradient all-reduce of a segment goes out the moment the segment is complete (NCCL kernels
        # on the optimizer stream, next to the rest of backward); the Adam update of segments 0-4 still waits for the
        # BPTT launch.  EKAID_B200_EARLY_REDUCE=0: one all-reduce at the BPTT launch (the round-1 behaviour).
        er = os.environ.get("EKAID_B200_EARLY_REDUCE", "auto")
        if er == "auto":
            # measured (profiles/r02_notes.md): with 2 ranks the early exchanges hide behind backward (3.95 vs 4.15 ms);
            # with 8 the ranks drift apart during backward and every early ring stalls the faster ones (4.34 vs 4.12 ms)
            import torch.distributed as dist
            self._early_reduce = process_group is not None and dist.get_world_size(process_group) <= 2
        else:
            self._early_reduce = er != "0"
        self._reduced = [False] * self._nseg
        wb = os.environ.get("EKAID_B200_AR_BF16", "auto")
        self._wire_bf16 = (change_detector.precision == "bf16") if wb == "auto" else (wb == "1")
        if self._hooked:
            for p, k in zip(params, self._seg_of_param):
                if k < 5:
                    p.register_post_accumulate_grad_hook(lambda _p, k=k: self._grad_ready(k))
        self._cot = None
        self._graph = None

    def _opt(self, wait_current: bool):
        dev = self.opt.flat.device
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(dev)
        st = self._opt_stream
        st.wait_stream(self._main)
        if wait_current:
            st.wait_stream(torch.cuda.current_stream(dev))
        return st

    def _reduce_run(self, k0: int, k1: int) -> None:
        """All-reduce the not yet reduced parts of segments k0 .. k1-1 (contiguous runs together); caller is on the
        optimizer stream."""
        k = k0
        while k < k1:
            if self._reduced[k] or not self._nonempty(k):
                k += 1
                continue
            j = k
            while j + 1 < k1 and not self._reduced[j + 1] and self._nonempty(j + 1):
                j += 1
            allreduce_mean_(self.opt.grad[self._seg_range[k][0]:self._seg_range[j][1]], self.pg, self._wire_bf16)
            for i in range(k, j + 1):
                self._reduced[i] = True
            k = j + 1

    def _launch_reduce(self, k: int) -> bool:
        """Early gradient exchange of one complete segment (no update yet)."""
        first, last = self._seg_pidx[k]
        if self.opt.sync_slots(first, last, dry_run=True):
            return False
        with torch.cuda.stream(self._opt(wait_current=True)):
            self._reduce_run(k, k + 1)
        return True

    def _launch_run(self, k0: int, k1: int, wait_current: bool) -> bool:
        """All-reduce (what has not been exchanged yet) + Adam of the contiguous segments k0 .. k1-1 on the optimizer
        stream (after everything enqueued so far on the main stream and, if asked, on the calling stream)."""
        first, last = self._seg_pidx[k0][0], self._seg_pidx[k1 - 1][1]
        lo, hi = self._seg_range[k0][0], self._seg_range[k1 - 1][1]
        if hi <= lo:
            return False
        if self.opt.sync_slots(first, last, dry_run=True):
            return False                                    # some gradient is not in its slot: take the late path
        with torch.cuda.stream(self._opt(wait_current)):
            if self.pg is not None:
                self._reduce_run(k0, k1)
            if not self._advanced:
                self.opt.advance()
                self._advanced = True
            self.opt.update_range(lo, hi, max_ctas=int(os.environ.get("EKAID_B200_BG_CTAS", "148")))
        for k in range(k0, k1):
            self._done[k] = True
        return True

    def _nonempty(self, k: int) -> bool:
        return self._seg_range[k][1] > self._seg_range[k][0]

    def _grad_ready(self, k: int):
        """autograd hook: one more gradient of segment k has been accumulated."""
        self._fired[k] += 1
        if not (self._armed and k < 5 and self._expected is not None and self._fired[k] == self._expected[k]
                and not self._done[k] and self._nonempty(k)):
            return
        if k < 4 and self._early_segments:
            self._launch_run(k, k + 1, wait_current=True)
        elif self.pg is not None and self._early_reduce and not self._reduced[k]:
            self._launch_reduce(k)

    def _bptt_starts(self):
        """functions.BPTT_HOOK: the question path is about to run its recurrence backwards (one long kernel on 64 SMs).
        By now every image-path gradient has been enqueued: what is still pending of segments 0-4 goes out (contiguous
        segments together) and runs next to the recurrence -- and not before the GPU reaches it: the GEMMs ahead of it
        need whole SMs."""
        if not self._armed or self._expected is None:
            return
        ready = [k for k in range(5) if not self._done[k] and self._nonempty(k) and self._fired[k] == self._expected[k]]
        i = 0
        while i < len(ready):
            j = i
            while j + 1 < len(ready) and ready[j + 1] == ready[j] + 1:
                j += 1
            self._launch_run(ready[i], ready[j] + 1, wait_current=True)
            i = j + 1

    def _cotangents(self, bef):
        # stand in for d(decoder NLL)/d(bef, aft, diff): fixed unit-scale cotangents
        if self._cot is None or self._cot[0].shape != bef.shape:
            g = torch.Generator(device="cpu").manual_seed(4242)
            self._cot = [torch.randn(bef.shape, generator=g).to(bef.device) / bef.shape[1] for _ in range(3)]
        return self._cot

    def _surrogate(self, bef, aft, diff):
        self._cotangents(bef)
        return (bef * self._cot[0]).sum() + (aft * self._cot[1]).sum() + (diff * self._cot[2]).sum()

    def loss(self, inputs, labels=None, masks=None):
        pred, att_bef, att_aft, bef, aft, diff = self.cd(*inputs, setting="mode2", graph=self.graph)
        if self.decoder_loss is None and bef.is_cuda:
            # same objective as below in one launch (and five tiny ones in backward instead of ~40)
            cot = self._cotangents(bef)
            bsz = bef.shape[0]
            return functions.WeightedSumsFn.apply((1.0, 1.0, 1.0, 2.5e-03 / (2 * bsz), 2.5e-03 / (2 * bsz)),
                                                  (cot[0], cot[1], cot[2], None, None), bef, aft, diff, att_bef, att_aft)
        if self.decoder_loss is not None:
            dec = self.decoder_loss(bef, aft, diff, labels, masks)
        else:
            dec = self._surrogate(bef, aft, diff)
        bsz = bef.shape[0]
        att_sum = (att_bef.sum() + att_aft.sum()) / (2 * bsz)                 # train_mimic.py:246
        return dec + 2.5e-03 * att_sum                                        # train_mimic.py:247

    def train_step(self, inputs, labels=None, masks=None) -> torch.Tensor:
        """optimizer.zero_grad -> forward -> backward -> [all-reduce(mean)] -> Adam  (train_mimic.py:220-269).
        Returns the (device) loss tensor; no host sync happens here."""
        self.opt.zero_grad()
        # (fresh dropout masks: ChangeDetector.forward advances the device-side seed itself in train mode -- a kernel, so
        # replays of the captured step draw new masks too)
        total = self.loss(inputs, labels, masks)
        self._fired = [0] * self._nseg
        self._done = [False] * self._nseg
        self._reduced = [False] * self._nseg
        self._advanced = False
        self._main = torch.cuda.current_stream() if total.is_cuda else None
        self._armed = self._hooked
        if self._hooked:
            functions.BPTT_HOOK = self._bptt_starts
        try:
            total.backward()
        finally:
            self._armed = False
            if self._hooked:
                functions.BPTT_HOOK = None
        if self._hooked:
            if self._expected is None:
                self._expected = list(self._fired)
            elif self._fired != self._expected:
                raise RuntimeError("the set of parameters receiving gradients changed between steps (%s -> %s): the "
                                   "early optimizer updates are no longer valid" % (self._expected, self._fired))
        # whatever was not updated early: (clone-instead-of-adopt gradients go into their slots first; normally none)
        pending = [k for k in range(self._nseg) if not self._done[k] and self._seg_range[k][1] > self._seg_range[k][0]]
        if pending:
            if self._opt_stream is not None and any(self._reduced):
                torch.cuda.current_stream().wait_stream(self._opt_stream)      # early exchanges of pending segments
            for k in pending:
                self.opt.sync_slots(*self._seg_pidx[k])
            if not self._advanced:
                self.opt.advance()
                self._advanced = True
            # contiguous runs of pending segments share one all-reduce / one launch
            runs, cur = [], None
            for k in pending:
                lo, hi = self._seg_range[k]
                if cur is not None and cur[1] == lo:
                    cur[1] = hi
                else:
                    cur = [lo, hi]
                    runs.append(cur)
            for lo, hi in runs:
                if self.pg is not None:
                    # (segments exchanged early keep their flag; a run may mix both, so go segment by segment)
                    for k in pending:
                        slo, shi = self._seg_range[k]
                        if lo <= slo and shi <= hi and not self._reduced[k]:
                            allreduce_mean_(self.opt.grad[slo:shi], self.pg, self._wire_bf16)
                            self._reduced[k] = True
                self.opt.update_range(lo, hi)
        if self._opt_stream is not None and any(self._done):
            torch.cuda.current_stream().wait_stream(self._opt_stream)
        return total.detach()

    # -- CUDA-graph replay of the whole step ------------------------------------------------------------------
    def capture(self, raw_example, train: bool = True, warmup: int = 3, profile: bool = False):
        """Capture  process_matrix x4 -> forward -> backward -> [all-reduce] -> Adam  (or the inference forward) into
        one CUDA graph.  `raw_example` fixes the shapes: the device tuple `to_device` returns.  The step is ~600 small
        launches at batch 64; replaying a graph removes the Python/launch overhead between them."""
        self._static = [t.clone() for t in raw_example]
        self._train = train
        if train and self.speaker is not None and self.decoder_steps is None:
            raise RuntimeError("capturing a step with the answer decoder needs a fixed decoder_steps (the reference's "
                               "data-dependent loop length cannot be part of a CUDA graph)")

        def body():
            inputs = expand_adjacency(self._static, self.cfg)
            if train:
                return self.train_step(inputs, self._static[9], self._static[10].float())
            return self.infer_step(inputs)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        before = lib.LAUNCHES
        self._graph = torch.cuda.CUDAGraph()
        self.profile_events = None
        if profile:
            # measurement build of the graph: every C-ABI call is bracketed by event-record nodes (external events), so
            # after a replay each kernel's duration inside the replayed step can be read with cudaEventElapsedTime
            lib.PROFILE, lib.PROFILE_EXTERNAL = [], True
        try:
            with lib.graph_capture(self._graph):
                self._static_out = body()
        finally:
            if profile:
                self.profile_events, lib.PROFILE, lib.PROFILE_EXTERNAL = lib.PROFILE, None, False
        self.launches_per_replay = lib.LAUNCHES - before
        return self

    def replay(self, raw):
        """Copy a new raw device (or pinned host) batch into the captured buffers and replay the step."""
        if all(s_.is_cuda and s_.is_contiguous() and s_.dtype == d.dtype and s_.shape == d.shape
               and (s_.numel() * s_.element_size()) % 16 == 0 and s_.data_ptr() % 16 == 0 for d, s_ in zip(self._static, raw)):
            functions.copy_many_bytes(list(zip(raw, self._static)))          # one launch instead of 11 memcpy nodes
        else:
            for dst, src in zip(self._static, raw):
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        lib.LAUNCHES += self.launches_per_replay
        return self._static_out

    def pipeline(self) -> "StepPipeline":
        """Input pipeline for the captured step (see StepPipeline)."""
        return StepPipeline(self)

    @torch.no_grad()
    def infer_decode(self, inputs, check_every: int = 16):
        """test_mimic.py:116-122: graph + fusion forward, then the greedy answer tokens [B, seq_length]."""
        if self.speaker is None:
            raise RuntimeError("infer_decode needs GraphFusionStep(speaker=...)")
        out = self.cd(*inputs, setting="mode2", graph=self.graph)
        return self.speaker._sample(out[3], out[4], out[5], None, self.cfg, sample_max=1, check_every=check_every)[0]

    @torch.no_grad()
    def infer_step(self, inputs):
        """test_mimic.py:116-117: the three vectors the decoder consumes, plus the attention maps."""
        return self.cd(*inputs, setting="mode2", graph=self.graph)


class StepPipeline:
    """What a data loader does around the captured step: while step i computes, the host->device copies of batch i+1
    run on a copy stream into one of two staging sets, and the scalar result of every step comes back through a pinned
    host buffer, read one step late.  Every batch still crosses PCIe and every result is still read on the host; they
    just no longer serialise with the compute.

        pipe = step.pipeline(); pipe.prefetch(host_batch_0)
        for i in ...: pipe.prefetch(host_batch_{i+1}); k = pipe.run(); ...; value = pipe.result(k)
    """

    def __init__(self, step: GraphFusionStep):
        if step._graph is None:
            raise RuntimeError("StepPipeline needs a captured step (GraphFusionStep.capture)")
        self.step = step
        dev = step._static[0].device
        self.dev = dev
        self.copy_stream = torch.cuda.Stream(dev)
        self.stage = [[torch.empty_like(t) for t in step._static] for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]       # staging set filled
        self.consumed = [torch.cuda.Event() for _ in range(2)]    # staging set copied into the captured buffers
        self.host_out = torch.zeros(2, dtype=torch.float32).pin_memory()
        self.out_done = [torch.cuda.Event() for _ in range(2)]
        self._filled = []          # staging sets that hold a prefetched batch, oldest first
        self._next = 0
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in step._static)

    def prefetch(self, host_batch) -> None:
        k = self._next
        self._next ^= 1
        self.copy_stream.wait_event(self.consumed[k])
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(self.stage[k], host_batch):
                dst.copy_(src, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self._filled.append(k)

    def run(self) -> int:
        """Run the step on the oldest prefetched batch; returns the slot to pass to result()."""
        k = self._filled.pop(0)
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.ready[k])
        out = self.step.replay(self.stage[k])
        self.consumed[k].record(cur)
        val = out if self.step._train else out[5].sum()
        self.host_out[k:k + 1].copy_(val.reshape(1), non_blocking=True)
        self.out_done[k].record(cur)
        return k

    def result(self, k: int) -> float:
        self.out_done[k].synchronize()
        return float(self.host_out[k])
