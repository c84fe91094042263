"""CPU-side checks: the C-ABI library loads without a GPU and exports every symbol the header declares; the drop-in
modules keep the reference's constructor contract and state_dict keys; host-side helpers."""
import contextlib
import ctypes
import io
import json
import os

import pytest
import torch

from helpers import GOLDEN, spec_for


def test_library_exports_every_declared_symbol():
    from ekaid_b200 import lib
    protos = lib.parse_header()
    assert len(protos) >= 35
    assert os.path.exists(lib.LIB_PATH), "build with python -m ekaid_b200.build"
    so = ctypes.CDLL(lib.LIB_PATH)
    for name in protos:
        assert hasattr(so, name), name
    so.ekaid_abi_version.restype = ctypes.c_int
    assert so.ekaid_abi_version() == 1
    lib.load()


def test_epilogue_struct_layout_matches_header(tmp_path):
    """ctypes mirror of struct ekaid_epilogue == what a C compiler lays out from the header."""
    import subprocess
    from ekaid_b200.lib import Epilogue, HEADER
    src = tmp_path / "lay.c"
    fields = [f[0] for f in Epilogue._fields_]
    body = "".join('printf("%%zu\\n", offsetof(ekaid_epilogue_t, %s));' % f for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu\\n", '
                   'sizeof(ekaid_epilogue_t));%s return 0;}' % (HEADER, body))
    exe = tmp_path / "lay"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    nums = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert nums[0] == ctypes.sizeof(Epilogue)
    assert nums[1:] == [getattr(Epilogue, f).offset for f in fields]


def test_graph_capture_holds_the_garbage_collector_off(monkeypatch):
    """lib.graph_capture: no cyclic collection may run while a step is being captured (an old CUDAGraph freed mid-capture
    invalidates it); the collector's previous state comes back afterwards, also when the capture raises."""
    import gc
    from ekaid_b200 import lib
    seen = {}

    @contextlib.contextmanager
    def fake_graph(g):
        seen["enabled_inside"] = gc.isenabled()
        yield

    monkeypatch.setattr(torch.cuda, "graph", fake_graph)
    assert gc.isenabled()
    with lib.graph_capture(object()):
        assert not gc.isenabled()
    assert gc.isenabled() and seen["enabled_inside"] is False
    with pytest.raises(RuntimeError):
        with lib.graph_capture(object()):
            raise RuntimeError("capture failed")
    assert gc.isenabled()
    gc.disable()
    try:
        with lib.graph_capture(object()):
            pass
        assert not gc.isenabled()                # was off before: stays off
    finally:
        gc.enable()


def test_no_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ekaid_b200 import lib
    with pytest.raises(lib.EkaidError):
        lib.require_device()
    from ekaid_b200.config import WORD_TO_IDX, default_cfg
    from ekaid_b200.modules import ChangeDetector
    with contextlib.redirect_stdout(io.StringIO()):
        m = ChangeDetector(default_cfg(), WORD_TO_IDX).eval()
    x = torch.zeros(1, 52, 1024)
    with pytest.raises(lib.EkaidError):
        m(x, x, torch.zeros(1, 52, 52, 11), torch.zeros(1, 52, 52, 11), torch.zeros(1, 52, 52, 3),
          torch.zeros(1, 52, 52, 3), torch.zeros(1, 52, 4).double(), torch.zeros(1, 52, 4).double(),
          torch.zeros(1, 20, dtype=torch.long))


@pytest.mark.parametrize("graph", ["all", "semantic", "spatial", "implicit", "i+s"])
def test_state_dict_contract(graph):
    from ekaid_b200.config import WORD_TO_IDX, default_cfg
    from ekaid_b200.modules import ChangeDetector
    with contextlib.redirect_stdout(io.StringIO()):
        m = ChangeDetector(default_cfg(graph), WORD_TO_IDX)
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ref = spec_for(graph)
    assert list(mine.keys()) == list(ref.keys())
    assert mine == ref
    assert not m.w_emb.emb_.weight.requires_grad          # frozen table (language_model.py:28-29)


def test_constructor_errors_like_reference():
    from ekaid_b200.config import WORD_TO_IDX, default_cfg
    from ekaid_b200.modules import ChangeDetector
    cfg = default_cfg()
    cfg.model.change_detector.att_head = 3                 # 1024 % 3 != 0 -> ValueError (modules.py:20-23)
    with pytest.raises(ValueError), contextlib.redirect_stdout(io.StringIO()):
        ChangeDetector(cfg, WORD_TO_IDX)
    cfg = default_cfg()
    cfg.model.change_detector.coef_sem = 0.8               # coef_sem + coef_spa > 1 -> AssertionError (:120)
    with pytest.raises(AssertionError), contextlib.redirect_stdout(io.StringIO()):
        ChangeDetector(cfg, WORD_TO_IDX)


def test_synthetic_loader_shapes_and_rules():
    from ekaid_b200.synthetic import spatial_labels_from_boxes, synthetic_batch
    b = synthetic_batch(4, 52, seed=1)
    assert len(b) == 13
    assert b[0].shape == (4, 52, 1024) and b[0].dtype == torch.float32 and float(b[0].min()) >= 0
    assert b[2].shape == (4, 1, 91) and b[4].shape == (4, 1, 91) and b[12].shape == (4, 20)
    assert b[6].shape == (4, 100, 100) and b[6].dtype == torch.float64 and b[10].dtype == torch.float64
    assert int(b[6].max()) <= 11 and int(b[8].max()) <= 2
    assert torch.equal(b[8], b[8].transpose(1, 2))          # semantic labels are symmetric
    # spatial: self edge is label 3 (IoU = 1), transposed entry is reverse_type
    bb = torch.tensor([[[0., 0., 100., 100.], [10., 10., 50., 50.], [400., 0., 500., 100.], [0., 900., 50., 1000.]]])
    lab = spatial_labels_from_boxes(bb)[0]
    assert lab.diagonal().tolist() == [3, 3, 3, 3]
    assert lab[0, 1] == 1 and lab[1, 0] == 2                # box 0 strictly contains box 1
    assert lab[0, 2] == 3 + 0 or lab[0, 2] >= 3             # right neighbour: angle 0 -> ceil(0/45)+3 = 3
    assert lab[0, 3] == 0 and lab[3, 0] == 0                # further than (1024+1024)/3 apart
    same = synthetic_batch(4, 52, seed=1)
    assert all(torch.equal(x, y) for x, y in zip(b, same))


def test_collate_narrows_label_matrices_to_int8_losslessly():
    """select_fields(compact=True): the four label matrices travel as int8 and hold the same labels."""
    from ekaid_b200.step import select_fields
    from ekaid_b200.synthetic import synthetic_batch
    b = synthetic_batch(3, 52, seed=5)
    wide, narrow = select_fields(b, compact=False), select_fields(b)
    for i in (2, 3, 4, 5):
        assert wide[i].dtype == torch.float64 and narrow[i].dtype == torch.int8
        assert torch.equal(narrow[i].double(), wide[i])
    for i in (0, 1, 6, 7, 8, 9, 10):
        assert narrow[i].dtype == wide[i].dtype and torch.equal(narrow[i], wide[i])
    per_sample = lambda t: sum(x.numel() * x.element_size() for x in t) // 3    # noqa: E731
    assert per_sample(wide) - per_sample(narrow) == 4 * 100 * 100 * 7


def test_loader_schema_round_trip_and_sample_rules(tmp_path):
    """ekaid_b200.loader: arrays with the reference's HDF5 schema -> the dataset's samples -> rcc_collate must reproduce the
    13-tuple the synthetic loader emits for the same seed (same tensors, same dtypes), through an .npz round trip; the mask
    rule of rcc_dataset_pos_mimic.py:258-263; int8 label matrices in compact mode."""
    from ekaid_b200.loader import FEATURE_KEYS, LABEL_KEYS, RCCArrays, RCCDataset, batches, rcc_collate
    from ekaid_b200.step import select_fields
    from ekaid_b200.synthetic import synthetic_batch
    arr = RCCArrays.synthetic(6, 52, seed=11)
    assert arr["image_features"].shape == (12, 52, 1024) and arr["image_features"].dtype.name == "float32"
    assert arr["image_adj_matrix"].shape == (12, 100, 100) and arr["image_adj_matrix"].dtype.name == "int64"
    assert arr["questions"].shape == (6, 20) and arr["answers"].shape == (6, 90) and arr["feature_idx"].shape == (6, 2)
    path = str(tmp_path / "rcc.npz")
    arr.save_npz(path)
    arr2 = RCCArrays.from_npz(path)
    assert set(FEATURE_KEYS + LABEL_KEYS) <= set(arr2.a)
    ds = RCCDataset(arr2)
    assert len(ds) == 6
    got = rcc_collate([ds[i] for i in range(6)])
    ref = synthetic_batch(6, 52, seed=11)
    assert len(got) == 13
    for i, (g, r) in enumerate(zip(got, ref)):
        assert g.shape == r.shape, i
        if i in (10, 11):
            assert g.dtype == torch.float64 and float((g - r.float().double()).abs().max()) == 0.0      # boxes: f32 on disk
        else:
            assert g.dtype == r.dtype and torch.equal(g, r), i
    seq, mask = got[2][:, 0], got[4][:, 0]
    assert torch.equal(mask.sum(1), (seq != 0).sum(1) + 1) and int(seq[:, -1].abs().max()) == 0
    comp = next(batches(ds, 4, compact=True))
    assert comp[6].dtype == torch.int8 and torch.equal(comp[6].double(), got[6][:4])
    assert select_fields(comp)[2] is comp[6]                      # already compact: passed through untouched
    with pytest.raises(KeyError):
        RCCArrays({"image_features": arr["image_features"]})
    with pytest.raises(ImportError):
        RCCArrays.from_hdf5("a.h5", "b.h5")                       # no h5py in this image: loud, not emulated


def test_speaker_constructor_state_dict_and_host_logic():
    """ekaid_b200.speaker.DynamicSpeaker on the CPU: the reference's parameter tree (tests/golden/speaker_spec.json, made from
    the reference's own module), the step count of its teacher-forcing loop (dynamic_speaker_change_pos.py:210-214), the
    entry points that are deliberately not implemented, and a loud failure without a GPU."""
    from helpers import speaker_spec
    from ekaid_b200 import lib
    from ekaid_b200.config import default_cfg
    from ekaid_b200.speaker import DynamicSpeaker, LanguageModelCriterion
    from oracle import ekaid_oracle as O
    cfg = default_cfg("all")
    with contextlib.redirect_stdout(io.StringIO()):
        sp = DynamicSpeaker(cfg, vocab_size=148)
    assert {k: tuple(v.shape) for k, v in sp.state_dict().items()} == speaker_spec()
    h, c = sp.init_hidden(5)
    assert h.shape == (2, 5, 512) and c.shape == (2, 5, 512) and float(h.abs().max()) == 0.0
    # the loop stops at the first all-empty column i >= 1, else runs seq_length steps
    seq = torch.zeros(3, 91, dtype=torch.long)
    seq[:, 0] = 1
    seq[0, 1:7] = 5
    seq[1, 1:4] = 9
    assert sp._steps(seq) == 7
    seq[2, 1:91] = 3
    assert sp._steps(seq) == 90
    seq[:, 0] = 0                      # column 0 is never a stop column (the reference tests i >= 1 only)
    assert sp._steps(seq) == 90
    # masked NLL restatement == oracle's on random log-probabilities
    g = torch.Generator().manual_seed(0)
    logp = torch.log_softmax(torch.randn(3, 90, 148, generator=g), 2)
    tgt = torch.randint(0, 148, (3, 90), generator=g)
    mask = (torch.rand(3, 90, generator=g) > 0.4).float()
    assert torch.allclose(LanguageModelCriterion()(logp, tgt, mask), O.lm_criterion(logp, tgt, mask))
    x = torch.zeros(3, 1024)
    sp.eval()
    cfg2 = default_cfg("all")
    cfg2.model.speaker.beam_size = 3
    with pytest.raises(NotImplementedError):
        sp._sample(x, x, x, seq, cfg2, sample_max=1)                # beam search
    sp.train()
    with pytest.raises(NotImplementedError):
        sp._sample(x, x, x, seq, cfg, sample_max=1)                 # inference entry point in train mode
    sp.ss_prob = 0.25
    with pytest.raises(NotImplementedError):
        sp._forward(x, x, x, seq)                                   # scheduled sampling
    sp.ss_prob = 0.0
    if not torch.cuda.is_available():
        with pytest.raises(lib.EkaidError):
            sp.eval()._forward(x, x, x, seq)                        # no CPU fallback
        with pytest.raises(lib.EkaidError):
            sp._sample(x, x, x, seq, cfg, sample_max=1)


def test_golden_manifest_complete():
    from helpers import CASES
    for c in CASES:
        assert os.path.exists(os.path.join(GOLDEN, c + ".npz")), c
    assert os.path.exists(os.path.join(GOLDEN, "make_golden.py"))
    json.load(open(os.path.join(GOLDEN, "state_dict_spec.json")))


def test_checkpoint_layout_round_trip(tmp_path):
    """train_mimic.py:283-287 checkpoint dict: written by the drop-in, read back strictly; a reference-shaped
    state dict (golden spec) loads strictly too; a foreign file is refused."""
    from ekaid_b200 import checkpoint as C
    from ekaid_b200.config import WORD_TO_IDX, default_cfg
    from ekaid_b200.modules import ChangeDetector
    from ekaid_b200.synthetic import synthetic_state_dict
    cfg = default_cfg("all")
    with contextlib.redirect_stdout(io.StringIO()):
        a = ChangeDetector(cfg, WORD_TO_IDX)
        b = ChangeDetector(cfg, WORD_TO_IDX)
    a.load_state_dict(synthetic_state_dict(spec_for("all"), 7), strict=True)
    path = str(tmp_path / "checkpoint_10000.pt")
    C.save_checkpoint(path, a, speaker=None, cfg=cfg)
    ckpt = C.load_checkpoint(path)
    assert set(C.KEYS) <= set(ckpt)
    assert list(ckpt["change_detector_state"].keys()) == list(spec_for("all").keys())
    C.restore(b, ckpt)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    assert ckpt["model_cfg"].model.change_detector.att_head == cfg.model.change_detector.att_head
    with contextlib.redirect_stdout(io.StringIO()):
        c = ChangeDetector(default_cfg("semantic"), WORD_TO_IDX)
    with pytest.raises(RuntimeError):
        C.restore(c, ckpt)                                  # graph='semantic' has no spatial/implicit encoders
    other = str(tmp_path / "other.pt")
    torch.save({"weights": 1}, other)
    with pytest.raises(KeyError):
        C.load_checkpoint(other)


def test_product_never_imports_the_oracle_or_reads_the_reference():
    """The oracle is test infrastructure: nothing under ekaid_b200/ may import it, and nothing the GPU box runs
    (package, bench.py, __graft_entry__.py, GPU tests) may read /root/reference at run time."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ekaid_b200")
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), (fn, names)
    runtime = [os.path.join(pkg, f) for f in os.listdir(pkg) if f.endswith(".py")]
    runtime += [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")]
    runtime += [os.path.join(root, "tests", f) for f in os.listdir(os.path.join(root, "tests")) if f.startswith("test_gpu")]
    for path in runtime:
        tree = ast.parse(open(path).read())
        docs = set()
        for node in ast.walk(tree):
            if isinstance(node, (ast.Module, ast.ClassDef, ast.FunctionDef, ast.AsyncFunctionDef)) and node.body and \
                    isinstance(node.body[0], ast.Expr) and isinstance(node.body[0].value, ast.Constant):
                docs.add(id(node.body[0].value))              # docstrings may cite reference paths
        for node in ast.walk(tree):
            if isinstance(node, ast.Constant) and isinstance(node.value, str) and id(node) not in docs:
                assert "/root/reference" not in node.value, (path, node.lineno)
