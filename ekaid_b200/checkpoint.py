"""Checkpoint files in the reference's layout (train_mimic.py:280-290 writes them, train_mimic.py:126-141 and
test_mimic.py:62-79 read them): a torch-pickled dict with `change_detector_state`, `speaker_state`, `model_cfg`.
The drop-in ChangeDetector keeps the reference's state_dict keys and shapes, so a checkpoint written by either side
loads into the other with strict=True; the dead parameters of the reference (SSRE.*, fc1.*, neighbor_net.0.*,
linear_out_.*) are carried as plain parameters so nothing is dropped on the way through."""
import os
from typing import Any, Dict, Optional

import torch

KEYS = ("change_detector_state", "speaker_state", "model_cfg")


def save_checkpoint(path: str, change_detector, speaker=None, cfg: Any = None) -> None:
    """utils/utils.py:14-18 `save_checkpoint(state, filename)` with the dict of train_mimic.py:283-287.
    Tensors are written from host copies so that the file loads on a machine without a GPU."""
    def host(sd):
        return {k: v.detach().to("cpu") for k, v in sd.items()}
    state = {
        "change_detector_state": host(change_detector.state_dict()),
        "speaker_state": host(speaker.state_dict()) if speaker is not None else {},
        "model_cfg": cfg,
    }
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)


def load_checkpoint(path: str) -> Dict[str, Any]:
    """utils/utils.py:8-12,21-23 `load_checkpoint` = torch.load onto the CPU.  `model_cfg` is the reference's AttrDict
    (a pickled python object), so the file is read with weights_only=False like the reference does."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    missing = [k for k in KEYS[:2] if k not in ckpt]
    if missing:
        raise KeyError("not an EKAID checkpoint, missing %s" % missing)
    return ckpt


def restore(change_detector, ckpt: Dict[str, Any], speaker: Optional[torch.nn.Module] = None) -> None:
    """train_mimic.py:136,140: strict load of both state dicts (a key or shape mismatch raises RuntimeError, as in
    the reference).  The bf16 operand copies of the weights are made inside every forward (cast_many), so there is no
    cache to invalidate after a load."""
    change_detector.load_state_dict(ckpt["change_detector_state"], strict=True)
    if speaker is not None:
        speaker.load_state_dict(ckpt["speaker_state"], strict=True)
