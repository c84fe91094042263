"""Config object accepted by the drop-in constructors.

The reference passes a Detectron-style global ``cfg`` AttrDict (reference
model/configs/config.py:7-157 merged with configs/dynamic/dynamic_change_pos_mimic.yaml:1-48).
The drop-in modules only read attributes, so any object with the same attribute tree works --
including the reference's own ``cfg``.  ``default_cfg()`` rebuilds the merged values the
reference's train script ends up with (train_mimic.py:52-58,73) without importing the reference.
"""
from __future__ import annotations

import copy


class AttrDict(dict):
    """dict with attribute access (same behaviour as reference utils/attr_dict.py:1-15)."""

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: _to_attr(v) for k, v in d.items()})
    return d


def default_cfg(graph: str = "all", setting: str = "mode2", nongt_dim: int = 52) -> AttrDict:
    """Merged yaml+defaults as seen by ChangeDetector / DynamicSpeaker in the reference."""
    return _to_attr({
        "model": {
            "change_detector": {
                "input_dim": 2052, "dim": 128, "feat_dim": 1026, "att_dim": 1024, "att_head": 4,
                "nongt_dim": nongt_dim, "spa_label_num": 11, "sem_label_num": 3, "dir_num": 2,
                "pos_emb_dim": 64, "coef_sem": 0.333, "coef_spa": 0.333,
            },
            "speaker": {
                "embed_input_dim": 3072, "embed_dim": 1024, "input_dim": 1024, "seq_length": 90,
                "rnn_size": 512, "drop_prob_lm": 0.5, "word_embed_size": 300, "beam_size": 1,
                "temperature": 1.0, "vocab_size": 60,
            },
        },
        "data": {"feature_mode": "both", "train": {"batch_size": 64, "empty_image": False}},
        "train": {"setting": setting, "graph": graph,
                  "optim": {"type": "adam", "lr": 1e-4, "weight_decay": 0.0, "alpha": 0.9,
                            "beta": 0.999, "epsilon": 1e-8}},
    })


# reference model/data/vocab_mimic_VQA.json has 147 words with ids 1..147
NTOKEN = 147
WORD_TO_IDX = {"w%d" % i: i for i in range(1, NTOKEN + 1)}
