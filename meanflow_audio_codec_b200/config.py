"""Reads the reference's ``configs/*.json`` files for the factories of this package.

The reference's ``TrainFlowConfig`` (configs/config.py: validation, v1 -> v2 migration, diff / merge, schema text) is control
plane and stays the reference's; what the hot path needs from a config is attribute access to its keys with ``None`` for the
ones a file leaves out -- the convention ``create_tokenization_strategy`` (preprocessing/tokenization_utils.py:15-48),
``create_flow_model`` (models/factories.py:106-148) and ``create_loss_strategy`` (trainers/train.py:52-153) rely on.
A ``TrainFlowConfig`` instance can be passed to those factories directly; this loader exists for boxes without the reference.
"""
from __future__ import annotations

import json
from pathlib import Path


class FlowConfig:
    """Attribute view of one config dict; keys the file does not carry read as ``None`` (the reference's defaults apply)."""

    def __init__(self, values: dict):
        self._values = dict(values)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return self._values.get(name)

    def to_dict(self) -> dict:
        return dict(self._values)

    def __repr__(self):
        return f"FlowConfig({self._values!r})"


def load_config(source) -> FlowConfig:
    """``source``: a path to a ``configs/*.json`` file, or a dict with the same keys."""
    if isinstance(source, dict):
        return FlowConfig(source)
    with Path(source).open("r", encoding="utf-8") as f:
        return FlowConfig(json.load(f))
