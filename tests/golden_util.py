"""Helpers shared by the parity tests: load golden cases and regenerate their inputs."""
import json
import os

import numpy as np

from pypevoc_b200 import signals

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(GOLD, "cases.json")) as _fh:
    CASES = json.load(_fh)


def case_signal(name):
    c = CASES[name]
    out = getattr(signals, c["generator"])(**c["gen_kwargs"])
    if isinstance(out, tuple):
        x, sr = out
    else:
        x, sr = out, c["gen_kwargs"]["sr"]
    return np.asarray(x, dtype=np.float32), sr


def case_golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=True)


def pv_kwargs(name):
    return dict(CASES[name]["pv_kwargs"])
