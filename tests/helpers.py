"""Shared helpers for the parity tests (golden fixtures, oracle access)."""
import json
import os

import numpy as np

from fireredtts2_b200.config import PRESETS
from fireredtts2_b200.weights import adversarial_state_dict, synthetic_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def manifest():
    with open(os.path.join(GOLDEN, "MANIFEST.json")) as f:
        return json.load(f)["cases"]


def cases(kind):
    return [c for c in manifest() if c["kind"] == kind]


def load_case(case):
    """-> (cfg, state_dict, npz)"""
    g = np.load(os.path.join(GOLDEN, case["name"] + ".npz"))
    cfg = PRESETS[case["preset"]]
    if case["kind"] == "reference_init":
        sd = {k[4:]: g[k] for k in g.files if k.startswith("sd::")}
    elif case.get("weights") == "adversarial":
        sd = adversarial_state_dict(cfg, case["wseed"])
    else:
        sd = synthetic_state_dict(cfg, case["wseed"])
    return cfg, sd, g
