"""Shared loader for tests/golden/*.npz (built by oracle/build_fixtures.py)."""

from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from oracle import oxdna_oracle as orc

GOLDEN = Path(__file__).resolve().parent / "golden"

# per-term absolute tolerances the reference's own integration tests use
TOL = {
    "dna1": [1e-6, 1e-6, 1e-6, 1e-6, 1e-3, 1e-3, 1e-6, 1e-3],
    "dna2": [1e-6, 1e-6, 1e-6, 1e-6, 1e-3, 1e-3, 1e-6, 1e-3],
    "rna2": [1e-6, 1e-6, 1e-6, 1e-6, 1e-3, 1e-3, 1e-6, 1e-3],
    # mythos/energy/na1/tests/test_integration.py: stacking 1e-3, cross/hb 1e-4, debye 1e-5
    "na1": [1e-6, 1e-6, 1e-3, 1e-6, 1e-4, 1e-4, 1e-6, 1e-5],
}

ALL_CASES = sorted(p.stem for p in GOLDEN.glob("*.npz"))


def load_case(name: str) -> dict:
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    c = {k: z[k] for k in z.files}
    c["model"] = str(c["model"])
    c["name"] = name
    c["bonded"] = orc.bonded_pairs(c["strand_counts"].tolist(), c["circular"].tolist())
    n = c["center"].shape[1]
    c["pairs"] = orc.all_unbonded_pairs(n, c["bonded"])
    c["kt"] = float(c["t_kelvin"]) * 0.1 / 300.0
    return c


def theta_for(case: dict) -> dict:
    """Independent parameters exactly as the reference's integration tests configure them."""
    model = case["model"]
    th = orc.default_theta(
        model,
        kt=case["kt"],
        salt_conc=float(case["salt_conc"]),
        half_charged_ends=bool(case["half_charged_ends"]),
    )
    if "ss_stack_weights" in case:  # dna1/tests/test_integration.py:262-268, 214-219
        th["stacking"]["ss_stack_weights"] = torch.as_tensor(case["ss_stack_weights"])
        th["stacking"]["eps_stack_kt_coeff"] = float(case["eps_stack_kt_coeff"])
        th["hydrogen_bonding"]["ss_hb_weights"] = torch.as_tensor(case["ss_hb_weights"])
    return th


def stack_nt_type(case: dict):
    """na1/tests/test_integration.py:252 reverses nt_type per strand for the stacking test only."""
    if case["model"] != "na1":
        return None
    nt = case["nt_type"]
    return np.concatenate([nt[:8][::-1], nt[8:][::-1]])
