"""Build the PRODUCT's energy functions (mythos_b200.energy.*) for the golden cases, the way the reference's
integration tests configure theirs, so the parity tests read like the reference's own."""

from __future__ import annotations

import numpy as np
import torch

import mythos_b200.energy.dna1 as dna1
import mythos_b200.energy.dna2 as dna2
import mythos_b200.energy.na1 as na1
import mythos_b200.energy.rna2 as rna2
from mythos_b200 import space
from mythos_b200.energy.base import ComposedEnergyFunction
from mythos_b200.input.topology import Topology
from tests.golden_cases import stack_nt_type


def topology_of(case: dict) -> Topology:
    return Topology(
        n_nucleotides=int(case["center"].shape[1]),
        strand_counts=np.asarray(case["strand_counts"]),
        bonded_neighbors=np.asarray(case["bonded"], dtype=np.int32).reshape(-1, 2),
        seq=np.asarray(case["seq"], dtype=np.int32),
        is_end=np.asarray(case["is_end"], dtype=np.int32),
        nt_type=np.asarray(case["nt_type"], dtype=np.int32),
    )


def energy_fn_of(case: dict, box=20.0) -> ComposedEnergyFunction:
    model = case["model"]
    top = topology_of(case)
    disp = space.periodic(box)[0] if box else space.free()[0]
    kt = torch.tensor(case["kt"], dtype=torch.float64)
    over = {"kT": kt, "salt_conc": torch.tensor(float(case["salt_conc"]), dtype=torch.float64),
            "half_charged_ends": bool(case["half_charged_ends"])}
    if model == "na1":
        cfgs = na1.default_energy_configs(top.nt_type, kt=kt, salt_conc=over["salt_conc"],
                                          half_charged_ends=over["half_charged_ends"], stack_nt_type=stack_nt_type(case))
        return ComposedEnergyFunction.from_lists(
            energy_fns=na1.default_energy_fns(), energy_configs=cfgs, transform_fn=na1.default_transform_fn(),
            displacement_fn=disp, topology=top,
        )
    mod = {"dna1": dna1, "dna2": dna2, "rna2": rna2}[model]
    efn = ComposedEnergyFunction.from_lists(
        energy_fns=mod.default_energy_fns(), energy_configs=mod.default_energy_configs(overrides=over),
        transform_fn=mod.default_transform_fn(), displacement_fn=disp, topology=top,
    )
    if "ss_stack_weights" in case:  # dna1/tests/test_integration.py:262-282, 214-228
        efn = efn.with_params(
            ss_stack_weights=torch.as_tensor(case["ss_stack_weights"]),
            eps_stack_kt_coeff=torch.tensor(float(case["eps_stack_kt_coeff"]), dtype=torch.float64),
            ss_hb_weights=torch.as_tensor(case["ss_hb_weights"]),
        )
    return efn
