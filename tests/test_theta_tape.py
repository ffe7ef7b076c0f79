"""theta -> parameter-bank chain replayed from the recorded tape (mythos_b200/energy/theta_tape.py, csrc/theta_tape.cu)
against the eager ``with_params`` / ``init_params`` chain it was recorded from (the reference's
``configuration.py:110-113`` + ``dna1/base_smoothing_functions.py:48-142`` formulas): values and reverse-mode gradients,
all four models, perturbed parameters (the tape is recorded at the defaults and replayed elsewhere)."""

import ctypes as C

import numpy as np
import pytest
import torch

from mythos_b200 import _lib
from mythos_b200.energy import dna1, dna2, na1, rna2, theta_tape
from mythos_b200.energy import model as kmodel
from mythos_b200.utils import synthetic

MODELS = {"dna1": dna1, "dna2": dna2, "rna2": rna2, "na1": na1}


@pytest.fixture(scope="module")
def topology():
    return synthetic.assembly(1, pitch=2.6, seed=1).topology


@pytest.mark.parametrize("name", list(MODELS))
def test_replay_matches_eager_chain(name, topology):
    efn = MODELS[name].create_default_energy_fn(topology)
    theta = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in efn.opt_params().items()}
    rng = np.random.default_rng(3)
    moved = {k: v * (1.0 + 0.02 * torch.as_tensor(rng.standard_normal(tuple(v.shape)))) for k, v in theta.items()}
    flat = theta_tape.FlatParams(moved)
    flat.flat.requires_grad_(True)
    bound = theta_tape.bind(efn, flat)
    assert isinstance(bound, theta_tape.BoundEnergyFunction), "the default compositions must take the replayed path"
    bank = bound._bank
    cot = torch.as_tensor(rng.standard_normal(bank.numel()))
    (g,) = torch.autograd.grad((bank * cot).sum(), [flat.flat])

    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in moved.items()}
    fns = efn.with_params(leaves).energy_fns
    want = kmodel.bank_vector(fns, any(fn.HYBRID for fn in fns))
    assert want.numel() == bank.numel() == _lib.param_count() * (3 if name == "na1" else 1)
    np.testing.assert_allclose(bank.detach().numpy(), want.detach().numpy(), rtol=1e-13, atol=1e-13)
    gw = torch.autograd.grad((want * cot).sum(), list(leaves.values()), allow_unused=True)
    got = flat.unflatten(g)
    for k, w in zip(leaves, gw):
        w = torch.zeros_like(leaves[k]) if w is None else w
        np.testing.assert_allclose(got[k].numpy(), w.numpy(), rtol=1e-9, atol=1e-9 * (1.0 + float(gw_scale(gw))))


def gw_scale(gs):
    return max(float(g.abs().max()) for g in gs if g is not None)


def test_bound_function_defers_everything_else(topology):
    efn = dna2.create_default_energy_fn(topology)
    theta = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in efn.opt_params().items()}
    theta["eps_hb"] = theta["eps_hb"] * 1.1
    bound = theta_tape.bind(efn, theta)
    real = efn.with_params(theta)
    assert float(bound.params_dict()["eps_hb"]) == float(real.params_dict()["eps_hb"])
    assert len(bound.energy_fns) == len(real.energy_fns)
    again = bound.with_params({"eps_hb": 2.0})
    assert float(again.params_dict()["eps_hb"]) == 2.0


def test_table_valued_parameters(topology):
    """ss_stack_weights (4,4) as an optimised parameter: a table input of the tape."""
    efn = dna1.create_default_energy_fn(topology)
    w = torch.rand(4, 4, dtype=torch.float64) + 0.5
    efn = efn.with_params({"ss_stack_weights": w})
    theta = {"ss_stack_weights": w * 1.05, "eps_stack_kt_coeff": torch.tensor(2.7, dtype=torch.float64)}
    flat = theta_tape.FlatParams(theta)
    flat.flat.requires_grad_(True)
    bound = theta_tape.bind(efn, flat)
    assert isinstance(bound, theta_tape.BoundEnergyFunction)
    fns = efn.with_params({k: v.clone() for k, v in theta.items()}).energy_fns
    np.testing.assert_allclose(bound._bank.detach().numpy(), kmodel.bank_vector(fns, False).detach().numpy(), rtol=1e-13, atol=0)
    (g,) = torch.autograd.grad(bound._bank.sum(), [flat.flat])
    assert flat.unflatten(g)["ss_stack_weights"].shape == (4, 4) and float(g.abs().sum()) > 0


def test_flat_params_is_a_mapping():
    fp = theta_tape.FlatParams({"b": 2.0, "a": torch.tensor([[1.0, 2.0], [3.0, 4.0]])})
    assert list(fp) == ["a", "b"] and len(fp) == 2 and dict(**fp)["b"].item() == 2.0
    assert fp["a"].shape == (2, 2) and fp.flat.tolist() == [1.0, 2.0, 3.0, 4.0, 2.0]
    with pytest.raises(KeyError):
        fp["c"]


def test_c_abi_rejects_malformed_tapes():
    lib = _lib.lib()
    op = np.array([theta_tape.OP_INPUT, theta_tape.OP_ADD], dtype=np.int32)
    a0 = np.array([0, 0], dtype=np.int32)
    a1 = np.array([-1, 1], dtype=np.int32)  # node 1 reads itself
    imm = np.zeros(2)
    out = np.array([1], dtype=np.int32)
    t = theta_tape._CTape(2, 1, 1, 0, op.ctypes.data, a0.ctypes.data, a1.ctypes.data, imm.ctypes.data, out.ctypes.data)
    x, vals, res = np.ones(1), np.zeros(2), np.zeros(1)
    assert lib.mythos_b200_theta_tape_forward(C.byref(t), x.ctypes.data, vals.ctypes.data, res.ctypes.data) == 1
    a1[1] = 0
    assert lib.mythos_b200_theta_tape_forward(C.byref(t), x.ctypes.data, vals.ctypes.data, res.ctypes.data) == 0
    assert res[0] == 2.0
