"""CPU tests of the host-side mirror of the reference's EnergyFunction / Configuration interface and of the C-ABI
library surface (no compute calls: there is no GPU here).

Modelled on mythos/energy/tests/test_base.py:41-436 and test_configuration.py:23-117: composition (+, *, weights),
with_params strict / qualified namespaces, without_terms, params_dict / opt_params, dependent-parameter
initialisation -- plus what is new here: the name-driven packing of configurations into the kernel bank, the model
description, the launch-group fusion and the loud failure without CUDA.
"""

import ctypes as C
import pickle
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import mythos_b200.energy.dna1 as dna1
import mythos_b200.energy.dna2 as dna2
import mythos_b200.energy.na1 as na1
import mythos_b200.energy.rna2 as rna2
from mythos_b200 import _lib, space
from mythos_b200.energy import model as kmodel
from mythos_b200.energy.base import ComposedEnergyFunction, QualifiedComposedEnergyFunction
from mythos_b200.energy.configuration import BaseConfiguration
from mythos_b200.input import toml, topology
from mythos_b200.rigid_body import Quaternion, RigidBody
from oracle import oxdna_oracle as orc

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def top():
    return topology.from_strands(["ACGTACGT", "ACGTACGT"])


# ------------------------------------------------------------------------------------------- C-ABI surface
def test_library_loads_and_exports_every_declared_symbol():
    header = (ROOT / "include" / "mythos_b200.h").read_text()
    declared = set(re.findall(r"\b(mythos_b200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no entry points parsed from the header"
    handle = C.CDLL(str(_lib.LIB_PATH))
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, missing
    assert set(_lib.EXPORTED_SYMBOLS) <= declared
    lib = _lib.lib()
    assert lib.mythos_b200_abi_version() == 3
    assert lib.mythos_b200_param_count() == 232
    names = _lib.param_names()
    assert len(names) == 231 and len(set(names)) == 231
    for i in (0, 57, 230):
        assert lib.mythos_b200_param_index(names[i].encode()) == i
    assert lib.mythos_b200_param_index(b"no.such_param") == -1
    assert lib.mythos_b200_sizeof_model() == C.sizeof(_lib.Model)
    assert lib.mythos_b200_sizeof_nl_args() == C.sizeof(_lib.NlArgs)
    assert lib.mythos_b200_nl_workspace_bytes(0, 1) == 0 and lib.mythos_b200_nl_workspace_bytes(100, 2) > 0


def test_argument_validation_happens_before_any_cuda_call():
    a = _lib.EnergyArgs()
    st = _lib.lib().mythos_b200_energy_f64(None, C.byref(a))
    assert st == 1 and b"null args" in _lib.lib().mythos_b200_last_error()
    m = _lib.Model()
    m.n_banks = 2
    a.model = C.pointer(m)
    assert _lib.lib().mythos_b200_energy_f64(None, C.byref(a)) == 2
    n = _lib.NlArgs()
    assert _lib.lib().mythos_b200_nl_build_f64(None, C.byref(n)) == 1
    lg = _lib.LangevinArgs()
    assert _lib.lib().mythos_b200_langevin_f32(None, C.byref(lg)) == 1


def test_no_cpu_fallback(top):
    efn = dna1.create_default_energy_fn(top)
    body = RigidBody(torch.zeros((16, 3), dtype=torch.float64), Quaternion(torch.ones((16, 4), dtype=torch.float64)))
    with pytest.raises(_lib.MythosB200Error, match="no CPU fallback"):
        efn(body)
    with pytest.raises(_lib.MythosB200Error):
        efn.map(RigidBody(body.center[None], Quaternion(body.orientation.vec[None])))


# ------------------------------------------------------------------------------------------- configuration
def test_configuration_required_dependent_and_merge():
    with pytest.raises(ValueError, match="Required properties"):
        dna1.FeneConfiguration(eps_backbone=2.0)
    cfg = dna1.FeneConfiguration(eps_backbone=2.0, r0_backbone=0.75, delta_backbone=0.25, fmax=500.0, finf=4.0)
    assert "eps_backbone" in cfg and "nope" not in cfg
    new = cfg | {"eps_backbone": 3.0}
    assert new.eps_backbone == 3.0 and cfg.eps_backbone == 2.0
    with pytest.raises(AttributeError):
        cfg.eps_backbone = 1.0
    with pytest.raises(ValueError, match="permitted for optimization"):
        dna1.FeneConfiguration.from_dict(dict(cfg.opt_params | {"eps_backbone": 2.0, "r0_backbone": 0.75, "delta_backbone": 0.25,
                                                                "fmax": 500.0, "finf": 4.0}), ("b_low",))
    allp = dna1.FeneConfiguration.from_dict({"eps_backbone": 2.0, "r0_backbone": 0.75, "delta_backbone": 0.25, "fmax": 500.0, "finf": 4.0}, ("*",))
    assert set(allp.opt_params) == set(cfg.required_params)
    some = allp.replace(params_to_optimize=("fmax",))
    assert list(some.opt_params) == ["fmax"]
    d = allp.to_dictionary(include_dependent=True, exclude_non_optimizable=False)
    assert set(d) == set(cfg.required_params)
    pickle.loads(pickle.dumps(allp))


def test_init_params_match_oracle_smoothing():
    cfgs = {c.term: c.init_params() for c in dna2.default_energy_configs()}
    want = orc.init_all("dna2", orc.default_theta("dna2"))
    for term, cfg in cfgs.items():
        for k in cfg.dependent_params:
            got, ref = getattr(cfg, k), want[term][k]
            np.testing.assert_allclose(np.asarray(got), np.asarray(ref), rtol=1e-13, err_msg=f"{term}.{k}")
    # seq-specific stacking weights (dna1/stacking.py:124-129 vs rna2/stacking.py:112-115)
    w = torch.arange(16, dtype=torch.float64).reshape(4, 4)
    sd = dna1.StackingConfiguration(**(dict(cfgs["stacking"].to_dictionary(include_dependent=False, exclude_non_optimizable=False)) | {"ss_stack_weights": w})).init_params()
    kt, co = float(sd.kt), float(sd.eps_stack_kt_coeff)
    np.testing.assert_allclose(sd.eps_stack.numpy(), w.numpy() * (1 - co + 9 * kt * co))


def test_toml_expressions():
    assert toml.parse_str("pi - 2.35") == pytest.approx(np.pi - 2.35, abs=0)
    assert toml.parse_str("296.15 * 0.1 / 300.0") == pytest.approx(296.15 * 0.1 / 300.0)
    assert toml.parse_str("abc") == "abc" and toml.parse_str("1.5") == 1.5
    d = toml.load_model_defaults("dna2")
    assert d["energy"]["hydrogen_bonding"]["theta0_hb_4"] == pytest.approx(np.pi)


# ------------------------------------------------------------------------------------------- composition
def test_composition_weights_and_namespaces(top):
    efn = dna1.create_default_energy_fn(top)
    assert len(efn.energy_fns) == 7 and np.allclose(np.asarray(efn.weights), 1.0)
    fene, stack = efn.energy_fns[0], efn.energy_fns[2]
    c = fene + stack
    assert isinstance(c, ComposedEnergyFunction) and c.weights is None and len(c.energy_fns) == 2
    c2 = c + (efn.energy_fns[4] * 2.0)
    assert len(c2.energy_fns) == 3 and np.allclose(np.asarray(c2.weights), [1, 1, 2])
    assert len(efn.without_terms("Fene", dna1.Stacking).energy_fns) == 5
    with pytest.raises(ValueError, match="same length"):
        ComposedEnergyFunction(energy_fns=[fene], weights=torch.ones(2))
    with pytest.raises(TypeError):
        ComposedEnergyFunction(energy_fns=[1, 2])
    # global namespace: eps_exc is shared by bonded and unbonded excluded volume (base.py:279-299)
    e2 = efn.with_params(eps_exc=3.0, kt=0.11)
    assert float(e2.energy_fns[1].params.eps_exc) == 3.0 and float(e2.energy_fns[3].params.eps_exc) == 3.0
    assert float(e2.energy_fns[2].params.kt) == 0.11
    with pytest.raises(ValueError, match="not used"):
        efn.with_params(not_a_param=1.0)
    efn.replace(strict_params=False).with_params(not_a_param=1.0)
    # dependent parameters are recomputed on every update
    e3 = efn.with_params(a_stack=7.0)
    assert float(e3.energy_fns[2].params.b_low_stack) != float(efn.energy_fns[2].params.b_low_stack)
    # qualified namespace (base.py:437-462)
    q = QualifiedComposedEnergyFunction(energy_fns=efn.energy_fns)
    q2 = q.with_params({"BondedExcludedVolume.eps_exc": 5.0})
    assert float(q2.energy_fns[1].params.eps_exc) == 5.0 and float(q2.energy_fns[3].params.eps_exc) == 2.0
    assert "Fene.eps_backbone" in q.opt_params()
    # params_dict / opt_params / with_noopt
    assert "b_low_stack" in efn.params_dict() and "b_low_stack" not in efn.params_dict(include_dependent=False)
    assert "eps_backbone" in efn.opt_params() and "eps_backbone" not in efn.with_noopt("eps_backbone").opt_params()
    assert set(efn.opt_params(from_fns=[dna1.Fene])) == {"eps_backbone", "r0_backbone", "delta_backbone", "fmax", "finf"}
    pickle.loads(pickle.dumps(efn))  # Ray transports energy functions by pickle


def test_missing_topology_information_raises():
    with pytest.raises(ValueError, match="Missing topology"):
        dna1.Fene(params=dna1.FeneConfiguration(eps_backbone=2.0, r0_backbone=0.75, delta_backbone=0.25, fmax=500.0, finf=4.0),
                  displacement_fn=space.free()[0])
    with pytest.raises(ValueError, match="is_end"):
        dna2.Debye(params=dna2.default_energy_configs()[7].init_params(), displacement_fn=space.free()[0], seq=np.zeros(4, np.int32),
                   bonded_neighbors=np.zeros((0, 2), np.int32), unbonded_neighbors=np.zeros((2, 0), np.int32))


# ------------------------------------------------------------------------------------------- kernel plan
def test_plan_model_forms_and_packing(top):
    cases = {
        "dna1": (dna1.create_default_energy_fn(top), (0, 0, 0, 0, 0)),
        "dna2": (dna2.create_default_energy_fn(top), (0, 0, 1, 1, 1)),
        "rna2": (rna2.create_default_energy_fn(top), (1, 1, 0, 1, 0)),
    }
    for name, (efn, (stack, cross, coax, debye, ubs)) in cases.items():
        assert kmodel.fusable_groups(efn.energy_fns) == [list(range(len(efn.energy_fns)))], name
        plan = kmodel.plan_for(efn.energy_fns)
        f = plan.model.forms[0]
        assert (f.stack_form, f.cross_form, f.coax_form, f.has_debye, plan.model.geom[0].use_back_stack) == (stack, cross, coax, debye, ubs)
        assert plan.model.n_banks == 1 and plan.term_mask == (0xFF if debye else 0x7F)
        vec = plan.params_vector()
        assert vec.shape == (232,) and vec.dtype == torch.float64
        want = orc.init_all(name, orc.default_theta(name))
        for idx, nm in enumerate(_lib.param_names()):
            term, field = nm.split(".")
            m = re.match(r"(\w+)\[(\d),(\d)\]", field)
            ref = want.get(term, {}).get(m.group(1) if m else field)
            if ref is None:
                assert float(vec[idx]) == 0.0, nm
            else:
                ref = torch.as_tensor(ref, dtype=torch.float64)
                ref = ref[int(m.group(2)), int(m.group(3))] if m else ref
                assert float(vec[idx]) == pytest.approx(float(ref), rel=1e-13), nm
    # periodic box reaches the model; a periodic and a free term cannot share a launch
    efn = dna1.create_default_energy_fn(top, displacement_fn=space.periodic(20.0)[0])
    assert tuple(kmodel.plan_for(efn.energy_fns).model.box) == (20.0, 20.0, 20.0)
    mixed = [efn.energy_fns[0], dna1.create_default_energy_fn(top).energy_fns[1]]
    assert len(kmodel.fusable_groups(mixed)) == 2
    # theta gradients chain through the pack
    a = torch.tensor(6.0, dtype=torch.float64, requires_grad=True)
    v = kmodel.plan_for(dna1.create_default_energy_fn(top).with_params(a_stack=a).energy_fns).params_vector()
    (g,) = torch.autograd.grad(v[_lib.param_names().index("stacking.b_low_stack")], a)
    assert torch.isfinite(g) and float(g) != 0.0


def test_hybrid_plan_three_banks(top):
    nt = np.array([1] * 8 + [2] * 8, dtype=np.int32)
    t2 = topology.Topology(n_nucleotides=16, strand_counts=top.strand_counts, bonded_neighbors=top.bonded_neighbors, seq=top.seq,
                           is_end=top.is_end, nt_type=nt)
    efn = na1.create_default_energy_fn(t2)
    plan = kmodel.plan_for(efn.energy_fns)
    assert plan.hybrid and plan.model.n_banks == 3 and plan.params_vector().shape == (696,)
    forms = [(plan.model.forms[b].stack_form, plan.model.forms[b].cross_form, plan.model.forms[b].coax_form) for b in range(3)]
    assert forms == [(0, 0, 1), (1, 1, 0), (0, 0, 0)]
    r = kmodel.interaction_range(plan)
    assert 3.0 < r < 3.6
    with pytest.raises(_lib.MythosB200Error, match="mix"):
        kmodel.plan_for([efn.energy_fns[0], dna1.create_default_energy_fn(top).energy_fns[1]])


def test_topology_parsing_and_all_pairs_sentinel(tmp_path, top):
    classic = tmp_path / "c.top"
    classic.write_text("4 2\n1 A -1 1\n1 C 0 -1\n2 G -1 3\n2 T 2 -1\n")
    t = topology.from_oxdna_file(classic)
    assert t.n_nucleotides == 4 and t.bonded_neighbors.tolist() == [[0, 1], [2, 3]] and t.seq.tolist() == [0, 1, 2, 3]
    assert t.is_end.tolist() == [1, 1, 1, 1] and t.nt_type.tolist() == [0, 0, 1, 1]
    assert sorted(map(tuple, t.unbonded_neighbors.tolist())) == [(0, 2), (0, 3), (1, 2), (1, 3)]
    new = tmp_path / "n.top"
    new.write_text("5 2 5->3\nACG type=DNA circular=true\nUU type=RNA\n")
    t, fmt = topology.from_oxdna_file(new, return_format=True)
    assert fmt == "new" and t.seq.tolist() == [2, 1, 0, 3, 3] and t.nt_type.tolist() == [1, 1, 1, 2, 2]
    assert t.bonded_neighbors.tolist() == [[0, 1], [1, 2], [0, 2], [3, 4]] and t.is_end.tolist() == [0, 0, 0, 1, 1]
    big = topology.from_strands(["A" * 400, "T" * 400])
    assert isinstance(big.unbonded_neighbors, topology.AllPairs) and big.unbonded_neighbors_t is big.unbonded_neighbors_t
    assert top.unbonded_neighbors_t.shape == (2, 16 * 15 // 2 - 14)
