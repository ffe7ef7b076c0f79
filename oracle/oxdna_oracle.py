"""CPU oracle for the oxDNA-family energy path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain torch-float64 on the CPU, the algorithm of the
reference's JAX energy functions so the CUDA kernels have something to be
checked against.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu-baseline legs may import it; nothing under ``mythos_b200/``
does.  It is deliberately written pair-array style (one row per listed pair,
``where`` selects, no early-outs), i.e. the way the reference evaluates, and
NOT the way the kernels do.

Parity status: energies are PINNED against the reference's oxDNA-standalone
golden files (``tests/golden/*.npz`` built by ``oracle/build_fixtures.py`` from
``/root/reference/data/test-data``; checked in ``tests/test_oracle_golden.py``).
Derivatives (forces, dE/dq, dE/dtheta) are UNPINNED by the reference (no
reference test differentiates a real term): they come from torch autograd on
this restatement and are self-checked by central finite differences.

Reference files followed (relative to /root/reference):
  axes            mythos/energy/utils.py:18-36
  sites           mythos/energy/dna1/nucleotide.py:29-53, dna2/nucleotide.py:30-58,
                  rna2/nucleotide.py:33-78, na1/nucleotide.py:23-78
  primitives      mythos/energy/potentials.py:11-70, dna1/base_functions.py:13-129,
                  dna2/base_functions.py:13-17, mythos/utils/math.py:68-81
  smoothing       mythos/energy/dna1/base_smoothing_functions.py:13-142
  terms           dna1/interactions.py, dna2/interactions.py, rna2/interactions.py and the
                  pairwise_energies of every term class (cited per function below)
  composition     mythos/energy/base.py:312-319
  displacement    jax_md.space.periodic / free (third party, jax_md==0.2.28, un-vendored):
                  periodic d(a,b) = mod(a-b+L/2, L) - L/2
"""

from __future__ import annotations

import math
from pathlib import Path

import torch

DT = torch.float64
TERMS = (
    "fene",
    "bonded_excluded_volume",
    "stacking",
    "unbonded_excluded_volume",
    "hydrogen_bonding",
    "cross_stacking",
    "coaxial_stacking",
    "debye",
)
DNA, RNA = 1, 2  # mythos/input/topology.py:44-49 NucleotideType

DEFAULTS_DIR = Path(__file__).resolve().parent.parent / "mythos_b200" / "energy" / "defaults"


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=DT)


# --------------------------------------------------------------------------- TOML defaults
def _parse_value(v):
    """mythos/input/toml.py:21-41 -- strings are sympy expressions evaluated to 32 digits."""
    if isinstance(v, str):
        try:
            return float(v)
        except ValueError:
            import sympy

            return float(sympy.parse_expr(v).evalf(n=32))
    if isinstance(v, bool):
        return v
    if isinstance(v, (int, float)):
        return float(v)
    if isinstance(v, list):
        return [_parse_value(x) for x in v]
    if isinstance(v, dict):
        return {k: _parse_value(x) for k, x in v.items()}
    return v


def load_defaults(model: str) -> dict:
    """Default constants of one model: {'energy': {term: {...}}, 'simulation': {...}}."""
    import tomllib

    with (DEFAULTS_DIR / f"{model}.toml").open("rb") as f:
        return _parse_value(tomllib.load(f))


# --------------------------------------------------------------------------- smoothing solvers
# mythos/energy/dna1/base_smoothing_functions.py:13-142
def _f1_b(x, a, x0, xc):
    e = torch.exp
    return (
        a**2
        * (-e(a * (3 * x0 + 2 * xc)) + 2 * e(a * (x + 2 * x0 + 2 * xc)) - e(a * (2 * x + x0 + 2 * xc)))
        * e(-2 * a * x)
        / (2 * e(a * (x + 2 * xc)) + e(a * (2 * x + x0)) - 2 * e(a * (2 * x + xc)) - e(a * (x0 + 2 * xc)))
    )


def _f1_xc(x, a, x0, xc):
    e = torch.exp
    return (
        (
            a * x * e(a * (x + 2 * xc))
            - a * x * e(a * (x0 + 2 * xc))
            + 2 * e(a * (x + 2 * xc))
            + e(a * (2 * x + x0))
            - 2 * e(a * (2 * x + xc))
            - e(a * (x0 + 2 * xc))
        )
        * e(-2 * a * xc)
        / (a * (e(a * x) - e(a * x0)))
    )


def f1_smoothing(x0, a, xc, x_low, x_high):
    x0, a, xc, x_low, x_high = map(_t, (x0, a, xc, x_low, x_high))
    return _f1_b(x_low, a, x0, xc), _f1_xc(x_low, a, x0, xc), _f1_b(x_high, a, x0, xc), _f1_xc(x_high, a, x0, xc)


def f2_smoothing(x0, xc, x_low, x_high):
    x0, xc, x_low, x_high = map(_t, (x0, xc, x_low, x_high))

    def b(x):
        return (x - x0) ** 2 / (2 * (x - xc) * (x - 2 * x0 + xc))

    def c(x):
        return (x * x0 - 2 * x0 * xc + xc**2) / (x - x0)

    return b(x_low), c(x_low), b(x_high), c(x_high)


def f3_smoothing(x, sigma):
    x, sigma = _t(x), _t(sigma)
    b = (
        -36
        * sigma**6
        * (-2 * sigma**6 + x**6) ** 2
        / (x**14 * (-sigma + x) * (sigma + x) * (sigma**2 - sigma * x + x**2) * (sigma**2 + sigma * x + x**2))
    )
    xc = x * (-7 * sigma**6 + 4 * x**6) / (3 * (-2 * sigma**6 + x**6))
    return b, xc


def f4_smoothing(a, x0, dstar):
    a, x0, dstar = _t(a), _t(x0), _t(dstar)
    x = x0 + dstar
    b = -(a**2) * (x - x0) ** 2 / (a * x**2 - 2 * a * x * x0 + a * x0**2 - 1)
    xc = (-a * x * x0 + a * x0**2 - 1) / (a * (-x + x0))
    return b, xc - x0


def f5_smoothing(a, xstar):
    a, x = _t(a), _t(xstar)
    x0 = 0.0
    b = -(a**2) * (x - x0) ** 2 / (a * x**2 - 2 * a * x * x0 + a * x0**2 - 1)
    xc = (a * x * x0 - a * x0**2 + 1) / (a * (x - x0))
    return b, xc


# --------------------------------------------------------------------------- primitives
def clamp(x):
    """mythos/utils/math.py:78-81"""
    return torch.where(x <= -1.0, torch.full_like(x, -1.0), torch.where(x >= 1.0, torch.full_like(x, 1.0), x))


def _safe_acos(x):
    # the reference calls arccos(clamp(x)); autograd at |x|==1 is singular there, the kernels define 0.
    return torch.acos(clamp(x))


def f1(r, r_low, r_high, r_c_low, r_c_high, eps, a, r0, r_c, b_low, b_high):
    """dna1/base_functions.py:13-37"""
    z = torch.zeros_like(r)
    morse = eps * (1 - torch.exp(-(r - r0) * a)) ** 2 - eps * (1 - torch.exp(-(r_c - r0) * a)) ** 2
    oob = torch.where(
        (r_c_low < r) & (r < r_low),
        eps * b_low * (r_c_low - r) ** 2,
        torch.where((r_high < r) & (r < r_c_high), eps * b_high * (r_c_high - r) ** 2, z),
    )
    return torch.where((r_low < r) & (r < r_high), morse, oob)


def f2(r, r_low, r_high, r_c_low, r_c_high, k, r0, r_c, b_low, b_high):
    """dna1/base_functions.py:40-63"""
    z = torch.zeros_like(r)
    harm = k / 2 * (r - r0) ** 2 - k / 2 * (r_c - r0) ** 2
    oob = torch.where(
        (r_c_low < r) & (r < r_low),
        k * b_low * (r_c_low - r) ** 2,
        torch.where((r_high < r) & (r < r_c_high), k * b_high * (r_c_high - r) ** 2, z),
    )
    return torch.where((r_low < r) & (r < r_high), harm, oob)


def f3(r, r_star, r_c, eps, sigma, b):
    """dna1/base_functions.py:66-79"""
    z = torch.zeros_like(r)
    rs = torch.where(r > 0, r, torch.ones_like(r))  # padded pairs never reach here; guard 0-division only
    lj = 4 * eps * ((sigma / rs) ** 12 - (sigma / rs) ** 6)
    oob = torch.where((r_star < r) & (r < r_c), eps * b * (r_c - r) ** 2, z)
    return torch.where(r < r_star, lj, oob)


def f4(theta, theta0, dstar, dc, a, b):
    """dna1/base_functions.py:82-107"""
    z = torch.zeros_like(theta)
    oob = torch.where(
        (theta0 - dc < theta) & (theta < theta0 - dstar),
        b * (theta0 - dc - theta) ** 2,
        torch.where((theta0 + dstar < theta) & (theta < theta0 + dc), b * (theta0 + dc - theta) ** 2, z),
    )
    return torch.where((theta0 - dstar < theta) & (theta < theta0 + dstar), 1 - a * (theta - theta0) ** 2, oob)


def f5(x, x_star, x_c, a, b):
    """dna1/base_functions.py:110-129"""
    z = torch.zeros_like(x)
    return torch.where(
        x > 0.0,
        torch.ones_like(x),
        torch.where(
            (x_star < x) & (x < 0.0),
            1 - a * x**2,
            torch.where((x_c < x) & (x < x_star), b * (x_c - x) ** 2, z),
        ),
    )


def f6(theta, a, b):
    """dna2/base_functions.py:13-17"""
    return torch.where(theta >= b, a / 2 * (theta - b) ** 2, torch.zeros_like(theta))


# --------------------------------------------------------------------------- geometry
def axes_from_quat(q):
    """mythos/energy/utils.py:18-36 -- a1 (back_base), a2 (cross_prod), a3 (base_normal)."""
    q0, q1, q2, q3 = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    a1 = torch.stack([q0**2 + q1**2 - q2**2 - q3**2, 2 * (q1 * q2 + q0 * q3), 2 * (q1 * q3 - q0 * q2)], -1)
    a2 = torch.stack([2 * (q1 * q2 - q0 * q3), q0**2 - q1**2 + q2**2 - q3**2, 2 * (q2 * q3 + q0 * q1)], -1)
    a3 = torch.stack([2 * (q1 * q3 + q0 * q2), 2 * (q2 * q3 - q0 * q1), q0**2 - q1**2 - q2**2 + q3**2], -1)
    return a1, a2, a3


def geometry_of(flavour: str) -> dict:
    """Site offsets of one nucleotide flavour from the [geometry] tables (not part of theta)."""
    if flavour == "dna1":
        g = load_defaults("dna1")["energy"]["geometry"]
        return {"back": (g["com_to_backbone"], 0.0, 0.0), "stack": g["com_to_stacking"], "base": g["com_to_hb"]}
    if flavour == "dna2":
        g = load_defaults("dna2")["energy"]["geometry"]
        return {
            "back": (g["com_to_backbone_x"], g["com_to_backbone_y"], 0.0),
            "back_stack": g["com_to_backbone_dna1"],
            "stack": g["com_to_stacking"],
            "base": g["com_to_hb"],
        }
    if flavour == "rna2":
        g = load_defaults("rna2")["energy"]["geometry"]
        # name mapping: mythos/energy/rna2/tests/test_integration.py:56-72
        return {
            "back": (g["pos_back_a1"], 0.0, g["pos_back_a3"]),
            "stack": g["pos_stack"],
            "base": g["pos_base"],
            "p3": (g["p3_x"], g["p3_y"], g["p3_z"]),
            "p5": (g["p5_x"], g["p5_y"], g["p5_z"]),
            "stack3": (g["pos_stack_3_a1"], g["pos_stack_3_a2"]),
            "stack5": (g["pos_stack_5_a1"], g["pos_stack_5_a2"]),
        }
    raise ValueError(flavour)


def sites(center, quat, geom: dict) -> dict:
    """from_rigid_body of dna1/dna2/rna2 Nucleotide (files cited in the module docstring)."""
    a1, a2, a3 = axes_from_quat(quat)
    bx, by, bz = geom["back"]
    s = {
        "a1": a1,
        "a2": a2,
        "a3": a3,
        "center": center,
        "back": center + bx * a1 + by * a2 + bz * a3,
        "stack": center + geom["stack"] * a1,
        "base": center + geom["base"] * a1,
    }
    # back site used by the stacking term: dna2 uses back_sites_dna1 (dna2/stacking.py:27-29)
    s["back_stack"] = center + geom["back_stack"] * a1 if "back_stack" in geom else s["back"]
    if "p3" in geom:
        p3, p5, s3, s5 = geom["p3"], geom["p5"], geom["stack3"], geom["stack5"]
        s["p3"] = p3[0] * a1 + p3[1] * a2 + p3[2] * a3
        s["p5"] = p5[0] * a1 + p5[1] * a2 + p5[2] * a3
        s["stack3"] = center + s3[0] * a1 + s3[1] * a2
        s["stack5"] = center + s5[0] * a1 + s5[1] * a2
    return s


def make_disp(box):
    """jax_md.space.periodic(L) / free() displacement (third party)."""
    if box is None:
        return lambda a, b: a - b
    L = _t(box)

    def disp(a, b):
        d = a - b
        return torch.remainder(d + L * 0.5, L) - L * 0.5

    return disp


def _dot(a, b):
    return (a * b).sum(-1)


def _norm(v):
    return torch.sqrt((v * v).sum(-1))


# --------------------------------------------------------------------------- dependent params
def init_fene(p):
    return dict(p)  # dna1/fene.py:26-28


def init_exc(p, unbonded: bool):
    """dna1/bonded_excluded_volume.py:56-76, dna1/unbonded_excluded_volume.py:67-97"""
    o = dict(p)
    for s in ("base", "back_base", "base_back") + (("backbone",) if unbonded else ()):
        o[f"b_{s}"], o[f"dr_c_{s}"] = f3_smoothing(p[f"dr_star_{s}"], p[f"sigma_{s}"])
    return o


def init_stacking(p, rna: bool = False):
    """dna1/stacking.py:120-183, rna2/stacking.py:107-175"""
    o = dict(p)
    kt = _t(p["kt"])
    ss = p.get("ss_stack_weights")
    ones = torch.ones(4, 4, dtype=DT)
    if ss is None:
        o["eps_stack"] = (_t(p["eps_stack_base"]) + _t(p["eps_stack_kt_coeff"]) * kt) * ones
    elif rna:
        o["eps_stack"] = _t(ss) * (1.0 + kt * _t(p["eps_stack_kt_coeff"]))
    else:
        o["eps_stack"] = _t(ss) * (1.0 - _t(p["eps_stack_kt_coeff"]) + kt * 9.0 * _t(p["eps_stack_kt_coeff"]))
    o["b_low_stack"], o["dr_c_low_stack"], o["b_high_stack"], o["dr_c_high_stack"] = f1_smoothing(
        p["dr0_stack"], p["a_stack"], p["dr_c_stack"], p["dr_low_stack"], p["dr_high_stack"]
    )
    for k in ("5", "6", "9", "10") if rna else ("4", "5", "6"):
        o[f"b_stack_{k}"], o[f"delta_theta_stack_{k}_c"] = f4_smoothing(
            p[f"a_stack_{k}"], p[f"theta0_stack_{k}"], p[f"delta_theta_star_stack_{k}"]
        )
    for k in ("1", "2"):
        o[f"b_neg_cos_phi{k}_stack"], o[f"neg_cos_phi{k}_c_stack"] = f5_smoothing(
            p[f"a_stack_{k}"], p[f"neg_cos_phi{k}_star_stack"]
        )
    return o


HB_SA = torch.tensor([[0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0], [1, 0, 0, 0]], dtype=DT)


def init_hb(p):
    """dna1/hydrogen_bonding.py:148-223"""
    o = dict(p)
    ss = p.get("ss_hb_weights")
    o["eps_hb_weights"] = HB_SA * _t(p["eps_hb"]) if ss is None else _t(ss)
    o["b_low_hb"], o["dr_c_low_hb"], o["b_high_hb"], o["dr_c_high_hb"] = f1_smoothing(
        p["dr0_hb"], p["a_hb"], p["dr_c_hb"], p["dr_low_hb"], p["dr_high_hb"]
    )
    for k in "123478":
        o[f"b_hb_{k}"], o[f"delta_theta_hb_{k}_c"] = f4_smoothing(
            p[f"a_hb_{k}"], p[f"theta0_hb_{k}"], p[f"delta_theta_star_hb_{k}"]
        )
    return o


def init_cross(p, rna: bool = False):
    """dna1/cross_stacking.py:110-183, rna2/cross_stacking.py (same without theta4)"""
    o = dict(p)
    o["b_low_cross"], o["dr_c_low_cross"], o["b_high_cross"], o["dr_c_high_cross"] = f2_smoothing(
        p["r0_cross"], p["dr_c_cross"], p["dr_low_cross"], p["dr_high_cross"]
    )
    for k in "12378" if rna else "123478":
        o[f"b_cross_{k}"], o[f"delta_theta_cross_{k}_c"] = f4_smoothing(
            p[f"a_cross_{k}"], p[f"theta0_cross_{k}"], p[f"delta_theta_star_cross_{k}"]
        )
    return o


def init_coax(p, dna2: bool):
    """dna1/coaxial_stacking.py:106-172, dna2/coaxial_stacking.py:101-130"""
    o = dict(p)
    o["b_low_coax"], o["dr_c_low_coax"], o["b_high_coax"], o["dr_c_high_coax"] = f2_smoothing(
        p["dr0_coax"], p["dr_c_coax"], p["dr_low_coax"], p["dr_high_coax"]
    )
    for k in "4156":
        o[f"b_coax_{k}"], o[f"delta_theta_coax_{k}_c"] = f4_smoothing(
            p[f"a_coax_{k}"], p[f"theta0_coax_{k}"], p[f"delta_theta_star_coax_{k}"]
        )
    if not dna2:
        o["b_cos_phi3_coax"], o["cos_phi3_c_coax"] = f5_smoothing(p["a_coax_3p"], p["cos_phi3_star_coax"])
        o["b_cos_phi4_coax"], o["cos_phi4_c_coax"] = f5_smoothing(p["a_coax_4p"], p["cos_phi4_star_coax"])
    return o


def init_debye(p):
    """dna2/debye.py:47-65"""
    o = dict(p)
    lam = _t(p["lambda_factor"]) * torch.sqrt(_t(p["kt"]) / 0.1) / torch.sqrt(_t(p["salt_conc"]))
    o["lambda_"] = lam
    o["kappa"] = 1.0 / lam
    rh = 3 * lam
    o["r_high"] = rh
    A = _t(p["prefactor_coeff"]) * _t(p["q_eff"]) ** 2
    o["prefactor"] = A
    o["smoothing_coeff"] = -(torch.exp(-rh / lam) * A * A * (rh + lam) * (rh + lam)) / (
        -4.0 * rh * rh * rh * lam * lam * A
    )
    o["r_cut"] = rh * (A * rh + 3.0 * A * lam) / (A * (rh + lam))
    return o


# --------------------------------------------------------------------------- pairwise terms
def e_fene(si, sj, bonded, p, disp):
    """dna1/fene.py:37-56 + dna1/interactions.py:16-41"""
    i, j = bonded[:, 0], bonded[:, 1]
    r = _norm(disp(si["back"][i], sj["back"][j]))
    eps, r0, delt, fmax, finf = (_t(p[k]) for k in ("eps_backbone", "r0_backbone", "delta_backbone", "fmax", "finf"))
    diff = torch.sqrt((r - r0) ** 2 + 1e-10)
    xmax = (-eps + torch.sqrt(eps**2 + 4 * fmax**2 * delt**2)) / (2 * fmax)
    fene_xmax = -(eps / 2.0) * torch.log(1.0 - xmax**2 / delt**2)
    long_xmax = (fmax - finf) * xmax * torch.log(xmax) + finf * xmax
    smoothed = (fmax - finf) * xmax * torch.log(diff) + finf * diff - long_xmax + fene_xmax
    x = (r - r0) ** 2 / delt**2
    xs = torch.where(diff > xmax, torch.zeros_like(x), x)  # keep log finite in the unselected branch
    plain = -eps / 2.0 * torch.log(1 - xs)
    return torch.where(diff > xmax, smoothed, plain)


def _exc3(r_base, r_back_base, r_base_back, p):
    """dna1/interactions.py:44-83"""
    return (
        f3(r_base, _t(p["dr_star_base"]), p["dr_c_base"], _t(p["eps_exc"]), _t(p["sigma_base"]), p["b_base"])
        + f3(
            r_back_base,
            _t(p["dr_star_back_base"]),
            p["dr_c_back_base"],
            _t(p["eps_exc"]),
            _t(p["sigma_back_base"]),
            p["b_back_base"],
        )
        + f3(
            r_base_back,
            _t(p["dr_star_base_back"]),
            p["dr_c_base_back"],
            _t(p["eps_exc"]),
            _t(p["sigma_base_back"]),
            p["b_base_back"],
        )
    )


def e_bonded_exc(si, sj, bonded, p, disp):
    """dna1/bonded_excluded_volume.py:84-114"""
    i, j = bonded[:, 0], bonded[:, 1]
    return _exc3(
        _norm(disp(si["base"][i], sj["base"][j])),
        _norm(disp(si["back"][i], sj["base"][j])),
        _norm(disp(si["base"][i], sj["back"][j])),
        p,
    )


def e_unbonded_exc(si, sj, pairs, p, disp):
    """dna1/unbonded_excluded_volume.py:105-146 + dna1/interactions.py:86-135"""
    i, j = pairs[0], pairs[1]
    r_back = _norm(disp(sj["back"][j], si["back"][i]))
    back = f3(
        r_back, _t(p["dr_star_backbone"]), p["dr_c_backbone"], _t(p["eps_exc"]), _t(p["sigma_backbone"]), p["b_backbone"]
    )
    return back + _exc3(
        _norm(disp(sj["base"][j], si["base"][i])),
        _norm(disp(si["back"][i], sj["base"][j])),
        _norm(disp(si["base"][i], sj["back"][j])),
        p,
    )


def _f4p(theta, p, fam, k):
    return f4(
        theta,
        _t(p[f"theta0_{fam}_{k}"]),
        _t(p[f"delta_theta_star_{fam}_{k}"]),
        p[f"delta_theta_{fam}_{k}_c"],
        _t(p[f"a_{fam}_{k}"]),
        p[f"b_{fam}_{k}"],
    )


def _f1_stack(r, p):
    one = torch.ones((), dtype=DT)
    return f1(
        r,
        _t(p["dr_low_stack"]),
        _t(p["dr_high_stack"]),
        p["dr_c_low_stack"],
        p["dr_c_high_stack"],
        one,
        _t(p["a_stack"]),
        _t(p["dr0_stack"]),
        _t(p["dr_c_stack"]),
        p["b_low_stack"],
        p["b_high_stack"],
    )


def _f5_stack(cosphi1, cosphi2, p):
    return f5(
        -cosphi1, _t(p["neg_cos_phi1_star_stack"]), p["neg_cos_phi1_c_stack"], _t(p["a_stack_1"]), p["b_neg_cos_phi1_stack"]
    ) * f5(
        -cosphi2, _t(p["neg_cos_phi2_star_stack"]), p["neg_cos_phi2_c_stack"], _t(p["a_stack_2"]), p["b_neg_cos_phi2_stack"]
    )


def e_stacking_dna(s, seq, bonded, p, disp):
    """dna1/stacking.py:192-289 (dna2/stacking.py:19-40 swaps in back_sites_dna1 == s['back_stack'])"""
    i, j = bonded[:, 0], bonded[:, 1]
    d_back = disp(s["back_stack"][i], s["back_stack"][j])
    r_back = _norm(d_back)
    d_st = disp(s["stack"][i], s["stack"][j])
    r_st = _norm(d_st)
    th4 = _safe_acos(_dot(s["a3"][i], s["a3"][j]))
    th5 = math.pi - _safe_acos(_dot(d_st, s["a3"][j]) / r_st)
    th6 = math.pi - _safe_acos(_dot(s["a3"][i], d_st) / r_st)
    cphi1 = -_dot(s["a2"][i], d_back) / r_back
    cphi2 = -_dot(s["a2"][j], d_back) / r_back
    v = _f1_stack(r_st, p) * _f4p(th4, p, "stack", "4") * _f4p(th5, p, "stack", "5") * _f4p(th6, p, "stack", "6")
    v = v * _f5_stack(cphi1, cphi2, p)
    return p["eps_stack"][seq[i], seq[j]] * v


def e_stacking_rna(s, seq, bonded, p, disp):
    """rna2/stacking.py:186-289 + rna2/interactions.py:14-138"""
    i, j = bonded[:, 0], bonded[:, 1]
    d_st = disp(s["stack5"][i], s["stack3"][j])
    r_st = _norm(d_st)
    th5 = math.pi - _safe_acos(_dot(d_st, s["a3"][j]) / r_st)
    th6 = math.pi - _safe_acos(_dot(s["a3"][i], d_st) / r_st)
    d_back = disp(s["back"][i], s["back"][j])
    r_back = _norm(d_back)
    th9 = _safe_acos(_dot(-s["p3"][j], d_back) / r_back)
    th10 = _safe_acos(_dot(-s["p5"][i], d_back) / r_back)
    cphi1 = -_dot(s["a2"][i], d_back) / r_back
    cphi2 = -_dot(s["a2"][j], d_back) / r_back
    v = _f1_stack(r_st, p) * _f4p(th5, p, "stack", "5") * _f4p(th6, p, "stack", "6")
    v = v * _f4p(th9, p, "stack", "9") * _f4p(th10, p, "stack", "10") * _f5_stack(cphi1, cphi2, p)
    return p["eps_stack"][seq[i], seq[j]] * v


def _hb_angles(si, sj, pairs, disp):
    """dna1/hydrogen_bonding.py:239-254 (shared verbatim by dna1/cross_stacking.py:204-217)"""
    i, j = pairs[0], pairs[1]
    d = disp(sj["base"][j], si["base"][i])
    r = _norm(d)
    rs = torch.where(r > 0, r, torch.ones_like(r))
    a1i, a1j, a3i, a3j = si["a1"][i], sj["a1"][j], si["a3"][i], sj["a3"][j]
    th1 = _safe_acos(_dot(-a1i, a1j))
    th2 = _safe_acos(_dot(-a1j, d) / rs)
    th3 = _safe_acos(_dot(a1i, d) / rs)
    th4 = _safe_acos(_dot(a3i, a3j))
    th7 = _safe_acos(_dot(-a3j, d) / rs)
    th8 = math.pi - _safe_acos(_dot(a3i, d) / rs)
    return r, th1, th2, th3, th4, th7, th8


def e_hb(si, sj, seq, pairs, p, disp):
    """dna1/hydrogen_bonding.py:232-335 + dna1/interactions.py:513-640"""
    r, th1, th2, th3, th4, th7, th8 = _hb_angles(si, sj, pairs, disp)
    one = torch.ones((), dtype=DT)
    v = f1(
        r,
        _t(p["dr_low_hb"]),
        _t(p["dr_high_hb"]),
        p["dr_c_low_hb"],
        p["dr_c_high_hb"],
        one,
        _t(p["a_hb"]),
        _t(p["dr0_hb"]),
        _t(p["dr_c_hb"]),
        p["b_low_hb"],
        p["b_high_hb"],
    )
    for th, k in ((th1, "1"), (th2, "2"), (th3, "3"), (th4, "4"), (th7, "7"), (th8, "8")):
        v = v * _f4p(th, p, "hb", k)
    return p["eps_hb_weights"][seq[pairs[0]], seq[pairs[1]]] * v


def e_cross(si, sj, pairs, p, disp, rna: bool):
    """dna1/cross_stacking.py:192-266 + dna1/interactions.py:253-385; rna2/cross_stacking.py:156-223 drops theta4"""
    r, th1, th2, th3, th4, th7, th8 = _hb_angles(si, sj, pairs, disp)
    v = f2(
        r,
        _t(p["dr_low_cross"]),
        _t(p["dr_high_cross"]),
        p["dr_c_low_cross"],
        p["dr_c_high_cross"],
        _t(p["k_cross"]),
        _t(p["r0_cross"]),
        _t(p["dr_c_cross"]),
        p["b_low_cross"],
        p["b_high_cross"],
    )
    v = v * _f4p(th1, p, "cross", "1") * _f4p(th2, p, "cross", "2") * _f4p(th3, p, "cross", "3")
    if not rna:
        v = v * (_f4p(th4, p, "cross", "4") + _f4p(math.pi - th4, p, "cross", "4"))
    v = v * (_f4p(th7, p, "cross", "7") + _f4p(math.pi - th7, p, "cross", "7"))
    v = v * (_f4p(th8, p, "cross", "8") + _f4p(math.pi - th8, p, "cross", "8"))
    return v


def e_coax(si, sj, pairs, p, disp, dna2: bool):
    """dna1/coaxial_stacking.py:181-260 + dna1/interactions.py:388-510;
    dna2/coaxial_stacking.py:138-201 + dna2/interactions.py:31-136"""
    i, j = pairs[0], pairs[1]
    d_st = disp(sj["stack"][j], si["stack"][i])
    r_st = _norm(d_st)
    rs = torch.where(r_st > 0, r_st, torch.ones_like(r_st))
    n_st = d_st / rs[:, None]
    a1i, a1j, a3i, a3j = si["a1"][i], sj["a1"][j], si["a3"][i], sj["a3"][j]
    th4 = _safe_acos(_dot(a3i, a3j))
    th1 = _safe_acos(_dot(-a1i, a1j))
    th5 = _safe_acos(_dot(a3i, n_st))
    th6 = _safe_acos(_dot(-a3j, n_st))
    v = f2(
        r_st,
        _t(p["dr_low_coax"]),
        _t(p["dr_high_coax"]),
        p["dr_c_low_coax"],
        p["dr_c_high_coax"],
        _t(p["k_coax"]),
        _t(p["dr0_coax"]),
        _t(p["dr_c_coax"]),
        p["b_low_coax"],
        p["b_high_coax"],
    )
    v = v * _f4p(th4, p, "coax", "4")
    v = v * (_f4p(th5, p, "coax", "5") + _f4p(math.pi - th5, p, "coax", "5"))
    v = v * (_f4p(th6, p, "coax", "6") + _f4p(math.pi - th6, p, "coax", "6"))
    if dna2:
        v = v * (_f4p(th1, p, "coax", "1") + f6(th1, _t(p["a_coax_1_f6"]), _t(p["b_coax_1_f6"])))
    else:
        v = v * (_f4p(th1, p, "coax", "1") + _f4p(2 * math.pi - th1, p, "coax", "1"))
        d_bb = disp(sj["back"][j], si["back"][i])
        r_bb = _norm(d_bb)
        n_bb = d_bb / torch.where(r_bb > 0, r_bb, torch.ones_like(r_bb))[:, None]
        cphi3 = _dot(n_st, torch.linalg.cross(n_bb, a1j))
        cphi4 = _dot(n_st, torch.linalg.cross(n_bb, a1i))
        v = v * f5(cphi3, _t(p["cos_phi3_star_coax"]), p["cos_phi3_c_coax"], _t(p["a_coax_3p"]), p["b_cos_phi3_coax"])
        v = v * f5(cphi4, _t(p["cos_phi4_star_coax"]), p["cos_phi4_c_coax"], _t(p["a_coax_4p"]), p["b_cos_phi4_coax"])
    return v


def e_debye(si, sj, pairs, is_end, p, disp):
    """dna2/debye.py:82-110 + dna2/interactions.py:15-28"""
    i, j = pairs[0], pairs[1]
    r = _norm(disp(sj["back"][j], si["back"][i]))
    rs = torch.where(r > 0, r, torch.ones_like(r))
    full = torch.exp(rs * -p["kappa"]) * (p["prefactor"] / rs)
    smooth = p["smoothing_coeff"] * (r - p["r_cut"]) ** 2
    e = torch.where(r < p["r_high"], full, smooth)
    e = torch.where(r < p["r_cut"], e, torch.zeros_like(e))
    if p["half_charged_ends"]:
        m = torch.where(is_end[i] > 0, 0.5, 1.0) * torch.where(is_end[j] > 0, 0.5, 1.0)
        e = e * m
    return e


# --------------------------------------------------------------------------- model assembly
def default_theta(model: str, kt=None, salt_conc=None, half_charged_ends=None, overrides=None) -> dict:
    """Independent parameters per term, as default_energy_configs builds them
    (dna1/__init__.py:27-61, dna2/__init__.py:33-71; rna2 and na1 as their integration tests do)."""
    overrides = overrides or {}
    if model == "na1":
        th = {}
        for pre, m in (("rna_", "rna2"), ("dna_", "dna2"), ("drh_", "na1")):
            for term, vals in load_defaults(m)["energy"].items():
                if term == "geometry":
                    continue
                th.setdefault(term, {}).update({pre + k: v for k, v in vals.items()})
        sim = load_defaults("dna2")["simulation"]
    else:
        d = load_defaults(model)
        th = {t: dict(v) for t, v in d["energy"].items() if t != "geometry"}
        sim = d.get("simulation") or load_defaults("dna2")["simulation"]
    kt = sim["kT"] if kt is None else kt
    th["stacking"]["kt"] = kt
    if "debye" in th:
        th["debye"]["kt"] = kt
        th["debye"]["salt_conc"] = sim.get("salt_conc", 0.5) if salt_conc is None else salt_conc
        hce = bool(sim.get("half_charged_ends", 0)) if half_charged_ends is None else half_charged_ends
        th["debye"]["half_charged_ends"] = hce
    for term, vals in overrides.items():
        th[term].update(vals)
    return th


def _strip(d, pre):
    return {k[len(pre) :]: v for k, v in d.items() if k.startswith(pre)}


def init_all(model: str, theta: dict) -> dict:
    """theta (independent) -> kernel-level parameters, per term (the init_params chain, a15)."""
    if model == "na1":
        out = {}
        shared = {k: theta["debye"][k] for k in ("kt", "salt_conc", "half_charged_ends")}
        for bank, pre in (("dna", "dna_"), ("rna", "rna_"), ("drh", "drh_")):
            o = {}
            if bank != "drh":
                o["fene"] = init_fene(_strip(theta["fene"], pre))
                o["bonded_excluded_volume"] = init_exc(_strip(theta["bonded_excluded_volume"], pre), False)
                st = _strip(theta["stacking"], pre) | {"kt": theta["stacking"]["kt"]}
                o["stacking"] = init_stacking(st, rna=(bank == "rna"))
            o["unbonded_excluded_volume"] = init_exc(_strip(theta["unbonded_excluded_volume"], pre), True)
            o["hydrogen_bonding"] = init_hb(_strip(theta["hydrogen_bonding"], pre))
            o["cross_stacking"] = init_cross(_strip(theta["cross_stacking"], pre), rna=(bank == "rna"))
            o["coaxial_stacking"] = init_coax(_strip(theta["coaxial_stacking"], pre), dna2=(bank == "dna"))
            o["debye"] = init_debye(_strip(theta["debye"], pre) | shared)
            out[bank] = o
        return out
    o = {
        "fene": init_fene(theta["fene"]),
        "bonded_excluded_volume": init_exc(theta["bonded_excluded_volume"], False),
        "stacking": init_stacking(theta["stacking"], rna=(model == "rna2")),
        "unbonded_excluded_volume": init_exc(theta["unbonded_excluded_volume"], True),
        "hydrogen_bonding": init_hb(theta["hydrogen_bonding"]),
        "cross_stacking": init_cross(theta["cross_stacking"], rna=(model == "rna2")),
        "coaxial_stacking": init_coax(theta["coaxial_stacking"], dna2=(model == "dna2")),
    }
    if model != "dna1":
        o["debye"] = init_debye(theta["debye"])
    return o


def energy_terms(
    model: str,
    center,
    quat,
    seq,
    bonded,
    pairs,
    params: dict,
    box=None,
    is_end=None,
    nt_type=None,
    stack_nt_type=None,
):
    """All 8 per-term energies of one frame -> tensor (8,) in TERMS order (0 where the model lacks a term).

    ``pairs`` is (2,U) with padding value N (masked as the reference does with ``op_i < N``).
    ``params`` is the output of :func:`init_all`.
    """
    center, quat = _t(center), _t(quat)
    seq = torch.as_tensor(seq, dtype=torch.long)
    bonded = torch.as_tensor(bonded, dtype=torch.long).reshape(-1, 2)
    pairs = torch.as_tensor(pairs, dtype=torch.long).reshape(2, -1)
    n = center.shape[0]
    valid = pairs[0] < n
    pairs = pairs[:, valid]  # identical to the reference's where(mask, v, 0) for the sum
    disp = make_disp(box)
    if is_end is None:
        is_end = torch.zeros(n, dtype=torch.long)
    is_end = torch.as_tensor(is_end, dtype=torch.long)
    z = torch.zeros((), dtype=DT)

    if model != "na1":
        s = sites(center, quat, geometry_of(model))
        st = e_stacking_rna if model == "rna2" else e_stacking_dna
        out = [
            e_fene(s, s, bonded, params["fene"], disp).sum(),
            e_bonded_exc(s, s, bonded, params["bonded_excluded_volume"], disp).sum(),
            st(s, seq, bonded, params["stacking"], disp).sum(),
            e_unbonded_exc(s, s, pairs, params["unbonded_excluded_volume"], disp).sum(),
            e_hb(s, s, seq, pairs, params["hydrogen_bonding"], disp).sum(),
            e_cross(s, s, pairs, params["cross_stacking"], disp, rna=(model == "rna2")).sum(),
            e_coax(s, s, pairs, params["coaxial_stacking"], disp, dna2=(model == "dna2")).sum(),
            e_debye(s, s, pairs, is_end, params["debye"], disp).sum() if model != "dna1" else z,
        ]
        return torch.stack(out)

    # NA1 hybrid: every variant is evaluated for every pair and selected by nt_type (a13)
    nt = torch.as_tensor(nt_type, dtype=torch.long)
    snt = nt if stack_nt_type is None else torch.as_tensor(stack_nt_type, dtype=torch.long)
    sd = sites(center, quat, geometry_of("dna2"))
    sr = sites(center, quat, geometry_of("rna2"))
    D, R, H = params["dna"], params["rna"], params["drh"]
    bi, bj = bonded[:, 0], bonded[:, 1]

    def sel_b(name, fd, fr, ntv=nt):
        rna_bond = (ntv[bi] == RNA) & (ntv[bj] == RNA)  # na1/utils.py:9-11
        return torch.where(rna_bond, fr(R[name]), fd(D[name])).sum()

    pi_, pj_ = pairs[0], pairs[1]
    is_rna = (nt[pi_] == RNA) & (nt[pj_] == RNA)
    is_drh = (nt[pi_] == DNA) & (nt[pj_] == RNA)  # na1/utils.py:14-16
    is_rdh = (nt[pj_] == DNA) & (nt[pi_] == RNA)

    def sel_u(fn_d, fn_r, fn_h):
        """fn_x(si, sj) -> (U,)"""
        dd, rr = fn_d(sd, sd), fn_r(sr, sr)
        dr, rd = fn_h(sd, sr), fn_h(sr, sd)
        return torch.where(is_rna, rr, torch.where(is_drh, dr, torch.where(is_rdh, rd, dd))).sum()

    out = [
        sel_b("fene", lambda p: e_fene(sd, sd, bonded, p, disp), lambda p: e_fene(sr, sr, bonded, p, disp)),
        sel_b(
            "bonded_excluded_volume",
            lambda p: e_bonded_exc(sd, sd, bonded, p, disp),
            lambda p: e_bonded_exc(sr, sr, bonded, p, disp),
        ),
        sel_b(
            "stacking",
            lambda p: e_stacking_dna(sd, seq, bonded, p, disp),
            lambda p: e_stacking_rna(sr, seq, bonded, p, disp),
            ntv=snt,
        ),
        sel_u(
            lambda a, b: e_unbonded_exc(a, b, pairs, D["unbonded_excluded_volume"], disp),
            lambda a, b: e_unbonded_exc(a, b, pairs, R["unbonded_excluded_volume"], disp),
            lambda a, b: e_unbonded_exc(a, b, pairs, H["unbonded_excluded_volume"], disp),
        ),
        sel_u(
            lambda a, b: e_hb(a, b, seq, pairs, D["hydrogen_bonding"], disp),
            lambda a, b: e_hb(a, b, seq, pairs, R["hydrogen_bonding"], disp),
            lambda a, b: e_hb(a, b, seq, pairs, H["hydrogen_bonding"], disp),
        ),
        sel_u(
            lambda a, b: e_cross(a, b, pairs, D["cross_stacking"], disp, rna=False),
            lambda a, b: e_cross(a, b, pairs, R["cross_stacking"], disp, rna=True),
            lambda a, b: e_cross(a, b, pairs, H["cross_stacking"], disp, rna=False),
        ),
        sel_u(
            lambda a, b: e_coax(a, b, pairs, D["coaxial_stacking"], disp, dna2=True),
            lambda a, b: e_coax(a, b, pairs, R["coaxial_stacking"], disp, dna2=False),
            lambda a, b: e_coax(a, b, pairs, H["coaxial_stacking"], disp, dna2=False),
        ),
        sel_u(
            lambda a, b: e_debye(a, b, pairs, is_end, D["debye"], disp),
            lambda a, b: e_debye(a, b, pairs, is_end, R["debye"], disp),
            lambda a, b: e_debye(a, b, pairs, is_end, H["debye"], disp),
        ),
    ]
    return torch.stack(out)


# --------------------------------------------------------------------------- topology helpers
def bonded_pairs(strand_counts, circular=None):
    """mythos/input/topology.py:166-183"""
    out, start = [], 0
    for k, n in enumerate(strand_counts):
        out += [(a, a + 1) for a in range(start, start + n - 1)]
        if circular is not None and circular[k]:
            out.append((start, start + n - 1))
        start += n
    return torch.tensor(out, dtype=torch.long).reshape(-1, 2)


def all_unbonded_pairs(n, bonded):
    """mythos/input/topology.py:186-190 -- every i<j that is not bonded, as (2,U)."""
    iu = torch.triu_indices(n, n, offset=1)
    b = torch.as_tensor(bonded, dtype=torch.long).reshape(-1, 2)
    lo, hi = torch.minimum(b[:, 0], b[:, 1]), torch.maximum(b[:, 0], b[:, 1])
    keep = torch.ones(iu.shape[1], dtype=torch.bool)
    key = iu[0] * n + iu[1]
    bkey = lo * n + hi
    keep &= ~torch.isin(key, bkey)
    return iu[:, keep]


def neighbor_pairs(center, bonded, r_cutoff, dr_threshold=0.0, box=None):
    """Pair-set semantics of mythos/utils/neighbors.py:12-59 over jax_md.partition.neighbor_list
    (OrderedSparse, disable_cell_list=True): all i<j, not bonded, with
    d2 = sum(disp(ci,cj)^2) < (r_cutoff + dr_threshold)^2 (strict).  O(N^2), oracle only."""
    center = _t(center)
    n = center.shape[0]
    cand = all_unbonded_pairs(n, bonded)
    d = make_disp(box)(center[cand[0]], center[cand[1]])
    d2 = (d * d).sum(-1)
    thr = _t(r_cutoff + dr_threshold) ** 2
    return cand[:, d2 < thr]


# --------------------------------------------------------------------------- DiffTRe pieces
def weights_and_neff(beta, new_energies, ref_energies):
    """mythos/optimization/objective.py:139-163"""
    diffs = new_energies - ref_energies
    boltz = torch.exp(-beta * diffs)
    w = boltz / boltz.sum()
    neff = torch.exp(-(w * torch.log(w)).sum())
    return w, neff / len(w)


# --------------------------------------------------------------------------- flat parameter helpers
def flatten_params(params: dict, prefix: str = "") -> dict:
    """{term:{name:tensor}} -> {"term.name": tensor} for differentiable leaves."""
    out = {}
    for k, v in params.items():
        if isinstance(v, dict):
            out.update(flatten_params(v, prefix + k + "."))
        elif isinstance(v, torch.Tensor) and v.dtype == DT:
            out[prefix + k] = v
    return out


# --------------------------------------------------------------------------- probabilistic sequences
BP_IDXS = ((0, 3), (3, 0), (2, 1), (1, 2))  # mythos/utils/constants.py:13-18: AT, TA, GC, CG over "ACGT"


def compute_seq_dep_weight(pseq, nt1, nt2, weights_table, is_unpaired, idx_to_unpaired_idx, idx_to_bp_idx):
    """mythos/energy/utils.py:45-132, case by case as the reference writes it (explicit sums, no marginals)."""
    up, bp = (torch.as_tensor(x, dtype=DT) for x in pseq)
    W = torch.as_tensor(weights_table, dtype=DT)
    u1, u2 = bool(is_unpaired[nt1]), bool(is_unpaired[nt2])
    if u1 and u2:  # case 1
        return torch.kron(up[int(idx_to_unpaired_idx[nt1])], up[int(idx_to_unpaired_idx[nt2])]) @ W.reshape(-1)
    if u1:  # case 2: nt1 unpaired, nt2 base paired
        p1 = up[int(idx_to_unpaired_idx[nt1])]
        k2, w2 = (int(x) for x in idx_to_bp_idx[nt2])
        return sum(p1[a] * bp[k2][t] * W[a, BP_IDXS[t][w2]] for a in range(4) for t in range(4))
    if u2:  # case 3
        p2 = up[int(idx_to_unpaired_idx[nt2])]
        k1, w1 = (int(x) for x in idx_to_bp_idx[nt1])
        return sum(p2[b] * bp[k1][t] * W[BP_IDXS[t][w1], b] for b in range(4) for t in range(4))
    k1, w1 = (int(x) for x in idx_to_bp_idx[nt1])
    k2, w2 = (int(x) for x in idx_to_bp_idx[nt2])
    if k1 == k2:  # case 4.I: the two members of one base pair
        return sum(bp[k1][t] * W[BP_IDXS[t][w1], BP_IDXS[t][w2]] for t in range(4))
    return sum(bp[k1][s] * bp[k2][t] * W[BP_IDXS[s][w1], BP_IDXS[t][w2]] for s in range(4) for t in range(4))  # case 4.II
