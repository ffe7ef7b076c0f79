"""Closed-form smoothing parameters of f1..f5 (host side of theta, differentiable with torch autograd).

Same formulas as ``mythos/energy/dna1/base_smoothing_functions.py:13-142`` (the oxDNA thesis, section 2.4.1):
they turn the independent parameters into the b / x_c constants that make each truncated function C^1.
They run on a few dozen scalars per parameter update, on the host, exactly where the reference runs them
(inside ``init_params``), so that d/dtheta chains through them outside the kernels.
"""

from __future__ import annotations

import torch

DT = torch.float64


def as_t(x) -> torch.Tensor:
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=DT)


def get_f1_smoothing_params(x0, a, xc, x_low, x_high):
    """-> (b_low, xc_low, b_high, xc_high)"""
    x0, a, xc, x_low, x_high = map(as_t, (x0, a, xc, x_low, x_high))
    e = torch.exp

    def denom(x):
        return 2 * e(a * (x + 2 * xc)) + e(a * (2 * x + x0)) - 2 * e(a * (2 * x + xc)) - e(a * (x0 + 2 * xc))

    def solve_b(x):
        num = -e(a * (3 * x0 + 2 * xc)) + 2 * e(a * (x + 2 * x0 + 2 * xc)) - e(a * (2 * x + x0 + 2 * xc))
        return a**2 * num * e(-2 * a * x) / denom(x)

    def solve_xc(x):
        lead = a * x * (e(a * (x + 2 * xc)) - e(a * (x0 + 2 * xc)))
        return (lead + denom(x)) * e(-2 * a * xc) / (a * (e(a * x) - e(a * x0)))

    if x_low.dim() == 0 and x_high.dim() == 0:  # both ends in one vectorised evaluation: half the autograd nodes
        x = torch.stack([x_low, x_high])
        b, c = solve_b(x), solve_xc(x)
        return b[0], c[0], b[1], c[1]
    return solve_b(x_low), solve_xc(x_low), solve_b(x_high), solve_xc(x_high)


def get_f2_smoothing_params(x0, xc, x_low, x_high):
    """-> (b_low, xc_low, b_high, xc_high)"""
    x0, xc, x_low, x_high = map(as_t, (x0, xc, x_low, x_high))

    def solve_b(x):
        return (x - x0) ** 2 / (2 * (x - xc) * (x - 2 * x0 + xc))

    def solve_xc(x):
        return (x * x0 - 2 * x0 * xc + xc**2) / (x - x0)

    if x_low.dim() == 0 and x_high.dim() == 0:
        x = torch.stack([x_low, x_high])
        b, c = solve_b(x), solve_xc(x)
        return b[0], c[0], b[1], c[1]
    return solve_b(x_low), solve_xc(x_low), solve_b(x_high), solve_xc(x_high)


def get_f3_smoothing_params(r_star, sigma):
    """-> (b, r_c)"""
    x, s = as_t(r_star), as_t(sigma)
    s6, x6 = s**6, x**6
    poly = (x - s) * (x + s) * (s**2 - s * x + x**2) * (s**2 + s * x + x**2)  # = x^6 - s^6, kept factored
    b = -36 * s6 * (x6 - 2 * s6) ** 2 / (x**14 * poly)
    xc = x * (4 * x6 - 7 * s6) / (3 * (x6 - 2 * s6))
    return b, xc


def get_f4_smoothing_params(a, x0, delta_x_star):
    """-> (b, delta_x_c)"""
    a, x0, d = as_t(a), as_t(x0), as_t(delta_x_star)
    x = x0 + d
    b = -(a**2) * (x - x0) ** 2 / (a * x**2 - 2 * a * x * x0 + a * x0**2 - 1)
    xc = (-a * x * x0 + a * x0**2 - 1) / (a * (x0 - x))
    return b, xc - x0


def get_f5_smoothing_params(a, x_star):
    """-> (b, x_c)"""
    a, x = as_t(a), as_t(x_star)
    b = -(a**2) * x**2 / (a * x**2 - 1)
    xc = 1 / (a * x)
    return b, xc
