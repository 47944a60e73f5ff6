"""Coefficient sets of the benchmark / parity configurations (numba-compatible callables, as ``add_data`` requires).

``adr``        README.md:89-102 and tools/_benchmarks.py:128-142  (a = I, b = (1,1), c = pi^2, u = sin(pi x) sin(pi y))
``diffusion``  tests/test_DGFEM.py:492-504 without the reaction term (a = I, f = 2 pi^2 u)
``hyperbolic`` tests/test_DGFEM.py:557-565 (b = (1,1), c = pi, no diffusion; inflow = left and bottom sides)
``adr_tensor`` tests/test_DGFEM.py:612-636 (a = [[2,1],[1,2]], b = (2,1), c = 2, u = exp(1-x) exp(y-1) / 2)
``variable``   spatially varying NON-symmetric diffusion tensor, varying advection (changes sign inside the domain, so
               both upwind branches and a partial inflow boundary are exercised), varying reaction -- parity stress case,
               not from the reference.
Each entry is a dict with the ``add_data`` keyword arguments plus ``exact`` / ``grad_exact`` where a closed form is known.
"""
from __future__ import annotations

import numpy as np


def _identity_tensor(x):
    out = np.zeros((x.shape[0], 2, 2), dtype=np.float64)
    for i in range(x.shape[0]):
        out[i, 0, 0] = 1.0
        out[i, 1, 1] = 1.0
    return out


def _full_tensor(x):
    out = np.zeros((x.shape[0], 2, 2), dtype=np.float64)
    for i in range(x.shape[0]):
        out[i, 0, 0] = 2.0
        out[i, 1, 0] = 1.0
        out[i, 0, 1] = 1.0
        out[i, 1, 1] = 2.0
    return out


def _variable_tensor(x):
    out = np.zeros((x.shape[0], 2, 2), dtype=np.float64)
    for i in range(x.shape[0]):
        out[i, 0, 0] = 1.0 + x[i, 0] * x[i, 0]
        out[i, 0, 1] = 0.3 * x[i, 1]
        out[i, 1, 0] = -0.2 * x[i, 0]
        out[i, 1, 1] = 2.0 + np.sin(x[i, 1])
    return out


def _sin_sin(x):
    return np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1])


def _grad_sin_sin(x):
    ux = np.pi * np.cos(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1])
    uy = np.pi * np.sin(np.pi * x[:, 0]) * np.cos(np.pi * x[:, 1])
    return np.vstack((ux, uy)).T


def _ones_vec(x):
    return np.ones(x.shape, dtype=np.float64)


def _adr_forcing(x):
    return np.pi * (np.cos(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) +
                    np.sin(np.pi * x[:, 0]) * np.cos(np.pi * x[:, 1])) + \
        3.0 * np.pi ** 2 * np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1])


def _hyp_forcing(x):
    return np.pi * (np.cos(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) +
                    np.sin(np.pi * x[:, 0]) * np.cos(np.pi * x[:, 1]) +
                    np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]))


def _exp_solution(x):
    return 0.5 * np.exp(1.0 - x[:, 0]) * np.exp(x[:, 1] - 1)


def _grad_exp_solution(x):
    u_x = -0.5 * np.exp(1.0 - x[:, 0]) * np.exp(x[:, 1] - 1)
    u_y = 0.5 * np.exp(1.0 - x[:, 0]) * np.exp(x[:, 1] - 1)
    return np.vstack((u_x, u_y)).T


def _tensor_advection(x):
    u_x = 2.0 * np.ones(x.shape[0], dtype=np.float64)
    u_y = np.ones(x.shape[0], dtype=np.float64)
    return np.vstack((u_x, u_y)).T


def _variable_advection(x):
    u_x = 0.6 - x[:, 1]
    u_y = x[:, 0] - 0.45 + 0.1 * np.cos(3.0 * x[:, 1])
    return np.vstack((u_x, u_y)).T


PROBLEMS = {
    'adr': dict(
        data=dict(diffusion=_identity_tensor, advection=_ones_vec,
                  reaction=lambda x: np.pi ** 2 * np.ones(x.shape[0], dtype=np.float64),
                  forcing=_adr_forcing, dirichlet_bcs=_sin_sin),
        exact=_sin_sin, grad_exact=_grad_sin_sin),
    'diffusion': dict(
        data=dict(diffusion=_identity_tensor,
                  forcing=lambda x: 2.0 * np.pi ** 2 * np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]),
                  dirichlet_bcs=_sin_sin),
        exact=_sin_sin, grad_exact=_grad_sin_sin),
    'hyperbolic': dict(
        data=dict(advection=_ones_vec, reaction=lambda x: np.pi * np.ones(x.shape[0], dtype=np.float64),
                  forcing=_hyp_forcing, dirichlet_bcs=_sin_sin),
        exact=_sin_sin, grad_exact=None),
    'adr_tensor': dict(
        data=dict(diffusion=_full_tensor, advection=_tensor_advection,
                  reaction=lambda x: 2.0 * np.ones(x.shape[0], dtype=np.float64),
                  forcing=lambda x: -0.5 * np.exp(1.0 - x[:, 0]) * np.exp(x[:, 1] - 1),
                  dirichlet_bcs=_exp_solution),
        exact=_exp_solution, grad_exact=_grad_exp_solution),
    'variable': dict(
        data=dict(diffusion=_variable_tensor, advection=_variable_advection,
                  reaction=lambda x: 1.0 + x[:, 0] * x[:, 1],
                  forcing=lambda x: np.exp(x[:, 0]) * np.cos(2.0 * x[:, 1]),
                  dirichlet_bcs=lambda x: np.cos(x[:, 0]) + x[:, 1] * x[:, 1]),
        exact=None, grad_exact=None),
    'variable_hyperbolic': dict(
        data=dict(advection=_variable_advection,
                  reaction=lambda x: 1.0 + x[:, 0] * x[:, 1],
                  forcing=lambda x: np.exp(x[:, 0]) * np.cos(2.0 * x[:, 1]),
                  dirichlet_bcs=lambda x: np.cos(x[:, 0]) + x[:, 1] * x[:, 1]),
        exact=None, grad_exact=None),
}


def problem(name: str) -> dict:
    return PROBLEMS[name]
