"""Boundary classification at edge midpoints -- host side, already vectorised in the reference.

Mirrors reyna/DGFEM/two_dimensional/_auxilliaries/boundary_information.py:32-93: a boundary edge is *elliptic* when
``n . a(mid) . n > tolerence`` and *inflow* when ``b(mid) . n <= tolerence`` (``tolerence = 1e-12``).  Plotting
(:95-131) is out of scope (matplotlib is not a dependency of the hot path).
"""
from __future__ import annotations

import typing

import numpy as np


class BoundaryInformation:

    def __init__(self, geometry, **kwargs):
        self.geometry = geometry
        self.tolerence = 1e-12
        if 'tolerence' in kwargs:
            self.tolerence = kwargs.pop('tolerence')
        self.elliptical_indecies: typing.Optional[np.ndarray] = None
        self.inflow_indecies: typing.Optional[np.ndarray] = None
        self.generated_boundary_information: bool = False

    def split_boundaries(self, advection, diffusion) -> None:
        g = self.geometry
        be = np.asarray(g.boundary_edges).reshape(-1, 2)
        bdmids = np.ascontiguousarray(0.5 * (g.nodes[be[:, 0], :] + g.nodes[be[:, 1], :]))
        normals = np.asarray(g.boundary_normals).reshape(-1, 2)
        if diffusion is not None:
            hyp = np.einsum('ni,nij,nj->n', normals, diffusion(bdmids), normals) <= self.tolerence
            self.elliptical_indecies = np.flatnonzero(~hyp)
        if advection is not None:
            inflow = np.sum(advection(bdmids) * normals, 1) <= self.tolerence
            self.inflow_indecies = np.flatnonzero(inflow)
        self.generated_boundary_information = True

    def plot_boundaries(self) -> None:
        if not self.generated_boundary_information:
            raise ValueError('The boundary information has not been generated: please run the '
                             '"split_boundaries" method.')
        raise NotImplementedError('plotting is outside the reyna_b200 hot path (needs matplotlib)')
