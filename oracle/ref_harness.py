"""Import shim for running the UNMODIFIED reference (mattevs24/reyna) in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``reyna_b200/`` may import this module.  It is used by
``oracle/make_golden.py`` (fixture generation) and by the ``not gpu`` tests that validate the numpy
restatement in ``oracle/dg_oracle.py`` against the reference itself.  ``/root/reference`` does not exist on
the GPU box, so everything here degrades to ``available() == False`` there.

The reference imports ``shapely`` (geometry/two_dimensional/DGFEM.py:7) and ``matplotlib``
(DGFEM/two_dimensional/_auxilliaries/boundary_information.py:4, tools.py:4-7, polymesher visualisation.py:1-2)
at module level; neither is installed in this image.  We inject minimal stand-ins:

* matplotlib / mpl_toolkits: empty modules (plotting is never called);
* shapely: ``Polygon(coords).area`` (shoelace), ``.minimum_rotated_rectangle.exterior.coords.xy`` (rotating
  calipers over hull edges) and ``Point(x, y).distance(Point)`` -- exactly the three things
  geometry/two_dimensional/DGFEM.py:134-140 uses.  The shoelace area agrees with GEOS to the last ulps, which is
  far below the 1e-12 parity tolerance; parity runs always hand ONE geometry object to both implementations
  anyway.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("REYNA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "reyna"))


class _Coords:
    def __init__(self, xy):
        self._xy = np.asarray(xy, dtype=float)

    @property
    def xy(self):
        return self._xy[:, 0].copy(), self._xy[:, 1].copy()


class _Ring:
    def __init__(self, xy):
        self.coords = _Coords(xy)


class _StubPolygon:
    def __init__(self, coords):
        self._c = np.asarray(coords, dtype=float)
        self.exterior = _Ring(np.vstack((self._c, self._c[:1])))

    @property
    def area(self) -> float:
        x, y = self._c[:, 0], self._c[:, 1]
        return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))

    @property
    def minimum_rotated_rectangle(self) -> "_StubPolygon":
        pts = self._c
        best, best_area = None, np.inf
        for k in range(pts.shape[0]):
            e = pts[(k + 1) % pts.shape[0]] - pts[k]
            ln = float(np.hypot(e[0], e[1]))
            if ln == 0.0:
                continue
            u = e / ln
            v = np.array([-u[1], u[0]])
            s, t = pts @ u, pts @ v
            a = (s.max() - s.min()) * (t.max() - t.min())
            if a < best_area:
                best_area = a
                best = np.array([s.min() * u + t.min() * v, s.max() * u + t.min() * v,
                                 s.max() * u + t.max() * v, s.min() * u + t.max() * v])
        return _StubPolygon(best)


class _StubPoint:
    def __init__(self, x, y):
        self.x, self.y = float(x), float(y)

    def distance(self, other) -> float:
        return float(np.hypot(self.x - other.x, self.y - other.y))


def _install_stubs() -> None:
    if "shapely" not in sys.modules:
        try:
            import shapely  # noqa: F401
        except Exception:
            m = types.ModuleType("shapely")
            m.Polygon, m.Point = _StubPolygon, _StubPoint
            sys.modules["shapely"] = m
    try:
        import matplotlib  # noqa: F401
    except Exception:
        names = ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.patches", "matplotlib.colors",
                 "matplotlib.collections", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d"]
        for n in names:
            if n not in sys.modules:
                sys.modules[n] = types.ModuleType(n)
        sys.modules["matplotlib.patches"].Polygon = object
        sys.modules["mpl_toolkits.mplot3d.art3d"].Poly3DCollection = object
        sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
        for parent, child in (("matplotlib", "pyplot"), ("matplotlib", "cm"), ("matplotlib", "patches"),
                              ("matplotlib", "colors"), ("matplotlib", "collections"),
                              ("mpl_toolkits", "mplot3d")):
            setattr(sys.modules[parent], child, sys.modules[parent + "." + child])
        sys.modules["mpl_toolkits.mplot3d"].art3d = sys.modules["mpl_toolkits.mplot3d.art3d"]


_ref = None


def load():
    """Return a namespace with the reference's public entry points (poly_mesher, domains, DGFEMGeometry, DGFEM)."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError("reference tree not present (expected on the GPU box); use tests/golden fixtures")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    from reyna.polymesher.two_dimensional.main import poly_mesher
    from reyna.polymesher.two_dimensional import domains
    from reyna.polymesher.two_dimensional._auxilliaries.abstraction import PolyMesh
    from reyna.geometry.two_dimensional.DGFEM import DGFEMGeometry
    from reyna.DGFEM.two_dimensional.main import DGFEM
    ns = types.SimpleNamespace(poly_mesher=poly_mesher, domains=domains, PolyMesh=PolyMesh,
                               DGFEMGeometry=DGFEMGeometry, DGFEM=DGFEM)
    _ref = ns
    return ns
