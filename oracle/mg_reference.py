"""scipy restatement of the multigrid preconditioner of reyna_b200/csrc/multigrid.cu (TEST INFRASTRUCTURE ONLY).

The reference has no iterative solver (it calls SuperLU, main.py:231); this file exists so that the CUDA hierarchy and
cycle can be checked operator-for-operator against an independent implementation of the same algorithm."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def block_inverse(B: sp.csr_matrix, d: int) -> sp.csr_matrix:
    E = B.shape[0] // d
    Bd = B.tobsr(blocksize=(d, d))
    inv = np.zeros((E, d, d))
    for e in range(E):
        for k in range(Bd.indptr[e], Bd.indptr[e + 1]):
            if Bd.indices[k] == e:
                inv[e] = np.linalg.inv(Bd.data[k])
                break
    return sp.bsr_matrix((inv, np.arange(E), np.arange(E + 1)), shape=B.shape).tocsr()


class CoarseCorrection:
    """z = P1^T V1(P1 r) (p >= 2) or z = P0^T V0(P0 r) (p = 1): the multigrid part of the preconditioner."""

    def __init__(self, B: sp.csr_matrix, p: int, nu: int = 3, omega: float = 0.7, alpha: float = 1.8, dense_max: int = 256):
        d = (p + 1) * (p + 2) // 2
        N = B.shape[0]
        E = N // d
        self.nu, self.omega, self.alpha = nu, omega, alpha
        self.has_p1 = d > 3
        if self.has_p1:
            sel = np.array([0, 1, p + 1])
            self.R1 = sp.csr_matrix((np.ones(3 * E), ((np.arange(E)[:, None] * d + sel[None]).ravel(), np.arange(3 * E))),
                                    shape=(N, 3 * E))
            self.A1 = (self.R1.T @ B @ self.R1).tocsr()
            self.M1 = block_inverse(self.A1, 3)
            self.R10 = sp.csr_matrix((np.ones(E), (np.arange(E) * 3, np.arange(E))), shape=(3 * E, E))
            A0 = (self.R10.T @ self.A1 @ self.R10).tocsr()
        else:
            self.R0 = sp.csr_matrix((np.ones(E), (np.arange(E) * d, np.arange(E))), shape=(N, E))
            A0 = (self.R0.T @ B @ self.R0).tocsr()
        self.levels = []
        A = A0
        while True:
            n = A.shape[0]
            if n <= dense_max:
                self.levels.append(dict(A=A, inv=np.linalg.inv(A.toarray())))
                break
            nc = (n + 3) // 4
            P = sp.csr_matrix((np.ones(n), (np.arange(n), np.arange(n) // 4)), shape=(n, nc))
            self.levels.append(dict(A=A.tocsr(), dinv=1.0 / A.diagonal(), P=P))
            A = (P.T @ A @ P).tocsr()

    def vcycle(self, k: int, b: np.ndarray) -> np.ndarray:
        L = self.levels[k]
        if "inv" in L:
            return L["inv"] @ b
        A, dinv, P = L["A"], L["dinv"], L["P"]
        x = self.omega * dinv * b
        for _ in range(1, self.nu):
            x = x + self.omega * dinv * (b - A @ x)
        x = x + self.alpha * (P @ self.vcycle(k + 1, P.T @ (b - A @ x)))
        for _ in range(self.nu):
            x = x + self.omega * dinv * (b - A @ x)
        return x

    def apply(self, r: np.ndarray) -> np.ndarray:
        if self.has_p1:
            r1 = self.R1.T @ r
            x1 = self.M1 @ r1
            x1 = x1 + self.R10 @ self.vcycle(0, self.R10.T @ (r1 - self.A1 @ x1))
            x1 = x1 + self.M1 @ (r1 - self.A1 @ x1)
            return self.R1 @ x1
        return self.R0 @ self.vcycle(0, self.R0.T @ r)

    def level_sizes(self):
        return [L["A"].shape[0] for L in self.levels]


class AgglomeratedP1:
    """z = R1^T V(R1 r) (p >= 2) / z = alpha P V(P^T r) (p = 1): the agglomerated-P1 chain of multigrid.cu (rb_mg_opts.elem_bbox
    given).  Every level keeps the modes (0,0), (0,1), (1,0) per (agglomerated) element; an aggregate is 4 consecutive
    elements and its basis the scaled Legendre polynomials of degree <= 1 on the bounding box of the group."""

    C0, C1 = 1.0 / np.sqrt(2.0), np.sqrt(1.5)

    def __init__(self, B: sp.csr_matrix, p: int, bbox: np.ndarray, nu: int = 2, omega: float = 0.8, alpha: float = 2.2,
                 dense_max: int = 256):
        d = (p + 1) * (p + 2) // 2
        N = B.shape[0]
        E = N // d
        self.nu, self.omega, self.alpha, self.p = nu, omega, alpha, p
        self.has_p1 = d > 3
        if self.has_p1:
            sel = np.array([0, 1, p + 1])
            self.R1 = sp.csr_matrix((np.ones(3 * E), ((np.arange(E)[:, None] * d + sel[None]).ravel(), np.arange(3 * E))),
                                    shape=(N, 3 * E))
            A = (self.R1.T @ B @ self.R1).tocsr()
        else:
            A = B.tocsr()
        bb = np.asarray(bbox, dtype=float).reshape(-1, 4)
        self.levels = []
        while True:
            n = A.shape[0]
            if n <= dense_max:
                self.levels.append(dict(A=A, inv=np.linalg.inv(A.toarray())))
                break
            P, bb_c = self._prolongation(bb)
            self.levels.append(dict(A=A, M=block_inverse(A, 3), P=P))
            A = (P.T @ A @ P).tocsr()
            bb = bb_c

    @classmethod
    def _coefs(cls, bb):
        hx, hy = 0.5 * (bb[:, 1] - bb[:, 0]), 0.5 * (bb[:, 3] - bb[:, 2])
        mx, my = 0.5 * (bb[:, 1] + bb[:, 0]), 0.5 * (bb[:, 3] + bb[:, 2])
        r = 1.0 / np.sqrt(hx * hy)
        return mx, my, cls.C0 * cls.C0 * r, cls.C0 * cls.C1 * r / hx, cls.C0 * cls.C1 * r / hy

    @classmethod
    def _prolongation(cls, bb_f):
        n = bb_f.shape[0]
        nc = (n + 3) // 4
        gid = np.arange(n) // 4
        starts = np.arange(0, n, 4)
        bb_c = np.stack([np.minimum.reduceat(bb_f[:, 0], starts), np.maximum.reduceat(bb_f[:, 1], starts),
                         np.minimum.reduceat(bb_f[:, 2], starts), np.maximum.reduceat(bb_f[:, 3], starts)], axis=1)
        mxf, myf, k00f, kxf, kyf = cls._coefs(bb_f)
        mxc, myc, k00c, kxc, kyc = [a[gid] for a in cls._coefs(bb_c)]
        f = np.arange(n)
        rows = np.concatenate([3 * f, 3 * f + 1, 3 * f, 3 * f + 2, 3 * f])
        cols = np.concatenate([3 * gid, 3 * gid + 1, 3 * gid + 1, 3 * gid + 2, 3 * gid + 2])
        vals = np.concatenate([k00c / k00f, kyc / kyf, kyc * (myf - myc) / k00f, kxc / kxf, kxc * (mxf - mxc) / k00f])
        return sp.csr_matrix((vals, (rows, cols)), shape=(3 * n, 3 * nc)), bb_c

    def vcycle(self, k: int, b: np.ndarray) -> np.ndarray:
        L = self.levels[k]
        if "inv" in L:
            return L["inv"] @ b
        A, M, P = L["A"], L["M"], L["P"]
        x = self.omega * (M @ b)
        for _ in range(1, self.nu):
            x = x + self.omega * (M @ (b - A @ x))
        x = x + self.alpha * (P @ self.vcycle(k + 1, P.T @ (b - A @ x)))
        for _ in range(self.nu):
            x = x + self.omega * (M @ (b - A @ x))
        return x

    def apply(self, r: np.ndarray) -> np.ndarray:
        if self.has_p1:
            return self.R1 @ self.vcycle(0, self.R1.T @ r)
        L0 = self.levels[0]
        if "inv" in L0:
            return L0["inv"] @ r
        return self.alpha * (L0["P"] @ self.vcycle(1, L0["P"].T @ r))

    def level_sizes(self):
        return [L["A"].shape[0] for L in self.levels]
