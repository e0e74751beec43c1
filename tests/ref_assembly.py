"""Independent numpy/scipy statement of the reference's adjoint MATRIX assembly, used only to
cross-check the oracle's back-substitution against a sparse LU (the reference solves with
Eigen::SparseLU, deps/CustomOps/Eikonal3D/Eikonal3D.cpp:186-193; Eikonal.h:187-195).
Pure-python loops: small grids only."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl


def assemble3d(u, u0):
    """Triplet rules of Eikonal3D.cpp:118-184.  Returns (A as csc, g_mask) where g_mask zeroes g on Z."""
    m, n, l = u.shape
    N = m * n * l
    idf = lambda i, j, k: (i * n + j) * l + k
    rows, cols, vals = [], [], []
    zero = set()
    for i in range(m):
        for j in range(n):
            for k in range(l):
                t = idf(i, j, k)
                if u[i, j, k] == u0[i, j, k]:
                    zero.add(t)
                    continue
                nb = []
                for ax, (c, lim) in enumerate(((i, m), (j, n), (k, l))):
                    def at(d):
                        p = [i, j, k]
                        p[ax] += d
                        return tuple(p)
                    if c == 0:
                        pid = at(+1)
                    elif c == lim - 1:
                        pid = at(-1)
                    else:
                        pid = at(-1) if u[at(+1)] > u[at(-1)] else at(+1)
                    nb.append(pid)
                nz = False
                for pid in nb:
                    a = u[pid]
                    if u[i, j, k] > a:
                        nz = True
                        rows += [t, t]
                        cols += [t, idf(*pid)]
                        vals += [2.0 * (u[i, j, k] - a), -2.0 * (u[i, j, k] - a)]
                if not nz:
                    zero.add(t)
    rows = np.array(rows, dtype=np.int64)
    cols = np.array(cols, dtype=np.int64)
    vals = np.array(vals, dtype=np.float64)
    if zero:
        z = np.zeros(N, dtype=bool)
        z[list(zero)] = True
        kill = z[rows] | z[cols]
        vals = np.where(kill, 0.0, vals)
        zl = np.array(sorted(zero), dtype=np.int64)
        rows = np.concatenate([rows, zl])
        cols = np.concatenate([cols, zl])
        vals = np.concatenate([vals, np.ones(len(zl))])
    A = sp.coo_matrix((vals, (rows, cols)), shape=(N, N)).tocsc()
    gmask = np.ones(N)
    if zero:
        gmask[list(zero)] = 0.0
    return A, gmask


def backward3d_lu(grad_u, u, u0, f, h):
    A, gmask = assemble3d(u, u0)
    g = grad_u.ravel() * gmask
    x = spl.splu(A.T.tocsc()).solve(g)
    rhs = -2 * f.ravel() * h * h
    grad_f = (-x * rhs).reshape(u.shape)
    grad_u0 = np.where(u == u0, grad_u, 0.0)
    return grad_u0, grad_f


def assemble2d(u, ix, jx):
    """Eikonal.h:106-185; u is (n+1, m+1) [row j, col i]."""
    n1, m1 = u.shape
    m, n = m1 - 1, n1 - 1
    N = n1 * m1
    rows, cols, vals = [], [], []
    for j in range(n1):
        for i in range(m1):
            t = j * m1 + i
            if i == ix and j == jx:
                rows.append(t); cols.append(t); vals.append(1.0)
                continue
            if i == 0:
                p = (j, 1)
            elif i == m:
                p = (j, m - 1)
            else:
                p = (j, i - 1) if u[j, i + 1] > u[j, i - 1] else (j, i + 1)
            if u[j, i] > u[p]:
                rows += [t, t]; cols += [t, p[0] * m1 + p[1]]
                vals += [2 * (u[j, i] - u[p]), 2 * (u[p] - u[j, i])]
            if j == 0:
                p = (1, i)
            elif j == n:
                p = (n - 1, i)
            else:
                p = (j - 1, i) if u[j + 1, i] > u[j - 1, i] else (j + 1, i)
            if u[j, i] > u[p]:
                rows += [t, t]; cols += [t, p[0] * m1 + p[1]]
                vals += [2 * (u[j, i] - u[p]), 2 * (u[p] - u[j, i])]
    return sp.coo_matrix((vals, (rows, cols)), shape=(N, N)).tocsc()


def backward2d_lu(grad_u, u, f, h, ix, jx):
    A = assemble2d(u, ix, jx)
    x = spl.splu(A.T.tocsc()).solve(grad_u.ravel().astype(np.float64))
    dFdf = -2 * f.ravel() * h * h
    dFdf[jx * u.shape[1] + ix] = 0.0
    return (-x * dFdf).reshape(u.shape)
