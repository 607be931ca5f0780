"""TEST INFRASTRUCTURE ONLY -- CPU oracle (numpy) for the Remhos RK-stage hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product path (remhos_b200/) never does.

1-D finite-element tables: Gauss-Legendre rules, Gauss-Lobatto nodes, Bernstein
(positive) basis and Lagrange basis, all on the reference interval [0,1].

MFEM semantics restated (MFEM itself is not in /root/reference):
  * DG_FECollection(p, dim, BasisType::Positive) (remhos.cpp:588-590) is the tensor
    Bernstein basis B_i^p(x) = C(p,i) x^i (1-x)^(p-i), lexicographic DOFs, x fastest.
  * IntRules.Get(Segment, order) is Gauss-Legendre with n = order/2 + 1 points.
  * mesh nodes are a degree-`mesh_order` Gauss-Lobatto nodal field (remhos.cpp:510-527).
"""
import numpy as np
from math import comb


def gauss_legendre_01(n):
    """n-point Gauss-Legendre rule on [0,1] (points ascending, weights sum to 1)."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_lobatto_01(n):
    """n Gauss-Lobatto points on [0,1] (n >= 2)."""
    if n == 2:
        return np.array([0.0, 1.0])
    if n == 3:
        return np.array([0.0, 0.5, 1.0])
    # interior points: roots of P'_{n-1}
    c = np.zeros(n)
    c[-1] = 1.0
    dc = np.polynomial.legendre.legder(c)
    r = np.sort(np.polynomial.legendre.legroots(dc))
    x = np.concatenate(([-1.0], r, [1.0]))
    return 0.5 * (x + 1.0)


def bernstein(p, x):
    """B[q, i] = B_i^p(x_q)."""
    x = np.asarray(x, dtype=np.float64)
    B = np.empty((x.size, p + 1))
    for i in range(p + 1):
        B[:, i] = comb(p, i) * x ** i * (1.0 - x) ** (p - i)
    return B


def bernstein_deriv(p, x):
    """G[q, i] = d/dx B_i^p(x_q) = p (B_{i-1}^{p-1} - B_i^{p-1})."""
    x = np.asarray(x, dtype=np.float64)
    G = np.zeros((x.size, p + 1))
    if p == 0:
        return G
    Bm = bernstein(p - 1, x)
    for i in range(p + 1):
        lo = Bm[:, i - 1] if i >= 1 else 0.0
        hi = Bm[:, i] if i <= p - 1 else 0.0
        G[:, i] = p * (lo - hi)
    return G


def lagrange(nodes, x):
    """L[q, i] = l_i(x_q) for the Lagrange basis on `nodes`."""
    nodes = np.asarray(nodes, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = nodes.size
    L = np.ones((x.size, n))
    for i in range(n):
        for j in range(n):
            if j != i:
                L[:, i] *= (x - nodes[j]) / (nodes[i] - nodes[j])
    return L


def lagrange_deriv(nodes, x):
    """dL[q, i] = l_i'(x_q)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = nodes.size
    dL = np.zeros((x.size, n))
    for i in range(n):
        for m in range(n):
            if m == i:
                continue
            term = np.ones(x.size) / (nodes[i] - nodes[m])
            for j in range(n):
                if j != i and j != m:
                    term *= (x - nodes[j]) / (nodes[i] - nodes[j])
            dL[:, i] += term
    return dL


def tensor_basis(mats):
    """Kronecker tensor of 1-D matrices, lexicographic with the FIRST matrix fastest.

    mats = [A_x, A_y(, A_z)], each [nq_d, nd_d]; returns [prod nq, prod nd] where the
    flat index is x + nx*(y + ny*z) on both sides."""
    out = mats[0]
    for A in mats[1:]:
        out = np.kron(A, out)
    return out
