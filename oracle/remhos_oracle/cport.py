"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of the C/OpenMP port (oracle/c/remhos_stage.c).

Builds a port context from a numpy-oracle Run (same geometry, velocity samples, index maps) or
from raw arrays; used by tests/test_oracle_c.py (port vs numpy oracle) and by bench.py's
cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import fe

_HERE = os.path.dirname(os.path.abspath(__file__))
CDIR = os.path.join(os.path.dirname(_HERE), 'c')
LIB = os.path.join(CDIR, 'libremhos_oracle_c.so')
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(['make', '-C', CDIR])
        _lib = C.CDLL(LIB)
        _lib.roc_create.restype = C.c_void_p
        _lib.roc_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Port:
    def __init__(self, p, mesh_order, exec_mode, nodes, nbr_dof, lat, n_ent, vel_nodes=None,
                 vel_quad=None, vel_face=None):
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self.p = p
        self.Q = (2 * p + 3 * mesh_order - 1) // 2 + 1
        xq, wq = fe.gauss_legendre_01(self.Q)
        B = fe.bernstein(p, xq); G = fe.bernstein_deriv(p, xq)
        M1 = np.einsum('q,qi,qj->ij', wq, B, B)
        gll = fe.gauss_lobatto_01(mesh_order + 1)
        ends = np.array([0.0, 1.0])
        tabs = [B, G, np.linalg.inv(M1), wq, fe.lagrange(gll, xq), fe.lagrange_deriv(gll, xq),
                fe.lagrange(gll, ends), fe.lagrange_deriv(gll, ends)]
        tabs = [f64(t) for t in tabs]
        self._keep = [f64(nodes), f64(vel_nodes), f64(vel_quad), f64(vel_face), i32(nbr_dof), i32(lat)]
        k = self._keep
        self.ne = k[0].shape[0]
        self.nd = (p + 1) ** 3
        self.h = C.c_void_p(lib().roc_create(
            int(p), int(mesh_order), int(self.Q), int(exec_mode), C.c_int64(self.ne), _p(k[0]),
            _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]), int(n_ent), *[_p(t) for t in tabs]))
        if not self.h:
            raise RuntimeError('roc_create failed (order / quadrature beyond the compiled limits)')

    @classmethod
    def from_run(cls, run):
        """Same inputs as tests/helpers.ctx_from_oracle gives the CUDA context."""
        sp, m, topo, d = run.space, run.mesh, run.topo, run.disc
        assert m.dim == 3
        kw = {}
        if run.exec_mode == 1:
            kw['vel_nodes'] = d.Vnodes
        else:
            kw['vel_quad'] = run.vel(sp.quad_points(m.X))
            kw['vel_face'] = np.stack([run.vel(sp.face_quad_points(m.X, f)) for f in range(sp.nf)],
                                      axis=1)
        return cls(sp.p, sp.g, run.exec_mode, m.X, d.nbr, topo.lat, topo.n_ent, **kw)

    @classmethod
    def periodic_cube(cls, n, order, problem=0, mesh_order=2):
        """The bench workload built with the oracle's own mesh code only (no product library):
        periodic Cartesian n^3 hexes on [-1,1]^3 (the periodic-cube mesh of the reference refined
        to n = 3 * 2^rs per direction), velocity of `problem` at the mesh nodes.
        Returns (port, u0 [ne, nd])."""
        from . import mesh as meshmod, dg, problems
        m = meshmod.cartesian_mesh([n, n, n], [2.0, 2.0, 2.0], origin=[-1.0, -1.0, -1.0], periodic=True)
        bb_min, bb_max = meshmod.bounding_box(m)
        m = meshmod.set_curvature(m, mesh_order)
        topo = meshmod.Topology(m)
        sp = dg.Space(3, order, mesh_order)
        nbr = dg.nbr_dof_map(topo, order)
        vel = problems.velocity(problem, m.X.reshape(-1, 3), bb_min, bb_max).reshape(m.X.shape)
        port = cls(order, mesh_order, 0, m.X, nbr, topo.lat, topo.n_ent, vel_nodes=vel)
        u0 = problems.u0(problem, sp.dof_points(m.X).reshape(-1, 3), bb_min, bb_max).reshape(m.ne, sp.nd)
        return port, u0

    @classmethod
    def from_mesh_file(cls, path, rs, order, problem=0, mesh_order=2, nodal_velocity=None):
        """Port on a mesh file refined rs times, built with the oracle's own mesh code (reader,
        refinement, topology, NbrDof map) -- the element order is the one of the product's mesh module
        (tests/test_mesh.py pins the maps bit for bit).  No dense element matrices are assembled, so
        this scales to the -rs 3/4 meshes the numpy oracle is too heavy for.
        Returns (port, u0 [ne, nd], (u0_min, u0_max))."""
        from . import mesh as meshmod, dg, problems
        m = meshmod.read_mesh(path)
        for _ in range(rs):
            m = meshmod.refine_uniform(m)
        bb_min, bb_max = meshmod.bounding_box(m)
        m = meshmod.set_curvature(m, mesh_order)
        topo = meshmod.Topology(m)
        sp = dg.Space(3, order, mesh_order)
        nbr = dg.nbr_dof_map(topo, order)

        def vel(pts):
            return problems.velocity(problem, pts.reshape(-1, 3), bb_min, bb_max).reshape(pts.shape)
        if nodal_velocity is None:
            nodal_velocity = (problem % 20) in (0, 1, 2, 4, 5, 6, 7)
        kw = {}
        if nodal_velocity:
            kw['vel_nodes'] = vel(m.X)
        else:
            kw['vel_quad'] = vel(sp.quad_points(m.X))
            kw['vel_face'] = np.stack([vel(sp.face_quad_points(m.X, f)) for f in range(sp.nf)], axis=1)
        port = cls(order, mesh_order, 0, m.X, nbr, topo.lat, topo.n_ent, **kw)
        u0 = problems.u0(problem, sp.dof_points(m.X).reshape(-1, 3), bb_min, bb_max).reshape(m.ne, sp.nd)
        return port, u0

    def close(self):
        if self.h:
            lib().roc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_time(self, t):
        lib().roc_set_time(self.h, C.c_double(t))

    def lumped_mass(self):
        m = np.zeros(self.ne * self.nd)
        lib().roc_lumped_mass(self.h, _p(m))
        return m.reshape(self.ne, self.nd)

    def stage(self, dt, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        k = np.zeros_like(u)
        rc = lib().roc_stage(self.h, C.c_double(dt), _p(u), _p(k))
        assert rc == 0
        return k

    def rk3_step(self, t, dt, u):
        """in place on the contiguous float64 array u"""
        rc = lib().roc_rk3_step(self.h, C.c_double(t), C.c_double(dt), _p(u))
        assert rc == 0

    @staticmethod
    def threads():
        return lib().roc_threads()

    @staticmethod
    def set_threads(n):
        """use n OpenMP threads (torchrun exports OMP_NUM_THREADS=1 to its workers)"""
        lib().roc_set_threads(int(n))
