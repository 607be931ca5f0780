"""Host-side set-up of a Remhos run on top of the C ABI (what remhos.cpp:438-1121 does before
the time loop): mesh load/refine/curvature, CFL dt, remap mesh velocity, DofInfo maps, device
context, projected initial condition.  All arithmetic is done by librmh_b200 (C++); this module
only sequences the calls.  Used by bench.py, __graft_entry__.smoke() and the tests.
"""
import ctypes as C
import numpy as np

from . import capi
from .capi import lib, check, _ptr


def _f64(n):
    return np.zeros(n, dtype=np.float64)


def mesh_eval(mesh, pts1d, face=-1):
    pts1d = np.ascontiguousarray(pts1d, dtype=np.float64)
    d = mesh.dim
    npts = pts1d.size ** (d if face < 0 else d - 1)
    out = np.zeros((mesh.ne, npts, d))
    check(lib().rmh_mesh_eval(mesh.h, int(pts1d.size), _ptr(pts1d), int(face), _ptr(out)))
    return out


def velocity(problem, x, bb_min, bb_max):
    x = np.ascontiguousarray(x, dtype=np.float64)
    dim = x.shape[-1]
    v = np.zeros_like(x)
    check(lib().rmh_velocity(int(problem), dim, C.c_int64(x.size // dim), _ptr(x), _ptr(bb_min),
                             _ptr(bb_max), _ptr(v)))
    return v


def u0(problem, x, bb_min, bb_max):
    x = np.ascontiguousarray(x, dtype=np.float64)
    dim = x.shape[-1]
    u = np.zeros(x.size // dim)
    check(lib().rmh_u0(int(problem), dim, C.c_int64(u.size), _ptr(x), _ptr(bb_min), _ptr(bb_max),
                       _ptr(u)))
    return u


class Problem:
    """mesh: a capi.Mesh already refined.  Mirrors the reference defaults (remhos.cpp:216-244)."""

    def __init__(self, mesh, problem=0, order=3, mesh_order=2, bounds_type=0, dt=0.005,
                 t_final=4.0, device=0, velocity_samples='auto', create_ctx=True):
        self.mesh = mesh
        self.problem = problem
        self.order = order
        self.exec_mode = 0 if problem < 10 else 1                     # remhos.cpp:438-440
        self.bb_min, self.bb_max = mesh.bounding_box()                # :457 (before SetCurvature)
        mesh.set_curvature(mesh_order)                                # :513
        dim = mesh.dim
        if dt < 0.0:                                                  # :538-553
            d = C.c_double(0.0)
            check(lib().rmh_cfl_dt(mesh.h, int(problem), _ptr(self.bb_min), _ptr(self.bb_max),
                                   C.byref(d)))
            dt = d.value
        self.dt = dt
        nodes = mesh.nodes()
        kw = {}
        if self.exec_mode == 1:                                       # :562-584
            v = np.zeros_like(nodes)
            check(lib().rmh_remap_mesh_velocity(mesh.h, int(problem), _ptr(self.bb_min),
                                                _ptr(self.bb_max), C.c_double(dt),
                                                C.c_double(t_final), _ptr(v)))
            kw['vel_nodes'] = v
            t_final = 1.0                                             # :1128-1134
        else:
            # velocities that are polynomials of degree <= 1 in x are reproduced exactly by their
            # nodal interpolant on the mesh nodes; anything else is sampled at the quadrature points
            nodal_ok = (problem % 20) in (0, 1, 2, 4, 5, 6, 7)
            if velocity_samples == 'nodes' or (velocity_samples == 'auto' and nodal_ok):
                kw['vel_nodes'] = velocity(problem, nodes, self.bb_min, self.bb_max)
            else:
                Q = (2 * order + dim * mesh_order - 1) // 2 + 1
                xq, _ = np.polynomial.legendre.leggauss(Q)
                xq = 0.5 * (xq + 1.0)
                kw['vel_quad'] = velocity(problem, mesh_eval(mesh, xq), self.bb_min, self.bb_max)
                kw['vel_face'] = np.stack(
                    [velocity(problem, mesh_eval(mesh, xq, f), self.bb_min, self.bb_max)
                     for f in range(2 * dim)], axis=1)
        self.t_final = t_final
        maps = mesh.dof_maps(order)
        self.maps = maps
        lat_pts = np.arange(order + 1) / max(order, 1)
        xdof = mesh_eval(mesh, lat_pts)
        infl = np.zeros(xdof.shape[0] * xdof.shape[1])
        check(lib().rmh_inflow_project(mesh.h, int(problem), int(order), _ptr(infl)))   # remhos.cpp:625-636
        # the host-side inputs of the stage path (also what bench.py hands to the CPU port)
        self.inputs = dict(dim=dim, order=order, mesh_order=mesh_order, exec_mode=self.exec_mode,
                           bounds_type=bounds_type, nodes=nodes, nbr_dof=maps['nbr_dof'],
                           lat=maps['lat'], n_ent=maps['n_ent'], nbr_elem=maps['nbr_elem'],
                           inflow=infl, **kw)
        self.ctx = capi.Context(device=device, **self.inputs) if create_ctx else None
        self.u0 = u0(problem, xdof, self.bb_min, self.bb_max)         # :883
        self.ne = mesh.ne
        self.nd = maps['nd']

    def close(self):
        if self.ctx is not None:
            self.ctx.close()
