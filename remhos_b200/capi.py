"""ctypes binding of librmh_b200.so (the C ABI declared in include/remhos_b200.h).

Thin marshalling only: numpy arrays for host data, raw device pointers (ints) for device
vectors -- torch tensors are passed as tensor.data_ptr().  Raises RmhError with the library's
message on any nonzero status.  There is no CPU fallback: importing works without a GPU (so
the mesh module and symbol checks run anywhere) but every rmh_ctx_* call needs a CUDA device.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'librmh_b200.so')


class RmhError(RuntimeError):
    pass


class Desc(C.Structure):
    _fields_ = [
        ('dim', C.c_int32), ('order', C.c_int32), ('mesh_order', C.c_int32),
        ('exec_mode', C.c_int32), ('bounds_type', C.c_int32), ('device', C.c_int32),
        ('ne', C.c_int64), ('ne_ghost', C.c_int64),
        ('nodes', C.c_void_p), ('vel_nodes', C.c_void_p), ('vel_quad', C.c_void_p),
        ('vel_face', C.c_void_p), ('nbr_dof', C.c_void_p), ('lat', C.c_void_p),
        ('n_ent', C.c_int32), ('nbr_elem', C.c_void_p), ('inflow', C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RmhError('librmh_b200.so is not built: run __graft_entry__.build() '
                           '(make -C remhos_b200/csrc)')
        _lib = C.CDLL(LIB_PATH)
        _lib.rmh_last_error.restype = C.c_char_p
        _lib.rmh_mesh_nodes.restype = C.POINTER(C.c_double)
        _lib.rmh_mesh_elem_vertices.restype = C.POINTER(C.c_int64)
        _lib.rmh_ctx_ndofs.restype = C.c_int64
        _lib.rmh_launch_count.restype = C.c_int64
        _lib.rmh_dplan_blob_bytes.restype = C.c_int64
        _lib.rmh_dist_blob_bytes.restype = C.c_int64
    return _lib


def check(status):
    if status != 0:
        raise RmhError(lib().rmh_last_error().decode())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dp(x):
    """device pointer from an int / torch tensor / None"""
    if x is None:
        return C.c_void_p(0)
    if hasattr(x, 'data_ptr'):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


class Mesh:
    """Host mesh handle (rmh_mesh_*)."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def load(cls, path):
        h = C.c_void_p()
        check(lib().rmh_mesh_load(path.encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def cartesian(cls, n, size, origin=None, periodic=False):
        dim = len(n)
        na = (C.c_int * dim)(*n)
        sa = (C.c_double * dim)(*size)
        oa = (C.c_double * dim)(*(origin if origin is not None else [0.0] * dim))
        h = C.c_void_p()
        check(lib().rmh_mesh_cartesian(dim, na, oa, sa, int(periodic), C.byref(h)))
        return cls(h)

    def __del__(self):
        try:
            if self.h:
                lib().rmh_mesh_free(self.h)
                self.h = None
        except Exception:
            pass

    def refine(self, levels):
        check(lib().rmh_mesh_refine(self.h, int(levels)))
        return self

    def set_curvature(self, order):
        check(lib().rmh_mesh_set_curvature(self.h, int(order)))
        return self

    @property
    def dim(self):
        return lib().rmh_mesh_dim(self.h)

    @property
    def ne(self):
        return lib().rmh_mesh_ne(self.h)

    @property
    def nv(self):
        return lib().rmh_mesh_nv(self.h)

    @property
    def geom_order(self):
        return lib().rmh_mesh_geom_order(self.h)

    def bounding_box(self):
        d = self.dim
        lo = np.zeros(d); hi = np.zeros(d)
        check(lib().rmh_mesh_bounding_box(self.h, _ptr(lo), _ptr(hi)))
        return lo, hi

    def nodes(self):
        d, g = self.dim, self.geom_order
        n = self.ne * (g + 1) ** d * d
        p = lib().rmh_mesh_nodes(self.h)
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(self.ne, (g + 1) ** d, d).copy()

    def elem_vertices(self):
        d = self.dim
        p = lib().rmh_mesh_elem_vertices(self.h)
        return np.ctypeslib.as_array(p, shape=(self.ne * 2 ** d,)).reshape(self.ne, 2 ** d).copy()

    def extract(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        h = C.c_void_p()
        check(lib().rmh_mesh_extract(self.h, C.c_int64(ids.size), _ptr(ids), C.byref(h)))
        return Mesh(h)

    def elem_sizes(self):
        """Mesh::GetElementSize(e) = |det J(centre)|^(1/dim) per element"""
        h = np.empty(self.ne, dtype=np.float64)
        check(lib().rmh_mesh_elem_sizes(self.h, _ptr(h)))
        return h

    def make_refined(self, factor, nodes=None):
        """Mesh::MakeRefined(mesh, factor, ClosedUniform): factor^dim linear sub-elements per element
        (the subcell mesh, meshLO_*.mesh); nodes: moved nodes like Mesh.nodes(), default the mesh's own"""
        h = C.c_void_p()
        if nodes is not None:
            nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        check(lib().rmh_mesh_make_refined(self.h, int(factor), _ptr(nodes) if nodes is not None else None,
                                          C.byref(h)))
        return Mesh(h)

    def partition(self, nparts):
        part = np.zeros(self.ne, dtype=np.int32)
        check(lib().rmh_mesh_partition(self.h, int(nparts), _ptr(part)))
        return part

    def halo(self, part, rank, interior_first=False):
        return Halo(self, part, rank, interior_first)

    def nbr_lattice(self, order=1):
        """(nbr [ne, 3^dim], structured) of rmh_nbr_lattice on this mesh's lattice-entity map"""
        maps = self.dof_maps(order)
        lat = np.ascontiguousarray(maps['lat'], dtype=np.int32)
        nbr = np.zeros_like(lat)
        ok = C.c_int(0)
        check(lib().rmh_nbr_lattice(self.dim, C.c_int64(lat.shape[0]), C.c_int32(maps['n_ent']), _ptr(lat),
                                    _ptr(nbr), C.byref(ok)))
        return nbr, bool(ok.value)

    def dof_maps(self, order):
        d, ne, p = self.dim, self.ne, order
        nf, nfd, nd = 2 * d, (p + 1) ** (d - 1), (p + 1) ** d
        bd = np.zeros((nfd, nf), dtype=np.int32)
        nbr = np.zeros((ne, nf, nfd), dtype=np.int32)
        s2i = np.zeros((max(p, 1) ** d, 2 ** d), dtype=np.int32)
        lat = np.zeros((ne, 3 ** d), dtype=np.int32)
        n_ent = C.c_int32(0)
        nbe = np.zeros((ne, nf), dtype=np.int32)
        check(lib().rmh_mesh_dof_maps(self.h, int(p), _ptr(bd), _ptr(nbr), _ptr(s2i), _ptr(lat),
                                      C.byref(n_ent), _ptr(nbe)))
        return dict(bdr_dofs=bd, nbr_dof=nbr, sub2ind=s2i, lat=lat, n_ent=n_ent.value,
                    nbr_elem=nbe, nd=nd)


class Halo:
    """Halo plan of one rank (rmh_halo_*): owned / ghost global element ids, peers, per-peer
    send lists (local owned indices) and receive offsets into the ghost ordering.
    interior_first=True orders the owned elements interior first (rmh_halo_interior_first)."""

    def __init__(self, mesh, part, rank, interior_first=False):
        part = np.ascontiguousarray(part, dtype=np.int32)
        h = C.c_void_p()
        check(lib().rmh_halo_create(mesh.h, _ptr(part), int(rank), C.byref(h)))
        self.h = h
        self.n_interior = None
        if interior_first:
            ni = C.c_int64(0)
            check(lib().rmh_halo_interior_first(h, C.byref(ni)))
            self.n_interior = ni.value
        no, ng, ns = C.c_int64(), C.c_int64(), C.c_int64()
        npeer = C.c_int32()
        check(lib().rmh_halo_sizes(h, C.byref(no), C.byref(ng), C.byref(npeer), C.byref(ns)))
        self.owned = np.zeros(no.value, dtype=np.int64)
        self.ghost = np.zeros(ng.value, dtype=np.int64)
        self.peers = np.zeros(npeer.value, dtype=np.int32)
        self.send_off = np.zeros(npeer.value + 1, dtype=np.int32)
        self.recv_off = np.zeros(npeer.value + 1, dtype=np.int32)
        self.send_local = np.zeros(ns.value, dtype=np.int32)
        check(lib().rmh_halo_get(h, _ptr(self.owned), _ptr(self.ghost), _ptr(self.peers),
                                 _ptr(self.send_off), _ptr(self.recv_off), _ptr(self.send_local)))

    def __del__(self):
        try:
            if self.h:
                lib().rmh_halo_free(self.h)
                self.h = None
        except Exception:
            pass


def _blob_args(blobs):
    n = len(blobs)
    keep = [C.create_string_buffer(bytes(b), len(b)) for b in blobs]
    ptrs = (C.c_void_p * n)(*[C.cast(k, C.c_void_p).value for k in keep])
    sizes = (C.c_int64 * n)(*[len(b) for b in blobs])
    return n, ptrs, sizes, keep


class DPlan:
    """Host-only exchange plan (rmh_dplan_*): ghost-face slots, requests, send tables."""

    def __init__(self, halo, rank, world, dim, order, nbr_dof_owned):
        self._nbr = np.ascontiguousarray(nbr_dof_owned, dtype=np.int32)
        self.h = C.c_void_p()
        check(lib().rmh_dplan_create(halo.h, int(rank), int(world), int(dim), int(order),
                                     _ptr(self._nbr), C.byref(self.h)))
        ne, ng, ns = C.c_int64(), C.c_int64(), C.c_int64()
        npeer = C.c_int32()
        check(lib().rmh_dplan_sizes(self.h, C.byref(ne), C.byref(ng), C.byref(ns), C.byref(npeer)))
        self.ne, self.ne_ghost, self.n_slots, self.n_peers = ne.value, ng.value, ns.value, npeer.value

    def export(self):
        n = lib().rmh_dplan_blob_bytes(self.h)
        buf = C.create_string_buffer(n)
        check(lib().rmh_dplan_export(self.h, buf))
        return buf.raw

    def connect(self, blobs):
        n, ptrs, sizes, keep = _blob_args(blobs)
        check(lib().rmh_dplan_connect(self.h, n, ptrs, sizes))

    def slot_ghosts(self):
        out = np.zeros(self.n_slots, dtype=np.int32)
        check(lib().rmh_dplan_slot_ghosts(self.h, _ptr(out)))
        return out

    def peer(self, k):
        """(rank, flag_slot, tr_src, tr_dst, mm_src, mm_dst) of peer k (after connect)"""
        r, fs = C.c_int32(), C.c_int32()
        ntr, nmm = C.c_int64(), C.c_int64()
        check(lib().rmh_dplan_peer(self.h, int(k), C.byref(r), C.byref(ntr), C.byref(nmm), C.byref(fs)))
        a = [np.zeros(ntr.value, dtype=np.int32), np.zeros(ntr.value, dtype=np.int32),
             np.zeros(nmm.value, dtype=np.int32), np.zeros(nmm.value, dtype=np.int32)]
        check(lib().rmh_dplan_peer_tables(self.h, int(k), *[_ptr(x) for x in a]))
        return (r.value, fs.value, *a)

    def close(self):
        if self.h:
            lib().rmh_dplan_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Dist:
    """Device layer of a decomposed run (rmh_dist_*)."""

    def __init__(self, ctx, plan, rank, world, n_interior):
        self.ctx, self.plan = ctx, plan
        self.h = C.c_void_p()
        check(lib().rmh_dist_create(ctx.h, plan.h, int(rank), int(world), C.c_int64(n_interior),
                                    C.byref(self.h)))

    def export(self):
        n = lib().rmh_dist_blob_bytes(self.h)
        buf = C.create_string_buffer(n)
        check(lib().rmh_dist_export(self.h, buf))
        return buf.raw

    def connect(self, blobs):
        n, ptrs, sizes, keep = _blob_args(blobs)
        check(lib().rmh_dist_connect(self.h, n, ptrs, sizes))

    def rk_stage(self, lo_type, dt, a, b, x0, y, out, s=0):
        check(lib().rmh_dist_rk_stage(self.h, int(lo_type), C.c_double(dt), C.c_double(a), C.c_double(b),
                                      _dp(x0), _dp(y), _dp(out), C.c_void_p(s)))

    def rk_step(self, ode_solver_type, lo_type, t, dt, u, s=0):
        tt = C.c_double(t)
        check(lib().rmh_dist_rk_step(self.h, int(ode_solver_type), int(lo_type), C.byref(tt),
                                     C.c_double(dt), _dp(u), C.c_void_p(s)))
        return tt.value

    def rk_step_host(self, ode_solver_type, lo_type, t, dt, u_host):
        tt = C.c_double(t)
        check(lib().rmh_dist_rk_step_host(self.h, int(ode_solver_type), int(lo_type), C.byref(tt),
                                          C.c_double(dt), C.c_void_p(int(u_host))))
        return tt.value

    def rk_step_host_async(self, ode_solver_type, lo_type, t, dt, u_in_host, u_out_host=None):
        """queued variant (rmh_dist_rk_step_host_async); Context.host_sync() waits"""
        check(lib().rmh_dist_rk_step_host_async(self.h, int(ode_solver_type), int(lo_type), C.c_double(t),
                                                C.c_double(dt), C.c_void_p(int(u_in_host)),
                                                C.c_void_p(int(u_in_host if u_out_host is None else u_out_host))))
        return t + dt

    def allreduce(self, values, op='sum', s=0):
        v = np.ascontiguousarray(np.atleast_1d(values), dtype=np.float64).copy()
        check(lib().rmh_dist_allreduce(self.h, {'sum': 0, 'min': 1, 'max': 2}[op], _ptr(v), int(v.size),
                                       C.c_void_p(s)))
        return v

    def close(self):
        if self.h:
            lib().rmh_dist_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """Device context (rmh_ctx_*): owns the stored quadrature data of one rank."""

    def __init__(self, *, dim, order, mesh_order, exec_mode, bounds_type, nodes, nbr_dof,
                 vel_nodes=None, vel_quad=None, vel_face=None, lat=None, n_ent=0,
                 nbr_elem=None, inflow=None, ne_ghost=0, device=0):
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        self._keep = [f64(nodes), f64(vel_nodes), f64(vel_quad), f64(vel_face), i32(nbr_dof),
                      i32(lat), i32(nbr_elem), f64(inflow)]
        k = self._keep
        d = Desc()
        d.dim, d.order, d.mesh_order = dim, order, mesh_order
        d.exec_mode, d.bounds_type, d.device = exec_mode, bounds_type, device
        d.ne = k[0].shape[0]
        d.ne_ghost = ne_ghost
        d.nodes, d.vel_nodes, d.vel_quad, d.vel_face = _ptr(k[0]), _ptr(k[1]), _ptr(k[2]), _ptr(k[3])
        d.nbr_dof, d.lat, d.n_ent, d.nbr_elem, d.inflow = _ptr(k[4]), _ptr(k[5]), n_ent, _ptr(k[6]), _ptr(k[7])
        self.h = C.c_void_p()
        check(lib().rmh_ctx_create(C.byref(d), C.byref(self.h)))
        self._keep = None
        self.ndofs = lib().rmh_ctx_ndofs(self.h)
        self.nd = lib().rmh_ctx_nd(self.h)
        self.ne = d.ne
        self.nq1d = lib().rmh_ctx_nq1d(self.h)
        self.path_flags = lib().rmh_ctx_path_flags(self.h)
        self.dim = dim

    def close(self):
        if self.h:
            lib().rmh_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trust_state(self, on=True):
        check(lib().rmh_ctx_trust_state(self.h, int(bool(on))))

    def quad_points_1d(self):
        x = np.zeros(self.nq1d); w = np.zeros(self.nq1d)
        check(lib().rmh_ctx_quad_points_1d(self.h, _ptr(x), _ptr(w)))
        return x, w

    # each method mirrors one C entry point; `s` is a cudaStream_t as int (0 = default)
    def set_time(self, t, s=0):
        check(lib().rmh_set_time(self.h, C.c_double(t), C.c_void_p(s)))

    def lumped_mass(self, m, s=0):
        check(lib().rmh_lumped_mass(self.h, _dp(m), C.c_void_p(s)))

    def ho_mult(self, u, rhs, s=0):
        check(lib().rmh_ho_mult(self.h, _dp(u), _dp(rhs), C.c_void_p(s)))

    def mass_inv(self, rhs, du, s=0):
        check(lib().rmh_mass_inv(self.h, _dp(rhs), _dp(du), C.c_void_p(s)))

    def ho_local_inverse(self, u, du, s=0):
        check(lib().rmh_ho_local_inverse(self.h, _dp(u), _dp(du), C.c_void_p(s)))

    def lo_mass_avg(self, dt, u, du_ho, du_lo, s=0):
        check(lib().rmh_lo_mass_avg(self.h, C.c_double(dt), _dp(u), _dp(du_ho), _dp(du_lo),
                                    C.c_void_p(s)))

    def lo_discrete_upwind(self, u, du_lo, s=0):
        check(lib().rmh_lo_discrete_upwind(self.h, _dp(u), _dp(du_lo), C.c_void_p(s)))

    def lo_discrete_upwind_prec(self, u, du_lo, s=0):
        check(lib().rmh_lo_discrete_upwind_prec(self.h, _dp(u), _dp(du_lo), C.c_void_p(s)))

    def ho_neumann(self, u, du, s=0):
        check(lib().rmh_ho_neumann(self.h, _dp(u), _dp(du), C.c_void_p(s)))

    def lo_res_dist(self, u, du_lo, s=0):
        check(lib().rmh_lo_res_dist(self.h, _dp(u), _dp(du_lo), C.c_void_p(s)))

    def fa_setup(self, s=0):
        check(lib().rmh_fa_setup(self.h, C.c_void_p(s)))

    def subcell_setup(self, xlat, vel, s=0):
        xlat = np.ascontiguousarray(xlat, dtype=np.float64)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        check(lib().rmh_subcell_setup(self.h, _ptr(xlat), _ptr(vel), C.c_void_p(s)))

    def lo_res_dist_subcell(self, u, du_lo, s=0):
        check(lib().rmh_lo_res_dist_subcell(self.h, _dp(u), _dp(du_lo), C.c_void_p(s)))

    def fct_project(self, dt, u, du_ho, du_lo, xmin, xmax, du, s=0):
        check(lib().rmh_fct_project(self.h, C.c_double(dt), _dp(u), _dp(du_ho), _dp(du_lo), _dp(xmin),
                                    _dp(xmax), _dp(du), C.c_void_p(s)))

    def fct_nonlinear_penalty(self, dt, eps_w, u, m, du_ho, du_lo, xmin, xmax, du, s=0):
        check(lib().rmh_fct_nonlinear_penalty(self.h, C.c_double(dt), C.c_double(eps_w), _dp(u), _dp(m), _dp(du_ho),
                                              _dp(du_lo), _dp(xmin), _dp(xmax), _dp(du), C.c_void_p(s)))

    def si_update_bounds(self, dt, u, du_ho, si, xmin, xmax, s=0):
        check(lib().rmh_si_update_bounds(self.h, C.c_double(dt), _dp(u), _dp(du_ho), _dp(si), _dp(xmin), _dp(xmax),
                                         C.c_void_p(s)))

    def dt_control(self, mode):
        check(lib().rmh_dt_control(self.h, int(mode)))

    def dt_ratio(self, reset=False):
        r = C.c_double(0.0)
        check(lib().rmh_dt_ratio(self.h, int(bool(reset)), C.byref(r)))
        return r.value

    def mono_setup(self, mono_type, mass_lim, scale, s=0):
        scale = np.ascontiguousarray(scale, dtype=np.float64) if scale is not None else None
        check(lib().rmh_mono_setup(self.h, int(mono_type), int(bool(mass_lim)), _ptr(scale),
                                   C.c_void_p(s)))

    def si_setup(self, si_type, s=0):
        check(lib().rmh_si_setup(self.h, int(si_type), C.c_void_p(s)))

    def si_values(self, u, out, s=0):
        check(lib().rmh_si_values(self.h, _dp(u), _dp(out), C.c_void_p(s)))

    def mono_rd(self, u, du, s=0):
        check(lib().rmh_mono_rd(self.h, _dp(u), _dp(du), C.c_void_p(s)))

    def fa_get(self, which):
        nd, ne = self.nd, self.ne
        dim = self.dim
        nf = 2 * dim
        nfd = round(nd ** ((dim - 1) / dim))
        shape = {0: (ne, nd, nd), 1: (ne, nd, nd), 2: (ne, nd, nd), 3: (ne, nf, nfd, nfd),
                 4: (ne, nf, nfd)}[which]
        out = np.zeros(shape)
        check(lib().rmh_fa_get(self.h, int(which), _ptr(out)))
        return out

    def fct_flux_based(self, dt, u, m, du_ho, du_lo, xi_min, xi_max, du, s=0):
        check(lib().rmh_fct_flux_based(self.h, C.c_double(dt), _dp(u), _dp(m), _dp(du_ho),
                                       _dp(du_lo), _dp(xi_min), _dp(xi_max), _dp(du),
                                       C.c_void_p(s)))

    # ---- product-field remap (-ps): flags are uint8 device arrays
    def product_enable(self, on=True):
        check(lib().rmh_product_enable(self.h, int(bool(on))))

    def idp_use_mask(self, on=True):
        check(lib().rmh_idp_use_mask(self.h, int(bool(on))))

    def compute_mask(self, state, mask, s=0):
        check(lib().rmh_compute_mask(self.h, _dp(state), _dp(mask), C.c_void_p(s)))

    def prod_bool_indicators(self, u, el, dof, s=0):
        check(lib().rmh_prod_bool_indicators(self.h, _dp(u), _dp(el), _dp(dof), C.c_void_p(s)))

    def prod_compute_ratio(self, us, u, sr, el, dof, s=0):
        check(lib().rmh_prod_compute_ratio(self.h, _dp(us), _dp(u), _dp(sr), _dp(el), _dp(dof), C.c_void_p(s)))

    def elem_min_max_masked(self, u, el, dof, xe_min, xe_max, s=0):
        check(lib().rmh_elem_min_max_masked(self.h, _dp(u), _dp(el), _dp(dof), _dp(xe_min), _dp(xe_max),
                                            C.c_void_p(s)))

    def prod_compatible_lo(self, dt, us, m, d_us_ho, s_min, s_max, u_new, el, dof, d_lo, s=0):
        check(lib().rmh_prod_compatible_lo(self.h, C.c_double(dt), _dp(us), _dp(m), _dp(d_us_ho), _dp(s_min),
                                           _dp(s_max), _dp(u_new), _dp(el), _dp(dof), _dp(d_lo), C.c_void_p(s)))

    def prod_zero_empty(self, el, dof, d_us, s=0):
        check(lib().rmh_prod_zero_empty(self.h, _dp(el), _dp(dof), _dp(d_us), C.c_void_p(s)))

    def fct_product(self, fct_type, dt, us, m, d_us_ho, d_us_lo, s_min, s_max, u_new, el, dof, d_us, s=0):
        check(lib().rmh_fct_product(self.h, int(fct_type), C.c_double(dt), _dp(us), _dp(m), _dp(d_us_ho),
                                    _dp(d_us_lo), _dp(s_min), _dp(s_max), _dp(u_new), _dp(el), _dp(dof),
                                    _dp(d_us), C.c_void_p(s)))

    def limit_mult(self, lo_type, fct_type, dt, u, k, s=0):
        check(lib().rmh_limit_mult(self.h, int(lo_type), int(fct_type), C.c_double(dt), _dp(u), _dp(k),
                                   C.c_void_p(s)))

    def mult_unlimited(self, ho_type, lo_type, fct_type, t, dt, u, k, s=0):
        check(lib().rmh_mult_unlimited(self.h, int(ho_type), int(lo_type), int(fct_type), C.c_double(t),
                                       C.c_double(dt), _dp(u), _dp(k), C.c_void_p(s)))

    def mult(self, ho_type, lo_type, fct_type, t, dt, u, k, s=0):
        check(lib().rmh_mult(self.h, int(ho_type), int(lo_type), int(fct_type), C.c_double(t),
                             C.c_double(dt), _dp(u), _dp(k), C.c_void_p(s)))

    def ode_step(self, ode_solver_type, ho_type, lo_type, fct_type, t, dt, u, s=0):
        tt = C.c_double(t)
        check(lib().rmh_ode_step(self.h, int(ode_solver_type), int(ho_type), int(lo_type),
                                 int(fct_type), C.byref(tt), C.c_double(dt), _dp(u), C.c_void_p(s)))
        return tt.value

    def lincomb(self, coef, xs, out, s=0):
        n = len(coef)
        ca = (C.c_double * n)(*coef)
        xa = (C.c_void_p * n)(*[_dp(x).value for x in xs])
        check(lib().rmh_lincomb(self.h, n, ca, xa, _dp(out), C.c_void_p(s)))

    def elem_min_max(self, u, xe_min, xe_max, s=0):
        check(lib().rmh_elem_min_max(self.h, _dp(u), _dp(xe_min), _dp(xe_max), C.c_void_p(s)))

    def bounds(self, xe_min, xe_max, xi_min, xi_max, s=0):
        check(lib().rmh_bounds(self.h, _dp(xe_min), _dp(xe_max), _dp(xi_min), _dp(xi_max),
                               C.c_void_p(s)))

    def fct_clip_scale(self, dt, u, m, du_ho, du_lo, xi_min, xi_max, du, s=0):
        check(lib().rmh_fct_clip_scale(self.h, C.c_double(dt), _dp(u), _dp(m), _dp(du_ho),
                                       _dp(du_lo), _dp(xi_min), _dp(xi_max), _dp(du),
                                       C.c_void_p(s)))

    def stage(self, lo_type, dt, u, k, s=0):
        check(lib().rmh_stage(self.h, int(lo_type), C.c_double(dt), _dp(u), _dp(k), C.c_void_p(s)))

    def rk_stage(self, lo_type, dt, a, b, x0, y, out, s=0):
        check(lib().rmh_rk_stage(self.h, int(lo_type), C.c_double(dt), C.c_double(a),
                                 C.c_double(b), _dp(x0), _dp(y), _dp(out), C.c_void_p(s)))

    def rk_step(self, ode_solver_type, lo_type, t, dt, u, s=0):
        tt = C.c_double(t)
        check(lib().rmh_rk_step(self.h, int(ode_solver_type), int(lo_type), C.byref(tt),
                                C.c_double(dt), _dp(u), C.c_void_p(s)))
        return tt.value

    def rk_step_host(self, ode_solver_type, lo_type, t, dt, u_host):
        tt = C.c_double(t)
        check(lib().rmh_rk_step_host(self.h, int(ode_solver_type), int(lo_type), C.byref(tt),
                                     C.c_double(dt), C.c_void_p(int(u_host))))
        return tt.value

    def rk_step_host_async(self, ode_solver_type, lo_type, t, dt, u_in_host, u_out_host=None):
        """queue H2D + step + D2H (pinned host pointers); host_sync() waits"""
        check(lib().rmh_rk_step_host_async(self.h, int(ode_solver_type), int(lo_type), C.c_double(t), C.c_double(dt),
                                           C.c_void_p(int(u_in_host)),
                                           C.c_void_p(int(u_in_host if u_out_host is None else u_out_host))))
        return t + dt

    def halo_wait_stats(self, reset=False):
        """in-kernel halo diagnostics since the last reset (rmh_halo_wait_stats): dict of counts and ns"""
        out = (C.c_ulonglong * 7)()
        check(lib().rmh_halo_wait_stats(self.h, out, int(bool(reset))))
        keys = ('warps_waited', 'wait_ns_sum', 'wait_ns_max', 'warps_shell', 'shell_ns_sum', 'warps', 'run_ns_sum')
        return {k: int(v) for k, v in zip(keys, out)}

    def host_sync(self):
        check(lib().rmh_host_sync(self.h))

    def stage_minmax(self, y, s=0):
        check(lib().rmh_stage_minmax(self.h, _dp(y), C.c_void_p(s)))

    def profile(self, enable):
        ms = C.c_double(0.0); n = C.c_int64(0)
        check(lib().rmh_profile(self.h, int(enable), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def reduce(self, op, a, b=None, s=0):
        out = C.c_double(0.0)
        check(lib().rmh_reduce(self.h, int(op), _dp(a), _dp(b), C.byref(out), C.c_void_p(s)))
        return out.value


def launch_count(reset=False):
    return lib().rmh_launch_count(int(reset))
