"""Domain-decomposed stepping, one process per GPU.  Everything on the data path lives under the C
ABI (rmh_dplan_* / rmh_dist_*: face-trace + (min,max) puts into the peers' windows over NVLink,
in-kernel halo wait, ncclAllReduce for scalars); this module only sequences the set-up calls and
carries the set-up blobs between the ranks with torch.distributed (any backend: the blobs are
bytes), which replaces ParGridFunction::ExchangeFaceNbrData / GroupCommunicator / MPI_Allreduce
(remhos.cpp:1813, remhos_tools.cpp:463-466, remhos.cpp:1073-1076) on the reference side.
"""
import numpy as np

from . import capi
from .setup_problem import mesh_eval, velocity, u0 as eval_u0


def allgather_blobs(blob, world):
    """every rank's bytes, ordered by rank (torch.distributed must be initialised when world > 1)"""
    if world == 1:
        return [blob]
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, blob)
    return out


class LocalPart:
    """One rank's share of a global mesh: halo plan (owned elements interior first), the local
    mesh of owned + ghost elements, its index maps and the host-side exchange plan.  CPU only."""

    def __init__(self, mesh, rank, world, order, mesh_order=2, part=None):
        self.rank, self.world, self.order = rank, world, order
        self.dim = mesh.dim
        if part is None:
            part = mesh.partition(world)
        self.halo = mesh.halo(part, rank, interior_first=True)
        self.n_interior = self.halo.n_interior
        ids = np.concatenate([self.halo.owned, self.halo.ghost])
        self.local = mesh.extract(ids)
        self.local.set_curvature(mesh_order)
        self.n_owned, self.n_ghost = self.halo.owned.size, self.halo.ghost.size
        self.maps = self.local.dof_maps(order)
        self.plan = capi.DPlan(self.halo, rank, world, self.dim, order,
                               self.maps['nbr_dof'][:self.n_owned])


class DistProblem:
    """Set-up of one rank's share of a run (transport problems; same reference defaults as
    setup_problem.Problem).  `mesh` is the GLOBAL mesh (geometry order 1 is enough), already
    refined; every rank builds the same RCB partition and keeps only its part plus a ghost ring."""

    def __init__(self, mesh, rank, world, problem=0, order=3, mesh_order=2, bounds_type=0,
                 dt=0.005, device=0, part=None):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.order = order
        self.bb_min, self.bb_max = mesh.bounding_box()
        dim = mesh.dim
        if problem >= 10:
            raise NotImplementedError('distributed remap set-up')
        nodal_ok = (problem % 20) in (0, 1, 2, 4, 5, 6, 7)
        if not nodal_ok:
            raise NotImplementedError('distributed set-up samples the velocity at the nodes')
        lp = LocalPart(mesh, rank, world, order, mesh_order, part)
        self.part = lp
        self.plan = lp.halo               # owned / ghost global ids (tests, bench)
        no, ng = lp.n_owned, lp.n_ghost
        maps = lp.maps
        nodes = lp.local.nodes()[:no]
        vel_nodes = velocity(problem, nodes, self.bb_min, self.bb_max)
        own_mesh = lp.local.extract(np.arange(no, dtype=np.int64))
        lat_pts = np.arange(order + 1) / max(order, 1)
        xdof = mesh_eval(own_mesh, lat_pts)
        self.ctx = capi.Context(dim=dim, order=order, mesh_order=mesh_order, exec_mode=0,
                                bounds_type=bounds_type, nodes=nodes,
                                nbr_dof=maps['nbr_dof'][:no], lat=maps['lat'], n_ent=maps['n_ent'],
                                nbr_elem=maps['nbr_elem'][:no], vel_nodes=vel_nodes,
                                ne_ghost=ng, device=device)
        self.u0 = eval_u0(problem, xdof, self.bb_min, self.bb_max)
        self.dt = dt
        self.bounds_type = bounds_type
        self.nd = maps['nd']
        self.n_owned, self.n_ghost = no, ng
        self.n_interior = lp.n_interior
        self.dist = capi.Dist(self.ctx, lp.plan, rank, world, lp.n_interior)
        self.dist.connect(allgather_blobs(self.dist.export(), world))

    def rk3_step(self, t, u, stream=0):
        """RK3-SSP step (remhos.cpp:490) on the decomposed mesh: per stage one put kernel and one
        stage kernel (rmh_dist_rk_step)"""
        return self.dist.rk_step(3, 5, t, self.dt, u, stream)

    def allreduce(self, value, op='sum'):
        return float(self.dist.allreduce([value], op)[0])

    def close(self):
        """collective: all ranks must have finished stepping"""
        if self.world > 1:
            import torch.distributed as dist
            self.torch.cuda.synchronize()
            dist.barrier()
        self.dist.close()
        self.ctx.close()
