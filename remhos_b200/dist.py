"""Domain-decomposed stepping: one process per GPU, torch.distributed (NCCL over NVLink) for the
face-neighbour halo exchange that replaces ParGridFunction::ExchangeFaceNbrData and the shared-DOF
min/max reduction of DofInfo.  Scalars (mass, min, max, dt) go through all_reduce.

The exchange itself (`exchange`) only moves torch tensors, so the same code runs under gloo on
CPU tensors in the tests.
"""
import ctypes as C
import numpy as np

from . import capi
from .capi import lib, check, _ptr
from .setup_problem import mesh_eval, velocity, u0 as eval_u0


def exchange(dist, plan, send_bufs, recv_bufs, widths):
    """Point-to-point exchange of per-peer contiguous slices.  send_bufs/recv_bufs: lists of 1-D
    tensors laid out [n_elements * width]; plan gives peers and element offsets."""
    if len(plan.peers) == 0:
        return
    ops = []
    for k, peer in enumerate(plan.peers):
        s0, s1 = int(plan.send_off[k]), int(plan.send_off[k + 1])
        r0, r1 = int(plan.recv_off[k]), int(plan.recv_off[k + 1])
        for sb, rb, w in zip(send_bufs, recv_bufs, widths):
            ops.append(dist.P2POp(dist.isend, sb[s0 * w:s1 * w], int(peer)))
            ops.append(dist.P2POp(dist.irecv, rb[r0 * w:r1 * w], int(peer)))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def interior_first(plan):
    """Reorder the owned elements of a halo plan in place: those that some peer needs (= those
    sharing a vertex with a ghost element -- the ghost ring is vertex adjacency, which is symmetric)
    go last, so that the stage kernel can run on the leading elements while the halo is still in
    flight.  Returns the number of leading (interior) elements."""
    bnd = np.zeros(plan.owned.size, dtype=bool)
    bnd[plan.send_local] = True
    perm = np.concatenate([np.flatnonzero(~bnd), np.flatnonzero(bnd)])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    plan.owned = plan.owned[perm]
    plan.send_local = inv[plan.send_local].astype(np.int32)
    return int((~bnd).sum())


class DistProblem:
    """Set-up of one rank's share of a run (transport problems; same reference defaults as
    setup_problem.Problem).  `mesh` is the GLOBAL mesh (geometry order 1 is enough), already
    refined; every rank builds the same RCB partition and keeps only its part plus a ghost ring."""

    def __init__(self, mesh, rank, world, problem=0, order=3, mesh_order=2, bounds_type=0,
                 dt=0.005, device=0):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.order = order
        self.bb_min, self.bb_max = mesh.bounding_box()
        dim = mesh.dim
        part = mesh.partition(world)
        plan = mesh.halo(part, rank)
        self.n_interior = interior_first(plan)
        self.plan = plan
        ids = np.concatenate([plan.owned, plan.ghost])
        local = mesh.extract(ids)
        local.set_curvature(mesh_order)
        no, ng = plan.owned.size, plan.ghost.size
        maps = local.dof_maps(order)
        nodes = local.nodes()[:no]
        if problem >= 10:
            raise NotImplementedError('distributed remap set-up')
        nodal_ok = (problem % 20) in (0, 1, 2, 4, 5, 6, 7)
        if not nodal_ok:
            raise NotImplementedError('distributed set-up samples the velocity at the nodes')
        vel_nodes = velocity(problem, nodes, self.bb_min, self.bb_max)
        own_mesh = local.extract(np.arange(no, dtype=np.int64))
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
        dev = torch.device('cuda', device)
        f64 = torch.float64
        ns = plan.send_local.size
        self.send_local = torch.tensor(plan.send_local, dtype=torch.int32, device=dev)
        self.send_u = torch.empty(max(ns, 1) * self.nd, dtype=f64, device=dev)
        self.send_mm = torch.empty(max(ns, 1) * 2, dtype=f64, device=dev)
        self.ghost_u = torch.zeros(max(ng, 1) * self.nd, dtype=f64, device=dev)
        self.ghost_mm = torch.zeros(max(ng, 1) * 2, dtype=f64, device=dev)
        self.w1 = torch.empty(self.ctx.ndofs, dtype=f64, device=dev)
        self.w2 = torch.empty(self.ctx.ndofs, dtype=f64, device=dev)
        self.n_send = ns
        # overlap of the exchange with the interior elements (rmh_rk_stage_part; constant-coefficient
        # stage kernel and overlap bounds only).  Default: on from 4 ranks (8 ranks, 2x2x2 bricks
        # exchanging their whole surface: 2.39 vs 2.48 ms/step measured; 4 ranks: 2.17 vs ~2.20),
        # off on 2: there the second, small
        # launch for the elements next to the ghost ring plus the NCCL kernels competing for SMs
        # cost more (2.19 ms/step) than the ~60 us of exchange they hide (2.12 ms/step sequential).
        # RMH_OVERLAP=0/1 forces either form.
        import os
        self.overlap = bool(world > 1 and ng > 0 and bounds_type == 0 and (self.ctx.path_flags & 8)
                            and os.environ.get('RMH_OVERLAP', '1' if world >= 4 else '0') == '1')
        if self.overlap:
            self.ctx.dist_split(self.n_interior)
            self.cs = torch.cuda.Stream(device=dev)

    def halo(self, y, stream=0):
        """pack -> NCCL send/recv -> install the ghosts of y (element min/max of y must already be
        in the context)."""
        import torch.distributed as dist
        self.ctx.halo_pack(y, self.send_local, self.n_send, self.send_u, self.send_mm, stream)
        if self.world > 1:
            exchange(dist, self.plan, [self.send_u, self.send_mm], [self.ghost_u, self.ghost_mm],
                     [self.nd, 2])
        self.ctx.halo_set(self.ghost_u, self.ghost_mm, stream)

    def stage(self, a, b, x0, y, out, stream=0):
        """one RK stage out = a x0 + b (y + dt F(y)) on the decomposed mesh, halo of y included"""
        ctx, dt = self.ctx, self.dt
        if not self.overlap:
            self.halo(y, stream)
            ctx.rk_stage_dist(5, dt, a, b, x0, y, out, stream)
            return
        import torch.distributed as dist
        torch = self.torch
        main = torch.cuda.current_stream()
        ms = main.cuda_stream
        ctx.halo_pack(y, self.send_local, self.n_send, self.send_u, self.send_mm, ms)
        self.cs.wait_stream(main)
        with torch.cuda.stream(self.cs):                # NCCL send/recv ordered behind the pack
            exchange(dist, self.plan, [self.send_u, self.send_mm], [self.ghost_u, self.ghost_mm],
                     [self.nd, 2])
        ctx.rk_stage_part(5, dt, a, b, x0, y, out, 1, ms)     # interior elements: no ghost dependence
        main.wait_stream(self.cs)
        ctx.halo_set(self.ghost_u, self.ghost_mm, ms)
        ctx.rk_stage_part(5, dt, a, b, x0, y, out, 2, ms)     # elements next to the ghost ring

    def rk3_step(self, t, u, stream=0):
        """RK3-SSP step (remhos.cpp:490) on the decomposed mesh: three fused stage launches, each
        preceded by one halo exchange."""
        ctx, dt = self.ctx, self.dt
        chain = self.bounds_type == 0     # stage kernel leaves min/max of its output in the context
        # trust_state (cf. rmh_ctx_trust_state): the caller does not touch u between steps, so the
        # element min/max the last stage left for its output u are still valid
        if not (chain and getattr(self, 'trust_state', False) and getattr(self, '_xe_for', None) == u.data_ptr()):
            ctx.stage_minmax(u, stream)
        self._xe_for = None
        self.stage(0.0, 1.0, u, u, self.w1, stream)
        if not chain:
            ctx.stage_minmax(self.w1, stream)
        self.stage(0.75, 0.25, u, self.w1, self.w2, stream)
        if not chain:
            ctx.stage_minmax(self.w2, stream)
        self.stage(1.0 / 3.0, 2.0 / 3.0, u, self.w2, u, stream)
        if chain:
            self._xe_for = u.data_ptr()
        return t + dt

    def allreduce(self, value, op='sum'):
        import torch.distributed as dist
        if self.world == 1:
            return value
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.w1.device)
        dist.all_reduce(t, op={'sum': dist.ReduceOp.SUM, 'min': dist.ReduceOp.MIN,
                               'max': dist.ReduceOp.MAX}[op])
        return float(t[0])

    def close(self):
        self.ctx.close()
