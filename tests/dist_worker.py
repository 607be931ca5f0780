"""Worker for the multi-rank tests (launched by torchrun / mp.spawn).

mode cpu: gloo, CPU tensors -- checks the halo plan + exchange against a known global field.
mode gpu: nccl, one GPU per rank -- RK3 steps on the decomposed mesh must match the single-GPU run.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_cpu(rank, world, port):
    import torch
    import torch.distributed as dist
    import remhos_b200 as rb
    from remhos_b200.dist import exchange
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank,
                            world_size=world)
    for dim, n, periodic in ((2, [6, 6], True), (3, [3, 3, 3], True), (3, [4, 3, 2], False)):
        m = rb.Mesh.cartesian(n, [2.0] * dim, origin=[-1.0] * dim, periodic=periodic).refine(1)
        part = m.partition(world)
        plan = m.halo(part, rank)
        nd = 5
        field = np.arange(m.ne * nd, dtype=np.float64).reshape(m.ne, nd) * 0.5 + 1.0
        mm = np.stack([-np.arange(m.ne, dtype=np.float64), np.arange(m.ne, dtype=np.float64)], 1)
        own = plan.owned[plan.send_local]
        send_u = torch.tensor(field[own].reshape(-1))
        send_mm = torch.tensor(mm[own].reshape(-1))
        ghost_u = torch.zeros(plan.ghost.size * nd, dtype=torch.float64)
        ghost_mm = torch.zeros(plan.ghost.size * 2, dtype=torch.float64)
        exchange(dist, plan, [send_u, send_mm], [ghost_u, ghost_mm], [nd, 2])
        assert np.array_equal(ghost_u.numpy().reshape(-1, nd), field[plan.ghost]), (rank, dim)
        assert np.array_equal(ghost_mm.numpy().reshape(-1, 2), mm[plan.ghost]), (rank, dim)
        # every face neighbour of an owned element is owned or in the ghost ring
        maps = m.dof_maps(1)
        nb = maps['nbr_elem'][plan.owned].reshape(-1)
        nb = nb[nb >= 0]
        known = np.concatenate([plan.owned, plan.ghost])
        assert np.isin(nb, known).all()
        # interior-first reordering (overlap of the exchange with interior work): the elements sent
        # stay the same, and no leading element shares a lattice entity (vertex, edge, face) with an
        # element of another rank
        from remhos_b200.dist import interior_first
        sent = plan.owned[plan.send_local].copy()
        owned_set = set(plan.owned.tolist())
        n_int = interior_first(plan)
        assert np.array_equal(plan.owned[plan.send_local], sent)
        assert set(plan.owned.tolist()) == owned_set and (plan.send_local >= n_int).all()
        lat = maps['lat']
        ent_foreign = np.zeros(maps['n_ent'], dtype=bool)
        foreign = np.setdiff1d(np.arange(m.ne), plan.owned)
        ent_foreign[lat[foreign].reshape(-1)] = True
        assert not ent_foreign[lat[plan.owned[:n_int]]].any(), (rank, dim)
    dist.barrier()
    dist.destroy_process_group()


def run_gpu():
    import torch
    import torch.distributed as dist
    import remhos_b200 as rb
    from remhos_b200.dist import DistProblem
    from remhos_b200.setup_problem import Problem
    rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    steps = 4
    for bt, problem, overlap in ((0, 0, '0'), (0, 0, '1'), (0, 1, '0'), (1, 0, '0'), (1, 1, '0')):
        os.environ['RMH_OVERLAP'] = overlap      # '1': exchange overlapped with the interior elements
        if True:
            mesh = rb.Mesh.cartesian([3, 3, 3], [2.0] * 3, origin=[-1.0] * 3, periodic=True).refine(1)
            dp = DistProblem(mesh, rank, world, problem=problem, order=3, bounds_type=bt, dt=0.01,
                             device=local)
            u = torch.tensor(dp.u0, device='cuda')
            t = 0.0
            for _ in range(steps):
                t = dp.rk3_step(t, u)
            torch.cuda.synchronize()
            # single-GPU reference of the same global problem (every rank computes it)
            mesh1 = rb.Mesh.cartesian([3, 3, 3], [2.0] * 3, origin=[-1.0] * 3, periodic=True).refine(1)
            p1 = Problem(mesh1, problem=problem, order=3, bounds_type=bt, dt=0.01, device=local)
            u1 = torch.tensor(p1.u0, device='cuda')
            t1 = 0.0
            for _ in range(steps):
                t1 = p1.ctx.rk_step(3, 5, t1, 0.01, u1)
            ref = u1.cpu().numpy().reshape(mesh1.ne, -1)[dp.plan.owned]
            got = u.cpu().numpy().reshape(ref.shape)
            err = np.abs(got - ref).max() / np.abs(ref).max()
            assert err < 1e-12, (rank, bt, problem, err)
            m = torch.empty_like(u)
            dp.ctx.lumped_mass(m)
            mass = dp.allreduce(dp.ctx.reduce(0, u, m), 'sum')
            m1 = torch.empty_like(u1)
            p1.ctx.lumped_mass(m1)
            mass1 = p1.ctx.reduce(0, u1, m1)
            assert abs(mass - mass1) < 1e-12 * abs(mass1), (mass, mass1)
            dp.close(); p1.close()
    dist.barrier()
    if rank == 0:
        print('DIST_GPU_OK world=%d' % world)
    dist.destroy_process_group()


if __name__ == '__main__':
    if sys.argv[1] == 'gpu':
        run_gpu()
