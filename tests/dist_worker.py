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
    """Exchange plan on CPU: every rank builds its LocalPart, the plan blobs travel through gloo,
    and the put kernel is emulated with numpy on a known global field: every ghost-face trace and
    every ghost (min,max) pair must equal the owner's value."""
    import torch.distributed as dist
    import remhos_b200 as rb
    from remhos_b200.dist import LocalPart, allgather_blobs
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank,
                            world_size=world)
    cases = ((2, [6, 6], True, 2, 'cart'), (3, [3, 3, 3], True, 3, 'cart'), (3, [4, 3, 2], False, 2, 'cart'),
             (3, [3, 3, 3], True, 4, 'cart'), (2, None, True, 3, 'periodic-hexagon.mesh'))
    for dim, n, periodic, order, kind in cases:
        if kind == 'cart':
            m = rb.Mesh.cartesian(n, [2.0] * dim, origin=[-1.0] * dim, periodic=periodic).refine(1)
        else:
            m = rb.Mesh.load(os.path.join(ROOT, 'tests', 'data', kind)).refine(2)
        lp = LocalPart(m, rank, world, order)
        halo, maps, plan = lp.halo, lp.maps, lp.plan
        no, ng, nd = lp.n_owned, lp.n_ghost, maps['nd']
        nf, nfd = 2 * dim, (order + 1) ** (dim - 1)
        ids = np.concatenate([halo.owned, halo.ghost])
        # interior-first ordering: no leading element shares a lattice entity with a foreign element
        gmaps = m.dof_maps(1)
        lat = gmaps['lat']
        ent_foreign = np.zeros(gmaps['n_ent'], dtype=bool)
        foreign = np.setdiff1d(np.arange(m.ne), halo.owned)
        ent_foreign[lat[foreign].reshape(-1)] = True
        assert not ent_foreign[lat[halo.owned[:lp.n_interior]]].any(), (rank, dim)
        assert ent_foreign[lat[halo.owned[lp.n_interior:]]].any(axis=1).all(), (rank, dim)
        assert (halo.send_local >= lp.n_interior).all()
        # every face neighbour of an owned element is owned or in the ghost ring
        nbg = gmaps['nbr_elem'][halo.owned].reshape(-1)
        assert np.isin(nbg[nbg >= 0], ids).all()
        # connect the plans
        plan.connect(allgather_blobs(plan.export(), world))
        peers_all = [None] * world
        dist.all_gather_object(peers_all, halo.peers.tolist())
        # global field: dof (G, loc) -> G * nd + loc + 0.25; element G -> (-G, G)
        y = (halo.owned[:, None] * nd + np.arange(nd)[None, :] + 0.25).reshape(-1)
        mm_own = np.stack([-halo.owned.astype(np.float64), halo.owned.astype(np.float64)], 1)
        out = {}
        for k in range(plan.n_peers):
            pr, fslot, tr_src, tr_dst, mm_src, mm_dst = plan.peer(k)
            assert pr == halo.peers[k] and peers_all[pr][fslot] == rank
            out[pr] = (tr_dst, y[tr_src], mm_dst, mm_own[mm_src])
        box = [None] * world
        dist.all_gather_object(box, out)
        gtr = np.full(max(plan.n_slots, 1) * nfd, np.nan)
        mm = np.full((no + ng, 2), np.nan)
        for r in range(world):
            if rank in box[r]:
                tr_dst, vals, mm_dst, mmv = box[r][rank]
                gtr[tr_dst] = vals
                mm[mm_dst] = mmv
        # expected ghost traces, slot by slot in scan order, natural face order
        nbr = maps['nbr_dof'][:no]
        bd = maps['bdr_dofs']
        slot = 0
        slot_ghost = plan.slot_ghosts()
        for e in range(no):
            for f in range(nf):
                row = nbr[e, f]
                if row[0] < 0 or row[0] // nd < no:
                    continue
                nat = np.argsort(bd[:, f], kind='stable')
                le = row[nat] // nd
                assert (le == le[0]).all() and slot_ghost[slot] == le[0] - no
                exp = ids[le] * nd + row[nat] % nd + 0.25
                assert np.array_equal(gtr[slot * nfd:(slot + 1) * nfd], exp), (rank, dim, e, f)
                slot += 1
        assert slot == plan.n_slots and not np.isnan(gtr[:plan.n_slots * nfd]).any()
        assert np.array_equal(mm[no:], np.stack([-halo.ghost.astype(np.float64),
                                                 halo.ghost.astype(np.float64)], 1)), (rank, dim)
    dist.barrier()
    dist.destroy_process_group()


def run_gpu():
    """nccl, one GPU per rank: RK3 steps on the decomposed mesh must match the single-GPU run."""
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
    worst = 0.0
    for bt, problem, order, nofold in ((0, 0, 3, '0'), (0, 0, 3, '1'), (0, 1, 3, '0'), (1, 0, 3, '0'),
                                       (1, 1, 3, '0'), (0, 0, 4, '0'), (1, 0, 4, '0'), (0, 0, 2, '0')):
        os.environ['RMH_NO_FOLD'] = nofold      # '1': entity pass instead of the in-kernel bounds
        mesh = rb.Mesh.cartesian([3, 3, 3], [2.0] * 3, origin=[-1.0] * 3, periodic=True).refine(1)
        dp = DistProblem(mesh, rank, world, problem=problem, order=order, bounds_type=bt, dt=0.01,
                         device=local)
        dp.ctx.trust_state(True)
        u = torch.tensor(dp.u0, device='cuda')
        t = 0.0
        for _ in range(steps):
            t = dp.rk3_step(t, u)
        torch.cuda.synchronize()
        # single-GPU reference of the same global problem (every rank computes it)
        mesh1 = rb.Mesh.cartesian([3, 3, 3], [2.0] * 3, origin=[-1.0] * 3, periodic=True).refine(1)
        p1 = Problem(mesh1, problem=problem, order=order, bounds_type=bt, dt=0.01, device=local)
        u1 = torch.tensor(p1.u0, device='cuda')
        t1 = 0.0
        for _ in range(steps):
            t1 = p1.ctx.rk_step(3, 5, t1, 0.01, u1)
        ref = u1.cpu().numpy().reshape(mesh1.ne, -1)[dp.plan.owned]
        got = u.cpu().numpy().reshape(ref.shape)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        worst = max(worst, err)
        assert err < 1e-12, (rank, bt, problem, order, err)
        m = torch.empty_like(u)
        dp.ctx.lumped_mass(m)
        mass = dp.allreduce(dp.ctx.reduce(0, u, m), 'sum')
        m1 = torch.empty_like(u1)
        p1.ctx.lumped_mass(m1)
        mass1 = p1.ctx.reduce(0, u1, m1)
        assert abs(mass - mass1) < 1e-12 * abs(mass1), (mass, mass1)
        umax = dp.allreduce(dp.ctx.reduce(2, u), 'max')
        assert umax == p1.ctx.reduce(2, u1)
        dp.close(); p1.close()
    dist.barrier()
    if rank == 0:
        print('DIST_GPU_OK world=%d worst_rel_err=%.3e' % (world, worst))
    dist.destroy_process_group()


if __name__ == '__main__':
    if sys.argv[1] == 'gpu':
        run_gpu()
