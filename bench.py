#!/usr/bin/env python
"""bench.py -- DOF*RK-stage updates/s of the Remhos RK-stage path on B200 (BASELINE.json metric).

Workload (config C2): 3D periodic cube [-1,1]^3, 3x3x3 coarse hexes refined
`--rs` times (default 5 -> 884 736 elements), order 3 (56.6 M DOFs per GPU), mesh order 2,
problem 0 (erfc bump, constant velocity), `-ho 3 -lo 5 -fct 2 -pa -s 3`: LocalInverse HO +
MassBasedAvg LO + ClipScale FCT, RK3-SSP.  A "step" is one RK3 time step = 3 fused stage
launches; value = DOFs * 3 * steps / time, summed over ranks (weak scaling: every rank owns a
(3*2^rs)^3 brick of a periodic box that grows with the rank count).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--order P] [--problem Q]

Prints one JSON line (see the task contract): value (state resident in HBM), e2e (host state,
H2D + D2H inside the timed region, through rmh_rk_step_host / rmh_dist_rk_step_host), roofline of
the fused stage kernel, cpu_baseline (the CPU oracle port timed on the host cores on a bounded
sample), and -- N = 1 -- `extra`: the same metric on the general-velocity and remap paths.
With N > 1 the line also carries check.dist_rel_err: decomposed vs single-GPU runs of a small
global problem (orders 3 and 4, both bounds types), which must agree to 1e-12.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG = 136.0          # algorithmic bytes per DOF*stage of five UNFUSED kernels
STAGES = 3             # RK3-SSP


def read_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


def read_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the stage kernel, from the ncu
    --set full captures summarised in profiles/traffic.json (keyed by kernel|order|rs|problem);
    None when no capture of this exact workload is committed."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            e = json.load(f).get(key)
        return (float(e['dram_bytes']), e.get('source')) if e else (None, None)
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi sampled every 50 ms from before the warm-up to the end of the e2e loop; the
    reported clocks are the samples whose timestamp falls inside [t0, t1] (the timed regions)."""
    Q = ('timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '50'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0, t1):
        import datetime
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, sm_all, smax, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f:
            c = [x.strip() for x in line.split(',')]
            if len(c) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                clk, cmax = float(c[2]), float(c[3])
            except ValueError:
                continue
            sm_all.append(clk); smax.append(cmax)
            if t0 - 0.05 <= ts <= t1 + 0.05:
                sm.append(clk)
                for n, v in zip(names, c[6:10]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
        if sm_all:
            use = sm if sm else sm_all
            out = {'sm_mhz': float(np.median(use)), 'sm_max_mhz': float(max(smax)),
                   'reasons': sorted(reasons), 'samples': len(sm),
                   'window': 'timed + e2e loops' if sm else 'whole run (no sample inside the timed window)'}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_port_run(order, rs, problem, steps, warmup, budget_s):
    """The C/OpenMP port of the stage path (oracle/c/remhos_stage.c: same algorithm as the
    reference's `-ho 3 -lo 5 -fct 2 -pa -s 3` path, sum-factorised, all host threads) on the bench
    workload at -rs `rs`, inputs built by the oracle's own mesh code (no product library is loaded
    on this path).  The reference's own MFEM/MPI build cannot be produced in this image (SURVEY.md
    8c), so this is a port, not the reference.  Times `steps` RK3 steps (fewer when `budget_s`
    runs out first)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from remhos_oracle.cport import Port
    cores = host_cores()
    Port.set_threads(cores)           # launchers such as torchrun export OMP_NUM_THREADS=1
    n = 3 * 2 ** rs
    port, u0 = Port.periodic_cube(n, order, problem)
    h = 2.0 / n
    dt = 0.25 * h / order
    u = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1).copy()
    ndof = u.size
    m = port.lumped_mass().reshape(-1)
    mass0 = float((m * u).sum())
    t0 = time.perf_counter()
    for _ in range(max(1, warmup)):
        port.rk3_step(0.0, dt, u)
        if time.perf_counter() - t0 > 0.25 * budget_s:
            break
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        port.rk3_step(0.0, dt, u)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    el = time.perf_counter() - t0
    drift = abs(float((m * u).sum()) - mass0) / abs(mass0)
    threads = Port.threads()
    port.close()
    return {'value': ndof * STAGES * done / el, 'unit': 'DOF*stage/s', 'cores': int(threads),
            'kind': 'port', 'steps_timed': done, 'ms_per_step': 1e3 * el / done,
            'sample': 'C/OpenMP port of the stage path (oracle/c), periodic cube -rs %d order %d '
                      '(%d DOFs), %d RK3 steps in %.1f s, %d threads, mass drift %.1e'
                      % (rs, order, ndof, done, el, threads, drift)}


def dist_parity_check(rank, world, local_rank):
    """Decomposed vs single-GPU: 2 RK3 steps of a small global problem (periodic 12^3 cube), orders 3
    and 4, overlap (-bt 0) and sparsity (-bt 1) bounds; every rank's owned part is compared on rank 0
    with the single-GPU run of the same global mesh.  Returns the worst relative error (max norm)."""
    import torch
    import torch.distributed as dist
    import remhos_b200 as rb
    from remhos_b200.dist import DistProblem
    from remhos_b200.setup_problem import Problem
    worst, cases = 0.0, []
    for order, bt in ((3, 0), (3, 1), (4, 0), (4, 1)):
        n = 12
        dt = 0.25 * (2.0 / n) / order
        mesh = rb.Mesh.cartesian([n, n, n], [2.0] * 3, origin=[-1.0] * 3, periodic=True)
        dp = DistProblem(mesh, rank, world, problem=0, order=order, bounds_type=bt, dt=dt, device=local_rank)
        dp.ctx.trust_state(True)
        u = torch.tensor(dp.u0, device='cuda')
        t = 0.0
        for _ in range(2):
            t = dp.rk3_step(t, u)
        torch.cuda.synchronize()
        mine = (dp.plan.owned.copy(), u.cpu().numpy().reshape(dp.n_owned, -1))
        box = [None] * world
        dist.all_gather_object(box, mine)
        err = 0.0
        if rank == 0:
            mesh1 = rb.Mesh.cartesian([n, n, n], [2.0] * 3, origin=[-1.0] * 3, periodic=True)
            p1 = Problem(mesh1, problem=0, order=order, bounds_type=bt, dt=dt, device=local_rank)
            u1 = torch.tensor(p1.u0, device='cuda')
            t1 = 0.0
            for _ in range(2):
                t1 = p1.ctx.rk_step(3, 5, t1, dt, u1)
            ref = u1.cpu().numpy().reshape(mesh1.ne, -1)
            seen = np.zeros(mesh1.ne, dtype=bool)
            for ids, vals in box:
                err = max(err, float(np.abs(vals - ref[ids]).max() / np.abs(ref).max()))
                seen[ids] = True
            assert seen.all()
            p1.close()
        dp.close()
        cases.append({'order': order, 'bounds_type': bt, 'rel_err': err})
        worst = max(worst, err)
    return worst, cases


def extra_runs(order, rs, local_rank, steps=20):
    """The general paths, same metric, driver-visible: problem 1 (rotation: velocity varies inside an
    element -> FP64 tensor-core kernel k_stage3w) on the bench mesh, and remap (problem 10 on the
    refined unit cube: quadrature data rebuilt for every new stage time, non-affine mass solve)."""
    import torch
    import remhos_b200 as rb
    from remhos_b200.setup_problem import Problem
    out = {}

    def timed(ctx, u, dt, nst):
        t = 0.0
        for _ in range(2):
            t = ctx.rk_step(3, 5, t, dt, u)
        torch.cuda.synchronize()
        ctx.profile(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nst):
            t = ctx.rk_step(3, 5, t, dt, u)
        e1.record()
        torch.cuda.synchronize()
        kms, kl = ctx.profile(0)
        return e0.elapsed_time(e1), kms / max(kl, 1)
    try:
        mesh = rb.Mesh.cartesian([3, 3, 3], [2.0, 2.0, 2.0], origin=[-1.0, -1.0, -1.0], periodic=True)
        mesh.refine(rs)
        dt = 0.25 * (2.0 / (3 * 2 ** rs)) / order
        prob = Problem(mesh, problem=1, order=order, mesh_order=2, bounds_type=0, dt=dt, device=local_rank)
        prob.ctx.trust_state(True)
        u = torch.tensor(prob.u0, device='cuda')
        ms, kms = timed(prob.ctx, u, dt, steps)
        n = prob.ctx.ndofs
        tr, src = read_traffic('k_stage3w|o%d|rs%d|p1' % (order, rs))
        peak, _ = read_peaks()
        out['problem1_rotation'] = {
            'workload': '3D periodic-cube transport, order %d, -rs %d, problem 1 (rotation)' % (order, rs),
            'value': n * STAGES * steps / (ms * 1e-3), 'unit': 'DOF*stage/s', 'ms_per_step': ms / steps,
            'steps': steps, 'kernel': 'k_stage3w (FP64 DMMA, stored quadrature data)', 'kernel_ms': kms,
            'traffic': tr, 'frac': (tr / (kms * 1e-3) / 1e9 / peak) if tr else None, 'traffic_source': src}
        prob.close()
        del u, prob, mesh
        torch.cuda.empty_cache()
    except Exception as ex:                                    # the headline line must survive
        out['problem1_rotation'] = {'error': repr(ex)[:300]}
    try:
        mesh = rb.Mesh.cartesian([2, 2, 2], [1.0, 1.0, 1.0])
        mesh.refine(min(rs, 5))
        prob = Problem(mesh, problem=10, order=order, mesh_order=2, bounds_type=0, dt=-1.0, t_final=0.5,
                       device=local_rank)
        u = torch.tensor(prob.u0, device='cuda')
        nst = max(3, steps // 4)
        ms, kms = timed(prob.ctx, u, prob.dt, nst)
        n = prob.ctx.ndofs
        out['remap'] = {
            'workload': 'remap -p 10 (Taylor-Green mesh motion), unit cube -rs %d, order %d, operators '
                        'rebuilt for every new stage time (2 of the 3 stage times per RK3 step; t + dt is reused by the next step)' % (min(rs, 5), order),
            'value': n * STAGES * nst / (ms * 1e-3), 'unit': 'DOF*stage/s', 'ms_per_step': ms / nst,
            'steps': nst, 'dofs': n, 'stage_kernel_ms': kms}
        prob.close()
    except Exception as ex:
        out['remap'] = {'error': repr(ex)[:300]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--rs', type=int, default=5)
    ap.add_argument('--order', type=int, default=3)
    ap.add_argument('--nloc', type=int, default=0,
                    help='elements per direction of the periodic Cartesian brick every GPU owns (default '
                         '3 * 2^rs, the refined periodic-cube mesh); BASELINE config 4 = --order 4 --nloc 74 '
                         '--gpus 8: 148^3 elements, 405 M DOFs')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true')
    ap.add_argument('--no-dist-check', action='store_true')
    ap.add_argument('--replicas', action='store_true', help='under torchrun: independent replicas, no halo (experiment)')
    ap.add_argument('--force-dist', action='store_true',
                    help='N=1 through the decomposed path (ghost-aware kernel, no peers): isolates its overhead')
    ap.add_argument('--problem', type=int, default=0,
                    help='0: constant velocity (BASELINE config); 1: rotation (velocity linear in x: '
                         'exercises the general FP64 tensor-core kernel)')
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    nloc = a.nloc if a.nloc > 0 else 3 * 2 ** a.rs
    if a.nloc > 0:
        workload = ('3D periodic Cartesian transport, order %d hex, %d^3 elements per GPU, -ho 3 -lo 5 -fct 2 '
                    '-pa -s 3 (problem %d)' % (a.order, nloc, a.problem))
    else:
        workload = ('3D periodic-cube transport, order %d hex, -rs %d, -ho 3 -lo 5 -fct 2 -pa -s 3 '
                    '(problem %d)' % (a.order, a.rs, a.problem))
    metric = 'DOF*RK-stage updates/sec (3D hex, order 3, FCT)'

    if a.impl == 'reference':
        # the reference's MPI CPU build cannot be produced here (needs MFEM/hypre/METIS/MPI,
        # SURVEY.md 8c): the reference arm times the CPU oracle port on the host cores, same -rs
        # as the GPU arm, all host threads (set explicitly: torchrun exports OMP_NUM_THREADS=1)
        if rank != 0:
            return
        cb = cpu_port_run(a.order, a.rs, a.problem, steps=a.steps, warmup=min(a.warmup, 2), budget_s=75.0)
        line = {'impl': 'reference', 'metric': metric, 'value': cb['value'], 'unit': 'DOF*stage/s',
                'n_gpus': a.gpus, 'steps': cb['steps_timed'], 'warmup': a.warmup,
                'ms_per_step': cb['ms_per_step'],
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
                'data': 'synthetic', 'config': {'workload': workload, 'sample': cb['sample'],
                                                'note': 'one host, %d threads, independent of --gpus' % cb['cores']},
                'cpu_baseline': cb,
                'e2e': {'value': cb['value'], 'unit': 'DOF*stage/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import remhos_b200 as rb
    from remhos_b200.setup_problem import Problem

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (remhos_b200 has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL's debug output (version banner included) -> stderr
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    if a.replicas and world > 1:
        # experiment: N independent single-rank problems side by side (process group initialised, no halo):
        # isolates what merely running next to busy peers costs; every rank prints its own line
        world, rank = 1, 0
    dist_err, dist_cases = None, None
    if world > 1 and not a.no_dist_check:
        dist_err, dist_cases = dist_parity_check(rank, world, local_rank)

    h = 2.0 / nloc
    dt = 0.25 * h / a.order            # fixed dt = 0.25 h/|v| /p, |v| = 1
    if world == 1 and not a.force_dist:
        if a.nloc > 0:
            mesh = rb.Mesh.cartesian([nloc] * 3, [2.0, 2.0, 2.0], origin=[-1.0, -1.0, -1.0], periodic=True)
        else:
            mesh = rb.Mesh.cartesian([3, 3, 3], [2.0, 2.0, 2.0], origin=[-1.0, -1.0, -1.0], periodic=True)
            mesh.refine(a.rs)
        prob = Problem(mesh, problem=a.problem, order=a.order, mesh_order=2, bounds_type=0, dt=dt,
                       device=local_rank)
        ctx = prob.ctx

        def step(t, u, stream):
            return ctx.rk_step(3, 5, t, dt, u, stream)

        def step_host(t, uh):
            return ctx.rk_step_host(3, 5, t, dt, uh.data_ptr())

        def step_host_async(t, uh):
            return ctx.rk_step_host_async(3, 5, t, dt, uh.data_ptr())
        pdims = [1, 1, 1]
    else:
        # weak scaling: every rank owns a (3*2^rs)^3 brick of a periodic box that grows with the
        # rank count; recursive coordinate bisection of the global Cartesian mesh yields the bricks
        from remhos_b200.dist import DistProblem
        pdims = {1: [1, 1, 1], 2: [2, 1, 1], 4: [2, 2, 1], 8: [2, 2, 2]}.get(world)
        if pdims is None:
            raise SystemExit('bench.py: --gpus must be 1, 2, 4 or 8')
        mesh = rb.Mesh.cartesian([nloc * d for d in pdims], [2.0 * d for d in pdims],
                                 origin=[-1.0 * d for d in pdims], periodic=True)
        prob = DistProblem(mesh, rank, world, problem=a.problem, order=a.order, mesh_order=2,
                           bounds_type=0, dt=dt, device=local_rank)
        del mesh
        ctx = prob.ctx

        def step(t, u, stream):
            return prob.dist.rk_step(3, 5, t, dt, u, stream)

        def step_host(t, uh):
            return prob.dist.rk_step_host(3, 5, t, dt, uh.data_ptr())

        def step_host_async(t, uh):
            return prob.dist.rk_step_host_async(3, 5, t, dt, uh.data_ptr())
    # the time loop below never touches the state between steps (neither does the reference's,
    # remhos.cpp:1146-1330): the element min/max the last stage computes for its output are
    # reused by the next step instead of a separate pass over the state
    ctx.trust_state(True)
    N = ctx.ndofs
    u = torch.tensor(prob.u0, device='cuda')
    m = torch.empty(N, dtype=torch.float64, device='cuda')
    ctx.lumped_mass(m)

    def gred(v, op='sum'):
        return v if world == 1 else prob.allreduce(v, op)      # ncclAllReduce under the C ABI
    mass0 = gred(ctx.reduce(0, u, m))
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    t = 0.0
    for _ in range(a.warmup):
        t = step(t, u, stream)
    barrier()
    rb.launch_count(reset=True)
    if world > 1:
        ctx.halo_wait_stats(reset=True)
    ctx.profile(1)
    t_wall0 = time.time()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        t = step(t, u, stream)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = rb.launch_count(reset=True)
    kms, klaunch = ctx.profile(0)
    hw = None
    if world > 1:
        st = ctx.halo_wait_stats()
        hwt = torch.tensor([st[k] for k in ('warps_waited', 'wait_ns_sum', 'wait_ns_max', 'warps_shell', 'shell_ns_sum',
                                            'warps', 'run_ns_sum')], dtype=torch.float64, device='cuda')
        hwl = [torch.empty_like(hwt) for _ in range(world)]
        dist.all_gather(hwl, hwt)
        hw = [[float(x) for x in h_] for h_ in hwl]
    mass1 = gred(ctx.reduce(0, u, m))
    umin, umax = gred(ctx.reduce(1, u), 'min'), gred(ctx.reduce(2, u), 'max')

    # end-to-end: state in pinned host memory, H2D + step + D2H every step, three ways:
    #  serial      rmh_rk_step_host, one blocking call per step (copy, stages, copy back to back);
    #  pipelined   rmh_rk_step_host_async on the SAME host buffer: every step still uploads its input and
    #              downloads its result in full, but slab k of the next upload follows slab k of the previous
    #              download, so both directions of the link are busy (the headline e2e);
    #  fields      the same call on three independent host states in turn (several fields transported by the
    #              same velocity): upload, stages and download of consecutive calls overlap fully.
    uh = u.cpu().pin_memory()
    e_steps = max(3, min(a.steps // 2, 30))

    def timed(fn, n):
        barrier()
        t0 = time.perf_counter()
        fn(n)
        barrier()
        return (time.perf_counter() - t0) * 1e3

    def run_serial(n):
        for _ in range(n):
            step_host(t, uh)

    def run_pipe(n):
        for _ in range(n):
            step_host_async(t, uh)
        ctx.host_sync()

    run_serial(2)
    es_ms = timed(run_serial, e_steps)
    # the pipelined steps must produce what the serial ones do: same start, same number of steps
    ua = uh.clone().pin_memory(); ub = uh.clone().pin_memory()
    for _ in range(3):
        step_host(t, ua)
    for _ in range(3):
        step_host_async(t, ub)
    ctx.host_sync()
    pipe_ok = bool(torch.equal(ua, ub))
    run_pipe(2)
    e_ms = timed(run_pipe, e_steps)
    fields = [uh, ua, ub]

    def run_fields(n):
        for k in range(n):
            step_host_async(t, fields[k % 3])
        ctx.host_sync()

    run_fields(3)
    ef_ms = timed(run_fields, e_steps)
    del ua, ub
    clocks = sampler.stop(t_wall0, time.time())

    tmax = torch.tensor([ms, e_ms, es_ms, ef_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms, e_ms, es_ms, ef_ms = (float(v) for v in tmax)
    total_dofs = N * world
    value = total_dofs * STAGES * a.steps / (ms * 1e-3)
    e2e = total_dofs * STAGES * e_steps / (e_ms * 1e-3)
    peak, which = read_peaks()
    k_ms = kms / max(klaunch, 1)
    # which fused stage kernel ran (rmh_ctx_path_flags): bit 3 = constant-coefficient kernel
    # (affine elements, element-wise constant velocity: stage3c.cuh), bit 4 = with the overlap
    # bounds formed in the kernel; else the FP64 DMMA kernel
    flags = ctx.path_flags
    if flags & 8:
        kid = 'k_stage3c_fold' if flags & 16 else 'k_stage3c'
        kname = 'k_stage3c<%d,...,%s,%s> (constant-coefficient line kernel, %d elements per warp)' % (
            a.order + 1, 'ghosts' if world > 1 else 'local', 'fold' if flags & 16 else 'entities',
            {1: 8, 2: 2, 3: 2, 4: 1}.get(a.order, 1))
    else:
        kid = 'k_stage3w'
        kname = 'k_stage3w<%d,%d> (FP64 DMMA, warp per element)' % (a.order + 1, a.order + 3)
    traffic, tsrc = read_traffic('%s|o%d|rs%d|p%d' % (kid, a.order, a.rs, a.problem)) if a.nloc == 0 else (None, None)
    roof = {'bound': 'hbm', 'kernel': kname, 'peak': peak, 'peak_source': which, 'unit': 'GB/s',
            # the kernel's own DRAM traffic (ncu) over its own time (CUDA events, this run)
            'traffic': traffic, 'traffic_unit': 'bytes per launch', 'traffic_source': tsrc,
            'achieved': (traffic / (k_ms * 1e-3) / 1e9) if (traffic and klaunch) else None,
            'frac': (traffic / (k_ms * 1e-3) / 1e9 / peak) if (traffic and klaunch) else None,
            'traffic_bytes_per_dof': (traffic / N) if traffic else None,
            # SURVEY.md 8d's algorithmic figure: 136 B/DOF*stage of five separate kernels -- what an
            # unfused implementation at HBM speed would need; > 1 means faster than that bound
            'alg_bytes_per_dof_stage': B_ALG,
            'frac_vs_unfused_136B': (B_ALG * N / (k_ms * 1e-3) / 1e9 / peak) if klaunch else None,
            'frac_of_fused_lower_bound_56B': (56.0 * N / (k_ms * 1e-3) / 1e9 / peak) if klaunch else None,
            'kernel_ms': k_ms, 'kernel_launches': int(klaunch),
            'kernel_share_of_step': (kms / ms) if klaunch else None}
    line = {
        'metric': metric, 'value': value, 'unit': 'DOF*stage/s', 'n_gpus': world,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms / a.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': workload, 'dofs_per_gpu': N, 'elements_per_gpu': ctx.ne,
                   'stages_per_step': STAGES, 'dt': dt,
                   'l2': 'state vectors (%d MB each) exceed the 126 MB L2' % (8 * N // 10 ** 6),
                   'problem': a.problem,
                   'parallelism': ('domain decomposition %dx%dx%d bricks; per stage one put kernel (face '
                                   'traces + (min,max) pairs stored into the peers\' windows over NVLink) '
                                   'and one stage kernel that waits for the halo before its shell elements'
                                   % tuple(pdims)) if world > 1 else 'single GPU'},
        'e2e': {'value': e2e, 'unit': 'DOF*stage/s', 'h2d_bytes_per_step': 8 * N,
                'd2h_bytes_per_step': 8 * N, 'steps': e_steps,
                'api': 'rmh_rk_step_host_async on one pinned host state + rmh_host_sync: every step uploads its '
                       'input and downloads its result in full; upload of step n+1 follows the download of '
                       'step n slab by slab (both PCIe directions busy)',
                'ms_per_step': e_ms / e_steps,
                'pipelined_equals_serial': pipe_ok,
                'serial_value': total_dofs * STAGES * e_steps / (es_ms * 1e-3),
                'serial_api': 'rmh_rk_step_host, one blocking call per step (H2D, stages, D2H back to back)',
                'independent_fields_value': total_dofs * STAGES * e_steps / (ef_ms * 1e-3),
                'independent_fields_api': 'rmh_rk_step_host_async on three independent pinned states in turn',
                'link_GBps_each_way': 8 * N / (e_ms / e_steps * 1e-3) / 1e9},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': roof,
        'check': {'mass_rel_drift': abs(mass1 - mass0) / abs(mass0),
                  'mass_rel_drift_per_stage': abs(mass1 - mass0) / abs(mass0) / (STAGES * a.steps),
                  'mass_note': 'ClipScale rescales an element only when |sum of clipped fluxes| > 1e-15 '
                               '(remhos_fct.cpp:517-538, absolute threshold): in the far field of the bump '
                               '(u < 1e-9, most elements) the fluxes are below it and stay unbalanced -- the '
                               'reference algorithm\'s own defect, about 1e-14 of the mass per stage here; '
                               'the C port of the path drifts at the same rate',
                  'u_min': umin, 'u_max': umax},
    }
    if world > 1:
        # in-kernel halo waits of rank 0 over the timed steps: warps that found a peer's flag unpublished
        line['halo_wait'] = {'per_rank': [{'warps_waited': int(h_[0]), 'mean_wait_us': (h_[1] / h_[0] / 1e3) if h_[0] else 0.0,
                                           'longest_wait_us': h_[2] / 1e3,
                                           'mean_warp_run_us': (h_[6] / h_[5] / 1e3) if h_[5] else None,
                                           'mean_warp_shell_phase_us': (h_[4] / h_[3] / 1e3) if h_[3] else None}
                                          for h_ in hw],
                             'warps_per_rank': int(hw[0][5]),
                             'note': 'a warp waits when it reaches its first shell group before every peer has '
                                     'published the halo of this stage; shell phase = from that point to the '
                                     'end of the warp'}
        line['check']['dist_rel_err'] = dist_err
        line['check']['dist_cases'] = dist_cases
        if dist_err is not None and not (dist_err < 1e-12):
            line['check']['dist_FAILED'] = True
    if world == 1:
        prob.close()
        del u, m, prob
        torch.cuda.empty_cache()
        if not a.no_extras:
            line['extra'] = extra_runs(a.order, a.rs, local_rank)
        if rank == 0 and not a.no_cpu_baseline:
            cb = cpu_port_run(a.order, min(a.rs, 4), a.problem, steps=10 ** 6, warmup=1, budget_s=12.0)
            line['cpu_baseline'] = cb
    else:
        prob.close()
    if rank == 0:
        print(json.dumps(line))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
