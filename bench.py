#!/usr/bin/env python
"""bench.py -- DOF*RK-stage updates/s of the Remhos RK-stage path on B200 (BASELINE.json metric).

Workload (config C2): 3D periodic cube [-1,1]^3, 3x3x3 coarse hexes refined
`--rs` times (default 5 -> 884 736 elements), order 3 (56.6 M DOFs per GPU), mesh order 2,
problem 0 (erfc bump, constant velocity), `-ho 3 -lo 5 -fct 2 -pa -s 3`: LocalInverse HO +
MassBasedAvg LO + ClipScale FCT, RK3-SSP.  A "step" is one RK3 time step = 3 fused stage
launches; value = DOFs * 3 * steps / time, summed over ranks (weak scaling: every rank owns a
full cube of the same size).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints one JSON line (see the task contract): value (state resident in HBM), e2e (host state,
H2D + D2H inside the timed region, through rmh_rk_step_host), roofline of the fused stage
kernel, cpu_baseline (the CPU oracle timed on the host cores on a bounded sample).
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

B_ALG = 136.0          # algorithmic bytes per DOF*stage (five kernels)
STAGES = 3             # RK3-SSP


def read_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


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


def cpu_baseline(order, seconds=12.0, rs=3):
    """The C/OpenMP port of the stage path (oracle/c/remhos_stage.c: same algorithm as the
    reference's `-ho 3 -lo 5 -fct 2 -pa -s 3` path, sum-factorised, all host threads) on the same
    workload shrunk to -rs `rs`, timed on this box's host cores.  The reference's own MFEM/MPI
    build cannot be produced in this image, so this is a port, not the reference."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import remhos_b200 as rb
    from remhos_b200.setup_problem import Problem
    from remhos_oracle.cport import Port
    mesh = rb.Mesh.cartesian([3, 3, 3], [2.0, 2.0, 2.0], origin=[-1.0, -1.0, -1.0], periodic=True)
    mesh.refine(rs)
    h = 2.0 / (3 * 2 ** rs)
    dt = 0.25 * h / order
    prob = Problem(mesh, problem=0, order=order, mesh_order=2, bounds_type=0, dt=dt,
                   create_ctx=False)
    i = prob.inputs
    port = Port(order, 2, 0, i['nodes'], i['nbr_dof'], i['lat'], i['n_ent'],
                vel_nodes=i.get('vel_nodes'), vel_quad=i.get('vel_quad'), vel_face=i.get('vel_face'))
    u = np.ascontiguousarray(prob.u0, dtype=np.float64).copy()
    n = u.size
    m = port.lumped_mass().reshape(-1)
    mass0 = float((m * u).sum())
    port.rk3_step(0.0, dt, u)               # warm-up
    t0 = time.perf_counter()
    steps = 0
    while True:
        port.rk3_step(0.0, dt, u)
        steps += 1
        if time.perf_counter() - t0 > seconds:
            break
    el = time.perf_counter() - t0
    drift = abs(float((m * u).sum()) - mass0) / abs(mass0)
    cores = Port.threads()
    port.close()
    return {'value': n * STAGES * steps / el, 'unit': 'DOF*stage/s', 'cores': int(cores),
            'kind': 'port',
            'sample': 'C/OpenMP port of the stage path (oracle/c), periodic cube -rs %d order %d '
                      '(%d DOFs), %d RK3 steps in %.1f s, %d threads, mass drift %.1e'
                      % (rs, order, n, steps, el, cores, drift)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--rs', type=int, default=5)
    ap.add_argument('--order', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--problem', type=int, default=0,
                    help='0: constant velocity (BASELINE config); 1: rotation (velocity linear in x: '
                         'exercises the general FP64 tensor-core kernel)')
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    workload = ('3D periodic-cube transport, order %d hex, -rs %d, -ho 3 -lo 5 -fct 2 -pa -s 3 '
                '(problem %d)' % (a.order, a.rs, a.problem))
    metric = 'DOF*RK-stage updates/sec (3D hex, order 3, FCT)'

    if a.impl == 'reference':
        # the reference's MPI CPU build cannot be produced here (needs MFEM/hypre/METIS/MPI,
        # SURVEY.md 8c): the reference arm times the CPU oracle port on the host cores
        if rank != 0:
            return
        cb = cpu_baseline(a.order, seconds=max(5.0, 2.0 * a.steps), rs=min(a.rs, 4))
        line = {'impl': 'reference', 'metric': metric, 'value': cb['value'], 'unit': 'DOF*stage/s',
                'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': None,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
                'data': 'synthetic', 'config': {'workload': workload, 'sample': cb['sample']},
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

    h = 2.0 / (3 * 2 ** a.rs)
    dt = 0.25 * h / a.order            # fixed dt = 0.25 h/|v| /p, |v| = 1
    if world == 1:
        mesh = rb.Mesh.cartesian([3, 3, 3], [2.0, 2.0, 2.0], origin=[-1.0, -1.0, -1.0], periodic=True)
        mesh.refine(a.rs)
        prob = Problem(mesh, problem=a.problem, order=a.order, mesh_order=2, bounds_type=0, dt=dt,
                       device=local_rank)
        ctx = prob.ctx
        # the time loop below never touches the state between steps (neither does the reference's,
        # remhos.cpp:1146-1330): the element min/max the last stage computes for its output are
        # reused by the next step instead of a separate pass over the state
        ctx.trust_state(True)

        def step(t, u, stream):
            return ctx.rk_step(3, 5, t, dt, u, stream)
        pdims = [1, 1, 1]
    else:
        # weak scaling: every rank owns a (3*2^rs)^3 brick of a periodic box that grows with the
        # rank count; recursive coordinate bisection of the global Cartesian mesh yields the bricks
        from remhos_b200.dist import DistProblem
        pdims = {2: [2, 1, 1], 4: [2, 2, 1], 8: [2, 2, 2]}.get(world)
        if pdims is None:
            raise SystemExit('bench.py: --gpus must be 1, 2, 4 or 8')
        nloc = 3 * 2 ** a.rs
        mesh = rb.Mesh.cartesian([nloc * d for d in pdims], [2.0 * d for d in pdims],
                                 origin=[-1.0 * d for d in pdims], periodic=True)
        prob = DistProblem(mesh, rank, world, problem=a.problem, order=a.order, mesh_order=2,
                           bounds_type=0, dt=dt, device=local_rank)
        del mesh
        ctx = prob.ctx
        prob.trust_state = True      # as in the single-GPU loop: the state is not touched between steps

        def step(t, u, stream):
            return prob.rk3_step(t, u, stream)
    N = ctx.ndofs
    u = torch.tensor(prob.u0, device='cuda')
    m = torch.empty(N, dtype=torch.float64, device='cuda')
    ctx.lumped_mass(m)

    def gsum(v, op='sum'):
        if world == 1:
            return v
        tt = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op={'sum': dist.ReduceOp.SUM, 'min': dist.ReduceOp.MIN,
                                'max': dist.ReduceOp.MAX}[op])
        return float(tt[0])
    mass0 = gsum(ctx.reduce(0, u, m))
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
    mass1 = gsum(ctx.reduce(0, u, m))
    umin, umax = gsum(ctx.reduce(1, u), 'min'), gsum(ctx.reduce(2, u), 'max')

    # end-to-end: state in pinned host memory, H2D + step + D2H every step
    uh = u.cpu().pin_memory()

    def step_host(t):
        if world == 1:
            return ctx.rk_step_host(3, 5, t, dt, uh.data_ptr())
        u.copy_(uh, non_blocking=True)
        prob._xe_for = None          # fresh state from the host: recompute its element min/max
        t = step(t, u, stream)
        uh.copy_(u, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return t
    for _ in range(2):
        step_host(t)
    barrier()
    t0 = time.perf_counter()
    e_steps = max(3, a.steps // 2)
    for _ in range(e_steps):
        step_host(t)
    barrier()
    e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop(t_wall0, time.time())

    tmax = torch.tensor([ms, e_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms, e_ms = float(tmax[0]), float(tmax[1])
    total_dofs = N * world
    value = total_dofs * STAGES * a.steps / (ms * 1e-3)
    e2e = total_dofs * STAGES * e_steps / (e_ms * 1e-3)
    peak, which = read_peaks()
    k_ms = kms / max(klaunch, 1)
    achieved = B_ALG * N / (k_ms * 1e-3) / 1e9 if klaunch else None
    # which fused stage kernel ran (rmh_ctx_path_flags): bit 3 = constant-coefficient kernel
    # (affine elements, element-wise constant velocity: stage3c.cuh), else the FP64 DMMA kernel
    const_op = bool(ctx.path_flags & 8)
    if const_op:
        kname = 'k_stage3c<%d> (constant-coefficient line kernel, %d elements per warp)' % (
            a.order + 1, {1: 8, 2: 2, 3: 2, 4: 1}.get(a.order, 1))
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this
        # workload (profiles/r01/ncu_stage_v12_const_rs5.txt)
        traffic = (1.255223e9 + 441.227776e6) if (a.order == 3 and a.rs == 5) else None
    else:
        kname = 'k_stage3w<%d,%d,%s> (FP64 DMMA, warp per element)' % (
            a.order + 1, a.order + 3, '8,2' if a.order <= 3 else '6,2')
        # profiles/r01/ncu_stage_v9_hoisted_rs5.txt
        traffic = (7.000267e9 + 458.27968e6) if (a.order == 3 and a.rs == 5) else None
    line = {
        'metric': metric, 'value': value, 'unit': 'DOF*stage/s', 'n_gpus': world,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms / a.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': workload, 'dofs_per_gpu': N, 'elements_per_gpu': ctx.ne,
                   'stages_per_step': STAGES, 'dt': dt,
                   'l2': 'state vectors (453 MB each at -rs 5) exceed the 126 MB L2',
                   'problem': a.problem,
                   'parallelism': ('domain decomposition %dx%dx%d bricks, NCCL halo exchange per stage'
                                   % tuple(pdims)) if world > 1 else 'single GPU'},
        'e2e': {'value': e2e, 'unit': 'DOF*stage/s', 'h2d_bytes_per_step': 8 * N,
                'd2h_bytes_per_step': 8 * N, 'steps': e_steps},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {'bound': 'hbm',
                     'kernel': kname,
                     'achieved': achieved, 'peak': peak, 'peak_source': which, 'unit': 'GB/s',
                     'frac': (achieved / peak) if achieved else None,
                     'traffic': traffic, 'traffic_unit': 'bytes per launch',
                     # the fused kernel moves far less than the 136 B/DOF of five separate kernels:
                     # its own DRAM traffic over its own time, as a fraction of the HBM peak
                     'traffic_frac_of_peak': (traffic / (k_ms * 1e-3) / 1e9 / peak) if traffic else None,
                     # the same against SURVEY.md 8d's fused lower bound (56 B/DOF*stage: the "stretch" figure)
                     'frac_of_fused_lower_bound_56B': (56.0 * N / (k_ms * 1e-3) / 1e9 / peak) if klaunch else None,
                     'alg_bytes_per_dof_stage': B_ALG, 'kernel_ms': k_ms,
                     'kernel_share_of_step': (kms / ms) if klaunch else None},
        'check': {'mass_rel_drift': abs(mass1 - mass0) / abs(mass0), 'u_min': umin, 'u_max': umax},
    }
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(a.order)
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
