#!/usr/bin/env python
"""Remap-mode throughput (BASELINE.json configs[2]): unit cube, Taylor-Green mesh motion (-p 10),
order 3, -ho 3 -lo 5 -fct 2 -pa -s 3; every stage re-assembles the quadrature data on the moved
mesh and solves the non-affine element mass systems.  usage: python tools/bench_remap.py [rs] [steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import remhos_b200 as rb
from remhos_b200.setup_problem import Problem

rs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
order = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mesh = rb.Mesh.cartesian([2, 2, 2], [1.0, 1.0, 1.0])
mesh.refine(rs)
prob = Problem(mesh, problem=10, order=order, mesh_order=2, bounds_type=0, dt=-1.0, t_final=0.5)
ctx = prob.ctx
u = torch.tensor(prob.u0, device='cuda')
m = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
ctx.lumped_mass(m)
mass0 = ctx.reduce(0, u, m)
t, dt = 0.0, prob.dt
t = ctx.rk_step(3, 5, t, dt, u)
torch.cuda.synchronize()
rb.launch_count(reset=True)
t0 = time.perf_counter()
for _ in range(steps):
    t = ctx.rk_step(3, 5, t, dt, u)
torch.cuda.synchronize()
el = time.perf_counter() - t0
ctx.set_time(t)
ctx.lumped_mass(m)
mass1 = ctx.reduce(0, u, m)
print(json.dumps({'workload': 'remap -p 10 unit cube -rs %d order %d' % (rs, order), 'dofs': ctx.ndofs,
                  'value': ctx.ndofs * 3 * steps / el, 'unit': 'DOF*stage/s', 'ms_per_step': 1e3 * el / steps,
                  'launches': rb.launch_count(), 'mass_rel_change': abs(mass1 - mass0) / abs(mass0)}))
