"""Queued host-state stepping (rmh_rk_step_host_async / rmh_host_sync): H2D, the stages and D2H of
consecutive calls overlap on three streams.  The result must be bit-identical to the blocking entry
point rmh_rk_step_host (same kernels, same order of operations) -- for one host buffer stepped
repeatedly (slab-wise dependence between download and next upload), for independent buffers in turn,
and with separate input and output buffers.  ODESolver::Step on host-resident vectors,
remhos.cpp:1146-1180."""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from helpers import DATA


def make(order=2, rs=1, problem=0, bt=0):
    import remhos_b200 as rb
    from remhos_b200.setup_problem import Problem
    mesh = rb.Mesh.load(os.path.join(DATA, 'periodic-cube.mesh')).refine(rs)
    h = 2.0 / (3 * 2 ** rs)
    dt = 0.25 * h / order
    prob = Problem(mesh, problem=problem, order=order, mesh_order=2, bounds_type=bt, dt=dt)
    return prob, dt


@pytest.mark.parametrize('order,problem,bt,trust', [(2, 0, 0, True), (3, 0, 0, False), (3, 1, 0, True), (1, 0, 1, False)])
def test_same_buffer_matches_blocking_steps(order, problem, bt, trust):
    prob, dt = make(order, 1, problem, bt)
    ctx = prob.ctx
    ctx.trust_state(trust)
    ref = torch.tensor(prob.u0).pin_memory()
    got = ref.clone().pin_memory()
    steps = 7                      # more than the ring of device buffers
    t = 0.0
    for _ in range(steps):
        t = ctx.rk_step_host(3, 5, t, dt, ref.data_ptr())
    t = 0.0
    for _ in range(steps):
        t = ctx.rk_step_host_async(3, 5, t, dt, got.data_ptr())
    ctx.host_sync()
    assert torch.equal(got, ref)
    assert not np.array_equal(ref.numpy(), prob.u0.reshape(-1))
    prob.close()


def test_independent_fields_and_separate_output():
    prob, dt = make(2, 1, 0, 0)
    ctx = prob.ctx
    ctx.trust_state(True)
    rng = np.random.default_rng(5)
    base = prob.u0.reshape(-1)
    fields = [torch.tensor(base * (0.5 + 0.25 * k) + 0.01 * rng.random(base.size)).pin_memory() for k in range(4)]
    ref = [f.clone().pin_memory() for f in fields]
    for f in ref:
        for _ in range(2):
            ctx.rk_step_host(3, 5, 0.0, dt, f.data_ptr())
    # round robin over the fields, two steps each, results into the same buffers
    for _ in range(2):
        for f in fields:
            ctx.rk_step_host_async(3, 5, 0.0, dt, f.data_ptr())
    ctx.host_sync()
    for f, r in zip(fields, ref):
        assert torch.equal(f, r)
    # separate output: in -> out, then out -> in again (a chain through two host buffers)
    a = torch.tensor(base).pin_memory(); b = torch.zeros_like(a).pin_memory()
    c = a.clone().pin_memory()
    for _ in range(4):
        ctx.rk_step_host(3, 5, 0.0, dt, c.data_ptr())
    for _ in range(2):
        ctx.rk_step_host_async(3, 5, 0.0, dt, a.data_ptr(), b.data_ptr())
        ctx.rk_step_host_async(3, 5, 0.0, dt, b.data_ptr(), a.data_ptr())
    ctx.host_sync()
    assert torch.equal(a, c)
    # the blocking call still works afterwards and the pipeline restarts cleanly
    ctx.rk_step_host(3, 5, 0.0, dt, c.data_ptr())
    ctx.rk_step_host_async(3, 5, 0.0, dt, a.data_ptr())
    ctx.host_sync()
    assert torch.equal(a, c)
    prob.close()


def test_null_buffer_is_an_error():
    import remhos_b200 as rb
    prob, dt = make(1, 0, 0, 0)
    with pytest.raises(rb.RmhError, match='null host buffer'):
        prob.ctx.rk_step_host_async(3, 5, 0.0, dt, 0)
    prob.close()


def test_halo_wait_stats_shape():
    """rmh_halo_wait_stats on a context without peers: the seven counters, all readable and resettable"""
    prob, dt = make(1, 0, 0, 0)
    st = prob.ctx.halo_wait_stats(reset=True)
    assert set(st) == {'warps_waited', 'wait_ns_sum', 'wait_ns_max', 'warps_shell', 'shell_ns_sum', 'warps', 'run_ns_sum'}
    u = torch.tensor(prob.u0, device='cuda')
    prob.ctx.rk_step(3, 5, 0.0, dt, u)
    st = prob.ctx.halo_wait_stats()
    assert all(v == 0 for v in st.values()), 'no ghost-aware launch, no halo wait on a single GPU'
    prob.close()
