"""Shared test helpers: build a product Context from the oracle's set-up on the same inputs."""
import os
import numpy as np

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


def oracle_run(mesh, **kw):
    from remhos_oracle import driver
    return driver.Run(driver.Options(mesh_file=os.path.join(DATA, mesh), **kw))


def ctx_from_oracle(run, bounds_type=None, use_nodal_velocity=False):
    """Context on cuda:0 with the oracle's geometry, velocity samples and index maps."""
    import remhos_b200 as rb
    from remhos_oracle import dg
    sp, m, topo, d = run.space, run.mesh, run.topo, run.disc
    bt = run.opt.bounds_type if bounds_type is None else bounds_type
    kw = dict(dim=m.dim, order=sp.p, mesh_order=sp.g, exec_mode=run.exec_mode, bounds_type=bt,
              nodes=m.X, nbr_dof=d.nbr, lat=topo.lat, n_ent=topo.n_ent, nbr_elem=topo.nbr_elem,
              inflow=d.inflow.reshape(-1))
    if run.exec_mode == 1:
        kw['vel_nodes'] = d.Vnodes
    elif use_nodal_velocity:
        kw['vel_nodes'] = run.vel(m.X)
    else:
        kw['vel_quad'] = run.vel(sp.quad_points(m.X))
        vf = np.stack([run.vel(sp.face_quad_points(m.X, f)) for f in range(sp.nf)], axis=1)
        kw['vel_face'] = vf
    return rb.Context(**kw)


def rel_err(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def subcell_setup_from_oracle(run, ctx):
    """Hand the oracle's low-order refined mesh data (lattice points, velocity samples) to the
    context, as the driver does after building its subcell mesh (remhos.cpp:797-868)."""
    from remhos_oracle import dg, subcell
    sp = run.space
    dim = sp.dim
    xlat = sp.dof_points(run.disc.X0)
    if run.exec_mode == 1:
        vel = run.vel(xlat)
        vel = np.where(subcell._boundary_lattice_mask(run)[:, :, None], 0.0, vel)
    else:
        s2i = dg.sub2ind(sp.p, dim)
        vel = run.vel(xlat[:, s2i, :].mean(axis=2))
    ctx.subcell_setup(xlat, vel)
