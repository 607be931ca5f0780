import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from helpers import oracle_run, ctx_from_oracle, rel_err
from remhos_oracle import dg
mesh, opt = sys.argv[1], eval(sys.argv[2])
run = oracle_run(mesh, ho_type=3, lo_type=1, fct_type=1, **opt)
ctx = ctx_from_oracle(run); ctx.fa_setup()
sp = run.space; d = run.disc; A = d.cur
K, KH, M, BI = ctx.fa_get(0), ctx.fa_get(1), ctx.fa_get(2), ctx.fa_get(3)
print('K', rel_err(K, A.K), 'M', rel_err(M, A.M))
# natural order -> BdrDofs order
lat = dg.dof_lattice(sp.p, sp.dim)
from remhos_oracle.mesh import FACE_AXIS
for f in range(sp.nf):
    axis, side = FACE_AXIS[sp.dim][f]
    rem = [a for a in range(sp.dim) if a != axis]
    l = lat[sp.bd[:, f]][:, rem]
    nat = sum(l[:, m] * (sp.p + 1) ** m for m in range(len(rem)))
    ref = A.bdrInt[:, f]
    mine = BI[:, f][:, nat][:, :, nat]
    print('BI face', f, rel_err(mine, ref))
rng = np.random.default_rng(20260102)
u = run.u + 0.05 * rng.standard_normal(run.u.shape)
dt = 0.01
du_ho = d.ho_local_inverse(u); du_lo = d.lo_discrete_upwind(u)
umin, umax = d.bounds(u, 0)
ref = d.fct_flux_based(u, A.ml, du_ho, du_lo, umin, umax, dt)
dev = lambda a: torch.tensor(np.ascontiguousarray(a).reshape(-1), device='cuda')
out = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
ctx.fct_flux_based(dt, dev(u), dev(A.ml), dev(du_ho), dev(du_lo), dev(umin), dev(umax), out)
g = out.cpu().numpy().reshape(u.shape)
err = np.abs(g - ref) / np.abs(ref).max()
print('flux err', err.max(), 'n bad', (err > 1e-10).sum(), 'of', err.size)
bad = np.argwhere(err > 1e-10)
print('bad local dofs histogram', np.bincount(bad[:, 1], minlength=sp.nd))
# dense KH check vs oracle's sparse K_HO
I, J, kij, kji, same, Mij = d.build_sparse_K_HO()
nd = sp.nd
e = I[same] // nd
print('KH ij', np.abs(KH[e, I[same] % nd, J[same] % nd] - kij[same]).max(), np.abs(KH[e, J[same] % nd, I[same] % nd] - kji[same]).max())
print('pairs same', same.sum(), 'expected', run.mesh.ne * nd * (nd - 1) // 2, 'cross', (~same).sum())
