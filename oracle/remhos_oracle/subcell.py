"""TEST INFRASTRUCTURE ONLY -- CPU oracle (numpy): subcell fluctuation weights of the
residual-distribution LO solver (`-lo 4`).

Restates Assembly::ComputeSubcellWeights (remhos_tools.cpp:860-874) with
MixedConvectionIntegrator::AssembleElementMatrix2 (:1033-1076) on the low-order refined mesh
the driver builds (remhos.cpp:797-868):
  * every element is split into p^dim straight-sided subcells whose vertices are the images of
    the uniform lattice points i/p (ParMesh::MakeRefined(pmesh, order, ClosedUniform) followed by
    SetCurvature(1));
  * SubcellWeights(k)(m, j) = alpha * grad_ref(phi_j)(centre) . adj(J_sub(centre)) . v(centre)
    (midpoint rule, weight 1; trial space = order-1 positive basis = multilinear vertex functions
    in lexicographic order, matching Sub2Ind; test space = constants);
  * transport: v = velocity_function at the physical centre, alpha = -1 (:859-862);
    remap: v = the Q1 field v_sub_gf = velocity_function sampled at the subcell vertices at t = 0
    and zeroed on the domain boundary (:838-853), alpha = +1 (:866-867), and the subcell vertices
    move as x0_sub + t v_sub_gf (remhos.cpp:1269-1272).
"""
import numpy as np
from . import dg


def _boundary_lattice_mask(run):
    """[NE, nd] True where the lattice point lies on a domain-boundary face."""
    sp, topo = run.space, run.topo
    mask = np.zeros((run.mesh.ne, sp.nd), dtype=bool)
    for f in range(sp.nf):
        on = topo.nbr_elem[:, f] < 0
        mask[np.ix_(on, sp.bd[:, f])] = True
    return mask


def subcell_weights(run, t=0.0):
    sp = run.space
    dim, p = sp.dim, sp.p
    s2i = dg.sub2ind(p, dim)                                  # [ns, nc] lexicographic corners
    xlat = sp.dof_points(run.disc.X0)                         # [NE, nd, dim] subcell vertices at t=0
    if run.exec_mode == 1:
        shp = xlat.shape
        vlat = run.vel(xlat.reshape(-1, dim).reshape(shp))
        vlat = np.where(_boundary_lattice_mask(run)[:, :, None], 0.0, vlat)
        xlat = xlat + t * vlat
        alpha = 1.0
    else:
        vlat = None
        alpha = -1.0
    xs = xlat[:, s2i, :]                                      # [NE, ns, nc, dim]
    nc = s2i.shape[1]
    # reference gradients of the multilinear vertex functions at the subcell centre
    cc = np.array([[(c >> a) & 1 for a in range(dim)] for c in range(nc)])      # corner coords
    dphi = (2.0 * cc - 1.0) * 0.5 ** (dim - 1)                # [nc, dim]
    J = np.einsum('emci,cj->emij', xs, dphi)                  # [NE, ns, dim(i), dim(j)]
    _, adj = sp.det_adj(J)
    if run.exec_mode == 1:
        vc = vlat[:, s2i, :].mean(axis=2)                     # Q1 interpolation at the centre
    else:
        xc = xs.mean(axis=2)
        vc = run.vel(xc)
    av = np.einsum('emij,emj->emi', adj, vc)                  # adj(J) v
    return alpha * np.einsum('cj,emj->emc', dphi, av)
