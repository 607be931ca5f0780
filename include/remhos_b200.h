/* remhos_b200 -- C ABI of the B200-native Remhos RK-stage hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The reference
 * (CEED/Remhos) has no FFI layer; its boundary is the solver classes HOSolver / LOSolver /
 * FCTSolver / AdvectionOperator.  Each entry point below names the reference
 * method it replaces (file:line relative to the Remhos source tree).  The C++ classes in
 * remhos_b200/host/ mirror those class names on top of this ABI; INTEGRATION.md shows the
 * binding a Remhos maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, nonzero on error; rmh_last_error() gives the text.
 *   - all vectors are FP64, element-major: dof = k*nd + j, j lexicographic, x fastest
 *     (remhos_fct.cpp:492, remhos_lo.cpp:153); index maps are int32.
 *   - pointers named *_dev are CUDA device pointers on the context's device; `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).
 *   - no hidden device allocation after rmh_ctx_create() on the fused path (rmh_stage, rmh_rk_*);
 *     the matrix-based solvers allocate in rmh_fa_setup, the unfused rmh_mult / rmh_ode_step
 *     allocate their work vectors on first use.
 *   - there is NO CPU fallback: without a CUDA device every rmh_ctx_* call fails.
 */
#ifndef REMHOS_B200_H
#define REMHOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char *rmh_last_error(void);
int rmh_version(void);

/* ------------------------------------------------------------------------------------------
 * Host mesh module (CPU only; replaces what remhos.cpp:448-463,510-513 obtains from MFEM's
 * Mesh: load, uniform refinement, SetCurvature, and the topology DofInfo needs).
 * ---------------------------------------------------------------------------------------- */
typedef struct rmh_mesh rmh_mesh;

/* Mesh::LoadFromFile for "MFEM mesh v1.0" / "MFEM INLINE mesh v1.0", quads and hexes
 * (remhos.cpp:448). */
int rmh_mesh_load(const char *path, rmh_mesh **out);
/* Cartesian n[0] x n[1] (x n[2]) mesh of [origin, origin+size]; periodic != 0 identifies
 * opposite sides (the "-m default" generated mesh, remhos.cpp:451-455). */
int rmh_mesh_cartesian(int dim, const int *n, const double *origin, const double *size,
                       int periodic, rmh_mesh **out);
int rmh_mesh_free(rmh_mesh *m);
/* Mesh::UniformRefinement (remhos.cpp:449,463). */
int rmh_mesh_refine(rmh_mesh *m, int levels);
/* Mesh::SetCurvature(order, periodic) (remhos.cpp:513). */
int rmh_mesh_set_curvature(rmh_mesh *m, int order);
/* Mesh::GetBoundingBox (remhos.cpp:457). */
int rmh_mesh_bounding_box(const rmh_mesh *m, double *bb_min, double *bb_max);
int rmh_mesh_dim(const rmh_mesh *m);
int rmh_mesh_ne(const rmh_mesh *m);
int rmh_mesh_nv(const rmh_mesh *m);
int rmh_mesh_geom_order(const rmh_mesh *m);
/* element-wise nodal coordinates [ne][(g+1)^dim][dim], Gauss-Lobatto lattice, x fastest */
const double *rmh_mesh_nodes(const rmh_mesh *m);
/* element vertices [ne][2^dim], lexicographic corner order */
const int64_t *rmh_mesh_elem_vertices(const rmh_mesh *m);
/* Keep only the listed elements (used by the domain decomposition); ids are global. */
int rmh_mesh_extract(const rmh_mesh *m, int64_t n, const int64_t *elem_ids, rmh_mesh **out);
/* Mesh::MakeRefined(mesh, factor, BasisType::ClosedUniform) (remhos.cpp:801): every element split into
 * factor^dim linear sub-elements on the uniform lattice; shared lattice points become shared vertices
 * (periodic meshes stay periodic).  nodes == NULL: the mesh's own nodes, else [ne][(g+1)^dim][dim].
 * This is the subcell mesh -save writes as meshLO_*.mesh (remhos.cpp:1021-1026,1371-1376). */
int rmh_mesh_make_refined(const rmh_mesh *m, int factor, const double *nodes, rmh_mesh **out);

/* On-disk formats (-save, -visit: remhos.cpp:1016-1043,1366-1380).  rmh_mesh_save = Mesh::Print in
 * "MFEM mesh v1.0" with the nodes as the element-wise Gauss-Lobatto field L2_T1_<dim>D_P<g> (nodes ==
 * NULL: the mesh's own; else e.g. the moved mesh of a remap run); readable by MFEM / GLVis and by
 * rmh_mesh_load.  rmh_gf_save = GridFunction::Save of a scalar DG field (basis_type 2 = Positive). */
int rmh_mesh_save(const rmh_mesh *m, const char *path, const double *nodes, int precision);
int rmh_gf_save(const char *path, int dim, int order, int basis_type, int64_t n, const double *vals,
                int precision);

/* Domain decomposition (replaces ParMesh's METIS / Cartesian partitioning, remhos.cpp:451-461):
 * recursive coordinate bisection of the element centroids; part[ne] receives the owner rank. */
int rmh_mesh_partition(const rmh_mesh *m, int nparts, int32_t *part);

/* Halo plan of one rank (replaces the face-neighbour tables behind
 * ParGridFunction::ExchangeFaceNbrData and the GroupCommunicator of DofInfo):
 * owned elements, the ghost ring (elements of other ranks sharing a vertex with an owned one,
 * ordered by owner then global id) and per peer the owned elements that peer needs. */
typedef struct rmh_halo rmh_halo;
int rmh_halo_create(const rmh_mesh *m, const int32_t *part, int rank, rmh_halo **out);
int rmh_halo_free(rmh_halo *h);
int rmh_halo_sizes(const rmh_halo *h, int64_t *n_owned, int64_t *n_ghost, int32_t *n_peers,
                   int64_t *n_send);
int rmh_halo_get(const rmh_halo *h, int64_t *owned, int64_t *ghost, int32_t *peers,
                 int32_t *send_off, int32_t *recv_off, int32_t *send_local);
/* reorder the owned elements: those no peer needs (sharing no vertex with a ghost element) first, in
 * their previous relative order; *n_interior receives their number.  Call before rmh_halo_get. */
int rmh_halo_interior_first(rmh_halo *h, int64_t *n_interior);

/* DofInfo integer maps (remhos_tools.cpp:356-379):
 *   bdr_dofs [nfd][nf]      ExtractBdrDofs        (:1356-1431)   (row-major nfd x nf)
 *   nbr_dof  [ne][nf][nfd]  FillNeighborDofs      (:525-676), -1 = domain boundary
 *   sub2ind  [p^dim][2^dim] FillSubcell2CellDof   (:678-734)
 *   lat      [ne][3^dim]    entity ids of the H1 space used by ComputeOverlapBounds (:432-495)
 *   nbr_elem [ne][nf]       face-neighbour elements used by ComputeMatrixSparsityBounds
 * Any output pointer may be NULL. */
int rmh_mesh_dof_maps(const rmh_mesh *m, int order, int32_t *bdr_dofs, int32_t *nbr_dof,
                      int32_t *sub2ind, int32_t *lat, int32_t *n_ent, int32_t *nbr_elem);
/* Neighbourhood lattice derived from `lat` (the [ne][3^dim] lattice-entity map above): nbr[e][d],
 * d = sum (1 + d_a) 3^a, d_a in {-1,0,1}: the element across the face / edge / vertex in direction d
 * (e itself for d = 0, -1 at the domain boundary).  *structured = 1 iff these neighbourhoods
 * reproduce every entity's element set, i.e. the overlap bounds of ComputeOverlapBounds
 * (remhos_tools.cpp:432-495) can be formed from 3^dim neighbour values without the entity table. */
int rmh_nbr_lattice(int dim, int64_t ne, int32_t n_ent, const int32_t *lat, int32_t *nbr, int *structured);

/* Mesh::GetElementSize(e): |det J(centre)|^(1/dim) per element (remhos.cpp:544, remhos_mono.cpp:55) */
int rmh_mesh_elem_sizes(const rmh_mesh *m, double *h_out);

/* Problem definitions and driver set-up (host): velocity_function (remhos.cpp:2001-2120),
 * u0_function (:2201-2355), inflow_function (:2363-2381); x is [n][dim]. */
int rmh_velocity(int problem, int dim, int64_t n, const double *x, const double *bb_min,
                 const double *bb_max, double *v);
int rmh_u0(int problem, int dim, int64_t n, const double *x, const double *bb_min,
           const double *bb_max, double *u);
int rmh_inflow(int problem, int dim, int64_t n, const double *x, double *u);
/* inflow_gf (remhos.cpp:625-636) in the order-p Bernstein space, [ne][(p+1)^dim]: lattice samples,
 * or -- problem 7 -- the Bernstein form of the interpolant at the tensor Gauss-Legendre points */
int rmh_inflow_project(const rmh_mesh *m, int problem, int order, double *infl_out);
/* physical coordinates of the tensor lattice pts1d[npts]^dim of every element (face < 0), or of
 * the lattice on local face `face`; out [ne][npts^d'][dim].  With pts = i/p this gives the
 * points ProjectCoefficient samples on the positive basis (remhos.cpp:883). */
int rmh_mesh_eval(const rmh_mesh *m, int npts, const double *pts1d, int face, double *out);
/* n-point Gauss-Legendre rule on [0,1] (IntRules.Get(Geometry::SEGMENT, 2n-1)) */
int rmh_gauss_legendre_01(int n, double *x, double *w);
/* CFL time step used for -dt < 0 (remhos.cpp:538-553) */
int rmh_cfl_dt(const rmh_mesh *m, int problem, const double *bb_min, const double *bb_max,
               double *dt);
/* remap mesh velocity v_gf = x_final - x0 (remhos.cpp:562-584); v_nodes like rmh_mesh_nodes */
int rmh_remap_mesh_velocity(const rmh_mesh *m, int problem, const double *bb_min,
                            const double *bb_max, double dt, double t_final, double *v_nodes);

/* ------------------------------------------------------------------------------------------
 * Device context for the RK-stage path.
 * ---------------------------------------------------------------------------------------- */
typedef struct rmh_ctx rmh_ctx;

typedef struct rmh_desc
{
   int32_t dim;          /* 2 or 3 */
   int32_t order;        /* p: DG_FECollection(order, dim, Positive), remhos.cpp:588-590 */
   int32_t mesh_order;   /* degree of the nodal geometry (remhos.cpp:222,513) */
   int32_t exec_mode;    /* 0 transport, 1 remap (remhos.cpp:438-440) */
   int32_t bounds_type;  /* 0 overlap, 1 matrix sparsity (remhos.cpp:228) */
   int32_t device;       /* CUDA device ordinal */
   int64_t ne;           /* owned elements */
   int64_t ne_ghost;     /* ghost elements whose DOF blocks follow the owned ones (halo) */
   const double *nodes;      /* host [ne][(mo+1)^dim][dim] element-wise nodal coordinates */
   const double *vel_nodes;  /* host, same shape: remap mesh velocity v_gf (remhos.cpp:561),
                                or transport velocity sampled at the nodes when
                                vel_quad == NULL (exact for velocities of degree <= mo) */
   const double *vel_quad;   /* host [ne][Q^dim][dim] transport velocity at the volume
                                quadrature points (VectorFunctionCoefficient, remhos.cpp:537),
                                may be NULL */
   const double *vel_face;   /* host [ne][nf][Q^(dim-1)][dim] same at face quadrature points */
   const int32_t *nbr_dof;   /* host [ne][nf][nfd]; ids >= ne*nd address DOFs of ghost elements (only their
                                face traces are ever exchanged: rmh_dplan_*) */
   const int32_t *lat;       /* host [ne][3^dim] (bounds_type 0), ids < n_ent */
   int32_t n_ent;
   const int32_t *nbr_elem;  /* host [ne][nf] (bounds_type 1), -1 boundary, >= ne ghost */
   const double *inflow;     /* host [ne*nd] inflow_gf (remhos.cpp:626-637) or NULL (= 0) */
} rmh_desc;

int rmh_ctx_create(const rmh_desc *desc, rmh_ctx **out);
int rmh_ctx_destroy(rmh_ctx *ctx);
int64_t rmh_ctx_ndofs(const rmh_ctx *ctx);   /* ne*nd (owned) */
int rmh_ctx_nd(const rmh_ctx *ctx);
int rmh_ctx_nq1d(const rmh_ctx *ctx);
/* which stage-kernel path this context takes (diagnostics, tests): bit 0 every element has a
 * constant Jacobian, bit 1 stored quadrature data is in tensor-core fragment order, bit 2 the
 * velocity is linear over every element (quadrature data rebuilt in-kernel from 12 doubles per
 * element instead of streamed; opt-in with RMH_LINEAR_OP=1), bit 3 the constant-coefficient stage kernel
 * applies (affine elements, element-wise constant velocity), bit 4 the overlap bounds are formed inside
 * that kernel from the 3x3x3 element neighbourhoods (structured vertex topology; RMH_NO_FOLD=1 disables) */
int rmh_ctx_path_flags(const rmh_ctx *ctx);
/* on != 0: the caller promises not to modify the state vector between consecutive rmh_rk_step()
 * calls on the same device pointer (the reference's time loop does not, remhos.cpp:1146-1330).
 * The element min/max the last RK stage computes for its output (ComputeElementsMinMax,
 * remhos_tools.cpp:497-523, fused into the stage kernel) are then reused by the next step's first
 * stage instead of being recomputed by a separate pass over the state.  Default off.  Any other
 * entry point that is handed a state vector recomputes what it needs. */
int rmh_ctx_trust_state(rmh_ctx *ctx, int on);
/* reference-element coordinates of the volume / face quadrature points, so a caller can
 * evaluate its velocity coefficient there: q1d[Q] Gauss-Legendre points on [0,1] */
int rmh_ctx_quad_points_1d(const rmh_ctx *ctx, double *q1d, double *w1d);

/* Remap: move the mesh to x0 + t*v and rebuild all quadrature data, mass inverse data and
 * the lumped mass (AdvectionOperator::MultUnlimited, remhos.cpp:1598-1677). No-op cost in
 * transport mode.  The operators depend on t only: the context keeps the data of the last two
 * distinct times, so asking for the current time again is free and an RK3 step (t, t + dt,
 * t + dt/2, next step t + dt again) rebuilds twice instead of three times.  With the
 * matrix-based solvers or subcell weights switched on, and with RMH_NO_GEOM_CACHE=1, every
 * call rebuilds. */
int rmh_set_time(rmh_ctx *ctx, double t, void *stream);

/* lumpedM (remhos.cpp:705-727; remap refresh :1625-1632) */
int rmh_lumped_mass(rmh_ctx *ctx, double *m_dev, void *stream);

/* K_HO.Mult(u, rhs) (remhos_ho.cpp:122): PA convection + transposed DG-trace face terms */
int rmh_ho_mult(rmh_ctx *ctx, const double *u_dev, double *rhs_dev, void *stream);
/* M_inv->Mult(rhs, du) (remhos_ho.cpp:126; exact-inverse semantics of :100-116) */
int rmh_mass_inv(rmh_ctx *ctx, const double *rhs_dev, double *du_dev, void *stream);
/* LocalInverseHOSolver::CalcHOSolution (remhos_ho.cpp:84-129) */
int rmh_ho_local_inverse(rmh_ctx *ctx, const double *u_dev, double *du_dev, void *stream);

/* MassBasedAvg::CalcLOSolution (remhos_lo.cpp:247-288) */
int rmh_lo_mass_avg(rmh_ctx *ctx, double dt, const double *u_dev, const double *du_ho_dev,
                    double *du_lo_dev, void *stream);
/* Assemble the dense element matrices the matrix-based ("full assembly", remhos.cpp:1088)
 * solvers need: convection blocks K (remhos.cpp:646-657,716-717), mass blocks M, face blocks
 * bdrInt (Assembly::ComputeFluxTerms, remhos_tools.cpp:788-858) and the diagonal blocks of K_HO.
 * In remap mode rmh_set_time re-assembles them afterwards (remhos.cpp:1614-1677). */
int rmh_fa_setup(rmh_ctx *ctx, void *stream);
/* copy assembled blocks to the host (what the reference reads through SparseMatrix::GetData /
 * Assembly::bdrInt): which = 0 K [ne][nd][nd], 1 diagonal blocks of K_HO, 2 M [ne][nd][nd],
 * 3 bdrInt [ne][nf][nfd][nfd], 4 its row sums [ne][nf][nfd]; face DOFs in natural order
 * (remaining axes ascending) */
int rmh_fa_get(rmh_ctx *ctx, int which, double *host_out);
/* DiscreteUpwind::CalcLOSolution (remhos_lo.cpp:43-100) with ComputeDiscreteUpwindingMatrix
 * (remhos_tools.cpp:1464-1487) and Assembly::LinearFluxLumping, alpha = 0 (:876-913).
 * Needs rmh_fa_setup. */
int rmh_lo_discrete_upwind(rmh_ctx *ctx, const double *u_dev, double *du_lo_dev, void *stream);
/* DiscreteUpwind on the preconditioned convection blocks M_L M^-1 K (-lo 2; remhos.cpp:749-771,
 * PrecondConvectionIntegrator remhos_tools.cpp:975-1031).  Needs rmh_fa_setup. */
int rmh_lo_discrete_upwind_prec(rmh_ctx *ctx, const double *u_dev, double *du_lo_dev, void *stream);
/* NeumannHOSolver::CalcHOSolution (-ho 1; remhos_ho.cpp:136-187): rhs = k u + Galerkin face terms
 * (LinearFluxLumping with alpha = 1, remhos_tools.cpp:876-913), then at most 20 Neumann sweeps
 * du -= (M du - rhs)/m_L, stopped at |res|_2 <= 1e-4.  FA only: needs rmh_fa_setup. */
int rmh_ho_neumann(rmh_ctx *ctx, const double *u_dev, double *du_dev, void *stream);
/* ResidualDistribution / PAResidualDistribution::CalcLOSolution without subcells
 * (remhos_lo.cpp:111-245, 967-1035; -lo 2 / -lo 3): matrix-free */
int rmh_lo_res_dist(rmh_ctx *ctx, const double *u_dev, double *du_lo_dev, void *stream);

/* DofInfo::ComputeElementsMinMax (remhos_tools.cpp:497-523) */
int rmh_elem_min_max(rmh_ctx *ctx, const double *u_dev, double *xe_min_dev,
                     double *xe_max_dev, void *stream);
/* DofInfo::ComputeBounds (remhos_tools.hpp:156-170 -> remhos_tools.cpp:381-495) */
int rmh_bounds(rmh_ctx *ctx, const double *xe_min_dev, const double *xe_max_dev,
               double *xi_min_dev, double *xi_max_dev, void *stream);

/* ClipScaleSolver::CalcFCTSolution (remhos_fct.cpp:449-541) */
int rmh_fct_clip_scale(rmh_ctx *ctx, double dt, const double *u_dev, const double *m_dev,
                       const double *du_ho_dev, const double *du_lo_dev,
                       const double *xi_min_dev, const double *xi_max_dev, double *du_dev,
                       void *stream);

/* Subcell residual distribution (-lo 4): ResidualDistribution::CalcLOSolution with
 * subcell_scheme = true (remhos_lo.cpp:164-239) / PAResidualDistributionSubcell (:1040-1802).
 * rmh_subcell_setup hands over what the driver builds on its low-order refined mesh
 * (remhos.cpp:797-868): xlat [ne][nd][dim] = the lattice points i/p of every element at t = 0
 * (the subcell vertices), and vel = velocity_function at the subcell centres [ne][p^dim][dim]
 * (transport) or at the lattice points with zeros on the domain boundary [ne][nd][dim] (remap,
 * v_sub_gf); the weights of Assembly::ComputeSubcellWeights (remhos_tools.cpp:860-874) are then
 * (re)computed on the device whenever the mesh moves. */
int rmh_subcell_setup(rmh_ctx *ctx, const double *xlat_host, const double *vel_host, void *stream);
int rmh_lo_res_dist_subcell(rmh_ctx *ctx, const double *u_dev, double *du_lo_dev, void *stream);

/* SmoothnessIndicator (remhos_tools.hpp; remhos_tools.cpp:24-354; created at remhos.cpp:905-911):
 * si_type 1 or 2 (-si), 0 removes it.  Any order: the H1 operators live on the subcell mesh (one dof per
 * distinct lattice point; order 1 = the configuration of the reference's monolithic-solver known
 * answers); while set, rmh_mono_rd uses it as
 * remhos_mono.cpp:132-153,300-324 do.  rmh_si_values = ComputeSmoothnessIndicator followed by the
 * DG2CG gather: one value per DG dof (1 on the domain boundary); out_dev may be NULL. */
int rmh_si_setup(rmh_ctx *ctx, int si_type, void *stream);
int rmh_si_values(rmh_ctx *ctx, const double *u_dev, double *out_dev, void *stream);

/* MonolithicSolver (remhos_mono.hpp:28-65; set up at remhos.cpp:997-1011).  mono_type 1 =
 * MonoRDSolver, 2 = with the subcell scheme (needs rmh_subcell_setup), 0 removes it.  mass_lim as at
 * remhos.cpp:999.  scale_host[ne] = vmax / (2 sqrt(dim) h_e / order) (MonoRDSolver constructor,
 * remhos_mono.cpp:40-57).  While a monolithic solver is set, rmh_mult / rmh_mult_unlimited /
 * rmh_ode_step evaluate it instead of HO/LO/FCT (remhos.cpp:1687) and rmh_limit_mult is a no-op.
 * Serial only, as in the reference (remhos_mono.cpp:283). */
int rmh_mono_setup(rmh_ctx *ctx, int mono_type, int mass_lim, const double *scale_host, void *stream);
/* MonoRDSolver::CalcSolution (remhos_mono.cpp:60-356) */
int rmh_mono_rd(rmh_ctx *ctx, const double *u_dev, double *du_dev, void *stream);

/* ElementFCTProjection::CalcFCTSolution (remhos_fct.cpp:613-733; -fct 4); after rmh_fa_setup */
int rmh_fct_project(rmh_ctx *ctx, double dt, const double *u_dev, const double *du_ho_dev,
                    const double *du_lo_dev, const double *xi_min_dev, const double *xi_max_dev,
                    double *du_dev, void *stream);
/* NonlinearPenaltySolver::CalcFCTSolution (remhos_fct.cpp:760-996; -fct 3): rate clipped to the bounds,
 * conservation restored per element by the penalised flux correction of CorrectFlux / get_lambda
 * (bisection, every sum in DOF order as the host loop; both loops capped at 200 rounds).
 * eps_w = Mesh::GetElementSize(0, 0) / order (:961). */
int rmh_fct_nonlinear_penalty(rmh_ctx *ctx, double dt, double eps_w, const double *u_dev, const double *m_dev,
                              const double *du_ho_dev, const double *du_lo_dev, const double *xi_min_dev,
                              const double *xi_max_dev, double *du_dev, void *stream);
/* SmoothnessIndicator::UpdateBounds (remhos_tools.cpp:183-190) on every dof: si_dev = rmh_si_values(u),
 * u_HO = u + dt du_HO; xi_min / xi_max relaxed in place.  While a smoothness indicator is set
 * (rmh_si_setup), rmh_limit_mult applies it in front of -fct 2 and -fct 3 (remhos_fct.cpp:498-504,780-795). */
int rmh_si_update_bounds(rmh_ctx *ctx, double dt, const double *u_dev, const double *du_ho_dev,
                         const double *si_dev, double *xi_min_dev, double *xi_max_dev, void *stream);
/* Automatic time step control, -dtc 1 (remhos.cpp:312-316): with mode 1 every rmh_limit_mult runs
 * AdvectionOperator::UpdateTimeStepEstimate (remhos.cpp:1968-1998) on the LO rate.  rmh_dt_ratio =
 * GetTimeStepRatio (minimum of dt_estimate / dt since the last reset), reset != 0 =
 * ResetTimeStepRatio; the caller repeats or grows the step as remhos.cpp:1178-1197 does. */
int rmh_dt_control(rmh_ctx *ctx, int mode);
int rmh_dt_ratio(rmh_ctx *ctx, int reset, double *ratio);

/* FluxBasedFCT::CalcFCTSolution, one FCT iteration as the driver fixes it (remhos_fct.cpp:155-181,
 * 295-446; remhos.cpp:1093).  Needs rmh_fa_setup; single-rank meshes only. */
int rmh_fct_flux_based(rmh_ctx *ctx, double dt, const double *u_dev, const double *m_dev,
                       const double *du_ho_dev, const double *du_lo_dev,
                       const double *xi_min_dev, const double *xi_max_dev, double *du_dev,
                       void *stream);

/* ------------------------------------------------------------------------------------------
 * Product-field remap (-ps; remap mode): the state is the block (u, us) of 2 N doubles
 * (BlockVector S, remhos.cpp:594-598,886-903).  Flags are bytes (Array<bool>).
 * ---------------------------------------------------------------------------------------- */
/* on != 0: rmh_mult_unlimited / rmh_limit_mult / rmh_mult / rmh_ode_step take and return (u, us)
 * blocks: MultUnlimited remaps us with the same HO operator (remhos.cpp:1714-1738), LimitMult runs its
 * second pass (:1848-1915).  "Products are processed only in remap mode." (:1850) */
int rmh_product_enable(rmh_ctx *ctx, int on);
/* RKIDPSolver::UseMask (remhos_solvers.hpp; the driver switches the masks off, remhos.cpp:502-507):
 * -s 12/13/14/16 with ComputeMask / UpdateMask / AddMasked (remhos_solvers.cpp:97-147,171-249) */
int rmh_idp_use_mask(rmh_ctx *ctx, int on);
/* AdvectionOperator::ComputeMask (remhos.cpp:1741-1796): mask_dev [state length] */
int rmh_compute_mask(rmh_ctx *ctx, const double *state_dev, uint8_t *mask_dev, void *stream);
/* ComputeBoolIndicators (remhos_sync.cpp:24-47): el_dev [ne], dof_dev [N]; EMPTY_ZONE_TOL = 1e-12 */
int rmh_prod_bool_indicators(rmh_ctx *ctx, const double *u_dev, uint8_t *el_dev, uint8_t *dof_dev, void *stream);
/* ComputeRatio (remhos_sync.cpp:50-94): s = us / u on active dofs, their average elsewhere in an
 * active element, 0 in empty elements */
int rmh_prod_compute_ratio(rmh_ctx *ctx, const double *us_dev, const double *u_dev, double *s_dev,
                           uint8_t *el_dev, uint8_t *dof_dev, void *stream);
/* DofInfo::ComputeElementsMinMax with active_el / active_dof (remhos_tools.cpp:497-523); either mask
 * may be NULL; inactive elements get (inf, -inf) and so drop out of rmh_bounds */
int rmh_elem_min_max_masked(rmh_ctx *ctx, const double *u_dev, const uint8_t *el_dev, const uint8_t *dof_dev,
                            double *xe_min_dev, double *xe_max_dev, void *stream);
/* FCTSolver::CalcCompatibleLOProduct (remhos_fct.cpp:26-118); s_min / s_max adjusted in place */
int rmh_prod_compatible_lo(rmh_ctx *ctx, double dt, const double *us_dev, const double *m_dev,
                           const double *d_us_ho_dev, double *s_min_dev, double *s_max_dev,
                           const double *u_new_dev, const uint8_t *el_dev, const uint8_t *dof_dev,
                           double *d_us_lo_new_dev, void *stream);
/* ZeroOutEmptyDofs (remhos_sync.cpp:96-114) */
int rmh_prod_zero_empty(rmh_ctx *ctx, const uint8_t *el_dev, const uint8_t *dof_dev, double *d_us_dev, void *stream);
/* FCTSolver::CalcFCTProduct of FluxBasedFCT (fct_type 1, remhos_fct.cpp:183-294), ClipScaleSolver (2,
 * :543-563), ElementFCTProjection (4, :735-758): compatible LO product, ScaleProductBounds (:120-153),
 * the solver's limiter on us, empty dofs zeroed.  d_us_lo_dev is read by fct_type 1 only
 * (NeedsLOProductInput, remhos.cpp:1865-1869). */
int rmh_fct_product(rmh_ctx *ctx, int fct_type, double dt, const double *us_dev, const double *m_dev,
                    const double *d_us_ho_dev, const double *d_us_lo_dev, double *s_min_dev, double *s_max_dev,
                    const double *u_new_dev, const uint8_t *el_dev, const uint8_t *dof_dev, double *d_us_dev,
                    void *stream);

/* LimitedTimeDependentOperator::Mult (remhos_solvers.hpp:46-50) for any supported combination of
 * -ho {0,1,3} -lo {0,1,2,3,4,5} -fct {0,1,2,3,4} at time t (remap: mesh moved to x0 + t v first,
 * remhos.cpp:1598-1677): k = F(u; t, dt).  Orchestrates the separate kernels exactly as
 * MultUnlimited / LimitMult do (remhos.cpp:1596-1739, 1798-1916); -ho 3 -lo 5 -fct 2 runs the
 * fused stage kernel. */
int rmh_mult(rmh_ctx *ctx, int ho_type, int lo_type, int fct_type, double t, double dt,
             const double *u_dev, double *k_dev, void *stream);
/* The two halves of the stage operator, as the IDP Runge-Kutta solvers call them
 * (remhos_solvers.cpp:29-38,171-249): AdvectionOperator::MultUnlimited (remhos.cpp:1596-1739) --
 * remap re-assembly at time t, then the HO rate when an FCT solver will limit it, else the LO or
 * HO rate -- and AdvectionOperator::LimitMult (remhos.cpp:1798-1916) -- k_dev holds the
 * (combined) HO rate on entry and the limited rate on exit. */
int rmh_mult_unlimited(rmh_ctx *ctx, int ho_type, int lo_type, int fct_type, double t, double dt,
                       const double *u_dev, double *k_dev, void *stream);
int rmh_limit_mult(rmh_ctx *ctx, int lo_type, int fct_type, double dt, const double *u_dev,
                   double *k_dev, void *stream);
/* out = sum_i coef[i] * x[i], 1 <= n <= 9 (the vector updates of the explicit RK solvers) */
int rmh_lincomb(rmh_ctx *ctx, int n, const double *coef, const double *const *x_dev,
                double *out_dev, void *stream);
/* ODESolver::Step for -s 1, 2, 3, 4, 6 (ForwardEuler, RK2Solver(1.0), RK3SSPSolver, RK4Solver,
 * RK6Solver; remhos.cpp:488-492) over rmh_mult, and for -s 11, 12, 13, 14, 16
 * (ForwardEulerIDPSolver, RK{2,3,4,6}IDPSolver; remhos.cpp:493-497, remhos_solvers.cpp) over
 * rmh_mult_unlimited / rmh_limit_mult with masks off (remhos.cpp:502-507); returns 3 for an
 * unknown type as remhos() does (remhos.cpp:499-500). */
int rmh_ode_step(rmh_ctx *ctx, int ode_solver_type, int ho_type, int lo_type, int fct_type,
                 double *t, double dt, double *u_dev, void *stream);

/* LimitedTimeDependentOperator::Mult = MultUnlimited + LimitMult for the configuration
 * -ho 3 -lo {1,3,5} -fct 2 (remhos_solvers.hpp:46-50; remhos.cpp:1596-1739,1798-1916):
 * k = F(u; dt), evaluated by the fused stage kernel.  lo_type follows remhos.cpp:76-77. */
int rmh_stage(rmh_ctx *ctx, int lo_type, double dt, const double *u_dev, double *k_dev,
              void *stream);
/* One fused RK stage: out = a*x0 + b*(y + dt*F(y; dt)); also refreshes the per-element
 * min/max of `out` kept in the context for the next stage's bounds. x0 may equal y. */
int rmh_rk_stage(rmh_ctx *ctx, int lo_type, double dt, double a, double b,
                 const double *x0_dev, const double *y_dev, double *out_dev, void *stream);
/* ODESolver::Step for -s 1/2/3 (ForwardEuler, RK2Solver(1.0), RK3SSPSolver;
 * remhos.cpp:488-490) built from rmh_rk_stage; u updated in place, t advanced. */
int rmh_rk_step(rmh_ctx *ctx, int ode_solver_type, int lo_type, double *t, double dt,
                double *u_dev, void *stream);
/* Same call with HOST state: H2D of u, one step, D2H of u (the end-to-end entry point). */
int rmh_rk_step_host(rmh_ctx *ctx, int ode_solver_type, int lo_type, double *t, double dt,
                     double *u_host);
/* The same step QUEUED on three streams (H2D, stages, D2H) with a ring of three device buffers: returns at
 * once, rmh_host_sync waits for everything queued.  Consecutive calls overlap -- independent host states
 * (several fields transported by the same velocity) run H2D of call n+1, the stages of call n and D2H of
 * call n-1 together; the same host buffer stepped again follows the previous D2H slab by slab, so that both
 * directions of the PCIe link are busy.  The state read is the one at u_in_host when the copy runs; t is the
 * time at the start of the step.  Host buffers must be pinned, and identical or disjoint between calls.
 * The queued steps use the context's stage intermediates: call rmh_host_sync before entering the context
 * through any other entry point on another stream (rmh_rk_step_host does so itself).
 * (ODESolver::Step on the reference's host-resident vectors, remhos.cpp:1146-1180.) */
int rmh_rk_step_host_async(rmh_ctx *ctx, int ode_solver_type, int lo_type, double t, double dt,
                           const double *u_in_host, double *u_out_host);
int rmh_host_sync(rmh_ctx *ctx);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU: the mesh decomposed over the GPUs of one node, one context per rank (one process per
 * GPU, or several contexts in one process).  Replaces ParGridFunction::ExchangeFaceNbrData
 * (remhos.cpp:1813; inside K.Mult), the GroupCommunicator min/max reduction of
 * DofInfo::ComputeOverlapBounds (remhos_tools.cpp:463-466) and the MPI_Allreduce of mass / min / max /
 * dt (remhos.cpp:538-553,1073-1076,1403-1415).
 *
 * Set-up per rank:
 *   rmh_mesh_partition -> rmh_halo_create -> rmh_halo_interior_first -> rmh_mesh_extract(owned + ghost)
 *   -> rmh_mesh_dof_maps -> rmh_ctx_create(ne = owned, ne_ghost = ghost ring)
 *   -> rmh_dplan_create -> rmh_dist_create -> rmh_dist_export
 *   -> [the caller all-gathers the blobs: any transport] -> rmh_dist_connect.
 * Per stage the exchange is one kernel that stores face traces (nfd values per shared face, in the
 * receiver's face order) and ring-element (min,max) pairs straight into the peers' windows over
 * NVLink and publishes an epoch flag; the stage kernel runs interior elements first and waits for
 * the flags before its first shell element.  On a decomposed context, step only through rmh_dist_*.
 * ---------------------------------------------------------------------------------------- */
/* element min/max of y into the context (needed before the first stage of a step when the state
 * was modified from outside) */
int rmh_stage_minmax(rmh_ctx *ctx, const double *y_dev, void *stream);

/* host-only exchange plan (dplan.cpp); nbr_dof = the owned rows [ne][nf][nfd] of rmh_mesh_dof_maps on
 * the local (owned + ghost) mesh */
typedef struct rmh_dplan rmh_dplan;
int rmh_dplan_create(const rmh_halo *h, int rank, int world, int dim, int order, const int32_t *nbr_dof,
                     rmh_dplan **out);
int rmh_dplan_free(rmh_dplan *p);
int64_t rmh_dplan_blob_bytes(const rmh_dplan *p);
int rmh_dplan_export(const rmh_dplan *p, void *blob);
/* blobs[r] = what rank r exported; fills this rank's send tables */
int rmh_dplan_connect(rmh_dplan *p, int n_blobs, const void *const *blobs, const int64_t *sizes);
int rmh_dplan_sizes(const rmh_dplan *p, int64_t *ne, int64_t *ne_ghost, int64_t *n_slots, int32_t *n_peers);
/* per ghost-face slot (scan order over (element, face) with a ghost neighbour): the ghost element */
int rmh_dplan_slot_ghosts(const rmh_dplan *p, int32_t *slot_ghost);
int rmh_dplan_peer(const rmh_dplan *p, int k, int32_t *rank, int64_t *n_tr, int64_t *n_mm, int32_t *flag_slot);
/* tr_src own DOF index -> tr_dst index in the peer's ghost trace array [n_slots][nfd];
 * mm_src own element -> mm_dst index in the peer's (min,max) pair array [ne + ne_ghost] */
int rmh_dplan_peer_tables(const rmh_dplan *p, int k, int32_t *tr_src, int32_t *tr_dst, int32_t *mm_src,
                          int32_t *mm_dst);

/* device layer.  n_interior: the leading owned elements that share no vertex with a ghost element
 * (rmh_halo_interior_first).  The plan must outlive the rmh_dist. */
typedef struct rmh_dist rmh_dist;
int rmh_dist_create(rmh_ctx *ctx, rmh_dplan *plan, int rank, int world, int64_t n_interior, rmh_dist **out);
int64_t rmh_dist_blob_bytes(const rmh_dist *d);
/* blob = window handle (CUDA IPC) + NCCL id (rank 0) + the plan's requests */
int rmh_dist_export(rmh_dist *d, void *blob);
int rmh_dist_connect(rmh_dist *d, int n_blobs, const void *const *blobs, const int64_t *sizes);
/* all ranks must have finished stepping (barrier) before any rank destroys its layer or context */
int rmh_dist_destroy(rmh_dist *d);
/* rmh_rk_stage / rmh_rk_step / rmh_rk_step_host on the decomposed mesh (-ho 3 -lo 5 -fct 2, transport) */
int rmh_dist_rk_stage(rmh_dist *d, int lo_type, double dt, double a, double b, const double *x0_dev,
                      const double *y_dev, double *out_dev, void *stream);
int rmh_dist_rk_step(rmh_dist *d, int ode_solver_type, int lo_type, double *t, double dt, double *u_dev,
                     void *stream);
int rmh_dist_rk_step_host(rmh_dist *d, int ode_solver_type, int lo_type, double *t, double dt, double *u_host);
/* queued variant, see rmh_rk_step_host_async; rmh_host_sync(ctx) waits */
int rmh_dist_rk_step_host_async(rmh_dist *d, int ode_solver_type, int lo_type, double t, double dt,
                                const double *u_in_host, double *u_out_host);
/* diagnostic: in-kernel halo waits on this device since the last reset -- out[0] warps that found a peer's
 * epoch flag unpublished, out[1] their summed and out[2] longest wait in ns; out[3] warps that reached a
 * shell group, out[4] their summed time from there to their end; out[5] all warps of the ghost-aware
 * launches, out[6] their summed run time (ns).  Synchronises the device. */
int rmh_halo_wait_stats(rmh_ctx *ctx, unsigned long long *out7, int reset);
/* op 0 sum, 1 min, 2 max over the ranks, in place on n <= 16 host doubles (ncclAllReduce); collective */
int rmh_dist_allreduce(rmh_dist *d, int op, double *vals, int n, void *stream);

/* reductions over owned DOFs: op 0 = sum(a*b) (b may be NULL -> sum a), 1 = min(a), 2 = max(a)
 * (remhos.cpp:1073-1076,1403-1415; GetMinMax remhos_tools.cpp:1433-1439); result on host */
int rmh_reduce(rmh_ctx *ctx, int op, const double *a_dev, const double *b_dev, double *out,
               void *stream);

/* per-launch CUDA-event timing of the fused stage kernel: enable != 0 starts recording; every
 * call returns and clears the accumulated kernel time [ms] and launch count */
int rmh_profile(rmh_ctx *ctx, int enable, double *total_ms, int64_t *launches);

/* device memory for callers without a CUDA toolchain (the host C++ layer is plain g++):
 * allocation, release, and synchronous copies on the context's device */
int rmh_dev_malloc(rmh_ctx *ctx, int64_t n_doubles, double **out_dev);
int rmh_dev_free(rmh_ctx *ctx, double *p_dev);
int rmh_copy_h2d(rmh_ctx *ctx, double *dst_dev, const double *src_host, int64_t n_doubles);
int rmh_copy_d2h(rmh_ctx *ctx, double *dst_host, const double *src_dev, int64_t n_doubles);
int rmh_copy_d2d(rmh_ctx *ctx, double *dst_dev, const double *src_dev, int64_t n_doubles);
int rmh_sync(rmh_ctx *ctx);

/* number of kernels this library has launched since the counter was last reset */
int64_t rmh_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif
