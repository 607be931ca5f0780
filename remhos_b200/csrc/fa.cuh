// Matrix-based ("full assembly") solver kernels of the RK-stage path: the components the
// reference runs from assembled element matrices (remhos.cpp:1088 -- FA only):
//   DiscreteUpwind::CalcLOSolution              remhos_lo.cpp:43-100
//   ResidualDistribution::CalcLOSolution        remhos_lo.cpp:111-245  (gamma-free variant, -lo 2/3)
//   Assembly::LinearFluxLumping (alpha = 0)     remhos_tools.cpp:876-913
//   FluxBasedFCT::CalcFCTSolution               remhos_fct.cpp:155-181, 295-446
// Everything is generic in (dim, order): the element matrices are built from the same stored
// quadrature data the sum-factorised kernels stream (Dvol, Dface, detJw), so transport and remap
// share one code path.  These are the small-mesh paths; the fused stage kernels are the hot ones.
#ifndef RMH_FA_CUH
#define RMH_FA_CUH

#include "kernels.cuh"

namespace rmh
{

// read access to the stored quadrature data in either layout (see Pre3::load)
struct OpData
{
   int dim, D1, Q, frag;
   int ND, NQ, NF, NFD, NQF;
   const double *Dvol, *Dface, *detJw;
   const double *B, *G;          // device 1-D tables [Q][D1]
   __device__ __forceinline__ double dvol(int64_t e, int c, int q) const
   {
      if (dim == 3 && frag)
      {
         const int RQ = (Q + 1) & ~1, QQ = Q * Q;
         const int col = q % QQ, qz = q / QQ;
         return Dvol[(size_t)e * QQ * RQ * 3 + ((size_t)col * RQ + qz) * 3 + c];
      }
      return Dvol[((size_t)e * dim + c) * NQ + q];
   }
   __device__ __forceinline__ double dface(int64_t e, int f, int qf) const
   {
      if (dim == 3)
      {
         const int qa = qf % Q, qb = qf / Q;
         if (frag)
         {
            const int RQ = (Q + 1) & ~1;
            return Dface[(size_t)e * NF * Q * RQ + ((size_t)f * Q + qa) * RQ + qb];
         }
         return Dface[(size_t)e * NF * NQF + (size_t)qb * NF * Q + f * Q + qa];
      }
      return Dface[((size_t)e * NF + f) * NQF + qf];
   }
};

// lattice coordinates of local DOF i
__device__ __forceinline__ void dof_lattice(int dim, int D1, int i, int (&l)[3])
{
   l[0] = l[1] = l[2] = 0;
   for (int a = 0; a < dim; a++) { l[a] = i % D1; i /= D1; }
}
// natural face index of a DOF with lattice coordinates l on a face normal to `axis`
__device__ __forceinline__ int face_nat_index(int dim, int D1, const int (&l)[3], int axis)
{
   int j = 0, mul = 1;
   for (int b = 0; b < dim; b++)
   {
      if (b == axis) { continue; }
      j += l[b] * mul; mul *= D1;
   }
   return j;
}
// local DOF of natural face DOF j on face f (runtime version of face_dof<>)
__device__ __forceinline__ int face_dof_rt(int dim, int D1, int f, int j)
{
   int axis, side;
   face_axis_side(dim, f, axis, side);
   int l[3] = {0, 0, 0};
   for (int a = 0; a < dim; a++)
   {
      if (a == axis) { l[a] = side * (D1 - 1); }
      else { l[a] = j % D1; j /= D1; }
   }
   return l[0] + D1 * (l[1] + D1 * l[2]);
}
// value of the tensor face basis function a at face quadrature point qf
__device__ __forceinline__ double face_phi(const OpData &o, int a, int qf)
{
   double v = 1.0;
   for (int b = 0; b < o.dim - 1; b++)
   {
      v *= o.B[(qf % o.Q) * o.D1 + (a % o.D1)];
      qf /= o.Q; a /= o.D1;
   }
   return v;
}

// BL[e][f][a] = sum_b bdrInt(a,b) = -sum_q Dface_q phi_a(q): the fully lumped face matrix of
// LinearFluxLumping with alpha = 0 (remhos_tools.cpp:833-857, 897-910; partition of unity)
__global__ void k_face_lump(OpData o, int64_t ne, double *BL)
{
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= ne * o.NF * o.NFD) { return; }
   const int64_t e = idx / (o.NF * o.NFD);
   const int r = (int)(idx - e * o.NF * o.NFD), f = r / o.NFD, a = r - f * o.NFD;
   double s = 0.0;
   for (int qf = 0; qf < o.NQF; qf++) { s += o.dface(e, f, qf) * face_phi(o, a, qf); }
   BL[idx] = -s;
}

// Dense element matrices (one block per element):
//   K  [e][i][j] = sum_q phi_i (D . grad phi_j)      ConvectionIntegrator block (remhos.cpp:646-657)
//   M  [e][i][j] = sum_q detJw phi_i phi_j           MassIntegrator block
//   BI [e][f][a][b] = -sum_q Dface phi_a phi_b       bdrInt (remhos_tools.cpp:788-858)
//   KH = K - sum_f scatter(BI_f): the diagonal block of K_HO (volume + own-side face terms)
__global__ void k_fa_dense(OpData o, int64_t ne, double *K, double *KH, double *M, double *BI)
{
   extern __shared__ double sh[];
   const int64_t e = blockIdx.x;
   const int ND = o.ND, NQ = o.NQ, dim = o.dim, D1 = o.D1, Q = o.Q;
   double *dv = sh;                       // [dim][NQ]
   double *dj = dv + dim * NQ;            // [NQ]
   double *tB = dj + NQ, *tG = tB + Q * D1;
   for (int t = threadIdx.x; t < dim * NQ; t += blockDim.x) { dv[t] = o.dvol(e, t / NQ, t % NQ); }
   for (int t = threadIdx.x; t < NQ; t += blockDim.x) { dj[t] = o.detJw[(size_t)e * NQ + t]; }
   for (int t = threadIdx.x; t < Q * D1; t += blockDim.x) { tB[t] = o.B[t]; tG[t] = o.G[t]; }
   __syncthreads();
   for (int ij = threadIdx.x; ij < ND * ND; ij += blockDim.x)
   {
      const int i = ij / ND, j = ij - i * ND;
      int li[3], lj[3];
      dof_lattice(dim, D1, i, li);
      dof_lattice(dim, D1, j, lj);
      double k = 0.0, m = 0.0;
      for (int q = 0; q < NQ; q++)
      {
         int qa[3] = {0, 0, 0}, r = q;
         for (int a = 0; a < dim; a++) { qa[a] = r % Q; r /= Q; }
         double pi = 1.0, pj = 1.0;
         for (int a = 0; a < dim; a++) { pi *= tB[qa[a] * D1 + li[a]]; pj *= tB[qa[a] * D1 + lj[a]]; }
         double g = 0.0;
         for (int c = 0; c < dim; c++)
         {
            double d = 1.0;
            for (int a = 0; a < dim; a++) { d *= (a == c) ? tG[qa[a] * D1 + lj[a]] : tB[qa[a] * D1 + lj[a]]; }
            g += dv[c * NQ + q] * d;
         }
         k += pi * g;
         m += (pi * pj) * dj[q];
      }
      K[(size_t)e * ND * ND + ij] = k;
      KH[(size_t)e * ND * ND + ij] = k;
      M[(size_t)e * ND * ND + ij] = m;
   }
   const int NFD = o.NFD, NF = o.NF;
   for (int t = threadIdx.x; t < NF * NFD * NFD; t += blockDim.x)
   {
      const int f = t / (NFD * NFD), r = t - f * NFD * NFD, a = r / NFD, b = r - a * NFD;
      double s = 0.0;
      for (int qf = 0; qf < o.NQF; qf++) { s += o.dface(e, f, qf) * (face_phi(o, a, qf) * face_phi(o, b, qf)); }
      BI[(size_t)e * NF * NFD * NFD + t] = -s;
   }
   __syncthreads();
   // own-side face blocks into KH (a DOF pair can share several faces: serial over faces)
   for (int f = 0; f < NF; f++)
   {
      for (int t = threadIdx.x; t < NFD * NFD; t += blockDim.x)
      {
         const int a = t / NFD, b = t - a * NFD;
         const int i = face_dof_rt(dim, D1, f, a), j = face_dof_rt(dim, D1, f, b);
         KH[(size_t)e * ND * ND + i * ND + j] -= BI[((size_t)e * NF + f) * NFD * NFD + t];
      }
      __syncthreads();
   }
}

struct FaArgs
{
   int64_t ne;
   int dim, D1, ND, NF, NFD;
   FaceNbr fn;
   const int16_t *pat_idx;    // [npat][NFD] natural index, on the neighbour's face, of pat[id][j]
   const uint8_t *pat_face;   // [npat] the neighbour's local face
   const double *K, *KH, *M, *BI, *BL, *ml, *inflow;
   // decomposed meshes, FluxBasedFCT: the neighbour-side face block of every ghost face [n_gslots][NFD][NFD]
   // (k_fa_ghost_blocks) and the ghost traces of the flux coefficients R+ / R- (remhos_fct.cpp:406-409)
   const double *BIg = nullptr, *gcp = nullptr, *gcn = nullptr;
   // product remap (FluxBasedFCT::CalcFCTProduct, remhos_fct.cpp:214-246): element-local fluxes
   // beta_j fel_i - beta_i fel_j added to the in-element couplings (null: none)
   const double *pbeta = nullptr, *pfel = nullptr;
};

// exterior state seen by face DOF (f, a) of element e: the neighbour's value, or `bval` on the
// domain boundary
__device__ __forceinline__ double nbr_value(const FaArgs &A, const double *u, int64_t e, int f,
                                            int a, double bval, int64_t *gj = nullptr)
{
   const int64_t nb = A.fn.nbr_elem[e * A.NF + f];
   if (nb < 0) { if (gj) { *gj = -1; } return bval; }
   const int loc = A.fn.pat[(int)A.fn.nbr_pat[e * A.NF + f] * A.NFD + a];
   if (gj) { *gj = nb * A.ND + loc; }
   return (nb < A.fn.ne_owned) ? u[nb * A.ND + loc] : A.fn.ughost[(nb - A.fn.ne_owned) * A.NFD + a];
}

// sum over the faces containing DOF i of BL (u_nbr - u_own)   (LinearFluxLumping, alpha = 0)
__device__ __forceinline__ double lumped_faces(const FaArgs &A, const double *u, int64_t e, int i,
                                               double ui)
{
   int l[3];
   dof_lattice(A.dim, A.D1, i, l);
   const double infl = A.inflow ? A.inflow[e * A.ND + i] : 0.0;
   double s = 0.0;
   for (int ax = 0; ax < A.dim; ax++)
   {
      for (int side = 0; side < 2; side++)
      {
         if (l[ax] != side * (A.D1 - 1)) { continue; }
         const int f = face_of(A.dim, ax, side);
         const int a = face_nat_index(A.dim, A.D1, l, ax);
         const double un = nbr_value(A, u, e, f, a, infl);
         s += A.BL[(e * A.NF + f) * A.NFD + a] * (un - ui);
      }
   }
   return s;
}

// DiscreteUpwind: du = (D u + lumped faces) / m with D = K + d, d_ij = max(0, -k_ij, -k_ji),
// D_ii = K_ii - sum_{j != i} d_ij  (remhos_lo.cpp:43-100; remhos_tools.cpp:1464-1487)
__global__ void k_lo_du(FaArgs A, const double *u, double *du)
{
   extern __shared__ double sh[];
   const int64_t e = blockIdx.x;
   const int ND = A.ND;
   for (int t = threadIdx.x; t < ND; t += blockDim.x) { sh[t] = u[e * ND + t]; }
   __syncthreads();
   const double *Ke = A.K + (size_t)e * ND * ND;
   for (int i = threadIdx.x; i < ND; i += blockDim.x)
   {
      const double ui = sh[i];
      double s = 0.0;
      for (int j = 0; j < ND; j++)
      {
         const double kij = Ke[i * ND + j];
         s += kij * sh[j];
         if (j != i)
         {
            const double dij = fmax(fmax(0.0, -kij), -Ke[j * ND + i]);
            s += dij * (sh[j] - ui);
         }
      }
      s += lumped_faces(A, u, e, i, ui);
      du[e * ND + i] = s / A.ml[e * ND + i];
   }
}

// ResidualDistribution without subcells (remhos_lo.cpp:111-245 with subcell_scheme = false):
// z = K u given; du = (faces + w+ rho+ + w- rho-) / m.  One warp per element.
__global__ void k_lo_rd(FaArgs A, const double *u, const double *z, double *du)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= A.ne) { return; }
   const int ND = A.ND;
   double xmax = -INFINITY, xmin = INFINITY, xsum = 0.0, rhoP = 0.0, rhoN = 0.0;
   for (int j = lane; j < ND; j += 32)
   {
      const double v = u[e * ND + j], zz = z[e * ND + j];
      xmax = fmax(xmax, v); xmin = fmin(xmin, v); xsum += v;
      rhoP += fmax(0.0, zz); rhoN += fmin(0.0, zz);
   }
   xmax = warp_max(xmax); xmin = warp_min(xmin);
   xsum = warp_sum(xsum); rhoP = warp_sum(rhoP); rhoN = warp_sum(rhoN);
   constexpr double eps = 1.0e-15;
   const double sumWP = ND * xmax - xsum + eps, sumWN = ND * xmin - xsum - eps;
   for (int j = lane; j < ND; j += 32)
   {
      const double ui = u[e * ND + j];
      const double wP = (xmax - ui) / sumWP, wN = (xmin - ui) / sumWN;
      const double s = lumped_faces(A, u, e, j, ui) + wP * rhoP + wN * rhoN;
      du[e * ND + j] = s / A.ml[e * ND + j];
   }
}

// ---- NeumannHOSolver pieces (remhos_ho.cpp:136-187)
// rhs_i += sum over the faces containing i of sum_b bdrInt(a,b) (u_nbr(b) - u_own(b)):
// LinearFluxLumping with alpha = 1 (the Galerkin face term), inflow exterior state
__global__ void k_face_galerkin(FaArgs A, const double *u, double *rhs)
{
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= A.ne * A.ND) { return; }
   const int64_t e = idx / A.ND;
   const int i = (int)(idx - e * A.ND);
   int l[3];
   dof_lattice(A.dim, A.D1, i, l);
   double s = 0.0;
   for (int ax = 0; ax < A.dim; ax++)
   {
      for (int side = 0; side < 2; side++)
      {
         if (l[ax] != side * (A.D1 - 1)) { continue; }
         const int f = face_of(A.dim, ax, side);
         const int a = face_nat_index(A.dim, A.D1, l, ax);
         const double *BIe = A.BI + ((size_t)e * A.NF + f) * A.NFD * A.NFD + (size_t)a * A.NFD;
         for (int b = 0; b < A.NFD; b++)
         {
            const int jb = face_dof_rt(A.dim, A.D1, f, b);
            const double infl = A.inflow ? A.inflow[e * A.ND + jb] : 0.0;
            const double un = nbr_value(A, u, e, f, b, infl);
            s += BIe[b] * (un - u[e * A.ND + jb]);
         }
      }
   }
   rhs[idx] += s;
}
// res = M du - rhs with the dense element mass blocks
__global__ void k_mass_residual(int64_t ne, int ND, const double *M, const double *du,
                                const double *rhs, double *res)
{
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= ne * ND) { return; }
   const int64_t e = idx / ND;
   const int i = (int)(idx - e * ND);
   const double *Me = M + (size_t)e * ND * ND + (size_t)i * ND, *de = du + e * ND;
   double s = 0.0;
   for (int j = 0; j < ND; j++) { s += Me[j] * de[j]; }
   res[idx] = s - rhs[idx];
}
__global__ void k_neumann_update(int64_t n, const double *res, const double *ml, double *du)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { du[i] -= res[i] / ml[i]; }
}
// column j of the dense element blocks <-> a DOF vector (used to form M_L M^-1 K column-wise)
__global__ void k_col_get(int64_t ne, int ND, int j, const double *K, double *v)
{
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= ne * ND) { return; }
   const int64_t e = idx / ND;
   const int i = (int)(idx - e * ND);
   v[idx] = K[(size_t)e * ND * ND + (size_t)i * ND + j];
}
__global__ void k_col_put(int64_t ne, int ND, int j, const double *v, const double *ml, double *KP)
{
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= ne * ND) { return; }
   const int64_t e = idx / ND;
   const int i = (int)(idx - e * ND);
   KP[(size_t)e * ND * ND + (size_t)i * ND + j] = ml[idx] * v[idx];
}

// ---- subcell residual distribution (-lo 4)
// SubcellWeights(k)(m, j) = alpha grad_ref(phi_j)(centre) . adj(J_sub) . v(centre) on the straight-
// sided subcells spanned by the lattice points (Assembly::ComputeSubcellWeights,
// remhos_tools.cpp:860-874, 1033-1076; set-up remhos.cpp:797-868).  xlat0 [ne][nd][dim] lattice
// points at t = 0; transport: vel [ne][ns][dim] at the subcell centres, alpha = -1; remap: vel
// [ne][nd][dim] at the lattice points (zero on the boundary), subcells move as x0 + t v, alpha = +1.
__global__ void k_subcell_weights(int dim, int p, int exec_mode, double t, int64_t ne,
                                  const double *xlat0, const double *vel, double *SW)
{
   int ns = 1, nc = 1, nd = 1;
   for (int a = 0; a < dim; a++) { ns *= p; nc *= 2; nd *= p + 1; }
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= ne * ns) { return; }
   const int64_t e = idx / ns;
   const int m = (int)(idx - e * ns);
   int sc[3] = {0, 0, 0}, r = m;
   for (int a = 0; a < dim; a++) { sc[a] = r % p; r /= p; }
   double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, v[3] = {0, 0, 0};
   const double gs = (dim == 2) ? 0.5 : 0.25;
   for (int c = 0; c < nc; c++)
   {
      int loc = 0, mul = 1;
      for (int a = 0; a < dim; a++) { loc += (sc[a] + ((c >> a) & 1)) * mul; mul *= p + 1; }
      const double *x0 = xlat0 + ((size_t)e * nd + loc) * dim;
      for (int i = 0; i < dim; i++)
      {
         double x = x0[i];
         if (exec_mode == 1)
         {
            const double vv = vel[((size_t)e * nd + loc) * dim + i];
            x += t * vv;
            v[i] += vv / nc;
         }
         for (int j = 0; j < dim; j++) { J[i][j] += x * (((c >> j) & 1) ? gs : -gs); }
      }
   }
   if (exec_mode != 1) { for (int i = 0; i < dim; i++) { v[i] = vel[((size_t)e * ns + m) * dim + i]; } }
   double adj[3][3];
   if (dim == 2)
   {
      adj[0][0] = J[1][1]; adj[0][1] = -J[0][1]; adj[1][0] = -J[1][0]; adj[1][1] = J[0][0];
   }
   else
   {
      adj[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      adj[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
      adj[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
      adj[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      adj[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
      adj[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
      adj[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      adj[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
      adj[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
   }
   double av[3] = {0, 0, 0};
   for (int i = 0; i < dim; i++) { for (int j = 0; j < dim; j++) { av[i] += adj[i][j] * v[j]; } }
   const double alpha = (exec_mode == 1) ? 1.0 : -1.0;
   for (int c = 0; c < nc; c++)
   {
      double w = 0.0;
      for (int j = 0; j < dim; j++) { w += (((c >> j) & 1) ? gs : -gs) * av[j]; }
      SW[((size_t)e * ns + m) * nc + c] = alpha * w;
   }
}

// ResidualDistribution with the subcell fluctuation redistribution, gamma = 1
// (remhos_lo.cpp:164-239).  One warp per element; per-subcell quantities are staged in shared
// memory and gathered per DOF in ascending subcell order (the reference's summation order).
__global__ void k_lo_rd_sub(FaArgs A, int p, const double *SW, const double *u, const double *z,
                            double *du)
{
   extern __shared__ double sh[];
   const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
   int ns = 1, nc = 1;
   for (int a = 0; a < A.dim; a++) { ns *= p; nc *= 2; }
   double *sub = sh + (size_t)wib * ns * 6;      // fP, fN, xmax, xmin, swP, swN per subcell
   if (e >= A.ne) { return; }
   const int ND = A.ND, D1 = A.D1;
   const double *ue = u + e * ND;
   double xmax = -INFINITY, xmin = INFINITY, xsum = 0.0, rhoP = 0.0, rhoN = 0.0;
   for (int j = lane; j < ND; j += 32)
   {
      const double v = ue[j], zz = z[e * ND + j];
      xmax = fmax(xmax, v); xmin = fmin(xmin, v); xsum += v;
      rhoP += fmax(0.0, zz); rhoN += fmin(0.0, zz);
   }
   xmax = warp_max(xmax); xmin = warp_min(xmin);
   xsum = warp_sum(xsum); rhoP = warp_sum(rhoP); rhoN = warp_sum(rhoN);
   constexpr double eps = 1.0e-15, gamma = 1.0;
   double sfP = 0.0, sfN = 0.0;
   for (int m = lane; m < ns; m += 32)
   {
      int sc[3] = {0, 0, 0}, r = m;
      for (int a = 0; a < A.dim; a++) { sc[a] = r % p; r /= p; }
      double fl = 0.0, smax = -INFINITY, smin = INFINITY, ssum = 0.0;
      for (int c = 0; c < nc; c++)
      {
         int loc = 0, mul = 1;
         for (int a = 0; a < A.dim; a++) { loc += (sc[a] + ((c >> a) & 1)) * mul; mul *= D1; }
         const double v = ue[loc];
         fl += SW[((size_t)e * ns + m) * nc + c] * v;
         smax = fmax(smax, v); smin = fmin(smin, v); ssum += v;
      }
      const double fP = fmax(0.0, fl), fN = fmin(0.0, fl);
      sub[m * 6 + 0] = fP; sub[m * 6 + 1] = fN; sub[m * 6 + 2] = smax; sub[m * 6 + 3] = smin;
      sub[m * 6 + 4] = nc * smax - ssum + eps; sub[m * 6 + 5] = nc * smin - ssum - eps;
      sfP += fP; sfN += fN;
   }
   sfP = warp_sum(sfP); sfN = warp_sum(sfN);
   __syncwarp();
   const double sumWP = ND * xmax - xsum + eps, sumWN = ND * xmin - xsum - eps;
   for (int j = lane; j < ND; j += 32)
   {
      int l[3];
      dof_lattice(A.dim, D1, j, l);
      const double ui = ue[j];
      // nodal weights: subcells containing this lattice point, ascending subcell index
      double nwP = 0.0, nwN = 0.0;
      for (int k = 0; k < nc; k++)
      {
         // k enumerates the offsets (dz, dy, dx) in {-1, 0}: ascending m means descending offset bits
         int m = 0, mul = 1;
         bool ok = true;
         for (int a = 0; a < A.dim; a++)
         {
            const int off = ((k >> a) & 1) ? 0 : -1;
            const int s = l[a] + off;
            if (s < 0 || s >= p) { ok = false; }
            m += s * mul; mul *= p;
         }
         if (!ok) { continue; }
         nwP += sub[m * 6 + 0] * ((sub[m * 6 + 2] - ui) / sub[m * 6 + 4]);   // eq. (58)
         nwN += sub[m * 6 + 1] * ((sub[m * 6 + 3] - ui) / sub[m * 6 + 5]);   // eq. (59)
      }
      double wP = (xmax - ui) / sumWP, wN = (xmin - ui) / sumWN;
      double aux = gamma / (rhoP + eps);
      wP *= 1.0 - fmin(aux * sfP, 1.0);
      wP += fmin(aux, 1.0 / (sfP + eps)) * nwP;
      aux = gamma / (rhoN - eps);
      wN *= 1.0 - fmin(aux * sfN, 1.0);
      wN += fmax(aux, 1.0 / (sfN - eps)) * nwN;
      const double s = lumped_faces(A, u, e, j, ui) + wP * rhoP + wN * rhoN;
      du[e * ND + j] = s / A.ml[e * ND + j];
   }
}

// ---- MonoRDSolver::CalcSolution (remhos_mono.cpp:60-356) without a smoothness indicator.
// One warp per element, everything element-local in shared memory:
//   alpha_j from the bound gaps (beta = 10), volume split du = alpha z, z <- (1 - alpha) z;
//   faces in ascending order: Assembly::NonlinFluxLumping (remhos_tools.cpp:915-973) into du (alpha)
//   and into d (alpha = 1);  residual distribution of the remaining z (subcell variant: gamma = 10);
//   fixed-point mass correction (eq. 27-29; <= 101 sweeps, |res|_2 <= 1e-8) with the dense element
//   mass block.  z = K u (volume-only) and the per-DOF bounds are inputs.
__global__ void k_mono_rd(FaArgs A, int p, int subcell, int mass_lim, const double *SW,
                          const double *scale, const double *u, const double *z,
                          const double *xi_min, const double *xi_max, const double *si_tmp,
                          double *du_out)
{
   extern __shared__ double sh[];
   const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
   const int64_t e = (int64_t)blockIdx.x * wpb + wib;
   int ns = 1, nc = 1;
   for (int a = 0; a < A.dim; a++) { ns *= p; nc *= 2; }
   const int ND = A.ND, NFD = A.NFD, NF = A.NF, D1 = A.D1;
   const int per = 8 * ND + 3 * NFD + ns * 6;
   double *U = sh + (size_t)wib * per, *DU = U + ND, *D = DU + ND, *ZR = D + ND, *AL = ZR + ND,
          *MI = AL + ND, *UD = MI + ND, *GAP = UD + ND, *XD = GAP + ND, *C1 = XD + NFD, *C2 = C1 + NFD,
          *sub = C2 + NFD;
   if (e >= A.ne) { return; }
   constexpr double eps = 1.0e-15, beta = 10.0, gamma = 10.0, tol = 1.0e-8;
   // ---- bound gaps, alpha, volume split (remhos_mono.cpp:123-160)
   double xmax = -INFINITY, xmin = INFINITY, xsum = 0.0;
   for (int j = lane; j < ND; j += 32)
   {
      const double ui = u[e * ND + j], zz = z[e * ND + j];
      const double up = xi_max[e * ND + j] - ui, dn = ui - xi_min[e * ND + j];
      const double gap = fmin(up, dn);
      double al = fmin(1.0, beta * gap / (fmax(up, dn) + eps));
      if (si_tmp)                                                // remhos_mono.cpp:132-153
      {
         const double tmp = si_tmp[e * ND + j], lo = xi_min[e * ND + j], hi = xi_max[e * ND + j];
         const double bndN = fmax(0.0, tmp * (2.0 * ui - hi) + (1.0 - tmp) * lo);
         const double bndP = fmin(1.0, tmp * (2.0 * ui - lo) + (1.0 - tmp) * hi);
         if (lo + hi > 2.0 * ui + eps) { al = fmin(1.0, beta * (ui - bndN) / (hi - ui + eps)); }
         else if (lo + hi < 2.0 * ui - eps) { al = fmin(1.0, beta * (bndP - ui) / (ui - lo + eps)); }
      }
      U[j] = ui; GAP[j] = gap; AL[j] = al; MI[j] = 0.0;
      DU[j] = al * zz; ZR[j] = zz - al * zz; D[j] = zz;
      xmax = fmax(xmax, ui); xmin = fmin(xmin, ui); xsum += ui;
   }
   xmax = warp_max(xmax); xmin = warp_min(xmin); xsum = warp_sum(xsum);
   __syncwarp();
   // ---- face contributions, faces in ascending order (remhos_mono.cpp:163-167)
   for (int f = 0; f < NF; f++)
   {
      for (int a = lane; a < NFD; a += 32)
      {
         const int ja = face_dof_rt(A.dim, D1, f, a);
         const double infl = A.inflow ? A.inflow[e * ND + ja] : 0.0;
         XD[a] = nbr_value(A, u, e, f, a, infl) - U[ja];
      }
      __syncwarp();
      double sP = 0.0, sN = 0.0, tP = 0.0, tN = 0.0;
      for (int a = lane; a < NFD; a += 32)
      {
         const int ja = face_dof_rt(A.dim, D1, f, a);
         const double *BIe = A.BI + (((size_t)e * NF + f) * NFD + a) * NFD;
         double y1 = DU[ja], y2 = D[ja], corr = 0.0;
         const double xa = XD[a];
         for (int b = 0; b < NFD; b++)
         {
            y1 += BIe[b] * xa; y2 += BIe[b] * xa;
            corr += BIe[b] * (XD[b] - xa);
         }
         DU[ja] = y1; D[ja] = y2;
         const double c1 = corr * AL[ja];
         C1[a] = c1; C2[a] = corr;
         sP += fmax(0.0, c1); sN += fmin(0.0, c1);
         tP += fmax(0.0, corr); tN += fmin(0.0, corr);
      }
      sP = warp_sum(sP); sN = warp_sum(sN); tP = warp_sum(tP); tN = warp_sum(tN);
      for (int a = lane; a < NFD; a += 32)
      {
         const int ja = face_dof_rt(A.dim, D1, f, a);
         double c1 = C1[a], c2 = C2[a];
         if (sP + sN > eps) { c1 = fmin(0.0, c1) - fmax(0.0, c1) * sN / sP; }
         else if (sP + sN < -eps) { c1 = fmax(0.0, c1) - fmin(0.0, c1) * sP / sN; }
         if (tP + tN > eps) { c2 = fmin(0.0, c2) - fmax(0.0, c2) * tN / tP; }
         else if (tP + tN < -eps) { c2 = fmax(0.0, c2) - fmin(0.0, c2) * tP / tN; }
         DU[ja] += c1; D[ja] += c2;
      }
      __syncwarp();
   }
   // ---- element contributions (remhos_mono.cpp:169-262)
   double rhoP = 0.0, rhoN = 0.0;
   for (int j = lane; j < ND; j += 32) { rhoP += fmax(0.0, ZR[j]); rhoN += fmin(0.0, ZR[j]); }
   rhoP = warp_sum(rhoP); rhoN = warp_sum(rhoN);
   double sfP = 0.0, sfN = 0.0;
   if (subcell)
   {
      for (int m = lane; m < ns; m += 32)
      {
         int sc[3] = {0, 0, 0}, r = m;
         for (int a = 0; a < A.dim; a++) { sc[a] = r % p; r /= p; }
         double fl = 0.0, smax = -INFINITY, smin = INFINITY, ssum = 0.0;
         for (int c = 0; c < nc; c++)
         {
            int loc = 0, mul = 1;
            for (int a = 0; a < A.dim; a++) { loc += (sc[a] + ((c >> a) & 1)) * mul; mul *= D1; }
            const double v = U[loc];
            fl += SW[((size_t)e * ns + m) * nc + c] * v;
            smax = fmax(smax, v); smin = fmin(smin, v); ssum += v;
         }
         const double fP = fmax(0.0, fl), fN = fmin(0.0, fl);
         sub[m * 6 + 0] = fP; sub[m * 6 + 1] = fN; sub[m * 6 + 2] = smax; sub[m * 6 + 3] = smin;
         sub[m * 6 + 4] = nc * smax - ssum + eps; sub[m * 6 + 5] = nc * smin - ssum - eps;
         sfP += fP; sfN += fN;
      }
      sfP = warp_sum(sfP); sfN = warp_sum(sfN);
      __syncwarp();
   }
   const double sumWP = ND * xmax - xsum + eps, sumWN = ND * xmin - xsum - eps;
   for (int j = lane; j < ND; j += 32)
   {
      const double ui = U[j];
      double wP = (xmax - ui) / sumWP, wN = (xmin - ui) / sumWN;
      if (subcell)
      {
         int l[3];
         dof_lattice(A.dim, D1, j, l);
         double nwP = 0.0, nwN = 0.0;
         for (int k = 0; k < nc; k++)
         {
            int m = 0, mul = 1;
            bool ok = true;
            for (int a = 0; a < A.dim; a++)
            {
               const int off = ((k >> a) & 1) ? 0 : -1;
               const int s_ = l[a] + off;
               if (s_ < 0 || s_ >= p) { ok = false; }
               m += s_ * mul; mul *= p;
            }
            if (!ok) { continue; }
            nwP += sub[m * 6 + 0] * ((sub[m * 6 + 2] - ui) / sub[m * 6 + 4]);
            nwN += sub[m * 6 + 1] * ((sub[m * 6 + 3] - ui) / sub[m * 6 + 5]);
         }
         double aux = gamma / (rhoP + eps);
         wP *= 1.0 - fmin(aux * sfP, 1.0);
         wP += fmin(aux, 1.0 / (sfP + eps)) * nwP;
         aux = gamma / (rhoN - eps);
         wN *= 1.0 - fmin(aux * sfN, 1.0);
         wN += fmax(aux, 1.0 / (sfN - eps)) * nwN;
      }
      DU[j] += wP * rhoP + wN * rhoN;
   }
   __syncwarp();
   // ---- time derivative and mass matrix: element-local fixed point (remhos_mono.cpp:264-346)
   if (mass_lim)
   {
      const double sck = scale[e];
      const double *Me = A.M + (size_t)e * ND * ND;
      for (int it = 0; it <= 100; it++)
      {
         double udmin = INFINITY, udmax = -INFINITY;
         for (int j = lane; j < ND; j += 32)
         {
            const double v = (DU[j] + MI[j]) / A.ml[e * ND + j];
            UD[j] = v;
            udmin = fmin(udmin, v); udmax = fmax(udmax, v);
         }
         udmin = warp_min(udmin); udmax = warp_max(udmax);
         __syncwarp();
         double mi[8];
         double MP = 0.0, MN = 0.0;
         int q = 0;
         for (int i = lane; i < ND; i += 32, q++)
         {
            const double udi = UD[i];
            double s = 0.0;
            for (int j = ND - 1; j >= 0; j--) { s += Me[(size_t)i * ND + j] * (udi - UD[j]); }
            const double diff = D[i] - DU[i];
            const double tmp = si_tmp ? si_tmp[e * ND + i] : 0.0;
            s += fmin(1.0, fmax(tmp, fabs(s) / (fabs(diff) + eps))) * diff;
            const double den = fmax(udmax - udi, udi - udmin) + eps;
            double al = fmin(1.0, beta * sck * GAP[i] / den);
            if (si_tmp)                                          // remhos_mono.cpp:316-324
            {
               const double aglob = fmin(1.0, beta * sck * fmin(1.0 - U[i], U[i] - 0.0) / den);
               al = fmin(fmax(tmp, al), aglob);
            }
            s *= al;
            mi[q] = s;
            MP += fmax(0.0, s); MN += fmin(0.0, s);
         }
         MP = warp_sum(MP); MN = warp_sum(MN);
         double res2 = 0.0;
         q = 0;
         for (int i = lane; i < ND; i += 32, q++)
         {
            double s = mi[q];
            if (MP + MN > eps) { s = fmin(0.0, s) - fmax(0.0, s) * MN / MP; }
            else if (MP + MN < -eps) { s = fmax(0.0, s) - fmin(0.0, s) * MP / MN; }
            const double r = s + DU[i] - A.ml[e * ND + i] * UD[i];
            res2 += r * r;
            MI[i] = s;
         }
         res2 = warp_sum(res2);
         __syncwarp();
         if (sqrt(res2) <= tol) { break; }
      }
   }
   for (int j = lane; j < ND; j += 32) { du_out[e * ND + j] = (DU[j] + MI[j]) / A.ml[e * ND + j]; }
}

// ---- ElementFCTProjection::CalcFCTSolution (remhos_fct.cpp:613-733): element-local Zalesak
// limiter on F_ij = M_ij (du_i - du_j) + (beta_j z_i - beta_i z_j), beta = M_L / sum M_L,
// z = M du_HO - M_L du_LO, started from the LO rate.  One warp per element; the fluxes are
// re-evaluated from both ends with identical operands, so F_ji = -F_ij bit for bit.
__global__ void k_fct_project(FaArgs A, double dt, const double *u, const double *du_ho,
                              const double *du_lo, const double *xi_min, const double *xi_max,
                              double *du)
{
   extern __shared__ double sh[];
   const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
   const int64_t e = (int64_t)blockIdx.x * wpb + wib;
   const int ND = A.ND;
   double *DH = sh + (size_t)wib * 6 * ND, *Z = DH + ND, *BE = Z + ND, *ML = BE + ND, *GP = ML + ND, *GM = GP + ND;
   if (e >= A.ne) { return; }
   const double *Me = A.M + (size_t)e * ND * ND;
   for (int j = lane; j < ND; j += 32) { DH[j] = du_ho[e * ND + j]; }
   __syncwarp();
   double mls = 0.0;
   for (int i = lane; i < ND; i += 32)
   {
      double rhs = 0.0, ml = 0.0;
      for (int j = 0; j < ND; j++) { rhs += Me[(size_t)i * ND + j] * DH[j]; ml += Me[(size_t)i * ND + j]; }
      ML[i] = ml; Z[i] = rhs - ml * du_lo[e * ND + i];
      mls += ml;
   }
   mls = warp_sum(mls);
   __syncwarp();
   for (int i = lane; i < ND; i += 32) { BE[i] = ML[i] / mls; }
   __syncwarp();
   auto flux = [&](int i, int j)
   { return Me[(size_t)i * ND + j] * (DH[i] - DH[j]) + (BE[j] * Z[i] - BE[i] * Z[j]); };
   for (int i = lane; i < ND; i += 32)
   {
      double sp = 0.0, sm = 0.0;
      for (int j = 0; j < ND; j++)
      {
         if (j == i) { continue; }
         const double f = flux(i, j);
         sp += fmax(0.0, f); sm += fmin(0.0, f);
      }
      const double ui = u[e * ND + i], dlo = du_lo[e * ND + i];
      const double rp = fmax(ML[i] * ((xi_max[e * ND + i] - ui) / dt - dlo), 0.0);
      const double rm = fmin(ML[i] * ((xi_min[e * ND + i] - ui) / dt - dlo), 0.0);
      GP[i] = (rp < sp) ? rp / sp : 1.0;
      GM[i] = (rm > sm) ? rm / sm : 1.0;
   }
   __syncwarp();
   for (int i = lane; i < ND; i += 32)
   {
      double acc = 0.0;
      for (int j = 0; j < ND; j++)
      {
         if (j == i) { continue; }
         const double f = flux(i, j);
         const double a = (f >= 0.0) ? fmin(GP[i], GM[j]) : fmin(GM[i], GP[j]);
         acc += a * f;
      }
      du[e * ND + i] = du_lo[e * ND + i] + acc / ML[i];
   }
}

// ---- SmoothnessIndicator::UpdateBounds (remhos_tools.cpp:183-190) for every dof: si = the indicator
// at the dof (1 on the domain boundary), u_HO = u + dt du_HO
__global__ void k_si_update_bounds(int64_t n, double dt, const double *u, const double *du_ho, const double *si,
                                   double *xi_min, double *xi_max)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   const double t = si[i], uh = u[i] + dt * du_ho[i];
   xi_min[i] = fmax(0.0, t * uh + (1.0 - t) * xi_min[i]);
   xi_max[i] = fmin(1.0, t * uh + (1.0 - t) * xi_max[i]);
}

// ---- NonlinearPenaltySolver (remhos_fct.cpp:760-996).  k_penalty_flux: the clipped rate and the two
// non-conservative fluxes per dof (:787-807); k_penalty_correct: CorrectFlux + get_lambda (:843-996)
// per element, one thread each, every sum in DOF order and every branch as in the reference, so the
// bisection takes the same path as the host loop it replaces.  Both loops of get_lambda are capped at
// 200 rounds (the reference has no cap; the bracket has collapsed to one double long before).
__global__ void k_penalty_flux(int64_t n, double dt, const double *u, const double *m, const double *du_ho,
                               const double *du_lo, const double *xi_min, const double *xi_max, double *fL,
                               double *fH)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   // note that this uses u(i) at the old time (:801)
   const double star = fmin((xi_max[i] - u[i]) / dt, fmax(du_ho[i], (xi_min[i] - u[i]) / dt));
   fL[i] = m[i] * (star - du_lo[i]);
   fH[i] = m[i] * (star - du_ho[i]);
}

__device__ __forceinline__ double penalty_sum_z(double lam, int nd, const double *w, const double *fl)
{
   double acc = 0.0;
   for (int j = 0; j < nd; j++) { acc += (fabs(fl[j]) >= lam * fabs(w[j])) ? lam * w[j] : fl[j]; }
   return acc;
}

__global__ void k_penalty_correct(int64_t ne, int nd, double eps_w, const double *m, const double *du_lo,
                                  const double *fL, const double *fH, double *w_all, double *du)
{
   const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= ne) { return; }
   const double *fl = fL + e * nd, *fh = fH + e * nd;
   double *w = w_all + e * nd;
   double fp = 0.0, fn = 0.0;
   for (int j = 0; j < nd; j++) { if (fl[j] >= 0.0) { fp += fl[j]; } else { fn += fl[j]; } }
   const double delta = fp + fn;
   if (delta == 0.0)
   {
      for (int j = 0; j < nd; j++) { du[e * nd + j] = du_lo[e * nd + j] + fl[j] / m[e * nd + j]; }
      return;
   }
   double mx = -1.0;                                           // get_max_on_cellNi
   for (int j = 0; j < nd; j++) { mx = fmax(fabs(fh[j]), mx); }
   for (int j = 0; j < nd; j++)
   {
      if (delta > 0.0) { w[j] = (fl[j] > 0.0) ? eps_w * fabs(fl[j]) + fabs(mx) : 0.0; }
      else { w[j] = (fl[j] < 0.0) ? -eps_w * fabs(fl[j]) - fabs(mx) : 0.0; }
   }
   // get_lambda
   const double tol = 1e-15;
   double lam = 1.0;
   double F = delta - penalty_sum_z(lam, nd, w, fl);
   double lo = 0.0, hi = 0.0, FL = 0.0, FU = 0.0, factor = 1.0;
   for (int it = 0; it < 200; it++)
   {
      factor *= 2.0;
      lo = lam / factor; hi = factor * lam;
      FL = delta - penalty_sum_z(lo, nd, w, fl);
      FU = delta - penalty_sum_z(hi, nd, w, fl);
      if (!(F * FL > 0 && F * FU > 0)) { break; }
   }
   if (F * FL < 0) { hi = lam; } else { lo = lam; }
   FL = delta - penalty_sum_z(lo, nd, w, fl);
   FU = delta - penalty_sum_z(hi, nd, w, fl);
   for (int it = 0; it < 200; it++)
   {
      lam = 0.5 * (lo + hi);
      F = delta - penalty_sum_z(lam, nd, w, fl);
      if (F * FL < 0) { hi = lam; FU = F; } else { lo = lam; FL = F; }
      if (!(fabs(F) > tol)) { break; }
   }
   lam = 0.5 * (lo + hi);                                       // (:924: the midpoint of the last bracket)
   (void)FU;
   for (int j = 0; j < nd; j++)
   {
      const double z = (fabs(fl[j]) >= lam * fabs(w[j])) ? lam * w[j] : fl[j];
      du[e * nd + j] = du_lo[e * nd + j] + (fl[j] + (-z)) / m[e * nd + j];
   }
}

// AdvectionOperator::UpdateTimeStepEstimate (remhos.cpp:1968-1998): ratio <- min(ratio, dt_est / dt)
__global__ void k_dt_estimate(int64_t n, double dt, const double *x, const double *dx,
                              const double *x_min, const double *x_max, double *ratio)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   double r = INFINITY;
   if (i < n)
   {
      const double d = dx[i];
      if (d > 1e-12) { r = (x_max[i] - x[i]) / d; }
      else if (d < -1e-12) { r = (x_min[i] - x[i]) / d; }
   }
   r = warp_min(r);
   if ((threadIdx.x & 31) == 0 && r < INFINITY)
   {
      r = (dt != 0.0) ? r / dt : 0.0;
      unsigned long long *p = reinterpret_cast<unsigned long long *>(ratio);
      unsigned long long old = *p;
      while (__longlong_as_double((long long)old) > r)
      {
         const unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(r));
         if (prev == old) { break; }
         old = prev;
      }
   }
}

// ---- SmoothnessIndicator (remhos_tools.cpp:24-354) on device: sparse H1 operators in CSR
// y = A x - sub (sub may be null)
__global__ void k_csr_spmv(int n, const int32_t *I, const int32_t *J, const double *A, const double *x,
                           const double *sub, double *y)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   double s = 0.0;
   for (int k = I[i]; k < I[i + 1]; k++) { s += A[k] * x[J[k]]; }
   y[i] = sub ? s - sub[i] : s;
}
// out[0] = sum x_i^2 (one block; the H1 spaces of the meshes this runs on are small)
__global__ void k_sum_sq(int n, const double *x, double *out)
{
   __shared__ double sh[32];
   double s = 0.0;
   for (int i = threadIdx.x; i < n; i += blockDim.x) { s += x[i] * x[i]; }
   s = warp_sum(s);
   if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = s; }
   __syncthreads();
   if (threadIdx.x < 32)
   {
      s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
      s = warp_sum(s);
      if (threadIdx.x == 0) { out[0] = s; }
   }
}
// one sweep of the truncated Neumann series (ApproximateLaplacian, :261-287): z = M y - rhs is
// given; the reference leaves the loop when |z|_2 <= 1e-10
__global__ void k_si_sweep(int n, const double *z, const double *ml, const double *nrm2, double *y)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   if (sqrt(nrm2[0]) <= 1.0e-10) { return; }
   y[i] -= z[i] / ml[i];
}
// min / max of g over the sparsity pattern of the H1 mass matrix, then the indicator (:153-184)
__global__ void k_si_value(int n, const int32_t *I, const int32_t *J, const double *g, int type,
                           double param, double *si)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   double gmin = INFINITY, gmax = -INFINITY;
   for (int k = I[i]; k < I[i + 1]; k++) { const double v = g[J[k]]; gmin = fmin(gmin, v); gmax = fmax(gmax, v); }
   if (type == 1)
   {
      const double eps = 1.0e-50;
      si[i] = 1.0 - pow((fabs(gmin - gmax) + eps) / (fabs(gmin) + fabs(gmax) + eps), param);
   }
   else
   {
      const double eps = 1.0e-15;
      si[i] = fmin(1.0, param * fmax(0.0, gmin * gmax) / (fmax(gmin * gmin, gmax * gmax) + eps));
   }
}
// per DG dof: the indicator at its H1 dof, 1 on the domain boundary (DG2CG < 0)
__global__ void k_si_gather(int64_t n, const int32_t *d2c, const double *si, double *tmp)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { tmp[i] = (d2c[i] < 0) ? 1.0 : si[d2c[i]]; }
}

// ---- FluxBasedFCT (Zalesak), gather form: every DOF visits all its couplings of K_HO
// (in-element via KH / M, across faces via BI of both sides); each flux is evaluated from both
// ends with the same operands in the same order, so f_ji = -f_ij bit for bit.  Across a rank
// boundary the neighbour's trace comes from the ghost array, its face block from BIg (formed on
// this side from the same quadrature points: equal to the owner's up to the summation order, so
// f_ji = -f_ij holds to round-off there) and the visit gets the index -1 - (slot * NFD + b).
template <typename Visit>
__device__ __forceinline__ void flux_visit(const FaArgs &A, const double *u, const double *du_ho,
                                           double dt, int64_t e, int i, const double *ue,
                                           const double *dhe, Visit visit)
{
   const int ND = A.ND, NFD = A.NFD, NF = A.NF;
   const double *KHe = A.KH + (size_t)e * ND * ND, *Me = A.M + (size_t)e * ND * ND;
   const double ui = ue[i], di = dhe[i];
   for (int j = 0; j < ND; j++)
   {
      if (j == i) { continue; }
      const double kij = KHe[i * ND + j], kji = KHe[j * ND + i];
      const double dij = fmax(fmax(0.0, -kij), -kji);
      // remhos_fct.cpp:313-318 + 334-338: dt d_ij (u_i - u_j) + dt M_ij (du_i - du_j)
      double f = dt * dij * (ui - ue[j]) + Me[i * ND + j] * dt * (di - dhe[j]);
      if (A.pbeta)
      {
         // explicit roundings: the same two products enter from either end, so f_ji = -f_ij still holds
         const double t1 = __dmul_rn(A.pbeta[e * ND + j], A.pfel[e * ND + i]);
         const double t2 = __dmul_rn(A.pbeta[e * ND + i], A.pfel[e * ND + j]);
         f = __dadd_rn(f, __dsub_rn(t1, t2));
      }
      visit(e * ND + j, f);
   }
   int l[3];
   dof_lattice(A.dim, A.D1, i, l);
   for (int ax = 0; ax < A.dim; ax++)
   {
      for (int side = 0; side < 2; side++)
      {
         if (l[ax] != side * (A.D1 - 1)) { continue; }
         const int f = face_of(A.dim, ax, side);
         const int64_t nb = A.fn.nbr_elem[e * NF + f];
         if (nb < 0) { continue; }
         const int a = face_nat_index(A.dim, A.D1, l, ax);
         const double *BIe = A.BI + ((size_t)e * NF + f) * NFD * NFD;
         if (nb >= A.fn.ne_owned)
         {
            const int64_t slot = nb - A.fn.ne_owned;
            const double *BIn = A.BIg + (size_t)slot * NFD * NFD;
            for (int b = 0; b < NFD; b++)
            {
               const double kij = BIe[a * NFD + b], kji = BIn[b * NFD + a];
               const double dij = fmax(fmax(0.0, -kij), -kji);
               visit(-1 - (slot * NFD + b), dt * dij * (ui - A.fn.ughost[slot * NFD + b]));
            }
            continue;
         }
         const int pid = A.fn.nbr_pat[e * NF + f];
         const int f2 = A.pat_face[pid];
         const int a2 = A.pat_idx[pid * NFD + a];
         const double *BIn = A.BI + ((size_t)nb * NF + f2) * NFD * NFD;
         for (int b = 0; b < NFD; b++)
         {
            const int b2 = A.pat_idx[pid * NFD + b];
            const int64_t gj = nb * ND + A.fn.pat[pid * NFD + b];
            const double kij = BIe[a * NFD + b], kji = BIn[b2 * NFD + a2];
            const double dij = fmax(fmax(0.0, -kij), -kji);
            visit(gj, dt * dij * (ui - u[gj]));
         }
      }
   }
}

// AddFluxesAtDofs + ComputeFluxCoefficients (remhos_fct.cpp:344-399): block per element
__global__ void k_flux_coeff(FaArgs A, double dt, const double *u, const double *du_ho,
                             const double *du_lo, const double *umin, const double *umax,
                             double *cp, double *cn)
{
   extern __shared__ double sh[];
   const int64_t e = blockIdx.x;
   const int ND = A.ND;
   double *ue = sh, *dhe = sh + ND;
   for (int t = threadIdx.x; t < ND; t += blockDim.x) { ue[t] = u[e * ND + t]; dhe[t] = du_ho[e * ND + t]; }
   __syncthreads();
   for (int i = threadIdx.x; i < ND; i += blockDim.x)
   {
      double gp = 0.0, gm = 0.0;
      flux_visit(A, u, du_ho, dt, e, i, ue, dhe, [&](int64_t, double f)
      {
         if (f >= 0.0) { gp += f; } else { gm += f; }
      });
      const int64_t g = e * ND + i;
      const double m = A.ml[g];
      const double u_lo = ue[i] + dt * du_lo[g];
      const double max_pos = fmax((umax[g] - u_lo) * m, 0.0);
      const double min_neg = fmin((umin[g] - u_lo) * m, 0.0);
      cp[g] = (gp > max_pos) ? max_pos / gp : 1.0;
      cn[g] = (gm < min_neg) ? min_neg / gm : 1.0;
   }
}

// UpdateSolutionAndFlux (remhos_fct.cpp:401-446), single FCT iteration
__global__ void k_flux_apply(FaArgs A, double dt, const double *u, const double *du_ho,
                             const double *du_lo, const double *cp, const double *cn, double *du)
{
   extern __shared__ double sh[];
   const int64_t e = blockIdx.x;
   const int ND = A.ND;
   double *ue = sh, *dhe = sh + ND;
   for (int t = threadIdx.x; t < ND; t += blockDim.x) { ue[t] = u[e * ND + t]; dhe[t] = du_ho[e * ND + t]; }
   __syncthreads();
   for (int i = threadIdx.x; i < ND; i += blockDim.x)
   {
      const int64_t g = e * ND + i;
      const double cpi = cp[g], cni = cn[g];
      double s = 0.0;
      flux_visit(A, u, du_ho, dt, e, i, ue, dhe, [&](int64_t gj, double f)
      {
         const double cnj = (gj < 0) ? A.gcn[-1 - gj] : cn[gj], cpj = (gj < 0) ? A.gcp[-1 - gj] : cp[gj];
         const double a = (f >= 0.0) ? fmin(cpi, cnj) : fmin(cni, cpj);
         s += f * a;
      });
      du[g] = du_lo[g] + s / A.ml[g] / dt;
   }
}

// out = sum_k c[k] * x[k]   (general explicit RK combinations; up to 9 terms)
struct LinComb
{
   int n;
   double c[9];
   const double *x[9];
};
__global__ void k_lincomb(int64_t N, LinComb L, double *out)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= N) { return; }
   double s = 0.0;
   for (int k = 0; k < L.n; k++) { s += L.c[k] * L.x[k][i]; }
   out[i] = s;
}

} // namespace rmh

#endif
