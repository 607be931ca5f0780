/* TEST INFRASTRUCTURE ONLY -- CPU port (C99 + OpenMP) of the Remhos RK-stage path, 3D hexahedra.
 *
 * Used (a) as the timed CPU baseline of bench.py (`cpu_baseline.kind = "port"`, and the
 * `--impl reference` arm: the reference's own MFEM/MPI build cannot be produced in this image,
 * SURVEY.md 8c) and (b) as a second, independently written checker next to the numpy oracle
 * (tests/test_oracle_c.py compares the two).  Nothing in the product (remhos_b200/) links it.
 *
 * Restates, for `-ho 3 -lo 5 -fct 2 -pa -s 3` on hexahedral meshes:
 *   PA set-up: ConvectionIntegrator / MassIntegrator / DGTraceIntegrator quadrature data
 *              (remhos.cpp:640-727; formula restated in-repo at remhos_lo.cpp:1155-1190;
 *               upwinded face velocity remhos_tools.cpp:833-845)
 *   LocalInverseHOSolver::CalcHOSolution   remhos_ho.cpp:84-129  (K_HO.Mult :122, M^-1 :126)
 *   MassBasedAvg::CalcLOSolution           remhos_lo.cpp:247-324
 *   DofInfo::ComputeElementsMinMax / ComputeOverlapBounds   remhos_tools.cpp:497-523, 432-495
 *   ClipScaleSolver::CalcFCTSolution       remhos_fct.cpp:449-541
 *   RK3SSPSolver::Step                     remhos.cpp:490 (MFEM)
 * The mass inverse follows the exact-inverse semantics of the FA path (remhos_ho.cpp:100-116)
 * with the cost profile of the PA one: CG preconditioned by the Kronecker inverse of the
 * reference-element mass matrix, exact in one application on constant-Jacobian elements.
 * 1-D tables come from the caller (oracle/remhos_oracle/fe.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 6      /* order + 1 */
#define MAXQ 8      /* 1-D quadrature points */
#define MAXG 4      /* geometry nodes per direction */

typedef struct
{
   int p, D1, Q, G1, exec_mode, n_ent;
   int64_t ne;
   double B[MAXQ][MAXD], G[MAXQ][MAXD], Minv[MAXD][MAXD], w[MAXQ];
   double L[MAXQ][MAXG], dL[MAXQ][MAXG], Ls[2][MAXG], dLs[2][MAXG];
   const double *X0, *V, *velq, *velf;    /* borrowed from the caller, must stay alive */
   const int32_t *nbr;                    /* [ne][6][nfd] NbrDof */
   const int32_t *lat;                    /* [ne][27] */
   int32_t *ent_off, *ent_el;
   double *Dvol, *detJw, *Dface, *ml, *einv, *xe_min, *xe_max, *ent_mm;
   double *w1, *w2;
   double t_cur;
} roc_t;

/* local face -> (axis, side): bottom south east north west top (remhos_tools.cpp:1086-1286) */
static void face_axis_side(int f, int *axis, int *side)
{
   *axis = (f == 0 || f == 5) ? 2 : ((f == 1 || f == 3) ? 1 : 0);
   *side = (f == 2 || f == 3 || f == 5) ? 1 : 0;
}

static void det_adj3(const double J[3][3], double *det, double adj[3][3])
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
   *det = J[0][0] * adj[0][0] + J[0][1] * adj[1][0] + J[0][2] * adj[2][0];
}

/* Jacobian and interpolated nodal velocity at a point given by its per-axis 1-D basis rows */
static void eval_geom(const roc_t *c, int64_t e, const double *l[3], const double *dl[3],
                      double J[3][3], double v[3])
{
   const int n1 = c->G1, nn = n1 * n1 * n1;
   const double *X = c->X0 + (size_t)e * nn * 3;
   const double *V = c->V ? c->V + (size_t)e * nn * 3 : NULL;
   memset(J, 0, 9 * sizeof(double));
   v[0] = v[1] = v[2] = 0.0;
   for (int k = 0; k < n1; k++)
      for (int j = 0; j < n1; j++)
         for (int i = 0; i < n1; i++)
         {
            const int n = i + n1 * (j + n1 * k);
            double x[3];
            for (int d = 0; d < 3; d++)
            {
               x[d] = X[n * 3 + d];
               if (c->exec_mode == 1) { x[d] += c->t_cur * V[n * 3 + d]; }
            }
            const double g0 = dl[0][i] * l[1][j] * l[2][k];
            const double g1 = l[0][i] * dl[1][j] * l[2][k];
            const double g2 = l[0][i] * l[1][j] * dl[2][k];
            const double lv = l[0][i] * l[1][j] * l[2][k];
            for (int d = 0; d < 3; d++)
            {
               J[d][0] += g0 * x[d]; J[d][1] += g1 * x[d]; J[d][2] += g2 * x[d];
               if (V) { v[d] += lv * V[n * 3 + d]; }
            }
         }
}

static void assemble(roc_t *c)
{
   const int Q = c->Q, D1 = c->D1, NQ = Q * Q * Q, NQF = Q * Q, ND = D1 * D1 * D1;
   const double alpha = (c->exec_mode == 1) ? 1.0 : -1.0;       /* remhos.cpp:648-657 */
#pragma omp parallel for schedule(static)
   for (int64_t e = 0; e < c->ne; e++)
   {
      double vol = 0.0, dmin = INFINITY, dmax = -INFINITY;
      for (int q = 0; q < NQ; q++)
      {
         const int qx = q % Q, qy = (q / Q) % Q, qz = q / (Q * Q);
         const double *l[3] = {c->L[qx], c->L[qy], c->L[qz]};
         const double *dl[3] = {c->dL[qx], c->dL[qy], c->dL[qz]};
         double J[3][3], v[3], det, adj[3][3];
         eval_geom(c, e, l, dl, J, v);
         det_adj3(J, &det, adj);
         if (c->velq) { for (int d = 0; d < 3; d++) { v[d] = c->velq[((size_t)e * NQ + q) * 3 + d]; } }
         const double wq = c->w[qx] * c->w[qy] * c->w[qz];
         for (int a = 0; a < 3; a++)
         {
            c->Dvol[((size_t)e * 3 + a) * NQ + q] =
               alpha * wq * (adj[a][0] * v[0] + adj[a][1] * v[1] + adj[a][2] * v[2]);
         }
         c->detJw[(size_t)e * NQ + q] = wq * det;
         vol += wq * det;
         dmin = fmin(dmin, det); dmax = fmax(dmax, det);
      }
      /* constant Jacobian determinant (up to round-off of the coordinate differences) */
      c->einv[e] = (vol > 0.0 && (dmax - dmin) <= 1e-11 * fabs(dmax)) ? 1.0 / vol : 0.0;
      for (int f = 0; f < 6; f++)
      {
         int axis, side;
         face_axis_side(f, &axis, &side);
         for (int qf = 0; qf < NQF; qf++)
         {
            const int qa = qf % Q, qb = qf / Q;
            const double *l[3], *dl[3];
            int m = 0;
            const int qq[2] = {qa, qb};
            for (int a = 0; a < 3; a++)
            {
               if (a == axis) { l[a] = c->Ls[side]; dl[a] = c->dLs[side]; }
               else { l[a] = c->L[qq[m]]; dl[a] = c->dL[qq[m]]; m++; }
            }
            double J[3][3], v[3], det, adj[3][3];
            eval_geom(c, e, l, dl, J, v);
            det_adj3(J, &det, adj);
            if (c->velf) { for (int d = 0; d < 3; d++) { v[d] = c->velf[(((size_t)e * 6 + f) * NQF + qf) * 3 + d]; } }
            const double sgn = side ? 1.0 : -1.0;
            const double vn = sgn * (adj[axis][0] * v[0] + adj[axis][1] * v[1] + adj[axis][2] * v[2]);
            /* remhos_tools.cpp:833-845: transport min(0, v.n), remap -max(0, v.n) */
            const double vs = (c->exec_mode == 1) ? -fmax(0.0, vn) : fmin(0.0, vn);
            c->Dface[((size_t)e * 6 + f) * NQF + qf] = c->w[qa] * c->w[qb] * vs;
         }
      }
      /* lumped mass m_i = sum_q phi_i detJw (M_HO * 1, remhos.cpp:721-727) */
      for (int i = 0; i < ND; i++)
      {
         const int ix = i % D1, iy = (i / D1) % D1, iz = i / (D1 * D1);
         double s = 0.0;
         for (int q = 0; q < NQ; q++)
         {
            const int qx = q % Q, qy = (q / Q) % Q, qz = q / (Q * Q);
            s += c->B[qx][ix] * c->B[qy][iy] * c->B[qz][iz] * c->detJw[(size_t)e * NQ + q];
         }
         c->ml[(size_t)e * ND + i] = s;
      }
   }
}

/* out[a][o][s] = sum_i M[o][i] in[a][i][s]; M row stride ldm; a < pre, s < post */
static inline __attribute__((always_inline)) void
contract(const double *M, int ldm, int transposed, int nout, int nin, const double *in, double *out,
         int pre, int post)
{
   for (int a = 0; a < pre; a++)
      for (int o = 0; o < nout; o++)
      {
         double *po = out + ((size_t)a * nout + o) * post;
         for (int s = 0; s < post; s++) { po[s] = 0.0; }
         for (int i = 0; i < nin; i++)
         {
            const double m = transposed ? M[i * ldm + o] : M[o * ldm + i];
            const double *pi = in + ((size_t)a * nin + i) * post;
            for (int s = 0; s < post; s++) { po[s] += m * pi[s]; }
         }
      }
}

/* rhs = K_HO u on one element: sum-factorised volume term + upwinded face terms */
static inline __attribute__((always_inline)) void
ho_apply(const roc_t *c, const int D1, const int Q, int64_t e, const double *ue, const double *ug,
         double *rhs, double *wk)
{
   const int ND = D1 * D1 * D1, NQ = Q * Q * Q, NFD = D1 * D1, NQF = Q * Q;
   double *t1 = wk, *t2 = t1 + 2 * NQ, *g0 = t2 + 3 * NQ, *g1 = g0 + NQ, *g2 = g1 + NQ;   /* 8 NQ */
   const double *B = &c->B[0][0], *G = &c->G[0][0];
   /* x: [z][y][x] -> Bu, Gu [z][y][qx] */
   double *Bu = t1, *Gu = t1 + D1 * D1 * Q;
   contract(B, MAXD, 0, Q, D1, ue, Bu, D1 * D1, 1);
   contract(G, MAXD, 0, Q, D1, ue, Gu, D1 * D1, 1);
   /* y */
   double *BB = t2, *BG = BB + D1 * Q * Q, *GB = BG + D1 * Q * Q;   /* 3*D1*Q*Q <= 3*NQ */
   contract(B, MAXD, 0, Q, D1, Bu, BB, D1, Q);
   contract(G, MAXD, 0, Q, D1, Bu, BG, D1, Q);
   contract(B, MAXD, 0, Q, D1, Gu, GB, D1, Q);
   /* z */
   contract(B, MAXD, 0, Q, D1, GB, g0, 1, Q * Q);
   contract(B, MAXD, 0, Q, D1, BG, g1, 1, Q * Q);
   contract(G, MAXD, 0, Q, D1, BB, g2, 1, Q * Q);
   const double *d = c->Dvol + (size_t)e * 3 * NQ;
   for (int q = 0; q < NQ; q++) { g0[q] = d[q] * g0[q] + d[NQ + q] * g1[q] + d[2 * NQ + q] * g2[q]; }
   contract(B, MAXD, 1, D1, Q, g0, t1, 1, Q * Q);
   contract(B, MAXD, 1, D1, Q, t1, t2, D1, Q);
   contract(B, MAXD, 1, D1, Q, t2, rhs, D1 * D1, 1);
   /* faces: rhs_i += sum_q phi_i Dface (u_own - u_nbr), exterior state 0 on the boundary */
   for (int f = 0; f < 6; f++)
   {
      int axis, side;
      face_axis_side(f, &axis, &side);
      const int st[3] = {1, D1, D1 * D1};
      const int sa = (axis == 0) ? st[1] : st[0], sb = (axis == 2) ? st[1] : st[2];
      const int base = side * (D1 - 1) * st[axis];
      double fd[MAXD * MAXD], f1[MAXD * MAXQ], f2[MAXQ * MAXQ];
      const int32_t *nb = c->nbr + ((size_t)e * 6 + f) * NFD;
      for (int j = 0; j < NFD; j++)
      {
         const int a = j % D1, b = j / D1;
         const double own = ue[base + a * sa + b * sb];
         fd[j] = own - (nb[j] >= 0 ? ug[nb[j]] : 0.0);
      }
      contract(B, MAXD, 0, Q, D1, fd, f1, D1, 1);          /* [b][qa] */
      contract(B, MAXD, 0, Q, D1, f1, f2, 1, Q);           /* [qb][qa] */
      const double *df = c->Dface + ((size_t)e * 6 + f) * NQF;
      for (int q = 0; q < NQF; q++) { f2[q] *= df[q]; }
      contract(B, MAXD, 1, D1, Q, f2, f1, 1, Q);           /* [b][qa] */
      contract(B, MAXD, 1, D1, Q, f1, fd, D1, 1);          /* [b][a] */
      for (int j = 0; j < NFD; j++)
      {
         const int a = j % D1, b = j / D1;
         rhs[base + a * sa + b * sb] += fd[j];
      }
   }
   (void)ND;
}

static inline __attribute__((always_inline)) void
kron_apply(const roc_t *c, const int D1, const double *r, double *z, double *wk)
{
   const double *M = &c->Minv[0][0];
   contract(M, MAXD, 0, D1, D1, r, wk, D1 * D1, 1);
   contract(M, MAXD, 0, D1, D1, wk, z, D1, D1);
   contract(M, MAXD, 0, D1, D1, z, wk, 1, D1 * D1);
   memcpy(z, wk, sizeof(double) * D1 * D1 * D1);
}

static inline __attribute__((always_inline)) void
mass_apply(const roc_t *c, const int D1, const int Q, int64_t e, const double *p, double *ap,
           double *wk)
{
   const int NQ = Q * Q * Q;
   const double *B = &c->B[0][0];
   double *t1 = wk, *t2 = wk + NQ, *t3 = t2 + NQ;
   contract(B, MAXD, 0, Q, D1, p, t1, D1 * D1, 1);
   contract(B, MAXD, 0, Q, D1, t1, t2, D1, Q);
   contract(B, MAXD, 0, Q, D1, t2, t3, 1, Q * Q);
   const double *d = c->detJw + (size_t)e * NQ;
   for (int q = 0; q < NQ; q++) { t3[q] *= d[q]; }
   contract(B, MAXD, 1, D1, Q, t3, t1, 1, Q * Q);
   contract(B, MAXD, 1, D1, Q, t1, t2, D1, Q);
   contract(B, MAXD, 1, D1, Q, t2, ap, D1 * D1, 1);
}

/* x = M_e^-1 r */
static inline __attribute__((always_inline)) void
mass_solve(const roc_t *c, const int D1, const int Q, int64_t e, double *r, double *x, double *wk)
{
   const int ND = D1 * D1 * D1;
   double *z = wk, *pp = z + ND, *ap = pp + ND, *wk2 = ap + ND;
   if (c->einv[e] > 0.0)
   {
      kron_apply(c, D1, r, x, wk2);
      for (int i = 0; i < ND; i++) { x[i] *= c->einv[e]; }
      return;
   }
   double scale = 0.0;
   for (int i = 0; i < ND; i++) { scale = fmax(scale, fabs(r[i])); }
   for (int i = 0; i < ND; i++) { x[i] = 0.0; }
   if (!(scale > 1e-290)) { return; }
   for (int i = 0; i < ND; i++) { r[i] /= scale; }
   kron_apply(c, D1, r, z, wk2);
   double rz = 0.0;
   for (int i = 0; i < ND; i++) { pp[i] = z[i]; rz += r[i] * z[i]; }
   const double rz0 = rz;
   for (int it = 0; it < 60 && rz0 > 0.0; it++)
   {
      mass_apply(c, D1, Q, e, pp, ap, wk2);
      double pap = 0.0;
      for (int i = 0; i < ND; i++) { pap += pp[i] * ap[i]; }
      const double al = rz / pap;
      for (int i = 0; i < ND; i++) { x[i] += al * pp[i]; r[i] -= al * ap[i]; }
      kron_apply(c, D1, r, z, wk2);
      double rzn = 0.0;
      for (int i = 0; i < ND; i++) { rzn += r[i] * z[i]; }
      if (!(rzn > 1e-28 * rz0)) { break; }
      const double be = rzn / rz;
      for (int i = 0; i < ND; i++) { pp[i] = z[i] + be * pp[i]; }
      rz = rzn;
   }
   for (int i = 0; i < ND; i++) { x[i] *= scale; }
}

static int lattice_class(int D1, int i)
{
   int t = 0, mul = 1;
   for (int a = 0; a < 3; a++)
   {
      const int l = i % D1; i /= D1;
      t += ((l == 0) ? 0 : ((l == D1 - 1) ? 2 : 1)) * mul; mul *= 3;
   }
   return t;
}

/* out = a*x0 + b*(y + dt*F(y)) (out_mode 1) or F(y) (out_mode 0) */
static inline __attribute__((always_inline)) void
stage_impl(roc_t *c, const int D1, const int Q, double dt, int out_mode, double a, double b,
           const double *x0, const double *y, double *out)
{
   const int ND = D1 * D1 * D1, NQ = Q * Q * Q;
   const int64_t ne = c->ne;
#pragma omp parallel for schedule(static)
   for (int64_t e = 0; e < ne; e++)                         /* remhos_tools.cpp:497-523 */
   {
      double mn = INFINITY, mx = -INFINITY;
      for (int i = 0; i < ND; i++) { mn = fmin(mn, y[e * ND + i]); mx = fmax(mx, y[e * ND + i]); }
      c->xe_min[e] = mn; c->xe_max[e] = mx;
   }
#pragma omp parallel for schedule(static)
   for (int32_t k = 0; k < c->n_ent; k++)                   /* remhos_tools.cpp:449-466 */
   {
      double mn = INFINITY, mx = -INFINITY;
      for (int32_t j = c->ent_off[k]; j < c->ent_off[k + 1]; j++)
      {
         mn = fmin(mn, c->xe_min[c->ent_el[j]]); mx = fmax(mx, c->xe_max[c->ent_el[j]]);
      }
      c->ent_mm[2 * k] = mn; c->ent_mm[2 * k + 1] = mx;
   }
   int cls[MAXD * MAXD * MAXD];
   for (int i = 0; i < ND; i++) { cls[i] = lattice_class(D1, i); }
#pragma omp parallel
   {
      /* ho_apply: 8 NQ; rhs, du: 2 ND; mass_solve: 3 ND + 3 NQ */
      double *wk = (double *)malloc(sizeof(double) * (11 * NQ + 5 * ND + 64));
      double *rhs = wk + 8 * NQ, *du = rhs + ND, *swk = du + ND;
#pragma omp for schedule(static)
      for (int64_t e = 0; e < ne; e++)
      {
         const double *ue = y + e * ND, *m = c->ml + e * ND;
         ho_apply(c, D1, Q, e, ue, y, rhs, wk);
         mass_solve(c, D1, Q, e, rhs, du, swk);
         /* MassBasedAvg, remhos_lo.cpp:278-285 */
         double s1 = 0.0, s0 = 0.0;
         for (int i = 0; i < ND; i++) { s1 += m[i] * (ue[i] + dt * du[i]); s0 += m[i]; }
         const double ubar = s1 / s0;
         /* ClipScale, remhos_fct.cpp:490-539 */
         double f[MAXD * MAXD * MAXD], lo[MAXD * MAXD * MAXD];
         double sumPos = 0.0, sumNeg = 0.0;
         for (int i = 0; i < ND; i++)
         {
            const int ent = c->lat[e * 27 + cls[i]];
            const double umin = c->ent_mm[2 * ent], umax = c->ent_mm[2 * ent + 1];
            lo[i] = (ubar - ue[i]) / dt;
            const double u_new_lo = ue[i] + dt * lo[i];
            const double fmn = m[i] / dt * (umin - u_new_lo), fmx = m[i] / dt * (umax - u_new_lo);
            double fc = m[i] * (du[i] - lo[i]);
            fc = fmin(fmx, fmax(fmn, fc));
            f[i] = fc;
            sumNeg += fmin(fc, 0.0); sumPos += fmax(fc, 0.0);
         }
         const double new_mass = sumNeg + sumPos, eps = 1.0e-15;
         for (int i = 0; i < ND; i++)
         {
            double fc = f[i];
            if (new_mass > eps) { fc = fmin(0.0, fc) - fmax(0.0, fc) * sumNeg / sumPos; }
            if (new_mass < -eps) { fc = fmax(0.0, fc) - fmin(0.0, fc) * sumPos / sumNeg; }
            const double k = lo[i] + fc / m[i];
            out[e * ND + i] = out_mode ? a * x0[e * ND + i] + b * (ue[i] + dt * k) : k;
         }
      }
      free(wk);
   }
}

#define ROC_DISPATCH(CALL)                                             \
   switch (c->D1 * 100 + c->Q)                                          \
   {                                                                    \
      case 204: { enum { D1 = 2, Q = 4 }; CALL; } break;                \
      case 305: { enum { D1 = 3, Q = 5 }; CALL; } break;                \
      case 406: { enum { D1 = 4, Q = 6 }; CALL; } break;                \
      case 507: { enum { D1 = 5, Q = 7 }; CALL; } break;                \
      default: return 2;                                                \
   }

int roc_stage(roc_t *c, double dt, const double *u, double *k)
{
   ROC_DISPATCH(stage_impl(c, D1, Q, dt, 0, 0.0, 0.0, u, u, k));
   return 0;
}

int roc_set_time(roc_t *c, double t)
{
   if (c->exec_mode == 1) { c->t_cur = t; assemble(c); }
   return 0;
}

/* RK3SSPSolver::Step: stage times t, t+dt, t+dt/2 */
int roc_rk3_step(roc_t *c, double t, double dt, double *u)
{
   roc_set_time(c, t);
   ROC_DISPATCH(stage_impl(c, D1, Q, dt, 1, 0.0, 1.0, u, u, c->w1));
   roc_set_time(c, t + dt);
   ROC_DISPATCH(stage_impl(c, D1, Q, dt, 1, 0.75, 0.25, u, c->w1, c->w2));
   roc_set_time(c, t + dt / 2);
   ROC_DISPATCH(stage_impl(c, D1, Q, dt, 1, 1.0 / 3.0, 2.0 / 3.0, u, c->w2, c->w1));
   memcpy(u, c->w1, sizeof(double) * (size_t)c->ne * c->D1 * c->D1 * c->D1);
   return 0;
}

void roc_lumped_mass(const roc_t *c, double *m)
{
   memcpy(m, c->ml, sizeof(double) * (size_t)c->ne * c->D1 * c->D1 * c->D1);
}

/* n > 0: use n OpenMP threads from now on (launchers such as torchrun export OMP_NUM_THREADS=1) */
void roc_set_threads(int n)
{
#ifdef _OPENMP
   if (n > 0) { omp_set_num_threads(n); }
#else
   (void)n;
#endif
}

int roc_threads(void)
{
#ifdef _OPENMP
   return omp_get_max_threads();
#else
   return 1;
#endif
}

/* tables: B, G [Q][D1]; Minv [D1][D1]; w [Q]; L, dL [Q][G1]; Ls, dLs [2][G1] */
roc_t *roc_create(int p, int mesh_order, int Q, int exec_mode, int64_t ne, const double *nodes,
                  const double *vel_nodes, const double *vel_quad, const double *vel_face,
                  const int32_t *nbr_dof, const int32_t *lat, int n_ent, const double *B,
                  const double *G, const double *Minv, const double *w, const double *L,
                  const double *dL, const double *Ls, const double *dLs)
{
   if (p + 1 > MAXD || Q > MAXQ || mesh_order + 1 > MAXG) { return NULL; }
   roc_t *c = (roc_t *)calloc(1, sizeof(roc_t));
   c->p = p; c->D1 = p + 1; c->Q = Q; c->G1 = mesh_order + 1; c->exec_mode = exec_mode;
   c->ne = ne; c->n_ent = n_ent;
   const int D1 = c->D1, G1 = c->G1;
   for (int q = 0; q < Q; q++)
   {
      c->w[q] = w[q];
      for (int i = 0; i < D1; i++) { c->B[q][i] = B[q * D1 + i]; c->G[q][i] = G[q * D1 + i]; }
      for (int i = 0; i < G1; i++) { c->L[q][i] = L[q * G1 + i]; c->dL[q][i] = dL[q * G1 + i]; }
   }
   for (int s = 0; s < 2; s++)
      for (int i = 0; i < G1; i++) { c->Ls[s][i] = Ls[s * G1 + i]; c->dLs[s][i] = dLs[s * G1 + i]; }
   for (int i = 0; i < D1; i++)
      for (int j = 0; j < D1; j++) { c->Minv[i][j] = Minv[i * D1 + j]; }
   c->X0 = nodes; c->V = vel_nodes; c->velq = vel_quad; c->velf = vel_face;
   c->nbr = nbr_dof; c->lat = lat;
   const size_t NQ = (size_t)Q * Q * Q, ND = (size_t)D1 * D1 * D1, N = ND * ne;
   c->Dvol = (double *)malloc(sizeof(double) * 3 * NQ * ne);
   c->detJw = (double *)malloc(sizeof(double) * NQ * ne);
   c->Dface = (double *)malloc(sizeof(double) * 6 * Q * Q * ne);
   c->ml = (double *)malloc(sizeof(double) * N);
   c->einv = (double *)malloc(sizeof(double) * ne);
   c->xe_min = (double *)malloc(sizeof(double) * ne);
   c->xe_max = (double *)malloc(sizeof(double) * ne);
   c->ent_mm = (double *)malloc(sizeof(double) * 2 * (size_t)n_ent);
   c->w1 = (double *)malloc(sizeof(double) * N);
   c->w2 = (double *)malloc(sizeof(double) * N);
   /* entity -> elements CSR */
   c->ent_off = (int32_t *)calloc((size_t)n_ent + 1, sizeof(int32_t));
   c->ent_el = (int32_t *)malloc(sizeof(int32_t) * 27 * (size_t)ne);
   for (size_t i = 0; i < 27 * (size_t)ne; i++) { c->ent_off[lat[i] + 1]++; }
   for (int k = 0; k < n_ent; k++) { c->ent_off[k + 1] += c->ent_off[k]; }
   int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_ent);
   memcpy(cur, c->ent_off, sizeof(int32_t) * (size_t)n_ent);
   for (int64_t e = 0; e < ne; e++)
      for (int t = 0; t < 27; t++) { c->ent_el[cur[lat[e * 27 + t]]++] = (int32_t)e; }
   free(cur);
   /* first touch in parallel so pages land near the threads that use them */
#pragma omp parallel for schedule(static)
   for (int64_t e = 0; e < ne; e++)
   {
      for (size_t i = 0; i < ND; i++) { c->w1[e * ND + i] = 0.0; c->w2[e * ND + i] = 0.0; }
   }
   c->t_cur = 0.0;
   assemble(c);
   return c;
}

void roc_destroy(roc_t *c)
{
   if (!c) { return; }
   free(c->Dvol); free(c->detJw); free(c->Dface); free(c->ml); free(c->einv); free(c->xe_min);
   free(c->xe_max); free(c->ent_mm); free(c->w1); free(c->w2); free(c->ent_off); free(c->ent_el);
   free(c);
}
