// 3D set-up kernel: all mesh-position dependent operator data of one element per block,
// sum-factorised.  This is what "re-assembly" costs in remap mode, where it runs before every RK
// stage (AdvectionOperator::MultUnlimited, remhos.cpp:1598-1677: move the mesh, drop the geometric
// factors, re-assemble M_HO, K_HO, lumpedM and the face terms):
//   Dvol   = alpha w_q adj(J) v          ConvectionIntegrator PA data (remhos_lo.cpp:1155-1190)
//   detJw  = w_q det J                   MassIntegrator PA data
//   Dface  = w_q min(0, v.n) | -w_q max(0, v.n)   upwinded face data (remhos_tools.cpp:833-845)
//   ml     = B^T detJw                   lumped mass M_HO 1 (remhos.cpp:721-727, 1625-1632)
//   einv   = 1/volume if det J is constant over the element, else 0
//   BL     = -B_f^T Dface                row sums of bdrInt (LinearFluxLumping, alpha = 0)
// The nodal fields (positions x0 + t v, velocity) are interpolated to the quadrature points axis
// by axis (15 k FMA per element at order 3 instead of 150 k for point-wise evaluation).
#ifndef RMH_GEOM3_CUH
#define RMH_GEOM3_CUH

#include "kernels.cuh"

namespace rmh
{

struct Geom3Args
{
   int Q, NG1, D1, exec_mode, frag;
   int64_t ne;
   double t;
   const double *X0, *V, *velq, *velf;      // nodes / nodal velocity / velocity samples (may be NULL)
   const double *L, *dL, *Ls, *dLs, *w, *B; // 1-D tables: [Q][NG1], [Q][NG1], [2][NG1], [2][NG1], [Q], [Q][D1]
   double *Dvol, *detJw, *Dface, *ml, *einv, *BL;
};

__device__ __forceinline__ double block_reduce(double v, double *red, bool is_max)
{
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
   v = is_max ? warp_max(v) : warp_sum(v);
   __syncthreads();
   if (lane == 0) { red[w] = v; }
   __syncthreads();
   double r = is_max ? -INFINITY : 0.0;
   for (int i = 0; i < nw; i++) { r = is_max ? fmax(r, red[i]) : r + red[i]; }
   return r;
}

// compile-time sizes: the index arithmetic (divisions by Q, n, D1) and the short contraction loops
// dominate the instruction count otherwise (48 k warp instructions per element with run-time sizes)
template <int n, int Q, int D1, int T>
__global__ void __launch_bounds__(T) k_geom3(Geom3Args g)
{
   extern __shared__ double sh[];
   constexpr int n2 = n * n, n3 = n2 * n, QQ = Q * Q, NQ = QQ * Q, ND = D1 * D1 * D1, NFD = D1 * D1;
   const int tid = threadIdx.x;
   const int64_t e = blockIdx.x;
   double *Xs = sh;                      // [3][n3] moved nodes
   double *Vs = Xs + 3 * n3;             // [3][n3] nodal velocity
   double *valx = Vs + 3 * n3;           // [6][n2][Q]
   double *derx = valx + 6 * n2 * Q;     // [3][n2][Q]
   double *P = derx + 3 * n2 * Q;        // [12][n][QQ]: X comp c: 3c + {dx, dy, value}; V comp c: 9 + c
   double *dJ = P + 12 * n * QQ;         // [NQ] detJw
   double *t1 = dJ + NQ;                 // [D1][QQ]
   double *t2 = t1 + D1 * QQ;            // [D1][D1][Q]
   double *Fv = t2 + D1 * D1 * Q;        // [6][3][n2] face value layer of X
   double *Fn = Fv + 18 * n2;            // [6][3][n2] normal derivative of X at the face
   double *Fw = Fn + 18 * n2;            // [6][3][n2] face value layer of V
   double *Df = Fw + 18 * n2;            // [6][QQ] face data
   double *red = Df + 6 * QQ;            // [32]
   double *tL = red + 32, *tdL = tL + Q * n, *tB = tdL + Q * n, *tw = tB + Q * D1;   // 1-D tables
   double *tLs = tw + Q;                 // [2][n]
   for (int i = tid; i < Q * n; i += T) { tL[i] = g.L[i]; tdL[i] = g.dL[i]; }
   for (int i = tid; i < Q * D1; i += T) { tB[i] = g.B[i]; }
   for (int i = tid; i < Q; i += T) { tw[i] = g.w[i]; }
   for (int i = tid; i < 2 * n; i += T) { tLs[i] = g.Ls[i]; }
   const bool has_v = (g.V != nullptr);
   double xmax = 0.0;
   for (int i = tid; i < 3 * n3; i += T)
   {
      const int c = i / n3, nd = i - c * n3;
      const size_t src = ((size_t)e * n3 + nd) * 3 + c;
      const double v = has_v ? g.V[src] : 0.0;
      double x = g.X0[src];
      if (g.exec_mode == 1) { x += g.t * v; }
      Xs[i] = x; Vs[i] = v;
      xmax = fmax(xmax, fabs(x));
   }
   __syncthreads();
   // ---- x-stage
   for (int i = tid; i < 6 * n2 * Q; i += T)
   {
      const int fld = i / (n2 * Q), r = i - fld * n2 * Q, line = r / Q, qx = r - line * Q;
      const double *src = (fld < 3 ? Xs + fld * n3 : Vs + (fld - 3) * n3) + line * n;
      double a = 0.0, d = 0.0;
#pragma unroll
      for (int k = 0; k < n; k++) { a += tL[qx * n + k] * src[k]; d += tdL[qx * n + k] * src[k]; }
      valx[i] = a;
      if (fld < 3) { derx[i] = d; }
   }
   __syncthreads();
   // ---- y-stage
   for (int i = tid; i < 12 * n * QQ; i += T)
   {
      const int arr = i / (n * QQ), r = i - arr * n * QQ, kz = r / QQ, rr = r - kz * QQ, qy = rr / Q, qx = rr - qy * Q;
      const double *src; const double *cf;
      if (arr < 9)
      {
         const int c = arr / 3, kind = arr - 3 * c;
         src = (kind == 0 ? derx : valx) + c * n2 * Q;
         cf = (kind == 1 ? tdL : tL);
      }
      else { src = valx + (arr - 6) * n2 * Q; cf = tL; }
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < n; k++) { a += cf[qy * n + k] * src[(kz * n + k) * Q + qx]; }
      P[i] = a;
   }
   __syncthreads();
   // ---- z-stage: one quadrature point per thread iteration
   const double alpha = (g.exec_mode == 1) ? 1.0 : -1.0;      // remhos.cpp:648-657
   constexpr int RQ = (Q + 1) & ~1;
   double vol = 0.0;
   for (int q = tid; q < NQ; q += T)
   {
      const int qz = q / QQ, col = q - qz * QQ;
      double J[3][3], v[3], det, adj[3][3];
#pragma unroll
      for (int c = 0; c < 3; c++)
      {
         double jx = 0.0, jy = 0.0, jz = 0.0, vv = 0.0;
#pragma unroll
         for (int k = 0; k < n; k++)
         {
            const double l = tL[qz * n + k], dl = tdL[qz * n + k];
            jx += l * P[((3 * c + 0) * n + k) * QQ + col];
            jy += l * P[((3 * c + 1) * n + k) * QQ + col];
            jz += dl * P[((3 * c + 2) * n + k) * QQ + col];
            vv += l * P[((9 + c) * n + k) * QQ + col];
         }
         J[c][0] = jx; J[c][1] = jy; J[c][2] = jz; v[c] = vv;
      }
      if (g.velq)
      {
#pragma unroll
         for (int c = 0; c < 3; c++) { v[c] = g.velq[((size_t)e * NQ + q) * 3 + c]; }
      }
      adj[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      adj[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
      adj[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
      adj[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      adj[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
      adj[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
      adj[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      adj[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
      adj[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      det = J[0][0] * adj[0][0] + J[0][1] * adj[1][0] + J[0][2] * adj[2][0];
      const int qy = col / Q, qx = col - qy * Q;
      const double wq = tw[qx] * tw[qy] * tw[qz];
#pragma unroll
      for (int c = 0; c < 3; c++)
      {
         const double s = adj[c][0] * v[0] + adj[c][1] * v[1] + adj[c][2] * v[2];
         if (g.frag) { g.Dvol[(size_t)e * QQ * RQ * 3 + ((size_t)col * RQ + qz) * 3 + c] = alpha * wq * s; }
         else { g.Dvol[((size_t)e * 3 + c) * NQ + q] = alpha * wq * s; }
      }
      const double d = wq * det;
      g.detJw[(size_t)e * NQ + q] = d;
      dJ[q] = d;
      vol += d;
   }
   vol = block_reduce(vol, red, false);
   // constant-Jacobian test (same criterion as k_elem_affine)
   {
      double dev = 0.0;
      for (int q = tid; q < NQ; q += T)
      {
         const int qz = q / QQ, col = q - qz * QQ, qy = col / Q, qx = col - qy * Q;
         dev = fmax(dev, fabs(dJ[q] / (tw[qx] * tw[qy] * tw[qz]) - vol));
      }
      dev = block_reduce(dev, red, true);
      xmax = block_reduce(xmax, red, true);
      if (tid == 0)
      {
         const double h = pow(fabs(vol), 1.0 / 3.0);
         const double tol = 100.0 * 2.220446049250313e-16 * fmax(1.0, xmax / h);
         g.einv[e] = (vol > 0.0 && dev <= tol * vol) ? 1.0 / vol : 0.0;
      }
   }
   // ---- lumped mass: back-contraction of detJw
   for (int i = tid; i < D1 * QQ; i += T)
   {
      const int iz = i / QQ, col = i - iz * QQ;
      double a = 0.0;
      for (int qz = 0; qz < Q; qz++) { a += tB[qz * D1 + iz] * dJ[qz * QQ + col]; }
      t1[i] = a;
   }
   __syncthreads();
   for (int i = tid; i < D1 * D1 * Q; i += T)
   {
      const int iz = i / (D1 * Q), r = i - iz * D1 * Q, iy = r / Q, qx = r - iy * Q;
      double a = 0.0;
      for (int qy = 0; qy < Q; qy++) { a += tB[qy * D1 + iy] * t1[iz * QQ + qy * Q + qx]; }
      t2[i] = a;
   }
   __syncthreads();
   for (int i = tid; i < ND; i += T)
   {
      const int ix = i % D1, r = i / D1;         // r = iz * D1 + iy
      double a = 0.0;
      for (int qx = 0; qx < Q; qx++) { a += tB[qx * D1 + ix] * t2[r * Q + qx]; }
      g.ml[(size_t)e * ND + i] = a;
   }
   // ---- faces: value layer of the nodal fields on each face (the face normal adj(J)[axis][:] only
   // involves the two tangential derivatives)
   for (int i = tid; i < 18 * n2; i += T)
   {
      const int f = i / (3 * n2), r = i - f * 3 * n2, c = r / n2, ab = r - c * n2, ia = ab % n, ib = ab / n;
      int axis, side;
      face_axis_side(3, f, axis, side);
      const int s0 = (axis == 0) ? n : 1, s1 = (axis == 2) ? n : n2, sx = (axis == 0) ? 1 : ((axis == 1) ? n : n2);
      double xv = 0.0, vv = 0.0;
      for (int k = 0; k < n; k++)
      {
         const int nd = k * sx + ia * s0 + ib * s1;
         xv += tLs[side * n + k] * Xs[c * n3 + nd];
         vv += tLs[side * n + k] * Vs[c * n3 + nd];
      }
      Fv[i] = xv; Fw[i] = vv;
   }
   __syncthreads();
   for (int i = tid; i < 6 * QQ; i += T)
   {
      const int f = i / QQ, qf = i - f * QQ, qa = qf % Q, qb = qf / Q;
      int axis, side;
      face_axis_side(3, f, axis, side);
      double ta[3] = {0.0, 0.0, 0.0}, tb[3] = {0.0, 0.0, 0.0}, v[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int ib = 0; ib < n; ib++)
#pragma unroll
         for (int ia = 0; ia < n; ia++)
         {
            const double la = tL[qa * n + ia], lb = tL[qb * n + ib];
            const double dla = tdL[qa * n + ia], dlb = tdL[qb * n + ib];
            const int ab = ib * n + ia;
#pragma unroll
            for (int c = 0; c < 3; c++)
            {
               const double xv = Fv[(f * 3 + c) * n2 + ab];
               ta[c] += dla * lb * xv;
               tb[c] += la * dlb * xv;
               v[c] += la * lb * Fw[(f * 3 + c) * n2 + ab];
            }
         }
      if (g.velf)
      {
#pragma unroll
         for (int c = 0; c < 3; c++) { v[c] = g.velf[(((size_t)e * 6 + f) * QQ + qf) * 3 + c]; }
      }
      // row `axis` of adj(J) = cross product of the other two Jacobian columns in cyclic order:
      // axis 0: ta x tb (columns y, z); axis 1: tb x ta (columns z, x); axis 2: ta x tb (columns x, y)
      const double sg = (axis == 1) ? -1.0 : 1.0;
      double ar[3];
      ar[0] = sg * (ta[1] * tb[2] - ta[2] * tb[1]);
      ar[1] = sg * (ta[2] * tb[0] - ta[0] * tb[2]);
      ar[2] = sg * (ta[0] * tb[1] - ta[1] * tb[0]);
      const double sgn = side ? 1.0 : -1.0;
      const double vn = sgn * (ar[0] * v[0] + ar[1] * v[1] + ar[2] * v[2]);
      const double vs = (g.exec_mode == 1) ? -fmax(0.0, vn) : fmin(0.0, vn);
      const double d = tw[qa] * tw[qb] * vs;
      Df[i] = d;
      if (g.frag) { g.Dface[(size_t)e * 6 * Q * RQ + ((size_t)f * Q + qa) * RQ + qb] = d; }
      else { g.Dface[(size_t)e * 6 * QQ + (size_t)qb * 6 * Q + f * Q + qa] = d; }
   }
   __syncthreads();
   for (int i = tid; i < 6 * NFD; i += T)
   {
      const int f = i / NFD, a = i - f * NFD, a0 = a % D1, a1 = a / D1;
      double s = 0.0;
      for (int qb = 0; qb < Q; qb++)
      {
         double sa = 0.0;
         for (int qa = 0; qa < Q; qa++) { sa += Df[f * QQ + qb * Q + qa] * tB[qa * D1 + a0]; }
         s += sa * tB[qb * D1 + a1];
      }
      g.BL[(size_t)e * 6 * NFD + i] = -s;
   }
}

// shared memory of k_geom3 in doubles
inline size_t geom3_smem_doubles(int Q, int n, int D1)
{
   const int n2 = n * n, n3 = n2 * n, QQ = Q * Q;
   return (size_t)6 * n3 + 9 * n2 * Q + (size_t)12 * n * QQ + (size_t)QQ * Q + (size_t)D1 * QQ +
          (size_t)D1 * D1 * Q + 54 * n2 + 6 * QQ + 32 + 2 * Q * n + Q * D1 + Q + 2 * n;
}

} // namespace rmh

#endif
