// 3D hexahedral fast path of the RK-stage kernels (sm_100a, FP64).
//
// Thread mapping: every element of the block's batch owns TPE = max(32, Q*Q) threads with fixed
// coordinates (a, b) in [0,Q)^2 for the whole kernel.  Each sum-factorisation stage maps (a, b)
// to the two tensor axes that are NOT contracted and keeps the contracted line in registers:
//
//   fwd-x  (a,b) = (y,z)     u[z][y][:]          -> Bx u, Gx u            [z][y][qx]
//   fwd-y  (a,b) = (qx,z)    [z][:][qx]          -> BB, BG, GB            [z][qy][qx]
//   z-fused(a,b) = (qx,qy)   [:][qy][qx]         -> grad at 6 qz points in registers,
//                            x D (prefetched from HBM into registers), Bz^T -> [iz][qy][qx]
//   bwd-y  (a,b) = (qx,iz)   [iz][:][qx]         -> [iz][iy][qx]
//   bwd-x  (a,b) = (iy,iz)   [iz][iy][:]         -> rhs line (+ face contributions)
//
// so consecutive threads touch consecutive shared-memory words in every stage (a is always the
// fastest index), no gradient array at the quadrature points is ever stored, and the stored
// operator data is loaded with fully coalesced accesses one stage ahead of its use.
#ifndef RMH_STAGE3D_CUH
#define RMH_STAGE3D_CUH

#include "kernels.cuh"

namespace rmh
{

template <int D1, int Q, int E>
struct Smem3
{
   static constexpr int ND = D1 * D1 * D1, NQ = Q * Q * Q, QQ = Q * Q;
   static constexpr int NF = 6, NFD = D1 * D1, NQF = Q * Q;
   static constexpr int TPE = QQ > 32 ? QQ : 32;            // threads per element
   static constexpr int T = ((TPE * E + 31) / 32) * 32;     // block size
   static constexpr int SZ_V = E * ND;
   static constexpr int SZ_C0 = 2 * E * D1 * D1 * Q, SZ_C1 = E * NF * D1 * Q;
   static constexpr int SZ_C = SZ_C0 > SZ_C1 ? SZ_C0 : SZ_C1;
   static constexpr int SZ_B0 = 3 * E * D1 * Q * Q;
   static constexpr int SZ_B = SZ_B0 > SZ_C1 ? SZ_B0 : SZ_C1;
   static constexpr int SZ_F = E * NF * NFD;
   static constexpr int OFF_U = 0;
   static constexpr int OFF_R = OFF_U + SZ_V;
   static constexpr int OFF_X = OFF_R + SZ_V;
   static constexpr int OFF_P = OFF_X + SZ_V;
   static constexpr int OFF_Z = OFF_P + SZ_V;
   static constexpr int OFF_C = OFF_Z + SZ_V;
   static constexpr int OFF_B = OFF_C + SZ_C;
   static constexpr int OFF_F = OFF_B + SZ_B;
   static constexpr int TOTAL = OFF_F + SZ_F;
   static constexpr size_t BYTES = (size_t)TOTAL * sizeof(double);
};

// thread coordinates
template <int D1, int Q, int E>
struct Tid3
{
   int e, a, b, r;
   bool ok;      // r < Q*Q and e < E
   __device__ __forceinline__ Tid3()
   {
      using S = Smem3<D1, Q, E>;
      e = threadIdx.x / S::TPE;
      r = threadIdx.x - e * S::TPE;
      b = r / Q;
      a = r - b * Q;
      ok = (r < S::QQ) && (e < E);
   }
};

// ---------------------------------------------------------------- face terms
// gather own - neighbour face DOF differences (global loads; call first, use after a sync)
template <int D1, int Q, int E>
__device__ __forceinline__ void face3_gather(double *sm, const double *__restrict__ ug,
                                             const FaceNbr &fn, int64_t e0, int ne,
                                             const Tid3<D1, Q, E> &t)
{
   using S = Smem3<D1, Q, E>;
   constexpr int ND = S::ND, NF = S::NF, NFD = S::NFD;
   double *FD = sm + S::OFF_F;
   if (t.e >= E) { return; }
   for (int idx = t.r; idx < NF * NFD; idx += S::TPE)
   {
      const int f = idx / NFD, j = idx - f * NFD;
      double d = 0.0;
      if (t.e < ne)
      {
         const int64_t ge = e0 + t.e;
         const double own = ug[ge * ND + face_dof<3, D1>(f, j)];
         const int64_t nb = fn.nbr_elem[ge * NF + f];
         double un = 0.0;
         if (nb >= 0)
         {
            const int loc = fn.pat[(int)fn.nbr_pat[ge * NF + f] * NFD + j];
            un = (nb < fn.ne_owned) ? ug[nb * ND + loc] : fn.ughost[(nb - fn.ne_owned) * ND + loc];
         }
         d = own - un;
      }
      FD[t.e * NF * NFD + idx] = d;
   }
}

// FD (differences) -> FD (face contributions to the rhs), Dface layout [e][qb][f][qa]
template <int D1, int Q, int E>
__device__ __forceinline__ void face3_apply(double *sm, const double *__restrict__ Dface, int ne,
                                            const Tab<D1, Q> &tab, const Tid3<D1, Q, E> &t)
{
   using S = Smem3<D1, Q, E>;
   constexpr int NF = S::NF, NFD = S::NFD, QQ = S::QQ;
   double *FD = sm + S::OFF_F + t.e * NF * NFD;
   double *F1 = sm + S::OFF_C + t.e * NF * D1 * Q;
   double *F1b = sm + S::OFF_B + t.e * NF * D1 * Q;
   const bool live = t.ok && (t.e < ne);
   // prefetch the face data of this thread's (qa = a) column: [qb][f][qa]
   double df[(NF + Q - 1) / Q][Q];
   if (live)
   {
#pragma unroll
      for (int k = 0; k < (NF + Q - 1) / Q; k++)
      {
         const int f = t.b + k * Q;
#pragma unroll
         for (int qb = 0; qb < Q; qb++)
         {
            df[k][qb] = (f < NF) ? Dface[(size_t)t.e * NF * QQ + qb * NF * Q + f * Q + t.a] : 0.0;
         }
      }
   }
   // F1: (a, b) = (jb, f): contract ja -> qa
   if (t.ok && t.a < D1)
   {
      for (int f = t.b; f < NF; f += Q)
      {
         double x[D1];
#pragma unroll
         for (int i = 0; i < D1; i++) { x[i] = FD[(f * D1 + t.a) * D1 + i]; }
#pragma unroll
         for (int q = 0; q < Q; q++)
         {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < D1; i++) { acc = fma(tab.B[q][i], x[i], acc); }
            F1[(f * D1 + t.a) * Q + q] = acc;
         }
      }
   }
   __syncthreads();
   // F2: (a, b) = (qa, f): contract jb -> qb, scale by the face data, contract qb -> ib
   if (t.ok)
   {
#pragma unroll
      for (int k = 0; k < (NF + Q - 1) / Q; k++)
      {
         const int f = t.b + k * Q;
         if (f < NF)
         {
            double x[D1];
#pragma unroll
            for (int i = 0; i < D1; i++) { x[i] = F1[(f * D1 + i) * Q + t.a]; }
            double y[D1];
#pragma unroll
            for (int i = 0; i < D1; i++) { y[i] = 0.0; }
#pragma unroll
            for (int q = 0; q < Q; q++)
            {
               double acc = 0.0;
#pragma unroll
               for (int i = 0; i < D1; i++) { acc = fma(tab.B[q][i], x[i], acc); }
               acc *= live ? df[k][q] : 0.0;
#pragma unroll
               for (int i = 0; i < D1; i++) { y[i] = fma(tab.B[q][i], acc, y[i]); }
            }
#pragma unroll
            for (int i = 0; i < D1; i++) { F1b[(f * D1 + i) * Q + t.a] = y[i]; }
         }
      }
   }
   __syncthreads();
   // F3: (a, b) = (ib, f): contract qa -> ia
   if (t.ok && t.a < D1)
   {
      for (int f = t.b; f < NF; f += Q)
      {
         double x[Q];
#pragma unroll
         for (int q = 0; q < Q; q++) { x[q] = F1b[(f * D1 + t.a) * Q + q]; }
#pragma unroll
         for (int i = 0; i < D1; i++)
         {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < Q; q++) { acc = fma(tab.B[q][i], x[q], acc); }
            FD[(f * D1 + t.a) * D1 + i] = acc;
         }
      }
   }
   __syncthreads();
}

// ---------------------------------------------------------------- volume term (+ face combine)
// R = B^T [D . grad U] + face contributions (FD).  Dvol [e][3][NQ].
template <int D1, int Q, int E, bool WITH_FACES>
__device__ __forceinline__ void vol3_apply(const double *Uall, double *Rall, double *sm,
                                           const double *__restrict__ Dvol, int ne,
                                           const Tab<D1, Q> &tab, const Tid3<D1, Q, E> &t)
{
   using S = Smem3<D1, Q, E>;
   constexpr int ND = S::ND, NQ = S::NQ, QQ = S::QQ, NF = S::NF, NFD = S::NFD;
   const double *U = Uall + t.e * ND;
   double *R = Rall + t.e * ND;
   double *BU = sm + S::OFF_C + t.e * 2 * D1 * D1 * Q, *GU = BU + D1 * D1 * Q;
   double *GB = sm + S::OFF_B + t.e * 3 * D1 * QQ, *BG = GB + D1 * QQ, *BB = BG + D1 * QQ;
   const double *FC = sm + S::OFF_F + t.e * NF * NFD;
   const bool live = t.ok && (t.e < ne);
   // V1 fwd-x: (a, b) = (y, z)
   if (t.ok && t.a < D1 && t.b < D1)
   {
      double x[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = U[(t.b * D1 + t.a) * D1 + i]; }
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
         double vb = 0.0, vg = 0.0;
#pragma unroll
         for (int i = 0; i < D1; i++)
         {
            vb = fma(tab.B[q][i], x[i], vb);
            vg = fma(tab.G[q][i], x[i], vg);
         }
         BU[(t.b * D1 + t.a) * Q + q] = vb;
         GU[(t.b * D1 + t.a) * Q + q] = vg;
      }
   }
   // prefetch this thread's column of the stored operator data (used two stages later)
   double d0[Q], d1[Q], d2[Q];
   if (live)
   {
      const double *dp = Dvol + (size_t)t.e * 3 * NQ + t.r;
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
         d0[q] = dp[q * QQ];
         d1[q] = dp[NQ + q * QQ];
         d2[q] = dp[2 * NQ + q * QQ];
      }
   }
   else
   {
#pragma unroll
      for (int q = 0; q < Q; q++) { d0[q] = 0.0; d1[q] = 0.0; d2[q] = 0.0; }
   }
   __syncthreads();
   // V2 fwd-y: (a, b) = (qx, z)
   if (t.ok && t.b < D1)
   {
      double x1[D1], x2[D1];
#pragma unroll
      for (int i = 0; i < D1; i++)
      {
         x1[i] = BU[(t.b * D1 + i) * Q + t.a];
         x2[i] = GU[(t.b * D1 + i) * Q + t.a];
      }
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
         double bb = 0.0, bg = 0.0, gb = 0.0;
#pragma unroll
         for (int i = 0; i < D1; i++)
         {
            bb = fma(tab.B[q][i], x1[i], bb);
            bg = fma(tab.G[q][i], x1[i], bg);
            gb = fma(tab.B[q][i], x2[i], gb);
         }
         BB[t.b * QQ + q * Q + t.a] = bb;
         BG[t.b * QQ + q * Q + t.a] = bg;
         GB[t.b * QQ + q * Q + t.a] = gb;
      }
   }
   __syncthreads();
   // V3 z-fused: (a, b) = (qx, qy): gradient at the qz points, x D, back-contract z
   if (t.ok)
   {
      double gb[D1], bg[D1], bb[D1];
#pragma unroll
      for (int i = 0; i < D1; i++)
      {
         gb[i] = GB[i * QQ + t.r];
         bg[i] = BG[i * QQ + t.r];
         bb[i] = BB[i * QQ + t.r];
      }
      double tz[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { tz[i] = 0.0; }
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
         double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
         for (int i = 0; i < D1; i++)
         {
            g0 = fma(tab.B[q][i], gb[i], g0);
            g1 = fma(tab.B[q][i], bg[i], g1);
            g2 = fma(tab.G[q][i], bb[i], g2);
         }
         const double s = d0[q] * g0 + d1[q] * g1 + d2[q] * g2;
#pragma unroll
         for (int i = 0; i < D1; i++) { tz[i] = fma(tab.B[q][i], s, tz[i]); }
      }
      // in place: this thread is the only one touching column r of GB
#pragma unroll
      for (int i = 0; i < D1; i++) { GB[i * QQ + t.r] = tz[i]; }
   }
   __syncthreads();
   // V4 bwd-y: (a, b) = (qx, iz)
   double *S2 = BU;   // [iz][iy][qx]
   if (t.ok && t.b < D1)
   {
      double x[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = GB[t.b * QQ + q * Q + t.a]; }
#pragma unroll
      for (int i = 0; i < D1; i++)
      {
         double acc = 0.0;
#pragma unroll
         for (int q = 0; q < Q; q++) { acc = fma(tab.B[q][i], x[q], acc); }
         S2[(t.b * D1 + i) * Q + t.a] = acc;
      }
   }
   __syncthreads();
   // V5 bwd-x + owner-computes face combine: (a, b) = (iy, iz)
   if (t.ok && t.a < D1 && t.b < D1)
   {
      double x[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = S2[(t.b * D1 + t.a) * Q + q]; }
      double rr[D1];
#pragma unroll
      for (int i = 0; i < D1; i++)
      {
         double acc = 0.0;
#pragma unroll
         for (int q = 0; q < Q; q++) { acc = fma(tab.B[q][i], x[q], acc); }
         rr[i] = acc;
      }
      if (WITH_FACES)
      {
         // faces: 0 z=0 (x,y)  1 y=0 (x,z)  2 x=p (y,z)  3 y=p (x,z)  4 x=0 (y,z)  5 z=p (x,y)
         rr[0] += FC[4 * NFD + t.b * D1 + t.a];
         rr[D1 - 1] += FC[2 * NFD + t.b * D1 + t.a];
         if (t.a == 0)
         {
#pragma unroll
            for (int i = 0; i < D1; i++) { rr[i] += FC[1 * NFD + t.b * D1 + i]; }
         }
         if (t.a == D1 - 1)
         {
#pragma unroll
            for (int i = 0; i < D1; i++) { rr[i] += FC[3 * NFD + t.b * D1 + i]; }
         }
         if (t.b == 0)
         {
#pragma unroll
            for (int i = 0; i < D1; i++) { rr[i] += FC[0 * NFD + t.a * D1 + i]; }
         }
         if (t.b == D1 - 1)
         {
#pragma unroll
            for (int i = 0; i < D1; i++) { rr[i] += FC[5 * NFD + t.a * D1 + i]; }
         }
      }
#pragma unroll
      for (int i = 0; i < D1; i++) { R[(t.b * D1 + t.a) * D1 + i] = rr[i]; }
   }
   __syncthreads();
}

// ---------------------------------------------------------------- mass apply  Z = M P
template <int D1, int Q, int E>
__device__ __forceinline__ void mass3_apply(const double *Pall, double *Zall, double *sm,
                                            const double *__restrict__ detJw, int ne,
                                            const Tab<D1, Q> &tab, const Tid3<D1, Q, E> &t)
{
   using S = Smem3<D1, Q, E>;
   constexpr int ND = S::ND, NQ = S::NQ, QQ = S::QQ;
   const double *P = Pall + t.e * ND;
   double *Z = Zall + t.e * ND;
   double *C1 = sm + S::OFF_C + t.e * 2 * D1 * D1 * Q;      // [z][y][qx]
   double *C2 = sm + S::OFF_B + t.e * 3 * D1 * QQ;          // [z][qy][qx]
   const bool live = t.ok && (t.e < ne);
   double dj[Q];
#pragma unroll
   for (int q = 0; q < Q; q++) { dj[q] = live ? detJw[(size_t)t.e * NQ + q * QQ + t.r] : 0.0; }
   if (t.ok && t.a < D1 && t.b < D1)
   {
      double x[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = P[(t.b * D1 + t.a) * D1 + i]; }
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
         double acc = 0.0;
#pragma unroll
         for (int i = 0; i < D1; i++) { acc = fma(tab.B[q][i], x[i], acc); }
         C1[(t.b * D1 + t.a) * Q + q] = acc;
      }
   }
   __syncthreads();
   if (t.ok && t.b < D1)
   {
      double x[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = C1[(t.b * D1 + i) * Q + t.a]; }
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
         double acc = 0.0;
#pragma unroll
         for (int i = 0; i < D1; i++) { acc = fma(tab.B[q][i], x[i], acc); }
         C2[t.b * QQ + q * Q + t.a] = acc;
      }
   }
   __syncthreads();
   if (t.ok)
   {
      double x[D1], tz[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = C2[i * QQ + t.r]; tz[i] = 0.0; }
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
         double acc = 0.0;
#pragma unroll
         for (int i = 0; i < D1; i++) { acc = fma(tab.B[q][i], x[i], acc); }
         acc *= dj[q];
#pragma unroll
         for (int i = 0; i < D1; i++) { tz[i] = fma(tab.B[q][i], acc, tz[i]); }
      }
#pragma unroll
      for (int i = 0; i < D1; i++) { C2[i * QQ + t.r] = tz[i]; }
   }
   __syncthreads();
   if (t.ok && t.b < D1)
   {
      double x[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = C2[t.b * QQ + q * Q + t.a]; }
#pragma unroll
      for (int i = 0; i < D1; i++)
      {
         double acc = 0.0;
#pragma unroll
         for (int q = 0; q < Q; q++) { acc = fma(tab.B[q][i], x[q], acc); }
         C1[(t.b * D1 + i) * Q + t.a] = acc;
      }
   }
   __syncthreads();
   if (t.ok && t.a < D1 && t.b < D1)
   {
      double x[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = C1[(t.b * D1 + t.a) * Q + q]; }
#pragma unroll
      for (int i = 0; i < D1; i++)
      {
         double acc = 0.0;
#pragma unroll
         for (int q = 0; q < Q; q++) { acc = fma(tab.B[q][i], x[q], acc); }
         Z[(t.b * D1 + t.a) * D1 + i] = acc;
      }
   }
   __syncthreads();
}

// Z = (Minv x Minv x Minv) R
template <int D1, int Q, int E>
__device__ __forceinline__ void kron3_apply(const double *Rall, double *Zall,
                                            const Tab<D1, Q> &tab, const Tid3<D1, Q, E> &t)
{
   using S = Smem3<D1, Q, E>;
   constexpr int ND = S::ND;
   const double *R = Rall + t.e * ND;
   double *Z = Zall + t.e * ND;
   const bool on = t.ok && t.a < D1 && t.b < D1;
   if (on)   // x axis: (a, b) = (y, z)
   {
      double x[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = R[(t.b * D1 + t.a) * D1 + i]; }
#pragma unroll
      for (int o = 0; o < D1; o++)
      {
         double acc = 0.0;
#pragma unroll
         for (int i = 0; i < D1; i++) { acc = fma(tab.Minv[o][i], x[i], acc); }
         Z[(t.b * D1 + t.a) * D1 + o] = acc;
      }
   }
   __syncthreads();
   if (on)   // y axis: (a, b) = (x, z), in place
   {
      double x[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = Z[(t.b * D1 + i) * D1 + t.a]; }
#pragma unroll
      for (int o = 0; o < D1; o++)
      {
         double acc = 0.0;
#pragma unroll
         for (int i = 0; i < D1; i++) { acc = fma(tab.Minv[o][i], x[i], acc); }
         Z[(t.b * D1 + o) * D1 + t.a] = acc;
      }
   }
   __syncthreads();
   if (on)   // z axis: (a, b) = (x, y), in place
   {
      double x[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = Z[(i * D1 + t.b) * D1 + t.a]; }
#pragma unroll
      for (int o = 0; o < D1; o++)
      {
         double acc = 0.0;
#pragma unroll
         for (int i = 0; i < D1; i++) { acc = fma(tab.Minv[o][i], x[i], acc); }
         Z[(o * D1 + t.b) * D1 + t.a] = acc;
      }
   }
   __syncthreads();
}

// Element mass solve X = M^-1 R (see mass_solve in kernels.cuh); one warp per element for the
// dot products, the (a,b) mapping for the operator applications.
template <int D1, int Q, int E>
__device__ __forceinline__ void mass3_solve(double *Rv, double *X, double *sm,
                                            const double *__restrict__ detJw, int ne, double tol2,
                                            int maxit, const Tab<D1, Q> &tab,
                                            const Tid3<D1, Q, E> &t)
{
   using S = Smem3<D1, Q, E>;
   constexpr int ND = S::ND;
   double *P = sm + S::OFF_P, *Z = sm + S::OFF_Z;
   const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const bool mine = (w < E);          // this warp owns element w in the reductions
   double scale = 0.0;
   if (mine) { for (int j = lane; j < ND; j += 32) { scale = fmax(scale, fabs(Rv[w * ND + j])); } }
   scale = warp_max(scale);
   const bool nonzero = mine && (w < ne) && (scale > 1e-290) && (scale < 1e290);
   const double iscale = nonzero ? 1.0 / scale : 0.0;
   if (mine) { for (int j = lane; j < ND; j += 32) { Rv[w * ND + j] *= iscale; } }
   __syncthreads();
   kron3_apply<D1, Q, E>(Rv, Z, tab, t);
   double rz = 0.0, rr0 = 0.0;
   if (mine)
   {
      for (int j = lane; j < ND; j += 32)
      {
         const double z = Z[w * ND + j], r = Rv[w * ND + j];
         P[w * ND + j] = z;
         X[w * ND + j] = 0.0;
         rz += r * z;
         rr0 += r * r;
      }
   }
   warp_sum2(rz, rr0);
   bool active = nonzero && (rz > 0.0);
   __syncthreads();
   for (int it = 0; it < maxit; it++)
   {
      if (!__syncthreads_or(active)) { break; }
      mass3_apply<D1, Q, E>(P, Z, sm, detJw, ne, tab, t);
      double pap = 0.0;
      if (mine) { for (int j = lane; j < ND; j += 32) { pap += P[w * ND + j] * Z[w * ND + j]; } }
      pap = warp_sum(pap);
      const double alpha = active ? rz / pap : 0.0;
      double rr = 0.0;
      if (mine)
      {
         for (int j = lane; j < ND; j += 32)
         {
            X[w * ND + j] += alpha * P[w * ND + j];
            const double r = Rv[w * ND + j] - alpha * Z[w * ND + j];
            Rv[w * ND + j] = r;
            rr += r * r;
         }
      }
      rr = warp_sum(rr);
      // converged when |r|_2 <= tol |r0|_2: skips the trailing preconditioner application
      if (active && !(rr > tol2 * rr0)) { active = false; }
      if (!__syncthreads_or(active)) { break; }
      kron3_apply<D1, Q, E>(Rv, Z, tab, t);
      double rzn = 0.0;
      if (mine) { for (int j = lane; j < ND; j += 32) { rzn += Rv[w * ND + j] * Z[w * ND + j]; } }
      rzn = warp_sum(rzn);
      const double beta = active ? rzn / rz : 0.0;
      if (mine)
      {
         for (int j = lane; j < ND; j += 32)
         {
            P[w * ND + j] = active ? Z[w * ND + j] + beta * P[w * ND + j] : 0.0;
         }
      }
      rz = rzn;
      __syncthreads();
   }
   if (mine) { for (int j = lane; j < ND; j += 32) { X[w * ND + j] *= scale; } }
   __syncthreads();
}

} // namespace rmh

#endif
