// 3D hexahedral fast path of the RK-stage kernels (sm_100a, FP64).
//
// Every sum-factorisation stage is a flat list of independent "line tasks" (one contracted line
// held in registers per task); tasks are dealt to threads in order, so the active threads of a
// stage are the first ntasks ones (whole warps idle at the barrier instead of executing masked
// instructions).  Shared-memory layouts always have the un-contracted fastest index mapped to
// consecutive threads:
//
//   fwd-x   (arr,e,z,y)      u[e][z][y][:]      -> BU | GU          [e][z][y][qx]
//   fwd-y   (arr,e,z,qx)     [e][z][:][qx]      -> BB | BG | GB     [e][z][qy][qx]
//   z-fused (e,qy,qx)        [e][:][qy][qx]     -> reference gradient at the Q points of the
//                            column in registers, x D (prefetched from HBM into registers one
//                            stage earlier), contracted back with Bz^T      -> [e][iz][qy][qx]
//   bwd-y   (e,iz,qx)        [e][iz][:][qx]     -> [e][iz][iy][qx]
//   bwd-x   (e,iz,iy)        [e][iz][iy][:]     -> rhs line + owner-computes face contributions
//
// No array at the quadrature points is ever stored; the stored operator data is read with fully
// coalesced loads.
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
   static constexpr int T = ((TPE * E + 31) / 32) * 32;     // block size (>= 32*E and >= QQ*E)
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
   static constexpr int OFF_G = OFF_F + SZ_F;          // face stage buffer F1 (merged path)
   static constexpr int TOTAL = OFF_G + SZ_C1;
   static constexpr size_t BYTES = (size_t)TOTAL * sizeof(double);
};

// out[o] = sum_i M(o,i) x[i] with compile-time indexed coefficients
#define RMH_LINE(NOUT, NIN, MAT, X, OUT)                                      \
   _Pragma("unroll") for (int o_ = 0; o_ < (NOUT); o_++)                      \
   {                                                                          \
      double acc_ = 0.0;                                                      \
      _Pragma("unroll") for (int i_ = 0; i_ < (NIN); i_++)                    \
      { acc_ = fma(MAT(o_, i_), (X)[i_], acc_); }                             \
      (OUT)[o_] = acc_;                                                       \
   }

// operator data of this thread's z-stage / fused-face-stage tasks, loaded into registers at the
// very start of the kernel so that one HBM latency covers every operator stream of the block
template <int D1, int Q, int E>
struct Pre3
{
   using S = Smem3<D1, Q, E>;
   static constexpr int K2 = (E * S::NF * Q + S::T - 1) / S::T;
   double d0[Q], d1[Q], d2[Q];
   double df[K2][Q];
   // frag = false: Dvol [e][3][NQ], Dface [e][qb][f][qa];  frag = true: the fragment-ordered
   // layout of stage3w.cuh, Dvol [e][col][qz (RQ)][3], Dface [e][f][qa][qb (RQ)]
   __device__ __forceinline__ void load(const double *__restrict__ Dvol,
                                        const double *__restrict__ Dface, int64_t e0, int ne,
                                        bool frag)
   {
      constexpr int NQ = S::NQ, QQ = S::QQ, NF = S::NF, T = S::T, RQ = (Q + 1) & ~1;
      const int zid = threadIdx.x, ze = zid / QQ, zr = zid - ze * QQ;
      const bool live = (zid < E * QQ) && (ze < ne);
      const size_t es = frag ? (size_t)QQ * RQ * 3 : (size_t)3 * NQ;
      const int sc = frag ? 1 : NQ, sq = frag ? 3 : QQ, scol = frag ? 3 * RQ : 1;
      const double *dp = Dvol + (size_t)(e0 + ze) * es + (size_t)zr * scol;
#pragma unroll
      for (int q = 0; q < Q; q++)
      {
         d0[q] = live ? __ldcs(dp + q * sq) : 0.0;
         d1[q] = live ? __ldcs(dp + sc + q * sq) : 0.0;
         d2[q] = live ? __ldcs(dp + 2 * sc + q * sq) : 0.0;
      }
      const size_t esf = frag ? (size_t)NF * Q * RQ : (size_t)NF * QQ;
      const int sqb = frag ? 1 : NF * Q, sr = frag ? RQ : 1;
#pragma unroll
      for (int k = 0; k < K2; k++)
      {
         const int id = threadIdx.x + k * T;
         const int e = id / (NF * Q), r = id - e * (NF * Q);
         const bool lv = (id < E * NF * Q) && (e < ne);
#pragma unroll
         for (int q = 0; q < Q; q++)
         {
            df[k][q] = lv ? __ldcs(Dface + (size_t)(e0 + e) * esf + (size_t)q * sqb + (size_t)r * sr) : 0.0;
         }
      }
   }
};

// ---------------------------------------------------------------- face terms
// gather own - neighbour face DOF differences (global loads; call first, use after a sync)
template <int D1, int Q, int E>
__device__ __forceinline__ void face3_gather(double *sm, const double *__restrict__ ug,
                                             const FaceNbr &fn, int64_t e0, int ne)
{
   using S = Smem3<D1, Q, E>;
   constexpr int ND = S::ND, NF = S::NF, NFD = S::NFD, T = S::T;
   double *FD = sm + S::OFF_F;
   for (int id = threadIdx.x; id < E * NF * NFD; id += T)
   {
      const int e = id / (NF * NFD), idx = id - e * (NF * NFD);
      const int f = idx / NFD, j = idx - f * NFD;
      double d = 0.0;
      if (e < ne)
      {
         const int64_t ge = e0 + e;
         const double own = ug[ge * ND + face_dof<3, D1>(f, j)];
         const int64_t nb = fn.nbr_elem[ge * NF + f];
         double un = 0.0;
         if (nb >= 0)
         {
            const int loc = fn.pat[(int)fn.nbr_pat[ge * NF + f] * NFD + j];
            un = (nb < fn.ne_owned) ? ug[nb * ND + loc] : fn.ughost[(nb - fn.ne_owned) * NFD + j];
         }
         d = own - un;
      }
      FD[id] = d;
   }
}

// FD (differences) -> FD (face contributions to the rhs); Dface layout [e][qb][f][qa]
template <int D1, int Q, int E>
__device__ __forceinline__ void face3_apply(double *sm, const Pre3<D1, Q, E> &pre,
                                            const Tab<D1, Q> &tab)
{
   using S = Smem3<D1, Q, E>;
   constexpr int NF = S::NF, T = S::T;
   constexpr int NT2 = E * NF * Q;                 // fused-stage tasks (e, f, qa)
   constexpr int K2 = (NT2 + T - 1) / T;
   double *FD = sm + S::OFF_F, *F1 = sm + S::OFF_C, *F1b = sm + S::OFF_B;
   auto mB = [&](int o, int i) { return tab.B[o][i]; };
   auto mBt = [&](int o, int i) { return tab.B[i][o]; };
   // F1: tasks (e, f, jb): contract ja -> qa
   for (int id = threadIdx.x; id < E * NF * D1; id += T)
   {
      double x[D1], y[Q];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = FD[id * D1 + i]; }
      RMH_LINE(Q, D1, mB, x, y)
#pragma unroll
      for (int q = 0; q < Q; q++) { F1[id * Q + q] = y[q]; }
   }
   __syncthreads();
   // F2: tasks (e, f, qa): contract jb -> qb, scale by the face data, contract qb -> ib
#pragma unroll
   for (int k = 0; k < K2; k++)
   {
      const int id = threadIdx.x + k * T;
      if (id < NT2)
      {
         const int ef = id / Q, qa = id - ef * Q;
         double x[D1], y[Q], z[D1];
#pragma unroll
         for (int i = 0; i < D1; i++) { x[i] = F1[(ef * D1 + i) * Q + qa]; }
         RMH_LINE(Q, D1, mB, x, y)
#pragma unroll
         for (int q = 0; q < Q; q++) { y[q] *= pre.df[k][q]; }
         RMH_LINE(D1, Q, mBt, y, z)
#pragma unroll
         for (int i = 0; i < D1; i++) { F1b[(ef * D1 + i) * Q + qa] = z[i]; }
      }
   }
   __syncthreads();
   // F3: tasks (e, f, ib): contract qa -> ia
   for (int id = threadIdx.x; id < E * NF * D1; id += T)
   {
      double x[Q], y[D1];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = F1b[id * Q + q]; }
      RMH_LINE(D1, Q, mBt, x, y)
#pragma unroll
      for (int i = 0; i < D1; i++) { FD[id * D1 + i] = y[i]; }
   }
   __syncthreads();
}

// ---------------------------------------------------------------- volume term (+ face combine)
// R = B^T [D . grad U] + face contributions (FD).  Dvol [e][3][NQ].
template <int D1, int Q, int E, bool WITH_FACES>
__device__ __forceinline__ void vol3_apply(const double *U, double *R, double *sm,
                                           const Pre3<D1, Q, E> &pre, const Tab<D1, Q> &tab)
{
   using S = Smem3<D1, Q, E>;
   constexpr int QQ = S::QQ, NFD = S::NFD, NF = S::NF, T = S::T;
   constexpr int NL = E * D1 * D1;                 // x-lines
   constexpr int NY = E * D1 * Q;                  // y-lines
   double *BU = sm + S::OFF_C, *GU = BU + NL * Q;
   double *GB = sm + S::OFF_B;                     // [e][3][z][qy][qx]: GB, BG, BB per element
   const double *FC = sm + S::OFF_F;
   auto mB = [&](int o, int i) { return tab.B[o][i]; };
   auto mG = [&](int o, int i) { return tab.G[o][i]; };
   auto mBt = [&](int o, int i) { return tab.B[i][o]; };
   // V1 fwd-x: tasks (arr, e, z, y)
   for (int id = threadIdx.x; id < 2 * NL; id += T)
   {
      const int arr = id / NL, l = id - arr * NL;
      double x[D1], y[Q];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = U[l * D1 + i]; }
      if (arr == 0) { RMH_LINE(Q, D1, mB, x, y) }
      else { RMH_LINE(Q, D1, mG, x, y) }
      double *o = (arr == 0 ? BU : GU) + l * Q;
#pragma unroll
      for (int q = 0; q < Q; q++) { o[q] = y[q]; }
   }
   const int zid = threadIdx.x;                    // z-stage task (e, qy, qx); T >= E*QQ
   const int ze = zid / QQ, zr = zid - ze * QQ;
   const bool zon = zid < E * QQ;
   __syncthreads();
   // V2 fwd-y: tasks (arr, e, z, qx); arr 0: GB = By Gx u, 1: BG = Gy Bx u, 2: BB = By Bx u
   for (int id = threadIdx.x; id < 3 * NY; id += T)
   {
      const int arr = id / NY, l = id - arr * NY;
      const int ez = l / Q, qx = l - ez * Q;
      const double *in = (arr == 0 ? GU : BU) + ez * D1 * Q + qx;
      double x[D1], y[Q];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = in[i * Q]; }
      if (arr == 1) { RMH_LINE(Q, D1, mG, x, y) }
      else { RMH_LINE(Q, D1, mB, x, y) }
      const int e = ez / D1, z = ez - e * D1;
      double *o = GB + ((e * 3 + arr) * D1 + z) * QQ + qx;
#pragma unroll
      for (int q = 0; q < Q; q++) { o[q * Q] = y[q]; }
   }
   __syncthreads();
   // V3 z-fused: task (e, qy, qx)
   if (zon)
   {
      double *col = GB + ze * 3 * D1 * QQ + zr;
      double gb[D1], bg[D1], bb[D1], tz[D1];
#pragma unroll
      for (int i = 0; i < D1; i++)
      {
         gb[i] = col[i * QQ];
         bg[i] = col[(D1 + i) * QQ];
         bb[i] = col[(2 * D1 + i) * QQ];
         tz[i] = 0.0;
      }
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
         const double s = pre.d0[q] * g0 + pre.d1[q] * g1 + pre.d2[q] * g2;
#pragma unroll
         for (int i = 0; i < D1; i++) { tz[i] = fma(tab.B[q][i], s, tz[i]); }
      }
      // in place: this task is the only one touching column zr of its element
#pragma unroll
      for (int i = 0; i < D1; i++) { col[i * QQ] = tz[i]; }
   }
   __syncthreads();
   // V4 bwd-y: tasks (e, iz, qx): [e][iz][qy][qx] -> S2 [e][iz][iy][qx]
   double *S2 = BU;
   for (int id = threadIdx.x; id < NY; id += T)
   {
      const int eiz = id / Q, qx = id - eiz * Q;
      const int e = eiz / D1, iz = eiz - e * D1;
      const double *in = GB + (e * 3 * D1 + iz) * QQ + qx;
      double x[Q], y[D1];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = in[q * Q]; }
      RMH_LINE(D1, Q, mBt, x, y)
#pragma unroll
      for (int i = 0; i < D1; i++) { S2[(eiz * D1 + i) * Q + qx] = y[i]; }
   }
   __syncthreads();
   // V5 bwd-x + owner-computes face combine: tasks (e, iz, iy)
   for (int id = threadIdx.x; id < NL; id += T)
   {
      double x[Q], rr[D1];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = S2[id * Q + q]; }
      RMH_LINE(D1, Q, mBt, x, rr)
      if (WITH_FACES)
      {
         const int e = id / (D1 * D1), r = id - e * D1 * D1, b = r / D1, a = r - b * D1;
         const double *fc = FC + e * NF * NFD;
         // faces: 0 z=0 (x,y)  1 y=0 (x,z)  2 x=p (y,z)  3 y=p (x,z)  4 x=0 (y,z)  5 z=p (x,y)
         rr[0] += fc[4 * NFD + b * D1 + a];
         rr[D1 - 1] += fc[2 * NFD + b * D1 + a];
         if (a == 0)
         {
#pragma unroll
            for (int i = 0; i < D1; i++) { rr[i] += fc[1 * NFD + b * D1 + i]; }
         }
         if (a == D1 - 1)
         {
#pragma unroll
            for (int i = 0; i < D1; i++) { rr[i] += fc[3 * NFD + b * D1 + i]; }
         }
         if (b == 0)
         {
#pragma unroll
            for (int i = 0; i < D1; i++) { rr[i] += fc[0 * NFD + a * D1 + i]; }
         }
         if (b == D1 - 1)
         {
#pragma unroll
            for (int i = 0; i < D1; i++) { rr[i] += fc[5 * NFD + a * D1 + i]; }
         }
      }
#pragma unroll
      for (int i = 0; i < D1; i++) { R[id * D1 + i] = rr[i]; }
   }
   __syncthreads();
}

// ---------------------------------------------------------------- mass apply  Z = M P
template <int D1, int Q, int E>
__device__ __forceinline__ void mass3_apply(const double *P, double *Z, double *sm,
                                            const double *__restrict__ detJw, int ne,
                                            const Tab<D1, Q> &tab)
{
   using S = Smem3<D1, Q, E>;
   constexpr int NQ = S::NQ, QQ = S::QQ, T = S::T;
   constexpr int NL = E * D1 * D1, NY = E * D1 * Q;
   double *C1 = sm + S::OFF_C;                     // [e][z][y][qx]
   double *C2 = sm + S::OFF_B;                     // [e][z][qy][qx]
   auto mB = [&](int o, int i) { return tab.B[o][i]; };
   auto mBt = [&](int o, int i) { return tab.B[i][o]; };
   const int zid = threadIdx.x, ze = zid / QQ, zr = zid - ze * QQ;
   const bool zon = zid < E * QQ;
   double dj[Q];
   {
      const bool live = zon && (ze < ne);
#pragma unroll
      for (int q = 0; q < Q; q++) { dj[q] = live ? detJw[(size_t)ze * NQ + q * QQ + zr] : 0.0; }
   }
   for (int id = threadIdx.x; id < NL; id += T)
   {
      double x[D1], y[Q];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = P[id * D1 + i]; }
      RMH_LINE(Q, D1, mB, x, y)
#pragma unroll
      for (int q = 0; q < Q; q++) { C1[id * Q + q] = y[q]; }
   }
   __syncthreads();
   for (int id = threadIdx.x; id < NY; id += T)
   {
      const int ez = id / Q, qx = id - ez * Q;
      double x[D1], y[Q];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = C1[(ez * D1 + i) * Q + qx]; }
      RMH_LINE(Q, D1, mB, x, y)
#pragma unroll
      for (int q = 0; q < Q; q++) { C2[ez * QQ + q * Q + qx] = y[q]; }
   }
   __syncthreads();
   if (zon)
   {
      double *col = C2 + ze * D1 * QQ + zr;
      double x[D1], y[Q], tz[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = col[i * QQ]; }
      RMH_LINE(Q, D1, mB, x, y)
#pragma unroll
      for (int q = 0; q < Q; q++) { y[q] *= dj[q]; }
      RMH_LINE(D1, Q, mBt, y, tz)
#pragma unroll
      for (int i = 0; i < D1; i++) { col[i * QQ] = tz[i]; }
   }
   __syncthreads();
   for (int id = threadIdx.x; id < NY; id += T)
   {
      const int eiz = id / Q, qx = id - eiz * Q;
      double x[Q], y[D1];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = C2[eiz * QQ + q * Q + qx]; }
      RMH_LINE(D1, Q, mBt, x, y)
#pragma unroll
      for (int i = 0; i < D1; i++) { C1[(eiz * D1 + i) * Q + qx] = y[i]; }
   }
   __syncthreads();
   for (int id = threadIdx.x; id < NL; id += T)
   {
      double x[Q], y[D1];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = C1[id * Q + q]; }
      RMH_LINE(D1, Q, mBt, x, y)
#pragma unroll
      for (int i = 0; i < D1; i++) { Z[id * D1 + i] = y[i]; }
   }
   __syncthreads();
}

// Z = (Minv x Minv x Minv) R, optionally scaled per element by einv[e] (affine mass inverse)
template <int D1, int Q, int E>
__device__ __forceinline__ void kron3_apply(const double *R, double *Z, const Tab<D1, Q> &tab,
                                            const double *escale = nullptr)
{
   using S = Smem3<D1, Q, E>;
   constexpr int ND = S::ND, T = S::T, NL = E * D1 * D1;
   auto mM = [&](int o, int i) { return tab.Minv[o][i]; };
   for (int id = threadIdx.x; id < NL; id += T)     // x axis
   {
      double x[D1], y[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = R[id * D1 + i]; }
      RMH_LINE(D1, D1, mM, x, y)
#pragma unroll
      for (int i = 0; i < D1; i++) { Z[id * D1 + i] = y[i]; }
   }
   __syncthreads();
   for (int id = threadIdx.x; id < NL; id += T)     // y axis, tasks (e, z, x), in place
   {
      const int ez = id / D1, ix = id - ez * D1;
      double *p = Z + ez * D1 * D1 + ix;
      double x[D1], y[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = p[i * D1]; }
      RMH_LINE(D1, D1, mM, x, y)
#pragma unroll
      for (int i = 0; i < D1; i++) { p[i * D1] = y[i]; }
   }
   __syncthreads();
   for (int id = threadIdx.x; id < NL; id += T)     // z axis, tasks (e, y, x), in place
   {
      const int e = id / (D1 * D1), r = id - e * D1 * D1;
      double *p = Z + e * ND + r;
      double x[D1], y[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = p[i * D1 * D1]; }
      RMH_LINE(D1, D1, mM, x, y)
      const double sc = escale ? escale[e] : 1.0;
#pragma unroll
      for (int i = 0; i < D1; i++) { p[i * D1 * D1] = y[i] * sc; }
   }
   __syncthreads();
}

// Element mass solve X = M^-1 R.
//  * affine batch (every element of the block has constant Jacobian): M = vol_e * M_ref exactly,
//    so X = (Minv x Minv x Minv) R / vol_e -- one Kronecker application, no quadrature data read.
//  * otherwise: CG preconditioned with that Kronecker inverse, iterated to round-off
//    (exact-inverse semantics of remhos_ho.cpp:100-116 at the cost profile of the PA
//    DGMassInverse of :79-80,126).   R is destroyed.
template <int D1, int Q, int E>
__device__ __forceinline__ void mass3_solve(double *Rv, double *X, double *sm,
                                            const double *__restrict__ detJw,
                                            const double *__restrict__ einv, int ne, double tol2,
                                            int maxit, const Tab<D1, Q> &tab)
{
   using S = Smem3<D1, Q, E>;
   constexpr int ND = S::ND;
   double *P = sm + S::OFF_P, *Z = sm + S::OFF_Z;
   const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const bool mine = (w < E);          // this warp owns element w in the reductions
   // einv[e] > 0 marks an affine element and holds 1/vol_e
   double *esc = P;                    // E doubles of scratch (P is unused on the affine path)
   bool aff = true;
   if (threadIdx.x < E)
   {
      const double v = (threadIdx.x < ne) ? einv[threadIdx.x] : 1.0;
      esc[threadIdx.x] = v;
      aff = v > 0.0;
   }
   if (__syncthreads_and(aff))
   {
      kron3_apply<D1, Q, E>(Rv, X, tab, esc);
      return;
   }
   __syncthreads();
   double scale = 0.0;
   if (mine) { for (int j = lane; j < ND; j += 32) { scale = fmax(scale, fabs(Rv[w * ND + j])); } }
   scale = warp_max(scale);
   const bool nonzero = mine && (w < ne) && (scale > 1e-290) && (scale < 1e290);
   const double iscale = nonzero ? 1.0 / scale : 0.0;
   if (mine) { for (int j = lane; j < ND; j += 32) { Rv[w * ND + j] *= iscale; } }
   __syncthreads();
   kron3_apply<D1, Q, E>(Rv, Z, tab);
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
      mass3_apply<D1, Q, E>(P, Z, sm, detJw, ne, tab);
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
      kron3_apply<D1, Q, E>(Rv, Z, tab);
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

// ---------------------------------------------------------------- fused affine HO path
// X = M^-1 K_HO u for a batch whose elements all have constant det J: the exact mass inverse is
// einv[e] * (Minv x Minv x Minv), and because contractions along different axes commute it is
// folded into the three back-contractions (C = Minv B^T replaces B^T) at zero cost; the face
// contributions get the two tangential Minv factors the same way and the normal one when they
// are combined into the lines.  Face and volume stages share barrier intervals:
//   A: fwd-x | face fwd-a     B: fwd-y | face fused     C: z-fused | face back-a
//   D: bwd-y                  E: bwd-x + face combine -> X
template <int D1, int Q, int E>
__device__ __forceinline__ void ho3_affine(const double *U, double *X, double *sm,
                                           const Pre3<D1, Q, E> &pre,
                                           const double *__restrict__ einv, int ne,
                                           const Tab<D1, Q> &tab)
{
   using S = Smem3<D1, Q, E>;
   constexpr int QQ = S::QQ, NFD = S::NFD, NF = S::NF, T = S::T;
   constexpr int NL = E * D1 * D1, NY = E * D1 * Q, NT1 = E * NF * D1, NT2 = E * NF * Q;
   constexpr int K2 = (NT2 + T - 1) / T;
   double *BU = sm + S::OFF_C, *GU = BU + NL * Q;
   double *GB = sm + S::OFF_B;
   double *FD = sm + S::OFF_F, *F1 = sm + S::OFF_G;
   auto mB = [&](int o, int i) { return tab.B[o][i]; };
   auto mG = [&](int o, int i) { return tab.G[o][i]; };
   auto mC = [&](int o, int i) { return tab.C[o][i]; };
   // ---- A: fwd-x tasks (arr, e, z, y)  |  face tasks (e, f, jb): contract ja -> qa
   for (int id = threadIdx.x; id < 2 * NL + NT1; id += T)
   {
      if (id < 2 * NL)
      {
         const int arr = id / NL, l = id - arr * NL;
         double x[D1], y[Q];
#pragma unroll
         for (int i = 0; i < D1; i++) { x[i] = U[l * D1 + i]; }
         if (arr == 0) { RMH_LINE(Q, D1, mB, x, y) }
         else { RMH_LINE(Q, D1, mG, x, y) }
         double *o = (arr == 0 ? BU : GU) + l * Q;
#pragma unroll
         for (int q = 0; q < Q; q++) { o[q] = y[q]; }
      }
      else
      {
         const int l = id - 2 * NL;
         double x[D1], y[Q];
#pragma unroll
         for (int i = 0; i < D1; i++) { x[i] = FD[l * D1 + i]; }
         RMH_LINE(Q, D1, mB, x, y)
#pragma unroll
         for (int q = 0; q < Q; q++) { F1[l * Q + q] = y[q]; }
      }
   }
   __syncthreads();
   // ---- B: fwd-y tasks (arr, e, z, qx)  |  face fused tasks (e, f, qa)
   for (int id = threadIdx.x; id < 3 * NY; id += T)
   {
      const int arr = id / NY, l = id - arr * NY;
      const int ez = l / Q, qx = l - ez * Q;
      const double *in = (arr == 0 ? GU : BU) + ez * D1 * Q + qx;
      double x[D1], y[Q];
#pragma unroll
      for (int i = 0; i < D1; i++) { x[i] = in[i * Q]; }
      if (arr == 1) { RMH_LINE(Q, D1, mG, x, y) }
      else { RMH_LINE(Q, D1, mB, x, y) }
      const int e = ez / D1, z = ez - e * D1;
      double *o = GB + ((e * 3 + arr) * D1 + z) * QQ + qx;
#pragma unroll
      for (int q = 0; q < Q; q++) { o[q * Q] = y[q]; }
   }
#pragma unroll
   for (int k = 0; k < K2; k++)
   {
      const int id = threadIdx.x + k * T;
      if (id < NT2)
      {
         const int ef = id / Q, qa = id - ef * Q;
         double x[D1], y[Q], z[D1];
#pragma unroll
         for (int i = 0; i < D1; i++) { x[i] = F1[(ef * D1 + i) * Q + qa]; }
         RMH_LINE(Q, D1, mB, x, y)
#pragma unroll
         for (int q = 0; q < Q; q++) { y[q] *= pre.df[k][q]; }
         RMH_LINE(D1, Q, mC, y, z)
         // in place: this task owns column (ef, :, qa)
#pragma unroll
         for (int i = 0; i < D1; i++) { F1[(ef * D1 + i) * Q + qa] = z[i]; }
      }
   }
   __syncthreads();
   // ---- C: z-fused task (e, qy, qx)  |  face tasks (e, f, ib): contract qa -> ia
   {
      const int zid = threadIdx.x, ze = zid / QQ, zr = zid - ze * QQ;
      if (zid < E * QQ)
      {
         double *col = GB + ze * 3 * D1 * QQ + zr;
         double gb[D1], bg[D1], bb[D1], tz[D1];
#pragma unroll
         for (int i = 0; i < D1; i++)
         {
            gb[i] = col[i * QQ];
            bg[i] = col[(D1 + i) * QQ];
            bb[i] = col[(2 * D1 + i) * QQ];
            tz[i] = 0.0;
         }
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
            const double s = pre.d0[q] * g0 + pre.d1[q] * g1 + pre.d2[q] * g2;
#pragma unroll
            for (int i = 0; i < D1; i++) { tz[i] = fma(tab.C[i][q], s, tz[i]); }
         }
#pragma unroll
         for (int i = 0; i < D1; i++) { col[i * QQ] = tz[i]; }
      }
   }
   for (int id = threadIdx.x; id < NT1; id += T)
   {
      double x[Q], y[D1];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = F1[id * Q + q]; }
      RMH_LINE(D1, Q, mC, x, y)
#pragma unroll
      for (int i = 0; i < D1; i++) { FD[id * D1 + i] = y[i]; }
   }
   __syncthreads();
   // ---- D: bwd-y tasks (e, iz, qx)
   double *S2 = BU;
   for (int id = threadIdx.x; id < NY; id += T)
   {
      const int eiz = id / Q, qx = id - eiz * Q;
      const int e = eiz / D1, iz = eiz - e * D1;
      const double *in = GB + (e * 3 * D1 + iz) * QQ + qx;
      double x[Q], y[D1];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = in[q * Q]; }
      RMH_LINE(D1, Q, mC, x, y)
#pragma unroll
      for (int i = 0; i < D1; i++) { S2[(eiz * D1 + i) * Q + qx] = y[i]; }
   }
   __syncthreads();
   // ---- E: bwd-x tasks (e, iz, iy) + face contributions with the normal-direction Minv
   for (int id = threadIdx.x; id < NL; id += T)
   {
      double x[Q], rr[D1];
#pragma unroll
      for (int q = 0; q < Q; q++) { x[q] = S2[id * Q + q]; }
      RMH_LINE(D1, Q, mC, x, rr)
      const int e = id / (D1 * D1), r = id - e * D1 * D1, b = r / D1, a = r - b * D1;
      const double *fc = FD + e * NF * NFD;
      // faces: 0 z=0 (x,y)  1 y=0 (x,z)  2 x=p (y,z)  3 y=p (x,z)  4 x=0 (y,z)  5 z=p (x,y)
      const double fx0 = fc[4 * NFD + b * D1 + a], fx1 = fc[2 * NFD + b * D1 + a];
      const double my0 = tab.Minv[a][0], my1 = tab.Minv[a][D1 - 1];
      const double mz0 = tab.Minv[b][0], mz1 = tab.Minv[b][D1 - 1];
      const double sc = (e < ne) ? einv[e] : 0.0;
#pragma unroll
      for (int i = 0; i < D1; i++)
      {
         double v = rr[i];
         v = fma(tab.Minv[i][0], fx0, v);
         v = fma(tab.Minv[i][D1 - 1], fx1, v);
         v = fma(my0, fc[1 * NFD + b * D1 + i], v);
         v = fma(my1, fc[3 * NFD + b * D1 + i], v);
         v = fma(mz0, fc[0 * NFD + a * D1 + i], v);
         v = fma(mz1, fc[5 * NFD + a * D1 + i], v);
         X[id * D1 + i] = v * sc;
      }
   }
   __syncthreads();
}

} // namespace rmh

#endif
