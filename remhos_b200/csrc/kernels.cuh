// Device code of the remhos_b200 RK-stage path (sm_100a, FP64).
//
// Design (DESIGN.md "Kernels"): one thread block owns a batch of E elements and T = 32*E
// threads.  Every tensor contraction is done "line-wise": a thread loads one line of the
// contracted axis into registers, applies the small 1-D matrix (held in the kernel-parameter
// constant bank, compile-time indexed so the coefficients are constant-bank FMA operands) and
// writes the output line -> NIN loads + NOUT stores per NIN*NOUT DFMAs.  All intermediates
// live in shared memory; operator data (stored quadrature data) is streamed from HBM with
// coalesced loads.  Element-wise phases (LO / FCT / RK) map one warp per element and reduce
// with shuffles in a fixed order (deterministic).
#ifndef RMH_KERNELS_CUH
#define RMH_KERNELS_CUH

#include <cstdint>
#include <cuda_runtime.h>

namespace rmh
{

__host__ __device__ constexpr int ipow(int b, int e) { return e <= 0 ? 1 : b * ipow(b, e - 1); }

// 1-D tables, passed by value as a kernel parameter.
template <int D1, int Q>
struct Tab
{
   double B[Q][D1];      // Bernstein values at the Gauss-Legendre points
   double G[Q][D1];      // derivatives
   double Minv[D1][D1];  // inverse of the 1-D Bernstein mass matrix (Kronecker preconditioner)
   double C[D1][Q];      // Minv * B^T: back-contraction with the mass inverse folded in
   double xq[Q], wq[Q];  // Gauss-Legendre points / weights on [0,1]
};

// local face -> fixed axis / side (quad: S E N W; hex: bottom south east north west top)
__device__ __forceinline__ void face_axis_side(int dim, int f, int &axis, int &side)
{
   if (dim == 2)
   {
      axis = (f == 0 || f == 2) ? 1 : 0;
      side = (f == 1 || f == 2) ? 1 : 0;
   }
   else
   {
      axis = (f == 0 || f == 5) ? 2 : ((f == 1 || f == 3) ? 1 : 0);
      side = (f == 2 || f == 3 || f == 5) ? 1 : 0;
   }
}
__device__ __forceinline__ int face_of(int dim, int axis, int side)
{
   if (dim == 2) { return axis == 1 ? (side ? 2 : 0) : (side ? 1 : 3); }
   return axis == 2 ? (side ? 5 : 0) : (axis == 1 ? (side ? 3 : 1) : (side ? 2 : 4));
}

// local DOF index of face DOF j (natural face parametrisation: remaining axes ascending)
template <int DIM, int D1>
__device__ __forceinline__ int face_dof(int f, int j)
{
   int axis, side;
   face_axis_side(DIM, f, axis, side);
   int l[3] = {0, 0, 0};
   int m = j;
#pragma unroll
   for (int a = 0; a < DIM; a++)
   {
      if (a == axis) { l[a] = side * (D1 - 1); }
      else { l[a] = m % D1; m /= D1; }
   }
   return DIM == 2 ? l[0] + D1 * l[1] : l[0] + D1 * (l[1] + D1 * l[2]);
}

// Contract one axis: in [A][NIN][S] -> out [A][NOUT][S], out(a,o,s) = sum_i mat(o,i) in(a,i,s).
// In-place is safe when NIN == NOUT (each thread owns whole lines).
template <int NIN, int NOUT, int S, int T, typename MatF>
__device__ __forceinline__ void contract(const double *in, double *out, int A, MatF mat)
{
   const int nl = A * S;
   for (int l = threadIdx.x; l < nl; l += T)
   {
      const int a = l / S, s = l - a * S;
      const double *pi = in + (size_t)a * NIN * S + s;
      double x[NIN];
#pragma unroll
      for (int i = 0; i < NIN; i++) { x[i] = pi[i * S]; }
      double *po = out + (size_t)a * NOUT * S + s;
#pragma unroll
      for (int o = 0; o < NOUT; o++)
      {
         double acc = 0.0;
#pragma unroll
         for (int i = 0; i < NIN; i++) { acc = fma(mat(o, i), x[i], acc); }
         po[o * S] = acc;
      }
   }
}

// Same with two matrices sharing the input line (forward B and G together).
template <int NIN, int NOUT, int S, int T, typename MatF1, typename MatF2>
__device__ __forceinline__ void contract2(const double *in, double *out1, double *out2, int A,
                                          MatF1 m1, MatF2 m2)
{
   const int nl = A * S;
   for (int l = threadIdx.x; l < nl; l += T)
   {
      const int a = l / S, s = l - a * S;
      const double *pi = in + (size_t)a * NIN * S + s;
      double x[NIN];
#pragma unroll
      for (int i = 0; i < NIN; i++) { x[i] = pi[i * S]; }
      double *p1 = out1 + (size_t)a * NOUT * S + s;
      double *p2 = out2 + (size_t)a * NOUT * S + s;
#pragma unroll
      for (int o = 0; o < NOUT; o++)
      {
         double a1 = 0.0, a2 = 0.0;
#pragma unroll
         for (int i = 0; i < NIN; i++)
         {
            a1 = fma(m1(o, i), x[i], a1);
            a2 = fma(m2(o, i), x[i], a2);
         }
         p1[o * S] = a1;
         p2[o * S] = a2;
      }
   }
}

// Shared-memory plan of one block (doubles).
template <int DIM, int D1, int Q, int E>
struct Smem
{
   static constexpr int ND = ipow(D1, DIM), NQ = ipow(Q, DIM);
   static constexpr int NF = 2 * DIM, NFD = ipow(D1, DIM - 1), NQF = ipow(Q, DIM - 1);
   // bufA: DIM gradient arrays at the quadrature points; bufB / bufC: partially contracted
   static constexpr int SZ_A = DIM * E * NQ;
   static constexpr int SZ_B = DIM * E * ipow(D1, 1) * ipow(Q, DIM - 1);
   static constexpr int SZ_C0 = 2 * E * ipow(D1, DIM - 1) * Q;
   static constexpr int SZ_C1 = E * NF * ipow(D1, DIM - 2) * Q;      // face work F1
   static constexpr int SZ_C = SZ_C0 > SZ_C1 ? SZ_C0 : SZ_C1;
   static constexpr int SZ_V = E * ND;
   // face work: FD/FC [E][NF][NFD], F1 [E][NF][D1^(DIM-2)][Q], F2 [E][NF][NQF]
   static constexpr int SZ_F = E * NF * NFD;
   static constexpr int OFF_A = 0;
   static constexpr int OFF_B = OFF_A + SZ_A;
   static constexpr int OFF_C = OFF_B + SZ_B;
   static constexpr int OFF_U = OFF_C + SZ_C;   // U
   static constexpr int OFF_R = OFF_U + SZ_V;   // R (rhs / residual)
   static constexpr int OFF_X = OFF_R + SZ_V;   // X (solution)
   static constexpr int OFF_P = OFF_X + SZ_V;   // P
   static constexpr int OFF_Z = OFF_P + SZ_V;   // Z / AP
   static constexpr int OFF_F = OFF_Z + SZ_V;   // FD / FC
   static constexpr int TOTAL = OFF_F + SZ_F;
   static constexpr size_t BYTES = (size_t)TOTAL * sizeof(double);
};

// ----------------------------------------------------------------------------------------
// Volume term: R = B^T [ D . grad_ref(U) ]   (PA ConvectionIntegrator apply; D = alpha w adj(J) v)
// U, R: smem [E][ND]; Dvol: global, already offset to the block's first element,
// layout [e][DIM][NQ]; ne = valid elements in the block.
template <int DIM, int D1, int Q, int E>
__device__ __forceinline__ void vol_apply(const double *U, double *R, double *sm,
                                          const double *__restrict__ Dvol, int ne,
                                          const Tab<D1, Q> &tab)
{
   using S = Smem<DIM, D1, Q, E>;
   constexpr int T = 32 * E, NQ = S::NQ;
   double *bufA = sm + S::OFF_A, *bufB = sm + S::OFF_B, *bufC = sm + S::OFF_C;
   auto fB = [&](int o, int i) { return tab.B[o][i]; };
   auto fG = [&](int o, int i) { return tab.G[o][i]; };
   auto fBt = [&](int o, int i) { return tab.B[i][o]; };
   if (DIM == 3)
   {
      constexpr int n1 = E * D1 * D1 * Q;           // size of an x-contracted array
      constexpr int n2 = E * D1 * Q * Q;
      double *BU = bufC, *GU = bufC + n1;
      contract2<D1, Q, 1, T>(U, BU, GU, E * D1 * D1, fB, fG);
      __syncthreads();
      double *GB = bufB, *BG = bufB + n2, *BB = bufB + 2 * n2;
      contract2<D1, Q, Q, T>(BU, BB, BG, E * D1, fB, fG);
      contract<D1, Q, Q, T>(GU, GB, E * D1, fB);
      __syncthreads();
      double *g0 = bufA, *g1 = bufA + E * NQ, *g2 = bufA + 2 * E * NQ;
      contract<D1, Q, Q * Q, T>(GB, g0, E, fB);
      contract<D1, Q, Q * Q, T>(BG, g1, E, fB);
      contract<D1, Q, Q * Q, T>(BB, g2, E, fG);
      __syncthreads();
      for (int t = threadIdx.x; t < E * NQ; t += T)
      {
         const int e = t / NQ, q = t - e * NQ;
         double s = 0.0;
         if (e < ne)
         {
            const double *d = Dvol + (size_t)e * 3 * NQ + q;
            s = d[0] * g0[t] + d[NQ] * g1[t] + d[2 * NQ] * g2[t];
         }
         g0[t] = s;
      }
      __syncthreads();
      contract<Q, D1, Q * Q, T>(g0, bufB, E, fBt);
      __syncthreads();
      contract<Q, D1, Q, T>(bufB, bufC, E * D1, fBt);
      __syncthreads();
      contract<Q, D1, 1, T>(bufC, R, E * D1 * D1, fBt);
      __syncthreads();
   }
   else
   {
      constexpr int n1 = E * D1 * Q;
      double *BU = bufC, *GU = bufC + n1;
      contract2<D1, Q, 1, T>(U, BU, GU, E * D1, fB, fG);
      __syncthreads();
      double *g0 = bufA, *g1 = bufA + E * NQ;
      contract<D1, Q, Q, T>(GU, g0, E, fB);
      contract<D1, Q, Q, T>(BU, g1, E, fG);
      __syncthreads();
      for (int t = threadIdx.x; t < E * NQ; t += T)
      {
         const int e = t / NQ, q = t - e * NQ;
         double s = 0.0;
         if (e < ne)
         {
            const double *d = Dvol + (size_t)e * 2 * NQ + q;
            s = d[0] * g0[t] + d[NQ] * g1[t];
         }
         g0[t] = s;
      }
      __syncthreads();
      contract<Q, D1, Q, T>(g0, bufB, E, fBt);
      __syncthreads();
      contract<Q, D1, 1, T>(bufB, R, E * D1, fBt);
      __syncthreads();
   }
}

// Neighbour data needed by the face terms.
struct FaceNbr
{
   const int32_t *nbr_elem;   // [NE][NF]   (-1 boundary; >= ne_owned: ne_owned + ghost-face slot)
   const uint8_t *nbr_pat;    // [NE][NF]   pattern id
   const int16_t *pat;        // [npat][NFD] neighbour-local DOF of own face DOF j (natural order)
   const double *ughost;      // ghost face traces [n_slots][NFD], natural face order of the reader (may be NULL)
   int64_t ne_owned;
};

// Face terms: R += sum_f Bf^T [ Dface . Bf (u_own - u_nbr) ]   (transposed DGTraceIntegrator,
// upwinded; exterior state 0 on the domain boundary).  ug = global u (owned DOFs).
template <int DIM, int D1, int Q, int E>
__device__ __forceinline__ void face_apply(const double *U, double *R, double *sm,
                                           const double *__restrict__ ug,
                                           const double *__restrict__ Dface, const FaceNbr &fn,
                                           int64_t e0, int ne, const Tab<D1, Q> &tab)
{
   using S = Smem<DIM, D1, Q, E>;
   constexpr int T = 32 * E, ND = S::ND, NF = S::NF, NFD = S::NFD, NQF = S::NQF;
   double *FD = sm + S::OFF_F;
   double *F1 = sm + S::OFF_C;     // [E*NF*D1^(DIM-2)][Q]
   double *F2 = sm + S::OFF_B;     // [E*NF][NQF]
   auto fB = [&](int o, int i) { return tab.B[o][i]; };
   auto fBt = [&](int o, int i) { return tab.B[i][o]; };
   for (int t = threadIdx.x; t < E * NF * NFD; t += T)
   {
      const int e = t / (NF * NFD), r = t - e * (NF * NFD), f = r / NFD, j = r - f * NFD;
      double d = 0.0;
      if (e < ne)
      {
         const double own = U[e * ND + face_dof<DIM, D1>(f, j)];
         const int64_t ge = e0 + e;
         const int64_t nb = fn.nbr_elem[ge * NF + f];
         double un = 0.0;
         if (nb >= 0)
         {
            const int loc = fn.pat[(int)fn.nbr_pat[ge * NF + f] * NFD + j];
            un = (nb < fn.ne_owned) ? ug[nb * ND + loc]
                 : fn.ughost[(nb - fn.ne_owned) * NFD + j];
         }
         d = own - un;
      }
      FD[t] = d;
   }
   __syncthreads();
   if (DIM == 3)
   {
      contract<D1, Q, 1, T>(FD, F1, E * NF * D1, fB);
      __syncthreads();
      contract<D1, Q, Q, T>(F1, F2, E * NF, fB);
      __syncthreads();
   }
   else
   {
      contract<D1, Q, 1, T>(FD, F2, E * NF, fB);
      __syncthreads();
   }
   for (int t = threadIdx.x; t < E * NF * NQF; t += T)
   {
      const int e = t / (NF * NQF);
      F2[t] = (e < ne) ? F2[t] * Dface[t] : 0.0;
   }
   __syncthreads();
   if (DIM == 3)
   {
      contract<Q, D1, Q, T>(F2, F1, E * NF, fBt);
      __syncthreads();
      contract<Q, D1, 1, T>(F1, FD, E * NF * D1, fBt);
      __syncthreads();
   }
   else
   {
      contract<Q, D1, 1, T>(F2, FD, E * NF, fBt);
      __syncthreads();
   }
   // owner-computes combine (deterministic)
   for (int t = threadIdx.x; t < E * ND; t += T)
   {
      const int e = t / ND, i = t - e * ND;
      int l[3];
      int m = i;
#pragma unroll
      for (int a = 0; a < DIM; a++) { l[a] = m % D1; m /= D1; }
      double acc = R[t];
#pragma unroll
      for (int a = 0; a < DIM; a++)
      {
         int j = 0, mul = 1;
#pragma unroll
         for (int b = 0; b < DIM; b++)
         {
            if (b == a) { continue; }
            j += l[b] * mul;
            mul *= D1;
         }
         if (l[a] == 0) { acc += FD[(e * NF + face_of(DIM, a, 0)) * NFD + j]; }
         if (l[a] == D1 - 1) { acc += FD[(e * NF + face_of(DIM, a, 1)) * NFD + j]; }
      }
      R[t] = acc;
   }
   __syncthreads();
}

// AP = M P with M = B^T diag(detJw) B   (PA MassIntegrator apply)
template <int DIM, int D1, int Q, int E>
__device__ __forceinline__ void mass_apply(const double *P, double *AP, double *sm,
                                           const double *__restrict__ detJw, int ne,
                                           const Tab<D1, Q> &tab)
{
   using S = Smem<DIM, D1, Q, E>;
   constexpr int T = 32 * E, NQ = S::NQ;
   double *bufA = sm + S::OFF_A, *bufB = sm + S::OFF_B, *bufC = sm + S::OFF_C;
   auto fB = [&](int o, int i) { return tab.B[o][i]; };
   auto fBt = [&](int o, int i) { return tab.B[i][o]; };
   if (DIM == 3)
   {
      contract<D1, Q, 1, T>(P, bufC, E * D1 * D1, fB);
      __syncthreads();
      contract<D1, Q, Q, T>(bufC, bufB, E * D1, fB);
      __syncthreads();
      contract<D1, Q, Q * Q, T>(bufB, bufA, E, fB);
      __syncthreads();
   }
   else
   {
      contract<D1, Q, 1, T>(P, bufC, E * D1, fB);
      __syncthreads();
      contract<D1, Q, Q, T>(bufC, bufA, E, fB);
      __syncthreads();
   }
   for (int t = threadIdx.x; t < E * NQ; t += T)
   {
      const int e = t / NQ;
      bufA[t] = (e < ne) ? bufA[t] * detJw[t] : 0.0;
   }
   __syncthreads();
   if (DIM == 3)
   {
      contract<Q, D1, Q * Q, T>(bufA, bufB, E, fBt);
      __syncthreads();
      contract<Q, D1, Q, T>(bufB, bufC, E * D1, fBt);
      __syncthreads();
      contract<Q, D1, 1, T>(bufC, AP, E * D1 * D1, fBt);
      __syncthreads();
   }
   else
   {
      contract<Q, D1, Q, T>(bufA, bufB, E, fBt);
      __syncthreads();
      contract<Q, D1, 1, T>(bufB, AP, E * D1, fBt);
      __syncthreads();
   }
}

// Z = (Minv (x) Minv (x) Minv) R : exact inverse of the reference-element mass matrix
template <int DIM, int D1, int Q, int E>
__device__ __forceinline__ void kron_apply(const double *Rv, double *Z, const Tab<D1, Q> &tab)
{
   constexpr int T = 32 * E;
   auto fM = [&](int o, int i) { return tab.Minv[o][i]; };
   contract<D1, D1, 1, T>(Rv, Z, E * ipow(D1, DIM - 1), fM);
   __syncthreads();
   contract<D1, D1, D1, T>(Z, Z, E * ipow(D1, DIM - 2), fM);
   __syncthreads();
   if (DIM == 3)
   {
      contract<D1, D1, D1 * D1, T>(Z, Z, E, fM);
      __syncthreads();
   }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
   return v;
}
// two independent sums in flight (halves the shuffle-latency chain)
__device__ __forceinline__ void warp_sum2(double &a, double &b)
{
#pragma unroll
   for (int o = 16; o > 0; o >>= 1)
   {
      const double ta = __shfl_xor_sync(0xffffffffu, a, o);
      const double tb = __shfl_xor_sync(0xffffffffu, b, o);
      a += ta; b += tb;
   }
}
__device__ __forceinline__ void warp_minmax(double &mn, double &mx)
{
#pragma unroll
   for (int o = 16; o > 0; o >>= 1)
   {
      const double ta = __shfl_xor_sync(0xffffffffu, mn, o);
      const double tb = __shfl_xor_sync(0xffffffffu, mx, o);
      mn = fmin(mn, ta); mx = fmax(mx, tb);
   }
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) { v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o)); }
   return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) { v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); }
   return v;
}

// Element-local mass solve X = M^{-1} R by CG preconditioned with the exact inverse of the
// reference mass matrix (exact in one iteration on affine elements; converges at the rate of
// the detJ variation otherwise).  Iterates to round-off (rel. tolerance `tol` on the
// preconditioned residual norm), i.e. the FA exact-inverse semantics of remhos_ho.cpp:100-116
// with the cost profile of the PA DGMassInverse of :79-80,126.   R is destroyed.
template <int DIM, int D1, int Q, int E>
__device__ __forceinline__ void mass_solve(double *Rv, double *X, double *sm,
                                           const double *__restrict__ detJw, int ne,
                                           double tol2, int maxit, const Tab<D1, Q> &tab)
{
   using S = Smem<DIM, D1, Q, E>;
   constexpr int ND = S::ND;
   double *P = sm + S::OFF_P, *Z = sm + S::OFF_Z;
   const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
   // normalise the right-hand side per element: keeps the dot products away from
   // under/overflow (fields with exact-zero regions carry values down to 1e-160)
   double scale = 0.0;
   for (int j = lane; j < ND; j += 32) { scale = fmax(scale, fabs(Rv[w * ND + j])); }
   scale = warp_max(scale);
   const bool nonzero = (w < ne) && (scale > 1e-290) && (scale < 1e290);
   const double iscale = nonzero ? 1.0 / scale : 0.0;
   for (int j = lane; j < ND; j += 32) { Rv[w * ND + j] *= iscale; }
   __syncthreads();
   kron_apply<DIM, D1, Q, E>(Rv, Z, tab);
   double rz = 0.0;
   for (int j = lane; j < ND; j += 32)
   {
      const double z = Z[w * ND + j];
      P[w * ND + j] = z;
      X[w * ND + j] = 0.0;
      rz += Rv[w * ND + j] * z;
   }
   rz = warp_sum(rz);
   const double rz0 = rz;
   bool active = nonzero && (rz0 > 0.0);
   __syncthreads();
   for (int it = 0; it < maxit; it++)
   {
      if (!__syncthreads_or(active)) { break; }
      mass_apply<DIM, D1, Q, E>(P, Z, sm, detJw, ne, tab);   // Z = M P
      double pap = 0.0;
      for (int j = lane; j < ND; j += 32) { pap += P[w * ND + j] * Z[w * ND + j]; }
      pap = warp_sum(pap);
      const double alpha = active ? rz / pap : 0.0;
      for (int j = lane; j < ND; j += 32)
      {
         X[w * ND + j] += alpha * P[w * ND + j];
         Rv[w * ND + j] -= alpha * Z[w * ND + j];
      }
      __syncthreads();
      kron_apply<DIM, D1, Q, E>(Rv, Z, tab);
      double rzn = 0.0;
      for (int j = lane; j < ND; j += 32) { rzn += Rv[w * ND + j] * Z[w * ND + j]; }
      rzn = warp_sum(rzn);
      if (active && !(rzn > tol2 * rz0)) { active = false; }
      const double beta = active ? rzn / rz : 0.0;
      for (int j = lane; j < ND; j += 32)
      {
         P[w * ND + j] = active ? Z[w * ND + j] + beta * P[w * ND + j] : 0.0;
      }
      rz = rzn;
      __syncthreads();
   }
   for (int j = lane; j < ND; j += 32) { X[w * ND + j] *= scale; }
   __syncthreads();
}

} // namespace rmh

#endif
