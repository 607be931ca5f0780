// Warp-per-element FP64 tensor-core (DMMA m8n8k4) RK-stage kernel for 3D meshes whose elements
// all have constant det J (sm_100a).
//
// One warp owns one element at a time and walks a static round-robin list of elements
// (persistent grid).  Everything the element needs lives in a warp-private shared-memory
// region, so the phases are separated by __syncwarp() only -- no block barrier exists in the
// element loop, warps drift freely and hide each other's latencies.
//
//   prefetch (cp.async, two stages per warp): y block, RK base x0, neighbour traces gathered
//     through (neighbour element, orientation pattern), entity (min,max) pairs, 1/volume; the
//     gather indices are fetched one element further ahead.  The stored quadrature data of the
//     next element is pulled into L2 with prefetch.global.L2 and read by plain 16-byte loads in
//     fragment order (layout below) right where it is consumed.
//   contractions: 8-line DMMA tiles with compile-time tile indices (all shared-memory offsets
//     are per-lane constants); the 1-D matrices are register fragments; the z-stage
//     (forward-z, D.grad u, backward-z) and the fused face stage chain D fragments as A operands
//     and never leave registers (fragment algebra below).
//   tail: MassBasedAvg, bounds gather, ClipScale (two shuffle reductions), RK combination and the
//     element min/max of the output, on the same warp.
//
// Why tensor cores here: profiles/r01 -- the DFMA kernel is issue-bound (20 % of issued
// instructions are DFMA; sm_100a DFMA takes no constant operand, so each coefficient costs an
// LDCU, each 4x6 line 10 LDS/STS).  DMMA has the same measured peak (36.9 TFLOP/s) at 1/8 of the
// issue slots.
//
// Fragment algebra.  Every contraction is a set of 8-line tiles  D[8 x 8] += A[8 x 4] B[4 x 8]
// (g = lane/4, c = lane%4):
//     A fragment: A[g][c]          B fragment: B[k=c][n=g]          D fragment: D[g][2c], D[g][2c+1]
// A 1-D matrix M (Q x D1 forward, D1 x Q backward) is held once per thread as fragment registers
// (the same register serves as A operand "rows = outputs" and as B operand "cols = outputs").
// A D fragment whose columns are the next contracted index is fed straight back as the A operand
// of the following DMMA (columns 2c -> k-step 1, 2c+1 -> k-step 2, coefficient rows permuted to
// match).  The stored quadrature data is laid out in HBM in fragment order (ctx.cu, run_geom):
//     Dvol [e][col = qy*Q+qx][qz (RQ)][3]     -> 3 x 16-byte loads per thread and tile, coalesced
//     Dface[e][f][qa][qb (RQ)]                -> 1 x 16-byte load per thread and tile
#ifndef RMH_STAGE3W_CUH
#define RMH_STAGE3W_CUH

#include "stage3p.cuh"

namespace rmh
{

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
   asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(d0), "+d"(d1)
                : "d"(a), "d"(b));
}


template <int D1, int Q>
struct SmemW
{
   static constexpr int ND = D1 * D1 * D1, QQ = Q * Q;
   static constexpr int NF = 6, NFD = D1 * D1, N3 = 27;
   static constexpr int RQ = (Q + 1) & ~1;
   static constexpr int KF = (D1 + 3) / 4, KB = (Q + 3) / 4;
   static constexpr int NL = D1 * D1;              // x-lines (z, y)
   static constexpr int NY = D1 * Q;               // y-lines (z, qx)
   static constexpr int NC = QQ;                   // z-columns (qy, qx)
   static constexpr int NT1 = NF * D1;             // face lines (f, jb)
   static constexpr int NT2 = NF * Q;              // face lines (f, qa)
   // z-plane stride: >= QQ and = 4 (mod 8) so that the four z-planes of a tile hit distinct banks
   static constexpr int PZ = (D1 == 4 && Q == 6) ? QQ : ((QQ + 3) / 8) * 8 + 4;
   static constexpr int PA = D1 * PZ;
   // ---- bank-conflict-free layouts (tools/bank_sim_w.py models every access below; 64-bit
   // accesses are served per half-warp over 16 bank pairs).  NEWL: order 3 (D1 = 4, Q = 6).
   //   U   (z,y,x)      z-plane stride 20, x ^= z: face traces of all six faces hit 16 distinct banks
   //   BU  (row,qx)     rows 8..15 swap the two columns of a pair (stride-6 rows only span 8 banks)
   //   G3  (arr,z,qy,qx) [arr][qy][qx][z] with the qx parity flipped for qy in {2,3}: 16-byte stores
   //                    in fwd-y and the z-stage, conflict-free loads in the z-stage and bwd-y
   //   S2  (iz,iy,qx)   [qx (stride 20)][iz][iy], lives in the dead G3 array 1
   static constexpr bool NEWL = (D1 == 4 && Q == 6);
   static constexpr int SZU = NEWL ? D1 * D1 + 4 : D1 * D1;
   static constexpr int FS = D1 * RQ;
   __device__ static __forceinline__ int posU(int z, int y, int x)
   { return NEWL ? z * SZU + y * D1 + (x ^ z) : (z * D1 + y) * D1 + x; }
   __device__ static __forceinline__ int posUj(int j)
   { return NEWL ? posU(j / (D1 * D1), (j / D1) % D1, j % D1) : j; }
   __device__ static __forceinline__ int posBU(int row, int qx)
   { return NEWL ? row * RQ + (qx ^ ((row >> 3) & 1)) : row * RQ + qx; }
   __device__ static __forceinline__ int posG3(int arr, int z, int qy, int qx)
   {
      return NEWL ? arr * PA + qy * (Q * D1) + ((D1 * qx + z) ^ (D1 * ((qy >> 1) & 1)))
                  : arr * PA + z * PZ + qy * Q + qx;
   }
   __device__ static __forceinline__ int posS2(int iz, int iy, int qx)
   { return NEWL ? (D1 * D1 + 4) * qx + D1 * iz + iy : (iz * D1 + iy) * RQ + qx; }
   __device__ static __forceinline__ int posF1(int f, int j, int q) { return f * FS + j * RQ + q; }
   // fwd-y / bwd-y line index -> (z, qx)
   __device__ static __forceinline__ void lineB(int line, int &z, int &qx)
   {
      if (NEWL) { qx = line / D1; z = line - qx * D1; }
      else { z = line / Q; qx = line - z * Q; }
   }
   // ---- one data stage (doubles)
   static constexpr int NDP = NEWL ? ((D1 * SZU + 1) & ~1) : ((ND + 1) & ~1);
   static constexpr int P_U = 0;
   static constexpr int P_X = P_U + NDP;
   static constexpr int P_N = P_X + ((ND + 1) & ~1);
   static constexpr int P_B = P_N + ((NF * NFD + 1) & ~1);
   static constexpr int P_E = P_B + N3 * 2;
   static constexpr int PSZ = P_E + 2;
   // ---- work arrays (doubles)
   static constexpr int SZ_C0 = 2 * NL * RQ;                      // BU | GU
   static constexpr int SZ_C1 = NL * RQ + NF * NFD;               // S2 | face results
   static constexpr int SZ_C = ((SZ_C0 > SZ_C1 ? SZ_C0 : SZ_C1) + 1) & ~1;
   static constexpr int SZ_B = 3 * PA;                            // GB | BG | BB, later T4
   static constexpr int SZ_G0 = NF * FS;                          // F1
   static constexpr int SZ_G = ((SZ_G0 > ND ? SZ_G0 : ND) + 1) & ~1;   // ... later X
   static constexpr int OFF_D = 0;
   static constexpr int OFF_C = OFF_D + 2 * PSZ;
   static constexpr int OFF_B = OFF_C + SZ_C;
   static constexpr int OFF_G = OFF_B + SZ_B;
   static constexpr int OFF_O = OFF_G + SZ_G;                     // 12 operator coefficients (LIN)
   static constexpr int WDBL = OFF_O + 12;                        // doubles per warp
   static constexpr int I_NE = 0, I_NP = NF, I_BI = 2 * NF, ISZ = (2 * NF + N3 + 1) & ~1;
   static constexpr int WINT = 2 * ISZ;                           // ints per warp
   static constexpr int WBYTES = WDBL * 8 + WINT * 4;
   // shared by the block
   static constexpr int PATMAX = 16;
   // block constants: Minv columns [0, 10), quadrature (point, weight) pairs [10, 26), patterns
   static constexpr int C_XW = 10, C_PAT = 26;
   static constexpr int CBYTES = C_PAT * 8 + ((PATMAX * NFD * 2 + 15) & ~15);
   static constexpr int TA_V = (NL + 7) / 8, TA_F = (NT1 + 7) / 8;
   static constexpr int TB_V = (NY + 7) / 8, TB_F = (NT2 + 7) / 8;
   static constexpr int TC_V = (NC + 7) / 8;
   static constexpr int ES_V = QQ * RQ * 3, ES_F = NF * Q * RQ;
   static constexpr size_t bytes(int nw) { return (size_t)CBYTES + (size_t)nw * WBYTES; }
};

template <int D1, int Q>
__device__ __forceinline__ void stagew_fetch_idx(const StagePArgs &a, int *ix, int64_t e, int lane)
{
   using S = SmemW<D1, Q>;
   constexpr int NF = S::NF, N3 = S::N3;
   if (lane < NF)
   {
      cp_async4(ix + S::I_NE + lane, a.fn.nbr_elem + e * NF + lane);
      cp_async4(ix + S::I_NP + lane, a.nbr_pat32 + e * NF + lane);
   }
   const int nb = (a.bounds_type == 0) ? N3 : NF;
   if (lane < nb) { cp_async4(ix + S::I_BI + lane, a.bidx + e * nb + lane); }
}

template <int D1, int Q>
__device__ __forceinline__ void stagew_fetch_data(const StagePArgs &a, double *dst, const int *ix,
                                                  const int16_t *spat, int64_t e, int lane)
{
   using S = SmemW<D1, Q>;
   constexpr int ND = S::ND, NF = S::NF, NFD = S::NFD, N3 = S::N3;
   {
      const double *gu = a.y + e * ND, *gx = a.x0 + e * ND;
      double *U = dst + S::P_U, *X = dst + S::P_X;
      if (S::NEWL)
      {
         // swizzled U: 8-byte granularity
#pragma unroll
         for (int c0 = 0; c0 < ND; c0 += 32)
         {
            const int c = c0 + lane;
            if (c < ND) { cp_async8(U + S::posUj(c), gu + c); }
         }
         if (a.has_x0)
         {
#pragma unroll
            for (int c0 = 0; c0 < ND / 2; c0 += 32)
            {
               const int c = c0 + lane;
               if (c < ND / 2) { cp_async16(X + 2 * c, gx + 2 * c); }
            }
         }
      }
      else if ((ND & 1) == 0)
      {
#pragma unroll
         for (int c0 = 0; c0 < ND / 2; c0 += 32)
         {
            const int c = c0 + lane;
            if (c < ND / 2)
            {
               cp_async16(U + 2 * c, gu + 2 * c);
               if (a.has_x0) { cp_async16(X + 2 * c, gx + 2 * c); }
            }
         }
      }
      else
      {
         // odd block length: element blocks are only 8-byte aligned
#pragma unroll
         for (int c0 = 0; c0 < ND; c0 += 32)
         {
            const int c = c0 + lane;
            if (c < ND)
            {
               cp_async8(U + c, gu + c);
               if (a.has_x0) { cp_async8(X + c, gx + c); }
            }
         }
      }
   }
   {
      double *NB = dst + S::P_N;
      const int *NE_ = ix + S::I_NE, *NP_ = ix + S::I_NP;
#pragma unroll
      for (int i0 = 0; i0 < NF * NFD; i0 += 32)
      {
         const int id = i0 + lane;
         if (id < NF * NFD)
         {
            const int f = id / NFD, j = id - f * NFD;
            const int nb = NE_[f];
            if (nb >= 0)
            {
               const int pid = NP_[f];
               const int loc = (pid < S::PATMAX) ? spat[pid * NFD + j] : a.fn.pat[pid * NFD + j];
               const double *src = (nb < a.fn.ne_owned)
                                      ? a.y + (int64_t)nb * ND + loc
                                      : a.fn.ughost + ((int64_t)nb - a.fn.ne_owned) * NFD + j;
               cp_async8(NB + id, src);
            }
            else { NB[id] = 0.0; }
         }
      }
   }
   if (lane == 0) { cp_async8(dst + S::P_E, a.einv + e); }
   {
      double *BD = dst + S::P_B;
      const int *BI = ix + S::I_BI;
      if (a.bounds_type == 0)
      {
         if (lane < N3) { cp_async16(BD + 2 * lane, a.ent_mm + 2 * (int64_t)BI[lane]); }
      }
      else if (lane <= NF)
      {
         const int64_t src = (lane == NF) ? e : (int64_t)BI[lane];
         double *d = BD + 2 * lane;
         if (src >= 0) { cp_async8(d, a.xe_min + src); cp_async8(d + 1, a.xe_max + src); }
         else { d[0] = INFINITY; d[1] = -INFINITY; }
      }
   }
}

// pull the stored quadrature data of element e into L2
template <int D1, int Q>
__device__ __forceinline__ void stagew_prefetch_op(const StagePArgs &a, int64_t e, int lane)
{
   using S = SmemW<D1, Q>;
   const char *pv = reinterpret_cast<const char *>(a.Dvol + (size_t)e * S::ES_V);
   const char *pf = reinterpret_cast<const char *>(a.Dface + (size_t)e * S::ES_F);
#pragma unroll
   for (int o = 0; o < S::ES_V * 8; o += 32 * 128)
   {
      const int off = o + lane * 128;
      if (off < S::ES_V * 8) { asm volatile("prefetch.global.L2 [%0];" ::"l"(pv + off)); }
   }
#pragma unroll
   for (int o = 0; o < S::ES_F * 8; o += 32 * 128)
   {
      const int off = o + lane * 128;
      if (off < S::ES_F * 8) { asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + off)); }
   }
}

// LIN: 12 coefficients of element e (see k_op_linear) -> the warp's single coefficient buffer
template <int D1, int Q>
__device__ __forceinline__ void stagew_fetch_opc(const StagePArgs &a, double *dst, int64_t e, int lane)
{
   if (lane < 6) { cp_async16(dst + 2 * lane, a.opc + e * 12 + 2 * lane); }
}

// LIN = true: every element's stored quadrature data is reproduced by a velocity that is linear
// over the (affine) element, adj(J) v = W0 + W1 x + W2 y + W3 z on the reference cube (detected at
// set-up by k_op_linear, ctx.cu).  The kernel then rebuilds Dvol / Dface at the quadrature points
// from 12 doubles per element instead of streaming 3 Q^3 + 6 Q^2 stored values (108 of the 134
// B/DOF the streamed variant moves at order 3).
template <int D1, int Q, int NW, int MINB, bool LIN>
__global__ void __launch_bounds__(NW * 32, MINB)
k_stage3w(StagePArgs a, const Tab<D1, Q> tab)
{
   using S = SmemW<D1, Q>;
   constexpr int ND = S::ND, NF = S::NF, NFD = S::NFD, RQ = S::RQ;
   constexpr int KF = S::KF, KB = S::KB, NL = S::NL, NY = S::NY, NC = S::NC, NT1 = S::NT1, NT2 = S::NT2;
   constexpr int PZ = S::PZ, PA = S::PA;
   constexpr int NK = (ND + 31) / 32;
   static_assert(Q <= 8 && D1 <= 8, "single DMMA tile per output index");
   extern __shared__ double sm[];
   const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int g = lane >> 2, c = lane & 3;
   // block-shared constants: Minv columns, pattern table
   double *MV = sm;
   const double2 *XW = reinterpret_cast<const double2 *>(sm + S::C_XW);
   int16_t *spat = reinterpret_cast<int16_t *>(sm + S::C_PAT);
   double *wsm = reinterpret_cast<double *>(reinterpret_cast<char *>(sm) + S::CBYTES) + (size_t)w * (S::WBYTES / 8);
   int *ismem = reinterpret_cast<int *>(wsm + S::WDBL);
   const double inv_dt = 1.0 / a.dt;
   // ---- coefficient fragments (fragment algebra: file header)
   double fB[KF], fG[KF], bC[KB];
#pragma unroll
   for (int ks = 0; ks < KF; ks++)
   {
      const int k = ks * 4 + c;
      const bool on = (g < Q) && (k < D1);
      fB[ks] = on ? tab.B[on ? g : 0][on ? k : 0] : 0.0;
      fG[ks] = on ? tab.G[on ? g : 0][on ? k : 0] : 0.0;
   }
#pragma unroll
   for (int ks = 0; ks < KB; ks++)
   {
      const int k = ks * 4 + c;
      const bool on = (g < D1) && (k < Q);
      bC[ks] = on ? tab.C[on ? g : 0][on ? k : 0] : 0.0;
   }
   const bool on1 = (g < D1) && (2 * c < Q), on2 = (g < D1) && (2 * c + 1 < Q);
   const double cC1 = on1 ? tab.C[on1 ? g : 0][on1 ? 2 * c : 0] : 0.0;
   const double cC2 = on2 ? tab.C[on2 ? g : 0][on2 ? 2 * c + 1 : 0] : 0.0;
   const int i0 = 2 * c, i1 = 2 * c + 1;
   const double mi00 = (i0 < D1) ? tab.Minv[i0 < D1 ? i0 : 0][0] : 0.0;
   const double mi01 = (i0 < D1) ? tab.Minv[i0 < D1 ? i0 : 0][D1 - 1] : 0.0;
   const double mi10 = (i1 < D1) ? tab.Minv[i1 < D1 ? i1 : 0][0] : 0.0;
   const double mi11 = (i1 < D1) ? tab.Minv[i1 < D1 ? i1 : 0][D1 - 1] : 0.0;
   if (threadIdx.x < D1)
   {
      MV[2 * threadIdx.x] = tab.Minv[threadIdx.x][0];
      MV[2 * threadIdx.x + 1] = tab.Minv[threadIdx.x][D1 - 1];
   }
   if (threadIdx.x < 8)
   {
      const bool on = threadIdx.x < Q;
      sm[S::C_XW + 2 * threadIdx.x] = on ? tab.xq[on ? threadIdx.x : 0] : 0.0;
      sm[S::C_XW + 2 * threadIdx.x + 1] = on ? tab.wq[on ? threadIdx.x : 0] : 0.0;
   }
   // LIN: points / weights of this lane's two z (or qb) indices 2c, 2c+1 (weight 0 beyond Q);
   // the convection sign alpha = -1 (transport, remhos.cpp:648-657) rides on the z weights
   const bool zon0 = (2 * c < Q), zon1 = (2 * c + 1 < Q);
   const double xz0 = zon0 ? tab.xq[zon0 ? 2 * c : 0] : 0.0, xz1 = zon1 ? tab.xq[zon1 ? 2 * c + 1 : 0] : 0.0;
   const double wz0 = zon0 ? tab.wq[zon0 ? 2 * c : 0] : 0.0, wz1 = zon1 ? tab.wq[zon1 ? 2 * c + 1 : 0] : 0.0;
   {
      const int np = a.npat < S::PATMAX ? a.npat : S::PATMAX;
      for (int i = threadIdx.x; i < np * NFD; i += NW * 32) { spat[i] = a.fn.pat[i]; }
   }
   int cls[NK];
#pragma unroll
   for (int k = 0; k < NK; k++)
   {
      int j = lane + 32 * k, t = 0, mul = 1;
      if (j >= ND) { j = 0; }
#pragma unroll
      for (int ax = 0; ax < 3; ax++)
      {
         const int l = j % D1; j /= D1;
         t += ((l == 0) ? 0 : ((l == D1 - 1) ? 2 : 1)) * mul; mul *= 3;
      }
      cls[k] = t;
   }
   __syncthreads();     // the only block barrier: shared constants are in place
   const int64_t GW = (int64_t)gridDim.x * NW;
   int64_t e = (int64_t)blockIdx.x * NW + w;
   if (e >= a.ne) { return; }
   // ---- prologue
   stagew_fetch_idx<D1, Q>(a, ismem, e, lane);
   cp_async_commit();
   cp_async_wait_all();
   __syncwarp();
   stagew_fetch_data<D1, Q>(a, wsm + S::OFF_D, ismem, spat, e, lane);
   if (e + GW < a.ne) { stagew_fetch_idx<D1, Q>(a, ismem + S::ISZ, e + GW, lane); }
   if (LIN) { stagew_fetch_opc<D1, Q>(a, wsm + S::OFF_O, e, lane); }
   cp_async_commit();
   if (!LIN) { stagew_prefetch_op<D1, Q>(a, e, lane); }
   const double *OPC = wsm + S::OFF_O;
   double *BU = wsm + S::OFF_C, *GU = BU + NL * RQ;
   double *G3 = wsm + S::OFF_B;
   double *F1 = wsm + S::OFF_G;
   double *FD = wsm + S::OFF_C + NL * RQ;   // face results (phase C on; GU is dead by then)
   double *S2 = S::NEWL ? wsm + S::OFF_B + PA : BU;   // phase D on (NEWL: the dead G3 array 1)
   double *X = wsm + S::OFF_G;              // HO result (phase E; F1 is dead by then)
   for (int it = 0; e < a.ne; e += GW, it++)
   {
      const int s = it & 1;
      double *dat = wsm + S::OFF_D + s * S::PSZ;
      const double *U = dat + S::P_U, *NB = dat + S::P_N;
      cp_async_wait_all();
      __syncwarp();      // data(e), idx(e+GW) landed; the previous element is fully consumed
      {
         const int64_t e1 = e + GW, e2 = e + 2 * GW;
         if (e1 < a.ne)
         {
            stagew_fetch_data<D1, Q>(a, wsm + S::OFF_D + (s ^ 1) * S::PSZ, ismem + (s ^ 1) * S::ISZ, spat,
                                     e1, lane);
            if (!LIN) { stagew_prefetch_op<D1, Q>(a, e1, lane); }
         }
         if (e2 < a.ne) { stagew_fetch_idx<D1, Q>(a, ismem + s * S::ISZ, e2, lane); }
         cp_async_commit();
      }
      const double *dvp = a.Dvol + (size_t)e * S::ES_V;
      const double *dfp = a.Dface + (size_t)e * S::ES_F;
      // operator data is consumed two phases later: issue the loads now (registers) so their L2
      // latency overlaps phases A and B -- all face tiles and the first z-stage tile; the other
      // z-stage tiles are fetched one tile ahead
      double dfv[LIN ? 1 : S::TB_F][2];
#pragma unroll
      for (int t = 0; t < (LIN ? 0 : S::TB_F); t++)
      {
         const int line = t * 8 + g;
         dfv[t][0] = 0.0; dfv[t][1] = 0.0;
         if (line < NT2 && 2 * c < Q)
         {
            const double2 v = __ldcs(reinterpret_cast<const double2 *>(dfp + line * RQ + 2 * c));
            dfv[t][0] = v.x; dfv[t][1] = v.y;
         }
      }
      double dvn[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      auto load_dv = [&](int t, double (&dv)[6])
      {
         const int col = t * 8 + g;
#pragma unroll
         for (int i = 0; i < 6; i++) { dv[i] = 0.0; }
         if (!LIN && col < NC && 2 * c < Q)
         {
            const double2 *p = reinterpret_cast<const double2 *>(dvp + (col * RQ + 2 * c) * 3);
            const double2 v0 = __ldcs(p), v1 = __ldcs(p + 1), v2 = __ldcs(p + 2);
            dv[0] = v0.x; dv[1] = v0.y; dv[2] = v1.x; dv[3] = v1.y; dv[4] = v2.x; dv[5] = v2.y;
         }
      };
      if (!LIN) { load_dv(0, dvn); }
      // ================= A: fwd-x (rows = lines (z,y), k = ix) | face fwd-a (rows = (f,jb), k = ja)
#pragma unroll
      for (int t = 0; t < S::TA_V; t++)
      {
         const int line = t * 8 + g;
         double bu0 = 0.0, bu1 = 0.0, gu0 = 0.0, gu1 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KF; ks++)
         {
            const int k = ks * 4 + c;
            const double x = (line < NL && k < D1) ? U[S::posU(line / D1, line % D1, k)] : 0.0;
            dmma884(bu0, bu1, x, fB[ks]);
            dmma884(gu0, gu1, x, fG[ks]);
         }
         if (line < NL && 2 * c < RQ)
         {
            const bool sw = S::NEWL && ((line >> 3) & 1);     // posBU: rows 8..15 hold swapped pairs
            *reinterpret_cast<double2 *>(BU + line * RQ + 2 * c) = sw ? make_double2(bu1, bu0) : make_double2(bu0, bu1);
            *reinterpret_cast<double2 *>(GU + line * RQ + 2 * c) = sw ? make_double2(gu1, gu0) : make_double2(gu0, gu1);
         }
      }
#pragma unroll
      for (int t = 0; t < S::TA_F; t++)
      {
         const int line = t * 8 + g;                        // (f, jb)
         const int f = line / D1, jb = line - f * D1;
         const int axis = (f == 0 || f == 5) ? 2 : ((f == 1 || f == 3) ? 1 : 0);
         const int side = (f == 2 || f == 3 || f == 5) ? 1 : 0;
         const int fx = side * (D1 - 1);
         double f0 = 0.0, f1 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KF; ks++)
         {
            const int k = ks * 4 + c;
            // natural face parametrisation: (ja, jb) = the two remaining axes, ascending
            const int ux = (axis == 0) ? fx : k;
            const int uy = (axis == 1) ? fx : ((axis == 0) ? k : jb);
            const int uz = (axis == 2) ? fx : jb;
            const double x = (line < NT1 && k < D1) ? U[S::posU(uz, uy, ux)] - NB[line * D1 + k] : 0.0;
            dmma884(f0, f1, x, fB[ks]);
         }
         if (line < NT1 && 2 * c < RQ)
         {
            *reinterpret_cast<double2 *>(F1 + S::posF1(f, jb, 2 * c)) = make_double2(f0, f1);
         }
      }
      __syncwarp();
      // ================= B: fwd-y (rows = qy, cols = lines (z,qx), k = iy) | fused face stage
#pragma unroll
      for (int t = 0; t < S::TB_V; t++)
      {
         const int line = t * 8 + g;
         int z, qx;
         S::lineB(line, z, qx);
         double gb0 = 0.0, gb1 = 0.0, bg0 = 0.0, bg1 = 0.0, bb0 = 0.0, bb1 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KF; ks++)
         {
            const int k = ks * 4 + c;
            const bool on = (line < NY) && (k < D1);
            const double xb = on ? BU[S::posBU(z * D1 + k, qx)] : 0.0;
            const double xg = on ? GU[S::posBU(z * D1 + k, qx)] : 0.0;
            dmma884(gb0, gb1, fB[ks], xg);     // GB = By Gx u
            dmma884(bg0, bg1, fG[ks], xb);     // BG = Gy Bx u
            dmma884(bb0, bb1, fB[ks], xb);     // BB = By Bx u
         }
         if (S::NEWL)
         {
            // lines 2c, 2c+1 of the tile = (qx, z), (qx, z+1): one 16-byte store per array
            const int ls = t * 8 + 2 * c;
            if (g < Q && ls < NY)
            {
               int zs, qxs;
               S::lineB(ls, zs, qxs);
               double *o = G3 + S::posG3(0, zs, g, qxs);
               *reinterpret_cast<double2 *>(o) = make_double2(gb0, gb1);
               *reinterpret_cast<double2 *>(o + PA) = make_double2(bg0, bg1);
               *reinterpret_cast<double2 *>(o + 2 * PA) = make_double2(bb0, bb1);
            }
         }
         else if (g < Q)
         {
#pragma unroll
            for (int h = 0; h < 2; h++)
            {
               const int ls = t * 8 + 2 * c + h;
               if (ls < NY)
               {
                  const int zs = ls / Q, qxs = ls - zs * Q;
                  double *o = G3 + zs * PZ + g * Q + qxs;
                  o[0] = h ? gb1 : gb0;
                  o[PA] = h ? bg1 : bg0;
                  o[2 * PA] = h ? bb1 : bb0;
               }
            }
         }
      }
#pragma unroll
      for (int t = 0; t < S::TB_F; t++)
      {
         const int line = t * 8 + g;                        // (f, qa)
         const int f = line / Q, qa = line - f * Q;
         double dfv0, dfv1;
         if constexpr (LIN)
         {
            // w_qa w_qb min(0, v.n), v.n = +-(adj(J) v)_axis on the face (k_geom_face, ctx.cu)
            const int fc_ = (line < NT2) ? f : 0;
            const int axis = (fc_ == 0 || fc_ == 5) ? 2 : ((fc_ == 1 || fc_ == 3) ? 1 : 0);
            const bool side = (fc_ == 2 || fc_ == 3 || fc_ == 5);
            const int ia = (axis == 0) ? 1 : 0, ib = (axis == 2) ? 1 : 2;
            const double2 ta = XW[(line < NT2) ? qa : 0];
            double base = OPC[axis];
            if (side) { base += OPC[3 * (1 + axis) + axis]; }
            base = fma(OPC[3 * (1 + ia) + axis], ta.x, base);
            const double cb = OPC[3 * (1 + ib) + axis];
            double vn0 = fma(cb, xz0, base), vn1 = fma(cb, xz1, base);
            if (!side) { vn0 = -vn0; vn1 = -vn1; }
            const double wa = (line < NT2) ? ta.y : 0.0;
            dfv0 = wa * wz0 * fmin(0.0, vn0);
            dfv1 = wa * wz1 * fmin(0.0, vn1);
         }
         else { dfv0 = dfv[t][0]; dfv1 = dfv[t][1]; }
         double y0 = 0.0, y1 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KF; ks++)
         {
            const int kk = ks * 4 + c;
            const double x = (line < NT2 && kk < D1) ? F1[S::posF1(f, kk, qa)] : 0.0;
            dmma884(y0, y1, x, fB[ks]);                     // [line][qb = 2c, 2c+1]
         }
         y0 *= dfv0; y1 *= dfv1;
         double z0 = 0.0, z1 = 0.0;
         dmma884(z0, z1, y0, cC1);                          // [line][ib = 2c, 2c+1]
         dmma884(z0, z1, y1, cC2);
         __syncwarp();
         if (line < NT2)
         {
            if (2 * c < D1) { F1[S::posF1(f, 2 * c, qa)] = z0; }
            if (2 * c + 1 < D1) { F1[S::posF1(f, 2 * c + 1, qa)] = z1; }
         }
      }
      __syncwarp();
      // ================= C: z-stage (rows = columns (qy,qx), k = iz; chained back) | face back-a
#pragma unroll
      for (int t = 0; t < S::TC_V; t++)
      {
         const int col = t * 8 + g;
         double dv[6];
         if constexpr (LIN)
         {
            // alpha w_q (adj(J) v)_c at (qx, qy, qz = 2c | 2c+1), alpha = -1
            const int cc_ = (col < NC) ? col : 0;
            const double2 tx = XW[cc_ % Q], ty = XW[cc_ / Q];
            const double wxy = (col < NC) ? -(tx.y * ty.y) : 0.0;
            const double w0_ = wxy * wz0, w1_ = wxy * wz1;
#pragma unroll
            for (int i = 0; i < 3; i++)
            {
               const double A = fma(OPC[6 + i], ty.x, fma(OPC[3 + i], tx.x, OPC[i]));
               dv[i] = w0_ * fma(OPC[9 + i], xz0, A);
               dv[3 + i] = w1_ * fma(OPC[9 + i], xz1, A);
            }
         }
         else
         {
#pragma unroll
            for (int i = 0; i < 6; i++) { dv[i] = dvn[i]; }
            if (t + 1 < S::TC_V) { load_dv(t + 1, dvn); }
         }
         double g00 = 0.0, g01 = 0.0, g10 = 0.0, g11 = 0.0, g20 = 0.0, g21 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KF; ks++)
         {
            const int kk = ks * 4 + c;
            const bool on = (col < NC) && (kk < D1);
            const int po = S::posG3(0, kk, col / Q, col % Q);
            const double x0 = on ? G3[po] : 0.0;
            const double x1 = on ? G3[PA + po] : 0.0;
            const double x2 = on ? G3[2 * PA + po] : 0.0;
            dmma884(g00, g01, x0, fB[ks]);                  // d/dx: Bz (By Gx u)
            dmma884(g10, g11, x1, fB[ks]);                  // d/dy: Bz (Gy Bx u)
            dmma884(g20, g21, x2, fG[ks]);                  // d/dz: Gz (By Bx u)
         }
         const double s0 = dv[0] * g00 + dv[1] * g10 + dv[2] * g20;
         const double s1 = dv[3] * g01 + dv[4] * g11 + dv[5] * g21;
         double z0 = 0.0, z1 = 0.0;
         dmma884(z0, z1, s0, cC1);                          // [column][iz = 2c, 2c+1]
         dmma884(z0, z1, s1, cC2);
         __syncwarp();
         if (S::NEWL)
         {
            if (col < NC && 2 * c < D1)
            {
               *reinterpret_cast<double2 *>(G3 + S::posG3(0, 2 * c, col / Q, col % Q)) = make_double2(z0, z1);
            }
         }
         else if (col < NC)
         {
            if (2 * c < D1) { G3[(2 * c) * PZ + col] = z0; }
            if (2 * c + 1 < D1) { G3[(2 * c + 1) * PZ + col] = z1; }
         }
      }
#pragma unroll
      for (int t = 0; t < S::TA_F; t++)
      {
         const int line = t * 8 + g;                        // (f, ib)
         double y0 = 0.0, y1 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KB; ks++)
         {
            const int kk = ks * 4 + c;
            const double x = (line < NT1 && kk < Q) ? F1[S::posF1(line / D1, line % D1, kk)] : 0.0;
            dmma884(y0, y1, x, bC[ks]);                     // [line][ia = 2c, 2c+1]
         }
         if (line < NT1)
         {
            if (2 * c < D1) { FD[line * D1 + 2 * c] = y0; }
            if (2 * c + 1 < D1) { FD[line * D1 + 2 * c + 1] = y1; }
         }
      }
      __syncwarp();
      if (LIN)
      {
         // the coefficient buffer is single: refill it for the next element now that phases B
         // and C have consumed it (lands long before the next element's phase B)
         if (e + GW < a.ne) { stagew_fetch_opc<D1, Q>(a, wsm + S::OFF_O, e + GW, lane); }
         cp_async_commit();
      }
      // ================= D: bwd-y (rows = iy, cols = lines (iz,qx), k = qy)
#pragma unroll
      for (int t = 0; t < S::TB_V; t++)
      {
         const int line = t * 8 + g;
         int iz, qx;
         S::lineB(line, iz, qx);
         double y0 = 0.0, y1 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KB; ks++)
         {
            const int kk = ks * 4 + c;
            const double x = (line < NY && kk < Q) ? G3[S::posG3(0, iz, kk, qx)] : 0.0;
            dmma884(y0, y1, bC[ks], x);                     // [iy = g][lines 2c, 2c+1]
         }
         if (g < D1)
         {
#pragma unroll
            for (int h = 0; h < 2; h++)
            {
               const int ls = t * 8 + 2 * c + h;
               if (ls < NY)
               {
                  int izs, qxs;
                  S::lineB(ls, izs, qxs);
                  S2[S::posS2(izs, g, qxs)] = h ? y1 : y0;
               }
            }
         }
      }
      __syncwarp();
      // ================= E: bwd-x (rows = lines (iz,iy), k = qx) + face combine -> X
      const double sc = dat[S::P_E];
#pragma unroll
      for (int t = 0; t < S::TA_V; t++)
      {
         const int line = t * 8 + g;
         double r0 = 0.0, r1 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KB; ks++)
         {
            const int kk = ks * 4 + c;
            const double x = (line < NL && kk < Q) ? S2[S::posS2(line / D1, line % D1, kk)] : 0.0;
            dmma884(r0, r1, x, bC[ks]);                     // [line][ix = 2c, 2c+1]
         }
         if (line < NL && i0 < D1)
         {
            const int bb_ = line / D1, aa_ = line - bb_ * D1;
            const double *fc = FD;
            // faces: 0 z=0 (x,y)  1 y=0 (x,z)  2 x=p (y,z)  3 y=p (x,z)  4 x=0 (y,z)  5 z=p (x,y)
            const double fx0 = fc[4 * NFD + bb_ * D1 + aa_], fx1 = fc[2 * NFD + bb_ * D1 + aa_];
            const double my0 = MV[2 * aa_], my1 = MV[2 * aa_ + 1];
            const double mz0 = MV[2 * bb_], mz1 = MV[2 * bb_ + 1];
            double v = r0;
            v = fma(mi00, fx0, v);
            v = fma(mi01, fx1, v);
            v = fma(my0, fc[1 * NFD + bb_ * D1 + i0], v);
            v = fma(my1, fc[3 * NFD + bb_ * D1 + i0], v);
            v = fma(mz0, fc[0 * NFD + aa_ * D1 + i0], v);
            v = fma(mz1, fc[5 * NFD + aa_ * D1 + i0], v);
            X[line * D1 + i0] = v * sc;
            if (i1 < D1)
            {
               v = r1;
               v = fma(mi10, fx0, v);
               v = fma(mi11, fx1, v);
               v = fma(my0, fc[1 * NFD + bb_ * D1 + i1], v);
               v = fma(my1, fc[3 * NFD + bb_ * D1 + i1], v);
               v = fma(mz0, fc[0 * NFD + aa_ * D1 + i1], v);
               v = fma(mz1, fc[5 * NFD + aa_ * D1 + i1], v);
               X[line * D1 + i1] = v * sc;
            }
         }
      }
      __syncwarp();
      // ================= element-wise tail (MassBasedAvg, bounds, ClipScale, RK); see stage3p.cuh
      {
         const double *X0 = dat + S::P_X;
         const double *BD = dat + S::P_B;
         const double dt = a.dt;
         const double inv_m = sc * (double)ND, m = 1.0 / inv_m, mdt = m * inv_dt;
         double u[NK], du_ho[NK], f[NK], lo[NK], bmn[NK], bmx[NK];
         double bmin1 = INFINITY, bmax1 = -INFINITY;
         if (a.bounds_type == 1)
         {
#pragma unroll
            for (int k = 0; k <= NF; k++) { bmin1 = fmin(bmin1, BD[2 * k]); bmax1 = fmax(bmax1, BD[2 * k + 1]); }
         }
         double s1 = 0.0;
#pragma unroll
         for (int k = 0; k < NK; k++)
         {
            const int j = lane + 32 * k;
            if (j < ND)
            {
               u[k] = U[S::posUj(j)];
               du_ho[k] = X[j];
               if (a.bounds_type == 0) { bmn[k] = BD[2 * cls[k]]; bmx[k] = BD[2 * cls[k] + 1]; }
               else { bmn[k] = bmin1; bmx[k] = bmax1; }
               s1 += u[k] + dt * du_ho[k];
            }
         }
         s1 = warp_sum(s1);
         const double ubar = s1 * (1.0 / ND);                // MassBasedAvg, remhos_lo.cpp:278-285
         double sumPos = 0.0, sumNeg = 0.0;
#pragma unroll
         for (int k = 0; k < NK; k++)
         {
            const int j = lane + 32 * k;
            if (j < ND)
            {
               lo[k] = (ubar - u[k]) * inv_dt;
               const double u_new_lo = u[k] + dt * lo[k];
               const double fmn = mdt * (bmn[k] - u_new_lo);
               const double fmx = mdt * (bmx[k] - u_new_lo);
               double fcl = m * (du_ho[k] - lo[k]);
               fcl = fmin(fmx, fmax(fmn, fcl));               // ClipScale, remhos_fct.cpp:490-515
               f[k] = fcl;
               sumNeg += fmin(fcl, 0.0);
               sumPos += fmax(fcl, 0.0);
            }
         }
         warp_sum2(sumNeg, sumPos);
         const double new_mass = sumNeg + sumPos;
         constexpr double eps = 1.0e-15;
         const bool sp = new_mass > eps, sn = new_mass < -eps;
         const double ratio = sp ? sumNeg / sumPos : (sn ? sumPos / sumNeg : 0.0);
         double omin = INFINITY, omax = -INFINITY;
#pragma unroll
         for (int k = 0; k < NK; k++)
         {
            const int j = lane + 32 * k;
            if (j < ND)
            {
               double fcl = f[k];
               if (sp) { fcl = fmin(0.0, fcl) - fmax(0.0, fcl) * ratio; }
               if (sn) { fcl = fmax(0.0, fcl) - fmin(0.0, fcl) * ratio; }
               const double du = lo[k] + fcl * inv_m;
               double o = du;
               if (a.out_mode == 1)
               {
                  const double base = a.has_x0 ? a.a * X0[j] : 0.0;
                  o = base + a.b * (u[k] + dt * du);
               }
               a.out[e * ND + j] = o;
               omin = fmin(omin, o); omax = fmax(omax, o);
            }
         }
         if (a.xe_min_out)
         {
            warp_minmax(omin, omax);
            if (lane == 0) { a.xe_min_out[e] = omin; a.xe_max_out[e] = omax; }
         }
      }
   }
}

} // namespace rmh

#endif
