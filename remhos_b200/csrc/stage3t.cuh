// FP64 tensor-core (DMMA m8n8k4) variant of the persistent pipelined 3D stage kernel.
//
// Why tensor cores: ncu on the DFMA kernel (profiles/r01/ncu_stage_v5_rs5.txt) shows the batched
// contractions are issue-bound, not HBM-bound: 54 % issue-slot utilisation with only 20 % of the
// issued instructions being DFMA -- sm_100a DFMA has no constant-bank operand, so every
// coefficient costs an LDCU into a uniform register, and every 4x6 line costs 10 LDS/STS.  DMMA
// runs on the same FP64 pipe at the same peak (measured 36.9 vs 36.6 TFLOP/s, tools/micro) but
// issues 256 FMAs per warp instruction with the 1-D matrices held as register fragments.
//
// Every contraction is a set of 8-line tiles  D[8 x 8] += A[8 x 4] B[4 x 8]  (g = lane/4, c = lane%4):
//     A fragment: A[g][c]          B fragment: B[k=c][n=g]          D fragment: D[g][2c], D[g][2c+1]
// A 1-D matrix M (Q x D1 forward, D1 x Q backward) is held once per thread as fragment registers
// (the same register serves as A operand "rows = outputs" and as B operand "cols = outputs").
// A D fragment whose columns are the next contracted index is fed straight back as the A operand
// of the following DMMA (columns 2c -> k-step 1, 2c+1 -> k-step 2, coefficient rows permuted to
// match), so the z-stage (forward-z, D.grad u, backward-z) and the fused face stage never leave
// registers.
//
// Pipeline, prefetch stages and the element-wise tail are those of stage3p.cuh.  The stored
// quadrature data is laid out in HBM in fragment order for this kernel (ctx.cu, OpLayout):
//     Dvol [e][col = qy*Q+qx][qz (RQ)][3]     -> 3 x 16-byte loads per thread and tile, coalesced
//     Dface[e][f][qa][qb (RQ)]                -> 1 x 16-byte load per thread and tile
#ifndef RMH_STAGE3T_CUH
#define RMH_STAGE3T_CUH

#include "stage3p.cuh"

namespace rmh
{

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
   asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(d0), "+d"(d1)
                : "d"(a), "d"(b));
}

template <int D1, int Q, int E, int NW>
struct SmemT
{
   static constexpr int ND = D1 * D1 * D1, NQ = Q * Q * Q, QQ = Q * Q;
   static constexpr int NF = 6, NFD = D1 * D1, N3 = 27;
   static constexpr int T = NW * 32;
   static constexpr int RQ = (Q + 1) & ~1;                        // row of Q values, even length
   static constexpr int KF = (D1 + 3) / 4, KB = (Q + 3) / 4;      // k-steps forward / backward
   static constexpr int NL = E * D1 * D1;                         // x-lines (e, z, y)
   static constexpr int NY = E * D1 * Q;                          // y-lines (e, z, qx)
   static constexpr int NC = E * QQ;                              // z-columns (e, qy, qx)
   static constexpr int NT1 = E * NF * D1;                        // face lines (e, f, jb)
   static constexpr int NT2 = E * NF * Q;                         // face lines (e, f, qa)
   static constexpr int PZ = ((NC + 15) / 16) * 16 + 4;           // z-plane stride, = 4 mod 16
   static constexpr int PA = D1 * PZ;                             // array stride of G3
   // ---- one data stage (doubles)
   static constexpr int P_U = 0;
   static constexpr int P_X = P_U + E * ND;
   static constexpr int P_N = P_X + E * ND;
   static constexpr int P_B = P_N + E * NF * NFD;
   static constexpr int P_E = P_B + E * N3 * 2;
   static constexpr int PSZ = ((P_E + E) + 1) & ~1;
   // ---- work arrays (doubles)
   static constexpr int SZ_C0 = 2 * NL * RQ;                      // BU | GU
   static constexpr int SZ_C1 = NL * RQ + E * NF * NFD;           // S2 | face results
   static constexpr int SZ_C = ((SZ_C0 > SZ_C1 ? SZ_C0 : SZ_C1) + 1) & ~1;
   static constexpr int SZ_B = 3 * PA;                            // G3: GB | BG | BB, later T4
   static constexpr int SZ_G0 = NT1 * RQ;                         // F1
   static constexpr int SZ_G = ((SZ_G0 > E * ND ? SZ_G0 : E * ND) + 1) & ~1;   // ... later X
   static constexpr int OFF_D = 0;
   static constexpr int OFF_C = OFF_D + 2 * PSZ;
   static constexpr int OFF_B = OFF_C + SZ_C;
   static constexpr int OFF_G = OFF_B + SZ_B;
   static constexpr int OFF_MV = OFF_G + SZ_G;                    // Minv[:, 0], Minv[:, D1-1]
   static constexpr int NDBL = OFF_MV + 2 * 8;
   static constexpr int I_NE = 0, I_NP = E * NF, I_BI = 2 * E * NF, ISZ = 2 * E * NF + E * N3;
   static constexpr int PATMAX = 32;
   static constexpr int NINT = 2 * ISZ + (PATMAX * NFD + 1) / 2;
   static constexpr size_t BYTES = (size_t)NDBL * 8 + (size_t)NINT * 4;
   // tiles
   static constexpr int TA_V = (NL + 7) / 8, TA_F = (NT1 + 7) / 8;
   static constexpr int TB_V = (NY + 7) / 8, TB_F = (NT2 + 7) / 8;
   static constexpr int TC_V = (NC + 7) / 8;
   // operator data element strides (doubles)
   static constexpr int ES_V = QQ * RQ * 3, ES_F = NF * Q * RQ;
};

// prefetch helpers shared with stage3p, parametrised on the block size
template <int ND, int NF, int NFD, int N3, int E, int T, typename S>
__device__ __forceinline__ void staget_fetch_idx(const StagePArgs &a, int *ix, int64_t e0, int ne)
{
   for (int id = threadIdx.x; id < E * NF; id += T)
   {
      if (id < ne * NF)
      {
         cp_async4(ix + S::I_NE + id, a.fn.nbr_elem + e0 * NF + id);
         cp_async4(ix + S::I_NP + id, a.nbr_pat32 + e0 * NF + id);
      }
   }
   const int nb = (a.bounds_type == 0) ? N3 : NF;
   for (int id = threadIdx.x; id < E * nb; id += T)
   {
      if (id < ne * nb) { cp_async4(ix + S::I_BI + id, a.bidx + e0 * nb + id); }
   }
}

template <int ND, int NF, int NFD, int N3, int E, int T, typename S>
__device__ __forceinline__ void staget_fetch_data(const StagePArgs &a, double *dst, const int *ix,
                                                  const int16_t *spat, int64_t e0, int ne)
{
   {
      const double *gu = a.y + e0 * ND, *gx = a.x0 + e0 * ND;
      double *U = dst + S::P_U, *X = dst + S::P_X;
      for (int c = threadIdx.x; c < E * ND / 2; c += T)
      {
         if (2 * c + 1 < ne * ND)
         {
            cp_async16(U + 2 * c, gu + 2 * c);
            if (a.has_x0) { cp_async16(X + 2 * c, gx + 2 * c); }
         }
         else
         {
#pragma unroll
            for (int h = 0; h < 2; h++)
            {
               const int i = 2 * c + h;
               const bool in = i < ne * ND;
               U[i] = in ? gu[i] : 0.0;
               X[i] = (in && a.has_x0) ? gx[i] : 0.0;
            }
         }
      }
   }
   {
      double *NB = dst + S::P_N;
      const int *NE_ = ix + S::I_NE, *NP_ = ix + S::I_NP;
      for (int id = threadIdx.x; id < E * NF * NFD; id += T)
      {
         const int ef = id / NFD, j = id - ef * NFD;
         const int nb = (ef < ne * NF) ? NE_[ef] : -1;
         if (nb >= 0)
         {
            const int pid = NP_[ef];
            const int loc = (pid < S::PATMAX) ? spat[pid * NFD + j] : a.fn.pat[pid * NFD + j];
            const double *src = (nb < a.fn.ne_owned)
                                   ? a.y + (int64_t)nb * ND + loc
                                   : a.fn.ughost + ((int64_t)nb - a.fn.ne_owned) * ND + loc;
            cp_async8(NB + id, src);
         }
         else { NB[id] = 0.0; }
      }
   }
   for (int e = threadIdx.x; e < E; e += T)
   {
      if (e < ne) { cp_async8(dst + S::P_E + e, a.einv + e0 + e); }
      else { dst[S::P_E + e] = 0.0; }
   }
   {
      double *BD = dst + S::P_B;
      const int *BI = ix + S::I_BI;
      if (a.bounds_type == 0)
      {
         for (int id = threadIdx.x; id < E * N3; id += T)
         {
            if (id < ne * N3) { cp_async16(BD + 2 * id, a.ent_mm + 2 * (int64_t)BI[id]); }
         }
      }
      else
      {
         for (int id = threadIdx.x; id < E * (NF + 1); id += T)
         {
            const int e = id / (NF + 1), k = id - e * (NF + 1);
            if (e < ne)
            {
               const int64_t src = (k == NF) ? (e0 + e) : (int64_t)BI[e * NF + k];
               double *d = BD + 2 * (e * N3 + k);
               if (src >= 0) { cp_async8(d, a.xe_min + src); cp_async8(d + 1, a.xe_max + src); }
               else { d[0] = INFINITY; d[1] = -INFINITY; }
            }
         }
      }
   }
}

template <int D1, int Q, int E, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
k_stage3t(StagePArgs a, const Tab<D1, Q> tab)
{
   using S = SmemT<D1, Q, E, NW>;
   constexpr int ND = S::ND, QQ = S::QQ, NF = S::NF, NFD = S::NFD, T = S::T, N3 = S::N3, RQ = S::RQ;
   constexpr int KF = S::KF, KB = S::KB, NL = S::NL, NY = S::NY, NC = S::NC, NT1 = S::NT1, NT2 = S::NT2;
   constexpr int PZ = S::PZ, PA = S::PA;
   constexpr int NK = (ND + 31) / 32;
   constexpr int KC = (S::TC_V + NW - 1) / NW;     // z-stage tiles per warp
   constexpr int KF2 = (S::TB_F + NW - 1) / NW;    // fused face tiles per warp
   static_assert(NW >= E, "one warp per element in the element-wise phase");
   static_assert(Q <= 8 && D1 <= 8, "single DMMA tile per output index");
   extern __shared__ double sm[];
   int *ismem = reinterpret_cast<int *>(sm + S::NDBL);
   int16_t *spat = reinterpret_cast<int16_t *>(ismem + 2 * S::ISZ);
   double *MV = sm + S::OFF_MV;
   const int64_t nbatch = (a.ne + E - 1) / E;
   const int G = gridDim.x;
   const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int g = lane >> 2, c = lane & 3;
   const double inv_dt = 1.0 / a.dt;
   // ---- coefficient fragments
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
   // chained backward: data columns 2c (k-step 1) and 2c+1 (k-step 2)
   const bool on1 = (g < D1) && (2 * c < Q), on2 = (g < D1) && (2 * c + 1 < Q);
   const double cC1 = on1 ? tab.C[on1 ? g : 0][on1 ? 2 * c : 0] : 0.0;
   const double cC2 = on2 ? tab.C[on2 ? g : 0][on2 ? 2 * c + 1 : 0] : 0.0;
   // normal-direction mass inverse columns for this lane's two output indices i = 2c, 2c+1
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
   {
      const int np = a.npat < S::PATMAX ? a.npat : S::PATMAX;
      for (int i = threadIdx.x; i < np * NFD; i += T) { spat[i] = a.fn.pat[i]; }
   }
   int64_t b = blockIdx.x;
   if (b >= nbatch) { return; }
   {
      const int64_t e0 = b * E;
      staget_fetch_idx<ND, NF, NFD, N3, E, T, S>(a, ismem, e0, (int)min((int64_t)E, a.ne - e0));
      cp_async_commit();
      cp_async_wait_all();
      __syncthreads();
      staget_fetch_data<ND, NF, NFD, N3, E, T, S>(a, sm + S::OFF_D, ismem, spat, e0,
                                                  (int)min((int64_t)E, a.ne - e0));
      const int64_t b1 = b + G;
      if (b1 < nbatch)
      {
         const int64_t e1 = b1 * E;
         staget_fetch_idx<ND, NF, NFD, N3, E, T, S>(a, ismem + S::ISZ, e1, (int)min((int64_t)E, a.ne - e1));
      }
      cp_async_commit();
   }
   for (int it = 0; b < nbatch; b += G, it++)
   {
      const int s = it & 1;
      const int64_t e0 = b * E;
      const int ne = (int)min((int64_t)E, a.ne - e0);
      double *dat = sm + S::OFF_D + s * S::PSZ;
      const double *U = dat + S::P_U, *NB = dat + S::P_N;
      cp_async_wait_all();
      __syncthreads();
      // ---- stored quadrature data of this warp's z-stage / fused-face tiles -> registers
      double dv[KC][6], df[KF2][2];
#pragma unroll
      for (int k = 0; k < KC; k++)
      {
         const int col = (w + k * NW) * 8 + g;
         const int e = col / QQ, cl = col - e * QQ;
         const bool on = (w + k * NW < S::TC_V) && (col < NC) && (e < ne) && (2 * c < Q);
         if (on)
         {
            const double2 *p = reinterpret_cast<const double2 *>(
               a.Dvol + (size_t)(e0 + e) * S::ES_V + (size_t)(cl * RQ + 2 * c) * 3);
            const double2 v0 = __ldcs(p), v1 = __ldcs(p + 1), v2 = __ldcs(p + 2);
            dv[k][0] = v0.x; dv[k][1] = v0.y; dv[k][2] = v1.x;
            dv[k][3] = v1.y; dv[k][4] = v2.x; dv[k][5] = v2.y;
         }
         else
         {
#pragma unroll
            for (int i = 0; i < 6; i++) { dv[k][i] = 0.0; }
         }
      }
#pragma unroll
      for (int k = 0; k < KF2; k++)
      {
         const int line = (w + k * NW) * 8 + g;          // (e, f, qa)
         const int e = line / (NF * Q);
         const bool on = (w + k * NW < S::TB_F) && (line < NT2) && (e < ne) && (2 * c < Q);
         if (on)
         {
            const double2 v = __ldcs(reinterpret_cast<const double2 *>(
               a.Dface + (size_t)e0 * S::ES_F + (size_t)line * RQ + 2 * c));
            df[k][0] = v.x; df[k][1] = v.y;
         }
         else { df[k][0] = 0.0; df[k][1] = 0.0; }
      }
      // ---- prefetch: data(b+G) through idx(b+G); idx(b+2G)
      {
         const int64_t b1 = b + G, b2 = b + 2 * (int64_t)G;
         if (b1 < nbatch)
         {
            const int64_t e1 = b1 * E;
            staget_fetch_data<ND, NF, NFD, N3, E, T, S>(a, sm + S::OFF_D + (s ^ 1) * S::PSZ,
                                                        ismem + (s ^ 1) * S::ISZ, spat, e1,
                                                        (int)min((int64_t)E, a.ne - e1));
         }
         if (b2 < nbatch)
         {
            const int64_t e2 = b2 * E;
            staget_fetch_idx<ND, NF, NFD, N3, E, T, S>(a, ismem + s * S::ISZ, e2,
                                                       (int)min((int64_t)E, a.ne - e2));
         }
         cp_async_commit();
      }
      double *BU = sm + S::OFF_C, *GU = BU + NL * RQ;
      double *G3 = sm + S::OFF_B;
      double *F1 = sm + S::OFF_G;
      double *FD = sm + S::OFF_C + NL * RQ;    // face results (phase C on; GU is dead by then)
      double *S2 = BU;                         // phase D on
      double *X = sm + S::OFF_G;               // HO result (phase E; F1 is dead by then)
      // ================= A: fwd-x (rows = lines (e,z,y), k = ix) | face fwd-a (rows = (e,f,jb), k = ja)
      for (int t = w; t < S::TA_V + S::TA_F; t += NW)
      {
         if (t < S::TA_V)
         {
            const int line = t * 8 + g;
            double bu0 = 0.0, bu1 = 0.0, gu0 = 0.0, gu1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KF; ks++)
            {
               const int k = ks * 4 + c;
               const double x = (line < NL && k < D1) ? U[line * D1 + k] : 0.0;
               dmma884(bu0, bu1, x, fB[ks]);
               dmma884(gu0, gu1, x, fG[ks]);
            }
            if (line < NL && 2 * c < RQ)
            {
               *reinterpret_cast<double2 *>(BU + line * RQ + 2 * c) = make_double2(bu0, bu1);
               *reinterpret_cast<double2 *>(GU + line * RQ + 2 * c) = make_double2(gu0, gu1);
            }
         }
         else
         {
            const int line = (t - S::TA_V) * 8 + g;         // (e, f, jb)
            const int ef = line / D1, jb = line - ef * D1;
            const int e = ef / NF, f = ef - e * NF;
            const int axis = (f == 0 || f == 5) ? 2 : ((f == 1 || f == 3) ? 1 : 0);
            const int side = (f == 2 || f == 3 || f == 5) ? 1 : 0;
            const int s1 = (axis == 0) ? D1 : 1;
            const int s2 = (axis == 2) ? D1 : D1 * D1;
            const int sa = (axis == 0) ? 1 : ((axis == 1) ? D1 : D1 * D1);
            const double *own = U + e * ND + side * (D1 - 1) * sa + jb * s2;
            double f0 = 0.0, f1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KF; ks++)
            {
               const int k = ks * 4 + c;
               const double x = (line < NT1 && k < D1) ? own[k * s1] - NB[line * D1 + k] : 0.0;
               dmma884(f0, f1, x, fB[ks]);
            }
            if (line < NT1 && 2 * c < RQ)
            {
               *reinterpret_cast<double2 *>(F1 + line * RQ + 2 * c) = make_double2(f0, f1);
            }
         }
      }
      __syncthreads();
      // ================= B: fwd-y (rows = qy, cols = lines (e,z,qx), k = iy) | fused face stage
      for (int t = w; t < S::TB_V; t += NW)
      {
         const int line = t * 8 + g;                        // B-operand column
         const int ez = line / Q, qx = line - ez * Q;
         double gb0 = 0.0, gb1 = 0.0, bg0 = 0.0, bg1 = 0.0, bb0 = 0.0, bb1 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KF; ks++)
         {
            const int k = ks * 4 + c;
            const bool on = (line < NY) && (k < D1);
            const double xb = on ? BU[(ez * D1 + k) * RQ + qx] : 0.0;
            const double xg = on ? GU[(ez * D1 + k) * RQ + qx] : 0.0;
            dmma884(gb0, gb1, fB[ks], xg);     // GB = By Gx u
            dmma884(bg0, bg1, fG[ks], xb);     // BG = Gy Bx u
            dmma884(bb0, bb1, fB[ks], xb);     // BB = By Bx u
         }
         // D fragment: row qy = g, columns = lines t*8 + 2c, +1
         if (g < Q)
         {
#pragma unroll
            for (int h = 0; h < 2; h++)
            {
               const int ls = t * 8 + 2 * c + h;
               if (ls < NY)
               {
                  const int ezs = ls / Q, qxs = ls - ezs * Q;
                  const int es = ezs / D1, zs = ezs - es * D1;
                  double *o = G3 + zs * PZ + es * QQ + g * Q + qxs;
                  o[0] = h ? gb1 : gb0;
                  o[PA] = h ? bg1 : bg0;
                  o[2 * PA] = h ? bb1 : bb0;
               }
            }
         }
      }
#pragma unroll
      for (int k = 0; k < KF2; k++)
      {
         const int t = w + k * NW;
         if (t < S::TB_F)
         {
            const int line = t * 8 + g;                     // (e, f, qa)
            const int ef = line / Q, qa = line - ef * Q;
            double y0 = 0.0, y1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KF; ks++)
            {
               const int kk = ks * 4 + c;
               const double x = (line < NT2 && kk < D1) ? F1[(ef * D1 + kk) * RQ + qa] : 0.0;
               dmma884(y0, y1, x, fB[ks]);                  // [line][qb = 2c, 2c+1]
            }
            y0 *= df[k][0]; y1 *= df[k][1];
            double z0 = 0.0, z1 = 0.0;
            dmma884(z0, z1, y0, cC1);                       // [line][ib = 2c, 2c+1]
            dmma884(z0, z1, y1, cC2);
            __syncwarp();                                   // all reads of the tile precede its writes
            if (line < NT2)
            {
               if (2 * c < D1) { F1[(ef * D1 + 2 * c) * RQ + qa] = z0; }
               if (2 * c + 1 < D1) { F1[(ef * D1 + 2 * c + 1) * RQ + qa] = z1; }
            }
         }
      }
      __syncthreads();
      // ================= C: z-stage (rows = columns (e,qy,qx), k = iz; chained back) | face back-a
#pragma unroll
      for (int k = 0; k < KC; k++)
      {
         const int t = w + k * NW;
         if (t < S::TC_V)
         {
            const int col = t * 8 + g;
            double g00 = 0.0, g01 = 0.0, g10 = 0.0, g11 = 0.0, g20 = 0.0, g21 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KF; ks++)
            {
               const int kk = ks * 4 + c;
               const bool on = (col < NC) && (kk < D1);
               const double x0 = on ? G3[kk * PZ + col] : 0.0;
               const double x1 = on ? G3[PA + kk * PZ + col] : 0.0;
               const double x2 = on ? G3[2 * PA + kk * PZ + col] : 0.0;
               dmma884(g00, g01, x0, fB[ks]);               // d/dx: Bz (By Gx u)
               dmma884(g10, g11, x1, fB[ks]);               // d/dy: Bz (Gy Bx u)
               dmma884(g20, g21, x2, fG[ks]);               // d/dz: Gz (By Bx u)
            }
            // D . grad u at (column, qz = 2c, 2c+1)
            const double s0 = dv[k][0] * g00 + dv[k][1] * g10 + dv[k][2] * g20;
            const double s1 = dv[k][3] * g01 + dv[k][4] * g11 + dv[k][5] * g21;
            double z0 = 0.0, z1 = 0.0;
            dmma884(z0, z1, s0, cC1);                       // [column][iz = 2c, 2c+1]
            dmma884(z0, z1, s1, cC2);
            __syncwarp();
            if (col < NC)
            {
               if (2 * c < D1) { G3[(2 * c) * PZ + col] = z0; }
               if (2 * c + 1 < D1) { G3[(2 * c + 1) * PZ + col] = z1; }
            }
         }
      }
      for (int t = w; t < S::TA_F; t += NW)
      {
         const int line = t * 8 + g;                        // (e, f, ib)
         double y0 = 0.0, y1 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KB; ks++)
         {
            const int kk = ks * 4 + c;
            const double x = (line < NT1 && kk < Q) ? F1[line * RQ + kk] : 0.0;
            dmma884(y0, y1, x, bC[ks]);                     // [line][ia = 2c, 2c+1]
         }
         if (line < NT1)
         {
            if (2 * c < D1) { FD[line * D1 + 2 * c] = y0; }
            if (2 * c + 1 < D1) { FD[line * D1 + 2 * c + 1] = y1; }
         }
      }
      __syncthreads();
      // ================= D: bwd-y (rows = iy, cols = lines (e,iz,qx), k = qy)
      for (int t = w; t < S::TB_V; t += NW)
      {
         const int line = t * 8 + g;
         const int eiz = line / Q, qx = line - eiz * Q;
         const int e = eiz / D1, iz = eiz - e * D1;
         double y0 = 0.0, y1 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KB; ks++)
         {
            const int kk = ks * 4 + c;
            const double x = (line < NY && kk < Q) ? G3[iz * PZ + e * QQ + kk * Q + qx] : 0.0;
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
                  const int eizs = ls / Q, qxs = ls - eizs * Q;
                  S2[(eizs * D1 + g) * RQ + qxs] = h ? y1 : y0;
               }
            }
         }
      }
      __syncthreads();
      // ================= E: bwd-x (rows = lines (e,iz,iy), k = qx) + face combine -> X
      for (int t = w; t < S::TA_V; t += NW)
      {
         const int line = t * 8 + g;
         double r0 = 0.0, r1 = 0.0;
#pragma unroll
         for (int ks = 0; ks < KB; ks++)
         {
            const int kk = ks * 4 + c;
            const double x = (line < NL && kk < Q) ? S2[line * RQ + kk] : 0.0;
            dmma884(r0, r1, x, bC[ks]);                     // [line][ix = 2c, 2c+1]
         }
         if (line < NL && i0 < D1)
         {
            const int e = line / (D1 * D1), r = line - e * D1 * D1, bb_ = r / D1, aa_ = r - bb_ * D1;
            const double *fc = FD + e * NF * NFD;
            // faces: 0 z=0 (x,y)  1 y=0 (x,z)  2 x=p (y,z)  3 y=p (x,z)  4 x=0 (y,z)  5 z=p (x,y)
            const double fx0 = fc[4 * NFD + bb_ * D1 + aa_], fx1 = fc[2 * NFD + bb_ * D1 + aa_];
            const double my0 = MV[2 * aa_], my1 = MV[2 * aa_ + 1];
            const double mz0 = MV[2 * bb_], mz1 = MV[2 * bb_ + 1];
            const double sc = dat[S::P_E + e];
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
      __syncthreads();
      // ================= element-wise part: one warp per element (see stage3p.cuh)
      if (w < ne)
      {
         const int64_t ge = e0 + w;
         const double *X0 = dat + S::P_X + w * ND;
         const double *BD = dat + S::P_B + w * N3 * 2;
         const double dt = a.dt;
         const double inv_m = dat[S::P_E + w] * (double)ND, m = 1.0 / inv_m, mdt = m * inv_dt;
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
               u[k] = U[w * ND + j];
               du_ho[k] = X[w * ND + j];
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
               a.out[ge * ND + j] = o;
               omin = fmin(omin, o); omax = fmax(omax, o);
            }
         }
         if (a.xe_min_out)
         {
            warp_minmax(omin, omax);
            if (lane == 0) { a.xe_min_out[ge] = omin; a.xe_max_out[ge] = omax; }
         }
      }
   }
}

} // namespace rmh

#endif
