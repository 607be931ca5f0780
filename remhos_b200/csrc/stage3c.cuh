// Constant-coefficient RK-stage kernel for 3D meshes whose elements are all affine AND carry a
// velocity that is constant over each element (sm_100a, FP64).  Detected at set-up by
// k_op_linear (ctx.cu); the headline transport configuration (periodic cube, problem 0) is of
// this kind.
//
// With adj(J) v = a (a 3-vector per element) the stored quadrature data is w_q a, and the
// sum-factorised operator collapses algebraically (tensor-product quadrature factorises exactly,
// M = vol (M1 x M1 x M1), C = M1^-1 B^T, C W B = I):
//
//     M^-1 K u = 1/vol * sum_axis [ -a_axis (I x I x T) u                      T = C W G   (D1 x D1)
//                                   + Minv[:,0] min(0,-a_axis) (u|lo - u_nbr|lo)   face at  0
//                                   + Minv[:,p] min(0,+a_axis) (u|hi - u_nbr|hi) ] face at  1
//
// i.e. along every grid line of the element the update is one (D1 x (D1+2)) matrix applied to the
// line's DOFs extended by the two neighbour trace values.  No quadrature-space intermediates, 18
// instead of ~580 FMAs per DOF at order 3: what is left is a streaming kernel (y, x0, out, indices:
// ~30 B/DOF of compulsory HBM traffic) whose element-wise tail (MassBasedAvg, bounds gather,
// ClipScale, RK combination, element min/max of the output) is the one of stage3w.cuh.
//
// Work split (v2; v1 spent 850 warp instructions per element, 72 % issue-slot utilisation):
//   * a warp owns E = 32 / LG consecutive elements at a time (LG = D1^2 rounded up to a power of
//     two: two elements at order 3), a lane owns one x-row of one element: the row's D1 DOFs stay
//     in registers from the x-contraction to the final store (16-byte row loads / stores, x0 and
//     out never touch shared memory, reductions run over LG lanes for E elements at once);
//   * the y- and z-contractions run line-wise (lane = one y-line, then one z-line, all lanes busy)
//     and hand their results to the row owners through two shared-memory arrays;
//   * inputs arrive through an NST-deep cp.async ring per warp (DOF blocks in a padded layout,
//     neighbour traces gathered through (neighbour element, orientation pattern), entity (min,max)
//     pairs, (a, 1/vol)); gather indices run NST-1 groups further ahead; x0 rows are prefetched
//     into registers one group ahead.  No block barrier in the loop.
// Shared-memory layout at order 3 (tools/bank_sim_c.py): plane stride 18, element stride 72 doubles:
// row (128-bit) and z-line accesses are conflict-free, y-lines use their own lane map (planes
// {0,2} / {1,3} per half-warp) which makes them conflict-free too.
//
// FOLD (v14): the overlap bounds (DofInfo::ComputeOverlapBounds, remhos_tools.cpp:432-495) are formed
// in the kernel from the (min,max) pairs of the 3x3x3 neighbourhood elements (rmh_nbr_lattice) by a
// separable min/max filter in registers (fold_bounds) -- the separate entity pass (k_ent_min_max +
// k_xe_interleave, 15 % of a stage in round 1) and its 2 x 113 MB entity array are gone; element
// (min,max) pairs ping-pong between two arrays (the kernel reads its neighbours' input pairs while
// other warps already write output pairs).
// GH + flags (multi-GPU): owned elements are ordered interior first; ghost neighbour traces and
// ghost (min,max) pairs are written into this rank's window by the peers' k_halo_put (dist.cuh);
// a warp polls the peers' epoch flags once, before it fetches its first shell group -- one launch
// per stage, the exchange hidden behind the interior elements.
#ifndef RMH_STAGE3C_CUH
#define RMH_STAGE3C_CUH

#include "stage3w.cuh"

namespace rmh
{

template <int D1>
struct TabC
{
   double T[D1][D1];        // C W G: mass-inverse-folded 1-D convection matrix
   double M0[D1], Mp[D1];   // Minv[:, 0], Minv[:, p]
};

template <int D1, int NST>
struct SmemC
{
   static constexpr int ND = D1 * D1 * D1, NF = 6, NFD = D1 * D1, N3 = 27, NL = D1 * D1;
   static constexpr int LG = NL <= 4 ? 4 : (NL <= 8 ? 8 : (NL <= 16 ? 16 : 32));   // lanes per element
   static constexpr int E = 32 / LG;                                                // elements per warp
   static constexpr bool V2 = (D1 % 2 == 0);                // rows move as 16-byte words
   // padded DOF block
   static constexpr int RS = D1, SZ = D1 * RS + (V2 ? 2 : 1), EL = (D1 * SZ + 1) & ~1;
   static constexpr int NEL = NF * NFD + 4;                 // neighbour traces of one element (+pad)
   static constexpr int BEL = N3 * 2;                       // (min,max) pairs of one element
   __device__ static __forceinline__ int posU(int z, int y, int x) { return z * SZ + y * RS + x; }
   // one data stage (doubles)
   static constexpr int P_U = 0;
   static constexpr int P_N = P_U + E * EL;
   static constexpr int P_B = P_N + ((E * NEL + 1) & ~1);
   static constexpr int P_A = P_B + E * BEL;                // a_x a_y a_z 1/vol per element
   static constexpr int PSZ = P_A + E * 4;
   static constexpr int OFF_X = NST * PSZ;                  // y- and z-line results [2][E * EL]
   static constexpr int WDBL = OFF_X + 2 * E * EL;
   // I_B64 (ghost-aware kernels): 64-bit source address of every neighbour face, filled from NE / NP once the
   // indices have landed
   static constexpr int I_NE = 0, I_NP = E * NF, I_BI = 2 * E * NF, I_B64 = (2 * E * NF + E * N3 + 1) & ~1;
   static constexpr int ISZ = (I_B64 + 2 * E * NF + 3) & ~3;
   static constexpr int WINT = NST * ISZ;
   static constexpr int WBYTES = ((WDBL * 8 + WINT * 4) + 15) & ~15;
   static constexpr int PATMAX = 16;
   static constexpr int CBYTES = ((PATMAX + 1) * NFD * 2 + 15) & ~15;     // + the identity pattern (row PATMAX)
   static constexpr size_t bytes(int nw) { return (size_t)CBYTES + (size_t)nw * WBYTES; }
};

template <int N>
__device__ __forceinline__ void cp_async_wait_group()
{
   asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// cp.async with 32-bit shared-window destinations (one generic->shared conversion per ring slot)
__device__ __forceinline__ void cps16(unsigned s, const void *g)
{
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cps8(unsigned s, const void *g)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cps4(unsigned s, const void *g)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(g) : "memory");
}
// 8-byte copy, or 8 bytes of zeros when nbytes == 0 (domain boundary: exterior state 0)
__device__ __forceinline__ void cps8z(unsigned s, const void *g, unsigned nbytes)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(g), "r"(nbytes) : "memory");
}
// min / max without the NaN fix-up of fmin / fmax (3 instead of ~7 instructions on sm_100a; the
// operands here are never NaN in a valid run, and a NaN still propagates into the output)
__device__ __forceinline__ double dmin(double x, double y) { return x < y ? x : y; }
__device__ __forceinline__ double dmax(double x, double y) { return x > y ? x : y; }

// gather indices of the group's nv elements (consecutive elements: contiguous index blocks)
template <int D1, int NST>
__device__ __forceinline__ void stagec_fetch_idx(const StagePArgs &a, unsigned six, int64_t e0, int nv, int lane)
{
   using S = SmemC<D1, NST>;
   constexpr int NF = S::NF, N3 = S::N3, E = S::E;
   const int nb = (a.bounds_type == 0) ? N3 : NF;
   if (nv == E && ((E * nb) & 1) == 0)
   {
      // full group: 8-byte copies (E*NF and E*nb are even, so every block starts 8-byte aligned)
      if (lane < E * NF / 2)
      {
         cps8(six + (S::I_NE + 2 * lane) * 4, a.fn.nbr_elem + e0 * NF + 2 * lane);
         cps8(six + (S::I_NP + 2 * lane) * 4, a.nbr_pat32 + e0 * NF + 2 * lane);
      }
#pragma unroll
      for (int i0 = 0; i0 < E * N3 / 2; i0 += 32)
      {
         const int i = i0 + lane;
         if (i < E * nb / 2) { cps8(six + (S::I_BI + 2 * i) * 4, a.bidx + e0 * nb + 2 * i); }
      }
   }
   else
   {
#pragma unroll
      for (int i0 = 0; i0 < E * NF; i0 += 32)
      {
         const int i = i0 + lane;
         if (i < nv * NF)
         {
            cps4(six + (S::I_NE + i) * 4, a.fn.nbr_elem + e0 * NF + i);
            cps4(six + (S::I_NP + i) * 4, a.nbr_pat32 + e0 * NF + i);
         }
      }
#pragma unroll
      for (int i0 = 0; i0 < E * N3; i0 += 32)
      {
         const int i = i0 + lane;
         if (i < nv * nb) { cps4(six + (S::I_BI + i) * 4, a.bidx + e0 * nb + i); }
      }
   }
}

// lane = (el, row): its own row of y; then the group's neighbour traces, bounds and coefficients.
// GH: some neighbours are ghost elements (multi-GPU): nbr_elem = ne_owned + slot, the neighbour's
// face trace in this element's natural face order is a.fn.ughost[slot][NFD].
template <int D1, int NST, bool GH>
__device__ __forceinline__ void stagec_fetch_data(const StagePArgs &a, double *dst, const int *ix,
                                                  const int16_t *spat, int64_t e0, int nv, int lane,
                                                  bool row_on, int row_src, int row_dst)
{
   using S = SmemC<D1, NST>;
   constexpr int ND = S::ND, NF = S::NF, NFD = S::NFD, N3 = S::N3, E = S::E;
   const unsigned sd = (unsigned)__cvta_generic_to_shared(dst);
   if (S::V2)
   {
      // 16-byte chunks in memory order: one instruction moves 512 contiguous bytes (4 cache lines;
      // a lane copying its own row would touch every 32-byte sector twice, 14 instead of 4
      // shared-memory wavefronts per instruction)
      constexpr int CPR = D1 / 2;                       // chunks per row
      const double *g = a.y + e0 * ND;
#pragma unroll
      for (int c0 = 0; c0 < E * ND / 2; c0 += 32)
      {
         const int ch = c0 + lane;
         const int row = ch / CPR, h = ch - row * CPR;  // row = el * NL + (iz * D1 + iy)
         const int el = row / S::NL, r = row - el * S::NL;
         if (ch < nv * (ND / 2))
         {
            cps16(sd + (S::P_U + el * S::EL + (r / D1) * S::SZ + (r % D1) * S::RS + 2 * h) * 8, g + 2 * ch);
         }
      }
   }
   else if (row_on)
   {
      const double *g = a.y + e0 * ND + row_src;
      const unsigned su = sd + (S::P_U + row_dst) * 8;
#pragma unroll
      for (int i = 0; i < D1; i++) { cps8(su + i * 8, g + i); }
   }
   {
      const int *NE_ = ix + S::I_NE, *NP_ = ix + S::I_NP;
      if (GH)
      {
         // per FACE, once: where its trace comes from -- the neighbour's block of y (its pattern row), the
         // ghost trace array (already in this element's face order: identity row) or, at a domain boundary,
         // a block of zeros -- so that the per-value loop below is free of case distinctions
         long long *B64 = reinterpret_cast<long long *>(const_cast<int *>(ix) + S::I_B64);
         int *NPw = const_cast<int *>(ix) + S::I_NP;
         const int ne_own = (int)a.fn.ne_owned;
#pragma unroll
         for (int f0 = 0; f0 < E * NF; f0 += 32)
         {
            const int f = f0 + lane;
            if (f < nv * NF)
            {
               const int nb = NE_[f], pid = NP_[f];
               const bool gh = nb >= ne_own;
               const double *p = (nb < 0) ? a.zeros
                                 : (gh ? a.fn.ughost + (int64_t)(nb - ne_own) * NFD : a.y + (int64_t)nb * ND);
               B64[f] = (long long)p;
               NPw[f] = ((nb < 0 || gh) ? S::PATMAX : pid) * NFD;
            }
         }
         __syncwarp();
      }
      // 32 % NFD == 0 (orders 1 and 3): face and face-DOF index of a lane differ by constants per slot
      constexpr bool P2 = (32 % NFD == 0);
      const int lq = lane / NFD, lj = lane % NFD;
#pragma unroll
      for (int i0 = 0; i0 < E * NF * NFD; i0 += 32)
      {
         int F, j;
         if (P2) { F = i0 / NFD + lq; j = lj; }
         else { const int id = i0 + lane; F = id / NFD; j = id - F * NFD; }       // F = el * NF + f
         if (F < nv * NF)
         {
            const int el = F / NF;
            const unsigned d = sd + (S::P_N + F * NFD + el * (S::NEL - NF * NFD) + j) * 8;
            if (GH)
            {
               const long long *B64 = reinterpret_cast<const long long *>(ix + S::I_B64);
               const int loc = spat[NP_[F] + j];
               cps8(d, reinterpret_cast<const double *>(B64[F]) + loc);
            }
            else
            {
               const int nb = NE_[F];
               const int pid = NP_[F];
               const int loc = spat[pid * NFD + j];
               const double *src = a.y + (unsigned)((nb < 0 ? 0 : nb) * ND + loc);
               cps8z(d, src, nb < 0 ? 0u : 8u);
            }
         }
      }
   }
   if (lane < 2 * nv) { cps16(sd + (S::P_A + 2 * lane) * 8, a.opa + e0 * 4 + 2 * lane); }
   {
      const int *BI = ix + S::I_BI;
      if (a.bounds_type == 0)
      {
#pragma unroll
         for (int i0 = 0; i0 < E * N3; i0 += 32)
         {
            const int i = i0 + lane;
            if (i < nv * N3) { cps16(sd + (S::P_B + 2 * i) * 8, a.ent_mm + 2 * (int64_t)BI[i]); }
         }
      }
      else
      {
         double *BD = dst + S::P_B;
#pragma unroll
         for (int i0 = 0; i0 < E * (NF + 1); i0 += 32)
         {
            const int i = i0 + lane;
            if (i < nv * (NF + 1))
            {
               const int el = i / (NF + 1), k = i - el * (NF + 1);
               const int64_t src = (k == NF) ? e0 + el : (int64_t)BI[el * NF + k];
               double *d = BD + el * S::BEL + 2 * k;
               if (src >= 0) { cp_async8(d, a.xe_min + src); cp_async8(d + 1, a.xe_max + src); }
               else { d[0] = INFINITY; d[1] = -INFINITY; }
            }
         }
      }
   }
}

// Separable min/max filter over the 3x3x3 neighbourhood pairs BD[el][27] (index dx + 3 dy + 9 dz,
// (min,max) as double2), in registers: lane L < 27 holds pair L; along each axis in turn the two
// outer values absorb the middle one -- (v0 ^ v1, v1, v1 ^ v2) -- through one shuffle from the middle
// lane of the line.  After the x-, y- and z-pass entry (cx,cy,cz) is the (min,max) over the elements
// sharing lattice entity (cx,cy,cz) of the element (c = 0 | 1 | 2: low face, interior, high face
// along that axis).  No shared-memory round trips and no barrier between the passes, so the chain
// overlaps with the line contractions that follow it.
template <int E, int BEL>
__device__ __forceinline__ void fold_bounds(double *BD, int lane)
{
   const int dx = lane % 3, dy = (lane / 3) % 3, dz = lane / 9;
   const int sx = lane - dx + 1, sy = lane - 3 * dy + 3, sz = lane - 9 * dz + 9;   // middle lane of the line
   const bool ex = (dx != 1), ey = (dy != 1), ez = (dz != 1);
   const int ld = lane < 27 ? lane : 26;
#pragma unroll
   for (int el = 0; el < E; el++)
   {
      double2 *p = reinterpret_cast<double2 *>(BD + el * BEL);
      double2 v = p[ld];
      double mn = __shfl_sync(0xffffffffu, v.x, sx), mx = __shfl_sync(0xffffffffu, v.y, sx);
      if (ex) { v.x = mn < v.x ? mn : v.x; v.y = mx > v.y ? mx : v.y; }
      mn = __shfl_sync(0xffffffffu, v.x, sy); mx = __shfl_sync(0xffffffffu, v.y, sy);
      if (ey) { v.x = mn < v.x ? mn : v.x; v.y = mx > v.y ? mx : v.y; }
      mn = __shfl_sync(0xffffffffu, v.x, sz); mx = __shfl_sync(0xffffffffu, v.y, sz);
      if (ez) { v.x = mn < v.x ? mn : v.x; v.y = mx > v.y ? mx : v.y; }
      if (lane < 27) { p[lane] = v; }
   }
}

// halo wait statistics (rmh_halo_wait_stats): [0] warps that found a flag not yet published, [1] their
// summed and [2] longest wait in ns (only the slow path touches them); phase clocks of the ghost-aware
// kernels, summed over the warps (differences are exact modulo 2^64): [3] time at the first shell group,
// [4] end of those warps, [5] their number, [6] start, [7] end and [8] number of all warps.
__device__ unsigned long long g_halo_wait[9];

__device__ __forceinline__ unsigned long long gtimer()
{
   unsigned long long t;
   asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
   return t;
}

// every lane waits until all n peers have published epoch (k_halo_put's release store, system scope)
__device__ __forceinline__ void wait_peer_flags(const unsigned long long *flags, int n, unsigned long long epoch)
{
   unsigned long long t0 = 0;
   bool spun = false;
   for (int p = 0; p < n; p++)
   {
      unsigned long long v;
      unsigned int spins = 0;
      while (true)
      {
         asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(flags + p) : "memory");
         if (v >= epoch) { break; }
         if (!spun) { spun = true; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); }
         __nanosleep(256);
         // a peer that died must not hang this GPU: give up after ~20 s and fail the launch loudly
         if (++spins > (1u << 26)) { __trap(); }
      }
   }
   if (spun && (threadIdx.x & 31) == 0)
   {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      atomicAdd(&g_halo_wait[0], 1ull); atomicAdd(&g_halo_wait[1], t1 - t0); atomicMax(&g_halo_wait[2], t1 - t0);
   }
}

template <int LG>
__device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
   for (int o = LG / 2; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
   return v;
}

template <int D1, int NW, int MINB, int NST, bool GH, bool FOLD, bool SEND>
__global__ void __launch_bounds__(NW * 32, MINB)
k_stage3c(StagePArgs a, const TabC<D1> tab)
{
   using S = SmemC<D1, NST>;
   constexpr int ND = S::ND, NF = S::NF, NFD = S::NFD, NL = S::NL, LG = S::LG, E = S::E;
   constexpr int RS = S::RS, SZ = S::SZ, EL = S::EL, NEL = S::NEL, BEL = S::BEL;
   static_assert(NST >= 2 && NST <= 4, "ring depth");
   extern __shared__ double sm[];
   const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
   int16_t *spat = reinterpret_cast<int16_t *>(sm);
   double *wsm = reinterpret_cast<double *>(reinterpret_cast<char *>(sm) + S::CBYTES) + (size_t)w * (S::WBYTES / 8);
   int *ismem = reinterpret_cast<int *>(wsm + S::WDBL);
   const unsigned six0 = (unsigned)__cvta_generic_to_shared(ismem);
   const double inv_dt = 1.0 / a.dt, dt = a.dt;
   {
      const int np = a.npat < S::PATMAX ? a.npat : S::PATMAX;
      for (int i = threadIdx.x; i < np * NFD; i += NW * 32) { spat[i] = a.fn.pat[i]; }
      if (GH) { for (int i = threadIdx.x; i < NFD; i += NW * 32) { spat[S::PATMAX * NFD + i] = (int16_t)i; } }
   }
   // ---- lane roles
   // row owner: element el, row r = iz * D1 + iy
   const int el = lane / LG, r = lane - el * LG;
   const bool lane_on = r < NL;
   const int iy = lane_on ? r % D1 : 0, iz = lane_on ? r / D1 : 0;
   const int row_src = el * ND + (iz * D1 + iy) * D1;               // in the group's y / x0 / out block
   const int row_dst = el * EL + S::posU(iz, iy, 0);                // in the padded shared block
   // y-line owner: element ely, line (x = ya, z = yb).  Order 3: planes {0,2} | {1,3} per half-warp
   int ely = el, ya = iy, yb = iz;
   if (D1 == 4)
   {
      const int t = (lane >> 2) & 3;
      ely = t >> 1; ya = lane & 3; yb = 2 * (t & 1) + (lane >> 4);
   }
   const int yl_base = ely * EL + S::posU(yb, 0, ya);               // + k * RS
   const int yl_nb = ely * NEL + ya + D1 * yb;                      // + f * NFD, f = 1 | 3
   // z-line owner: element el, line (x = iy, y = iz) in row-owner numbering
   const int zl_base = el * EL + S::posU(0, iz, iy);                // + k * SZ
   const int zl_nb = el * NEL + iy + D1 * iz;                       // + f * NFD, f = 0 | 5
   // lattice class of the row (bounds): base + {0, 1, 2} for x = 0 | interior | p
   const int cy = (iy == 0) ? 0 : ((iy == D1 - 1) ? 2 : 1), cz = (iz == 0) ? 0 : ((iz == D1 - 1) ? 2 : 1);
   const int cbase = el * BEL + 2 * (3 * cy + 9 * cz);
   __syncthreads();     // the only block barrier: the pattern table is in place
   const int64_t NG = (a.ne + E - 1) / E;                           // element groups
   const int64_t GW = (int64_t)gridDim.x * NW;
   int64_t gi = (int64_t)blockIdx.x * NW + w;       // LOGICAL group index
   if (gi >= NG) { return; }
   if (GH) { if (a.flags != nullptr && lane == 0) { atomicAdd(&g_halo_wait[6], gtimer()); atomicAdd(&g_halo_wait[8], 1ull); } }
   const int last_nv = (int)(a.ne - (NG - 1) * E);
   auto nvalid = [&](int64_t g) { return (g + 1 == NG) ? last_nv : E; };
   // Fused halo send (multi-GPU, a.send): the shell groups [G0, NG) come FIRST, so that their face traces and
   // (min,max) pairs are on their way to the peers while the interior groups are still being computed:
   // logical group g is physical group (g + G0) mod NG.  Everything below that touches memory uses the
   // physical index.
   // (SEND is its own instantiation: the rotation defeats the strength reduction of every group base address)
   const int64_t G0 = SEND ? a.shell_begin / E : 0;
   unsigned int n_sent = 0;                      // shell groups this warp has sent
   auto ph = [&](int64_t g) { if (SEND) { g += G0; return g >= NG ? g - NG : g; } return g; };
   // multi-GPU: the halo of y must have landed before the first group at or behind shell_begin is fetched
   bool halo_ok = !(GH && a.flags != nullptr);
   auto need_halo = [&](int64_t g)
   {
      if (GH)
      {
         if (!halo_ok && g * E + E > a.shell_begin)
         {
            if (lane == 0) { atomicAdd(&g_halo_wait[3], gtimer()); atomicAdd(&g_halo_wait[5], 1ull); }
            wait_peer_flags(a.flags, a.n_wait, a.epoch);
            halo_ok = true;
         }
      }
   };
   // ---- prologue: indices of the first NST-1 groups, then their data and the next NST-1 index sets
#pragma unroll
   for (int m = 0; m < NST - 1; m++)
   {
      const int64_t g = gi + m * GW;
      if (g < NG) { const int64_t gp = ph(g); stagec_fetch_idx<D1, NST>(a, six0 + (m % NST) * S::ISZ * 4, gp * E, nvalid(gp), lane); }
   }
   cp_async_commit();
   cp_async_wait_all();
   __syncwarp();
#pragma unroll
   for (int m = 0; m < NST - 1; m++)
   {
      const int64_t g = gi + m * GW;
      if (g < NG)
      {
         const int64_t gp = ph(g);
         const int nv = nvalid(gp);
         need_halo(gp);
         stagec_fetch_data<D1, NST, GH>(a, wsm + (m % NST) * S::PSZ, ismem + (m % NST) * S::ISZ, spat, gp * E, nv, lane,
                                    lane_on && el < nv, row_src, row_dst);
      }
   }
   __syncwarp();
#pragma unroll
   for (int m = NST - 1; m < 2 * (NST - 1); m++)
   {
      const int64_t g = gi + m * GW;
      if (g < NG) { const int64_t gp = ph(g); stagec_fetch_idx<D1, NST>(a, six0 + (m % NST) * S::ISZ * 4, gp * E, nvalid(gp), lane); }
   }
   cp_async_commit();
#pragma unroll
   for (int m = 0; m < NST - 2; m++) { cp_async_commit(); }    // keep the group count of the steady state
   double *XY = wsm + S::OFF_X, *XZ = XY + E * EL;
   // x0 row of the first group
   double x0n[D1];
#pragma unroll
   for (int i = 0; i < D1; i++) { x0n[i] = 0.0; }
   auto load_x0 = [&](int64_t g)
   {
      if (a.has_x0 && lane_on && el < nvalid(g))
      {
         const double *p = a.x0 + g * (E * ND) + row_src;
         if (S::V2)
         {
#pragma unroll
            for (int i = 0; i < D1; i += 2)
            {
               const double2 v = __ldcs(reinterpret_cast<const double2 *>(p + i));
               x0n[i] = v.x; x0n[i + 1] = v.y;
            }
         }
         else
         {
#pragma unroll
            for (int i = 0; i < D1; i++) { x0n[i] = __ldcs(p + i); }
         }
      }
   };
   load_x0(ph(gi));
   int slot = 0;                 // it % NST
   for (; gi < NG; gi += GW)
   {
      double *dat = wsm + slot * S::PSZ;
      const double *U = dat + S::P_U, *NB = dat + S::P_N, *A = dat + S::P_A, *BD = dat + S::P_B;
      const int64_t gp = ph(gi);                 // physical group
      const int nv = nvalid(gp);
      const bool on = lane_on && el < nv;
      cp_async_wait_group<NST - 2>();
      __syncwarp();      // data(g) and idx(g + (NST-1) GW) landed; the previous group is fully consumed
      double x0[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { x0[i] = x0n[i]; }
      {
         const int s1 = (slot == 0) ? NST - 1 : slot - 1;      // (it + NST - 1) % NST
         const int s2 = (NST == 2) ? slot : ((slot >= 2) ? slot - 2 : slot + NST - 2);   // (it + 2 NST - 2) % NST
         const int64_t g1 = gi + (int64_t)(NST - 1) * GW, g2 = gi + (int64_t)(2 * (NST - 1)) * GW;
         if (g1 < NG)
         {
            const int64_t gp1 = ph(g1);
            const int nv1 = nvalid(gp1);
            need_halo(gp1);
            stagec_fetch_data<D1, NST, GH>(a, wsm + s1 * S::PSZ, ismem + s1 * S::ISZ, spat, gp1 * E, nv1, lane,
                                       lane_on && el < nv1, row_src, row_dst);
         }
         if (g2 < NG) { const int64_t gp2 = ph(g2); stagec_fetch_idx<D1, NST>(a, six0 + s2 * S::ISZ * 4, gp2 * E, nvalid(gp2), lane); }
         cp_async_commit();
         if (gi + GW < NG) { load_x0(ph(gi + GW)); }
      }
      if (FOLD) { fold_bounds<E, BEL>(dat + S::P_B, lane); }     // visible to the tail after the __syncwarp below
      // ================= y-lines and z-lines -> XY, XZ (every lane owns one line of each kind)
      //   out_i = sum_k c_k v_k - (Minv[i][0] vs_lo) nbr_lo - (Minv[i][p] vs_hi) nbr_hi,
      //   c = -a T[i][:], c_0 += Minv[i][0] vs_lo, c_p += Minv[i][p] vs_hi
      if (lane_on && ely < nv)
      {
         const double ay = A[ely * 4 + 1];
         const double vlo = dmin(0.0, -ay), vhi = dmin(0.0, ay);
         double v[D1];
#pragma unroll
         for (int k = 0; k < D1; k++) { v[k] = U[yl_base + k * RS]; }
         const double jl = vlo * (v[0] - NB[yl_nb + 1 * NFD]), jh = vhi * (v[D1 - 1] - NB[yl_nb + 3 * NFD]);
#pragma unroll
         for (int i = 0; i < D1; i++)
         {
            double sacc = 0.0;
#pragma unroll
            for (int k = 0; k < D1; k++) { sacc = fma(tab.T[i][k], v[k], sacc); }
            XY[yl_base + i * RS] = fma(tab.Mp[i], jh, fma(tab.M0[i], jl, -ay * sacc));
         }
      }
      if (on)
      {
         const double az = A[el * 4 + 2];
         const double vlo = dmin(0.0, -az), vhi = dmin(0.0, az);
         double v[D1];
#pragma unroll
         for (int k = 0; k < D1; k++) { v[k] = U[zl_base + k * SZ]; }
         const double jl = vlo * (v[0] - NB[zl_nb + 0 * NFD]), jh = vhi * (v[D1 - 1] - NB[zl_nb + 5 * NFD]);
#pragma unroll
         for (int i = 0; i < D1; i++)
         {
            double sacc = 0.0;
#pragma unroll
            for (int k = 0; k < D1; k++) { sacc = fma(tab.T[i][k], v[k], sacc); }
            XZ[zl_base + i * SZ] = fma(tab.Mp[i], jh, fma(tab.M0[i], jl, -az * sacc));
         }
      }
      // ================= x-row in registers
      double u[D1], ho[D1];
#pragma unroll
      for (int i = 0; i < D1; i++) { u[i] = 0.0; ho[i] = 0.0; }
      double sc = 0.0;
      if (on)
      {
         if (S::V2)
         {
#pragma unroll
            for (int i = 0; i < D1; i += 2)
            {
               const double2 t = *reinterpret_cast<const double2 *>(U + row_dst + i);
               u[i] = t.x; u[i + 1] = t.y;
            }
         }
         else
         {
#pragma unroll
            for (int i = 0; i < D1; i++) { u[i] = U[row_dst + i]; }
         }
         const double ax = A[el * 4 + 0];
         sc = A[el * 4 + 3];
         const double vlo = dmin(0.0, -ax), vhi = dmin(0.0, ax);
         const int xnb = el * NEL + iy + D1 * iz;
         const double jl = vlo * (u[0] - NB[xnb + 4 * NFD]), jh = vhi * (u[D1 - 1] - NB[xnb + 2 * NFD]);
#pragma unroll
         for (int i = 0; i < D1; i++)
         {
            double sacc = 0.0;
#pragma unroll
            for (int k = 0; k < D1; k++) { sacc = fma(tab.T[i][k], u[k], sacc); }
            ho[i] = fma(tab.Mp[i], jh, fma(tab.M0[i], jl, -ax * sacc));
         }
      }
      __syncwarp();      // XY, XZ complete
      // ================= element-wise tail on the row owner (MassBasedAvg, bounds, ClipScale, RK)
      {
         const double inv_m = sc * (double)ND, m = on ? 1.0 / inv_m : 0.0, mdt = m * inv_dt;
         double bmn[3], bmx[3];
         if (on)
         {
            if (S::V2)
            {
#pragma unroll
               for (int i = 0; i < D1; i += 2)
               {
                  const double2 ty = *reinterpret_cast<const double2 *>(XY + row_dst + i);
                  const double2 tz = *reinterpret_cast<const double2 *>(XZ + row_dst + i);
                  ho[i] = (ho[i] + ty.x + tz.x) * sc; ho[i + 1] = (ho[i + 1] + ty.y + tz.y) * sc;
               }
            }
            else
            {
#pragma unroll
               for (int i = 0; i < D1; i++) { ho[i] = (ho[i] + XY[row_dst + i] + XZ[row_dst + i]) * sc; }
            }
            if (a.bounds_type == 0)
            {
#pragma unroll
               for (int q = 0; q < 3; q++)
               {
                  const double2 t = *reinterpret_cast<const double2 *>(BD + cbase + 2 * q);
                  bmn[q] = t.x; bmx[q] = t.y;
               }
            }
            else
            {
               double bmin1 = INFINITY, bmax1 = -INFINITY;
#pragma unroll
               for (int k = 0; k <= NF; k++)
               {
                  bmin1 = dmin(bmin1, BD[el * BEL + 2 * k]); bmax1 = dmax(bmax1, BD[el * BEL + 2 * k + 1]);
               }
#pragma unroll
               for (int q = 0; q < 3; q++) { bmn[q] = bmin1; bmx[q] = bmax1; }
            }
         }
         else
         {
#pragma unroll
            for (int q = 0; q < 3; q++) { bmn[q] = 0.0; bmx[q] = 0.0; }
         }
         double s1 = 0.0;
#pragma unroll
         for (int i = 0; i < D1; i++) { s1 += u[i] + dt * ho[i]; }
         s1 = group_sum<LG>(s1);
         const double ubar = s1 * (1.0 / ND);                // MassBasedAvg, remhos_lo.cpp:278-285
         double lo[D1], fp[D1], fn[D1];
         double sumPos = 0.0, sumNeg = 0.0;
#pragma unroll
         for (int i = 0; i < D1; i++)
         {
            const int q = (i == 0) ? 0 : ((i == D1 - 1) ? 2 : 1);
            lo[i] = (ubar - u[i]) * inv_dt;
            const double u_new_lo = u[i] + dt * lo[i];
            const double fmn = mdt * (bmn[q] - u_new_lo);
            const double fmx = mdt * (bmx[q] - u_new_lo);
            double fcl = m * (ho[i] - lo[i]);
            fcl = dmin(fmx, dmax(fmn, fcl));               // ClipScale, remhos_fct.cpp:490-515
            fp[i] = dmax(fcl, 0.0);
            fn[i] = fcl - fp[i];                           // = fmin(fcl, 0), exactly
            sumNeg += fn[i];
            sumPos += fp[i];
         }
         sumNeg = group_sum<LG>(sumNeg);
         sumPos = group_sum<LG>(sumPos);
         const double new_mass = sumNeg + sumPos;
         constexpr double eps = 1.0e-15;
         const bool sp = new_mass > eps, sn = new_mass < -eps;
         // positive excess: scale the positive fluxes by -sumNeg/sumPos; negative excess: the converse
         const double rat = (sp || sn) ? -((sp ? sumNeg : sumPos) / (sp ? sumPos : sumNeg)) : 1.0;
         const double cpos = sp ? rat : 1.0, cneg = sn ? rat : 1.0;
         double o[D1];
         double omin = INFINITY, omax = -INFINITY;
#pragma unroll
         for (int i = 0; i < D1; i++)
         {
            const double fcl = fma(cneg, fn[i], cpos * fp[i]);      // one of fn, fp is zero: exact
            const double du = lo[i] + fcl * inv_m;
            double v = du;
            if (a.out_mode == 1)
            {
               const double base = a.has_x0 ? a.a * x0[i] : 0.0;
               v = base + a.b * (u[i] + dt * du);
            }
            o[i] = v;
            if (on) { omin = dmin(omin, v); omax = dmax(omax, v); }
         }
         if (on)
         {
            double *p = a.out + gp * (E * ND) + row_src;
            if (S::V2)
            {
#pragma unroll
               for (int i = 0; i < D1; i += 2) { *reinterpret_cast<double2 *>(p + i) = make_double2(o[i], o[i + 1]); }
            }
            else
            {
#pragma unroll
               for (int i = 0; i < D1; i++) { p[i] = o[i]; }
            }
         }
         if (FOLD || a.xe_min_out)
         {
#pragma unroll
            for (int s = LG / 2; s > 0; s >>= 1)
            {
               omin = dmin(omin, __shfl_xor_sync(0xffffffffu, omin, s));
               omax = dmax(omax, __shfl_xor_sync(0xffffffffu, omax, s));
            }
            if (r == 0 && el < nv)
            {
               if (FOLD) { a.xe_mm_out[gp * E + el] = make_double2(omin, omax); }
               else { a.xe_min_out[gp * E + el] = omin; a.xe_max_out[gp * E + el] = omax; }
            }
         }
         // ================= fused halo send: this group's output IS the next stage's input -- store the face
         // traces the peers asked for (receiver's face order) and the (min,max) pair of every ring element
         // straight into their windows; the last shell group of the grid publishes the next epoch
         if (SEND)
         {
            if (gp >= G0)
            {
               const StageSend &S_ = *a.send;
               if (on)
               {
                  const int64_t es = gp * E + el - a.shell_begin;            // shell-local element
                  const int2 *sf = S_.face + es * NF;
                  const uint8_t *sp = S_.perm + es * NF;
                  auto put = [&](int f, int j, double v)
                  {
                     const int2 ps = sf[f];
                     if (ps.x >= 0) { S_.peer[ps.x].gtr[a.send_par][(int64_t)ps.y * NFD + S_.rperm[(int)sp[f] * NFD + j]] = v; }
                  };
                  put(4, iy + D1 * iz, o[0]);
                  put(2, iy + D1 * iz, o[D1 - 1]);
#pragma unroll
                  for (int i = 0; i < D1; i++)
                  {
                     if (iy == 0) { put(1, i + D1 * iz, o[i]); }
                     if (iy == D1 - 1) { put(3, i + D1 * iz, o[i]); }
                     if (iz == 0) { put(0, i + D1 * iy, o[i]); }
                     if (iz == D1 - 1) { put(5, i + D1 * iy, o[i]); }
                  }
                  if (r == 0)
                  {
                     const double2 mm = make_double2(omin, omax);
                     for (int k = S_.mm_off[es]; k < S_.mm_off[es + 1]; k++)
                     {
                        const int2 pd = S_.mm[k];
                        S_.peer[pd.x].mm[a.send_par][pd.y] = mm;
                     }
                  }
               }
               // one system fence per WARP, behind its last shell group (they are its first groups): a fence
               // per group would stall the warp for an NVLink round trip every time
               n_sent++;
               __syncwarp();
               if (lane == 0 && gi + GW >= (int64_t)a.send_groups)
               {
                  __threadfence_system();
                  const unsigned int t = atomicAdd(S_.counter, n_sent);
                  if (t + n_sent == a.send_groups)
                  {
                     *S_.counter = 0;
                     __threadfence_system();
                     for (int q = 0; q < S_.npeers; q++)
                     {
                        asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(S_.peer[q].flag), "l"(a.send_epoch) : "memory");
                     }
                  }
               }
            }
         }
      }
      slot = (slot + 1 == NST) ? 0 : slot + 1;
   }
   if (GH)
   {
      if (a.flags != nullptr && lane == 0)
      {
         const unsigned long long t = gtimer();
         atomicAdd(&g_halo_wait[7], t);
         if (halo_ok) { atomicAdd(&g_halo_wait[4], t); }
      }
   }
}

} // namespace rmh

#endif
