// Product-field remap (-ps): the second pass of AdvectionOperator::LimitMult (remhos.cpp:1848-1915)
// with ComputeBoolIndicators / ComputeRatio / ZeroOutEmptyDofs (remhos_sync.cpp:24-114),
// FCTSolver::CalcCompatibleLOProduct / ScaleProductBounds (remhos_fct.cpp:26-153) and the three
// CalcFCTProduct variants (FluxBasedFCT :183-294, ClipScaleSolver :543-563, ElementFCTProjection
// :735-758).  Everything is element-local apart from the bounds on s = us / u, which reuse the
// bounds kernels with masked element min/max.  One warp per element; flags are bytes.
#ifndef RMH_PRODUCT_CUH
#define RMH_PRODUCT_CUH

namespace rmh
{

constexpr double EMPTY_ZONE_TOL = 1e-12;     // remhos_sync.hpp

// ComputeBoolIndicators (remhos_sync.cpp:24-47)
__global__ void k_bool_indicators(int64_t ne, int nd, const double *u, uint8_t *el, uint8_t *dof)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= ne) { return; }
   bool any = false;
   for (int j = lane; j < nd; j += 32)
   {
      const bool on = u[e * nd + j] > EMPTY_ZONE_TOL;
      dof[e * nd + j] = on ? 1 : 0;
      any = any || on;
   }
   any = __any_sync(0xffffffffu, any);
   if (lane == 0) { el[e] = any ? 1 : 0; }
}

// ComputeRatio (remhos_sync.cpp:50-94): s = us / u on the active dofs, their average on the other
// dofs of an active element, 0 in empty elements (s may be null) -- fused with the masked
// DofInfo::ComputeElementsMinMax (remhos_tools.cpp:497-523) over the active dofs (xe may be null)
__global__ void k_prod_ratio(int64_t ne, int nd, const double *us, const double *u, double *s, uint8_t *el,
                             uint8_t *dof, double *xe_min, double *xe_max)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= ne) { return; }
   double sum = 0.0, mn = INFINITY, mx = -INFINITY;
   int n = 0;
   for (int j = lane; j < nd; j += 32)
   {
      const double uj = u[e * nd + j];
      const bool on = uj > EMPTY_ZONE_TOL;
      dof[e * nd + j] = on ? 1 : 0;
      if (on)
      {
         const double r = us[e * nd + j] / uj;
         sum += r; n++;
         mn = fmin(mn, r); mx = fmax(mx, r);
      }
   }
   sum = warp_sum(sum);
   for (int o = 16; o > 0; o >>= 1) { n += __shfl_xor_sync(0xffffffffu, n, o); }
   mn = warp_min(mn); mx = warp_max(mx);
   if (lane == 0)
   {
      el[e] = n > 0 ? 1 : 0;
      if (xe_min) { xe_min[e] = mn; xe_max[e] = mx; }     // inactive element: (inf, -inf), no contribution
   }
   if (s)
   {
      const double avg = n > 0 ? sum / n : 0.0;
      for (int j = lane; j < nd; j += 32)
      {
         const double uj = u[e * nd + j];
         s[e * nd + j] = (n == 0) ? 0.0 : ((uj > EMPTY_ZONE_TOL) ? us[e * nd + j] / uj : avg);
      }
   }
}

// masked ComputeElementsMinMax alone (remhos_tools.cpp:497-523)
__global__ void k_elem_min_max_masked(int64_t ne, int nd, const double *u, const uint8_t *el, const uint8_t *dof,
                                      double *xe_min, double *xe_max)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= ne) { return; }
   double mn = INFINITY, mx = -INFINITY;
   if (!el || el[e])
   {
      for (int j = lane; j < nd; j += 32)
      {
         if (dof && !dof[e * nd + j]) { continue; }
         const double v = u[e * nd + j];
         mn = fmin(mn, v); mx = fmax(mx, v);
      }
   }
   mn = warp_min(mn); mx = warp_max(mx);
   if (lane == 0) { xe_min[e] = mn; xe_max[e] = mx; }
}

// FCTSolver::CalcCompatibleLOProduct (remhos_fct.cpp:26-118): s_min / s_max are adjusted in place,
// d_lo = (u_new s_avg - us) / dt in active elements, 0 elsewhere; followed -- us_min != null -- by
// ScaleProductBounds (:120-153)
__global__ void k_prod_compatible_lo(int64_t ne, int nd, double dt, const double *us, const double *m,
                                     const double *d_us_ho, double *s_min, double *s_max, const double *u_new,
                                     const uint8_t *act_el, const uint8_t *act_dof, double *d_lo, double *us_min,
                                     double *us_max)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= ne) { return; }
   const double eps = 1e-12;
   const bool ael = act_el[e] != 0;
   if (!ael)
   {
      for (int j = lane; j < nd; j += 32)
      {
         d_lo[e * nd + j] = 0.0;
         if (us_min) { us_min[e * nd + j] = 0.0; us_max[e * nd + j] = 0.0; }
      }
      return;
   }
   double mass_us = 0.0, mass_u = 0.0, smin = INFINITY, smax = -INFINITY;
   bool any = false;
   for (int j = lane; j < nd; j += 32)
   {
      const int64_t g = e * nd + j;
      mass_us += (us[g] + dt * d_us_ho[g]) * m[g];
      mass_u += u_new[g] * m[g];
      if (act_dof[g]) { any = true; smin = fmin(smin, s_min[g]); smax = fmax(smax, s_max[g]); }
   }
   warp_sum2(mass_us, mass_u);
   smin = warp_min(smin); smax = warp_max(smax);
   any = __any_sync(0xffffffffu, any);
   double s_avg = mass_us / mass_u;
   if (any)
   {
      // round-off fixes (:69-77): s_avg outside the stencil bounds only by the inflated round-off of the division
      if (s_avg < smin && mass_us + eps > smin * mass_u) { s_avg = smin; }
      if (s_avg > smax && mass_us - eps < smax * mass_u) { s_avg = smax; }
   }
   for (int j = lane; j < nd; j += 32)
   {
      const int64_t g = e * nd + j;
      double lo = s_min[g], hi = s_max[g];
      const bool on = act_dof[g] != 0;
      if (on)
      {
         if (s_avg + eps < lo) { lo = s_avg; s_min[g] = lo; }
         if (s_avg - eps > hi) { hi = s_avg; s_max[g] = hi; }
      }
      d_lo[g] = (u_new[g] * s_avg - us[g]) / dt;
      if (us_min)
      {
         us_min[g] = on ? lo * u_new[g] : 0.0;
         us_max[g] = on ? hi * u_new[g] : 0.0;
      }
   }
}

// ZeroOutEmptyDofs (remhos_sync.cpp:96-114)
__global__ void k_zero_empty(int64_t n, int nd, const uint8_t *el, const uint8_t *dof, double *v)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && !el[i / nd] && !dof[i]) { v[i] = 0.0; }
}

// FluxBasedFCT::CalcFCTProduct (remhos_fct.cpp:214-246): the element-local fluxes that turn the LO
// product rate into the compatible one: fel = m dt (d_us_LO - d_lo_c), beta = m u_new / sum (both 0
// in empty elements, so that beta_j fel_i - beta_i fel_j vanishes there)
__global__ void k_prod_flux_el(int64_t ne, int nd, double dt, const double *m, const double *d_us_lo,
                               const double *d_lo_c, const double *u_new, const uint8_t *act_el, double *fel,
                               double *beta)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= ne) { return; }
   const bool ael = act_el[e] != 0;
   double sum = 0.0;
   for (int j = lane; j < nd; j += 32) { sum += m[e * nd + j] * u_new[e * nd + j]; }
   sum = warp_sum(sum);
   for (int j = lane; j < nd; j += 32)
   {
      const int64_t g = e * nd + j;
      fel[g] = ael ? m[g] * dt * (d_us_lo[g] - d_lo_c[g]) : 0.0;
      beta[g] = ael ? (m[g] * u_new[g]) / sum : 0.0;
   }
}

// out = a + dt b
__global__ void k_axpy_out(int64_t n, double dt, const double *a, const double *b, double *out)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { out[i] = a[i] + dt * b[i]; }
}

// AdvectionOperator::ComputeMask (remhos.cpp:1741-1796): an element of the product state is updated
// with the full stage combination only if all of its dofs are active in u; mask[d * N + i] for both fields
__global__ void k_compute_mask(int64_t ne, int nd, const double *u, uint8_t *mask, int nfields)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= ne) { return; }
   bool all = true;
   for (int j = lane; j < nd; j += 32) { all = all && (u[e * nd + j] > EMPTY_ZONE_TOL); }
   all = __all_sync(0xffffffffu, all);
   for (int d = 0; d < nfields; d++)
      for (int j = lane; j < nd; j += 32) { mask[(int64_t)d * ne * nd + e * nd + j] = all ? 1 : 0; }
}

// RKIDPSolver::UpdateMask (remhos_solvers.cpp:127-147): mask &= mask_new
__global__ void k_mask_and(int64_t n, uint8_t *mask, const uint8_t *mask_new)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { mask[i] = (mask[i] && mask_new[i]) ? 1 : 0; }
}

// AddMasked (remhos_solvers.cpp:97-125): dst += a * src where mask is set
__global__ void k_add_masked(int64_t n, const uint8_t *mask, double a, const double *src, double *dst)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && mask[i]) { dst[i] += a * src[i]; }
}

} // namespace rmh

#endif
