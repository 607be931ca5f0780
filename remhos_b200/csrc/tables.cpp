// 1-D finite-element tables and error state for remhos_b200 (host only).
#include "common.hpp"
#include "../../include/remhos_b200.h"

#include <cmath>
#include <cstdlib>

namespace rmh
{

static thread_local std::string g_error;

void set_error(const std::string &msg) { g_error = msg; }

void gauss_legendre_01(int n, std::vector<double> &x, std::vector<double> &w)
{
   x.assign(n, 0.0); w.assign(n, 0.0);
   for (int i = 0; i < (n + 1) / 2; i++)
   {
      double z = std::cos(M_PI * (i + 0.75) / (n + 0.5));
      double pp = 1.0;
      for (int it = 0; it < 100; it++)
      {
         double p0 = 1.0, p1 = z;
         for (int k = 2; k <= n; k++)
         {
            const double pk = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k;
            p0 = p1; p1 = pk;
         }
         if (n == 1) { p0 = 1.0; p1 = z; }
         pp = n * (z * p1 - p0) / (z * z - 1.0);
         const double dz = p1 / pp;
         z -= dz;
         if (std::fabs(dz) < 1e-16) { break; }
      }
      // recompute derivative at the converged root
      {
         double p0 = 1.0, p1 = z;
         for (int k = 2; k <= n; k++)
         {
            const double pk = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k;
            p0 = p1; p1 = pk;
         }
         pp = n * (z * p1 - p0) / (z * z - 1.0);
      }
      const double wt = 2.0 / ((1.0 - z * z) * pp * pp);
      // z is the i-th largest root on [-1,1]
      x[n - 1 - i] = 0.5 * (1.0 + z); w[n - 1 - i] = 0.5 * wt;
      x[i] = 0.5 * (1.0 - z);         w[i] = 0.5 * wt;
   }
   if (n % 2) { x[n / 2] = 0.5; }
}

std::vector<double> lagrange_deriv(const std::vector<double> &nodes, const std::vector<double> &x)
{
   const int n = (int)nodes.size(), nq = (int)x.size();
   std::vector<double> D((size_t)nq * n, 0.0);
   for (int q = 0; q < nq; q++)
      for (int i = 0; i < n; i++)
      {
         double s = 0.0;
         for (int m = 0; m < n; m++)
         {
            if (m == i) { continue; }
            double t = 1.0 / (nodes[i] - nodes[m]);
            for (int j = 0; j < n; j++)
               if (j != i && j != m) { t *= (x[q] - nodes[j]) / (nodes[i] - nodes[j]); }
            s += t;
         }
         D[(size_t)q * n + i] = s;
      }
   return D;
}

static double binom(int n, int k)
{
   double r = 1.0;
   for (int i = 1; i <= k; i++) { r = r * (n - k + i) / i; }
   return r;
}

std::vector<double> bernstein(int p, const std::vector<double> &x)
{
   const int n = p + 1, nq = (int)x.size();
   std::vector<double> B((size_t)nq * n);
   for (int q = 0; q < nq; q++)
      for (int i = 0; i <= p; i++)
      { B[(size_t)q * n + i] = binom(p, i) * std::pow(x[q], i) * std::pow(1.0 - x[q], p - i); }
   return B;
}

std::vector<double> bernstein_deriv(int p, const std::vector<double> &x)
{
   const int n = p + 1, nq = (int)x.size();
   std::vector<double> G((size_t)nq * n, 0.0);
   if (p == 0) { return G; }
   const std::vector<double> Bm = bernstein(p - 1, x);
   for (int q = 0; q < nq; q++)
      for (int i = 0; i <= p; i++)
      {
         const double lo = (i >= 1) ? Bm[(size_t)q * p + i - 1] : 0.0;
         const double hi = (i <= p - 1) ? Bm[(size_t)q * p + i] : 0.0;
         G[(size_t)q * n + i] = p * (lo - hi);
      }
   return G;
}

std::vector<double> invert_small(const std::vector<double> &A, int n)
{
   std::vector<long double> a((size_t)n * 2 * n, 0.0L);
   for (int i = 0; i < n; i++)
   {
      for (int j = 0; j < n; j++) { a[(size_t)i * 2 * n + j] = A[(size_t)i * n + j]; }
      a[(size_t)i * 2 * n + n + i] = 1.0L;
   }
   for (int c = 0; c < n; c++)
   {
      int piv = c;
      for (int r = c + 1; r < n; r++)
         if (fabsl(a[(size_t)r * 2 * n + c]) > fabsl(a[(size_t)piv * 2 * n + c])) { piv = r; }
      if (piv != c)
         for (int j = 0; j < 2 * n; j++) { std::swap(a[(size_t)c * 2 * n + j], a[(size_t)piv * 2 * n + j]); }
      const long double d = a[(size_t)c * 2 * n + c];
      for (int j = 0; j < 2 * n; j++) { a[(size_t)c * 2 * n + j] /= d; }
      for (int r = 0; r < n; r++)
      {
         if (r == c) { continue; }
         const long double f = a[(size_t)r * 2 * n + c];
         if (f == 0.0L) { continue; }
         for (int j = 0; j < 2 * n; j++) { a[(size_t)r * 2 * n + j] -= f * a[(size_t)c * 2 * n + j]; }
      }
   }
   std::vector<double> inv((size_t)n * n);
   for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) { inv[(size_t)i * n + j] = (double)a[(size_t)i * 2 * n + n + j]; }
   return inv;
}

} // namespace rmh

extern "C" const char *rmh_last_error(void) { return rmh::g_error.c_str(); }
extern "C" int rmh_version(void) { return 100; }
