// Problem definitions and driver set-up helpers (host, OpenMP).
//
// velocity / u0 / inflow restate remhos.cpp velocity_function (:2001-2120), u0_function
// (:2201-2355) and inflow_function (:2363-2381); the set-up helpers restate the CFL time step
// (:538-553), the remap mesh-velocity integration (:562-584) and ProjectCoefficient on the
// positive basis (:883, sampling at the uniform lattice).
#include "../../include/remhos_b200.h"
#include "common.hpp"

#include <cmath>
#include <cstring>
#include <vector>

namespace rmh
{
struct Mesh;
}

// accessors implemented in mesh.cpp
extern "C" int rmh_mesh_dim(const rmh_mesh *m);
extern "C" int rmh_mesh_ne(const rmh_mesh *m);
extern "C" int rmh_mesh_geom_order(const rmh_mesh *m);
extern "C" const double *rmh_mesh_nodes(const rmh_mesh *m);

using namespace rmh;

static inline void normalise(int dim, const double *x, const double *bmin, const double *bmax,
                             double *X)
{
   for (int i = 0; i < dim; i++)
   {
      const double center = (bmin[i] + bmax[i]) * 0.5;
      X[i] = 2 * (x[i] - center) / (bmax[i] - bmin[i]);
   }
}

static int velocity_pt(int problem, int dim, const double *x, const double *bmin,
                       const double *bmax, double *v)
{
   double X[3] = {0, 0, 0};
   normalise(dim, x, bmin, bmax, X);
   for (int i = 0; i < dim; i++) { v[i] = 0.0; }
   switch (problem % 20)
   {
      case 0:
         if (dim == 1) { v[0] = 1.0; }
         else if (dim == 2) { v[0] = sqrt(2. / 3.); v[1] = sqrt(1. / 3.); }
         else { v[0] = sqrt(3. / 6.); v[1] = sqrt(2. / 6.); v[2] = sqrt(1. / 6.); }
         return 0;
      case 1: case 2: case 4:
      {
         const double w = M_PI / 2;
         if (dim == 1) { v[0] = 1.0; }
         else { v[0] = -w * X[1]; v[1] = w * X[0]; }
         return 0;
      }
      case 3:
      {
         const double w = M_PI / 2;
         double d = fmax((X[0] + 1.) * (1. - X[0]), 0.) * fmax((X[1] + 1.) * (1. - X[1]), 0.);
         d = d * d;
         if (dim == 1) { v[0] = 1.0; }
         else { v[0] = d * w * X[1]; v[1] = -d * w * X[0]; }
         return 0;
      }
      case 5:
         for (int i = 0; i < dim; i++) { v[i] = 1.0; }
         return 0;
      case 6: case 7:
         if (dim == 1) { v[0] = 1.0; }
         else { v[0] = x[1]; v[1] = -x[0]; }
         return 0;
      case 11:
      {
         const double r = sqrt(x[0] * x[0] + x[1] * x[1]);
         if (r < 0.2) { v[0] = 5.0 * x[1]; v[1] = -5.0 * x[0]; }
         else if (r < 0.4)
         {
            v[0] = 2.0 * x[1] / r - 5.0 * x[1];
            v[1] = -2.0 * x[0] / r + 5.0 * x[0];
         }
         return 0;
      }
      case 10: case 12: case 13: case 14: case 15: case 16: case 17:
      {
         double Y[3];
         for (int d = 0; d < dim; d++) { Y[d] = X[d] * 0.5 + 0.5; }
         v[0] = sin(M_PI * Y[0]) * cos(M_PI * Y[1]);
         v[1] = -cos(M_PI * Y[0]) * sin(M_PI * Y[1]);
         if (dim == 3)
         {
            v[0] *= cos(M_PI * Y[2]);
            v[1] *= cos(M_PI * Y[2]);
            v[2] = 0.0;
         }
         return 0;
      }
   }
   return 1;
}

static double box2(double x1, double y1, double x2, double y2, double theta, double ox, double oy,
                   double x, double y)
{
   const double s = sin(theta * M_PI / 180), c = cos(theta * M_PI / 180);
   const double xn = c * (x - ox) - s * (y - oy) + ox, yn = s * (x - ox) + c * (y - oy) + oy;
   return (xn > x1 && xn < x2 && yn > y1 && yn < y2) ? 1.0 : 0.0;
}
static double box3(double xmin, double xmax, double ymin, double ymax, double zmin, double zmax,
                   double theta, double ox, double oy, double x, double y, double z)
{
   const double s = sin(theta * M_PI / 180), c = cos(theta * M_PI / 180);
   const double xn = c * (x - ox) - s * (y - oy) + ox, yn = s * (x - ox) + c * (y - oy) + oy;
   return (xn > xmin && xn < xmax && yn > ymin && yn < ymax && z > zmin && z < zmax) ? 1.0 : 0.0;
}
static double cross_u(double a, double b) { return a + b - a * b; }
static double ring(int dim, double rin, double rout, const double *c, const double *y)
{
   double r = 0.0;
   for (int i = 0; i < dim; i++) { r += (y[i] - c[i]) * (y[i] - c[i]); }
   r = sqrt(r);
   return (r > rin && r < rout) ? 1.0 : 0.0;
}

static double u0_pt(int problem, int dim, const double *x, const double *bmin, const double *bmax)
{
   double X[3] = {0, 0, 0};
   normalise(dim, x, bmin, bmax, X);
   switch (problem % 10)
   {
      case 0: case 1:
      {
         if (dim == 1) { return exp(-40. * pow(X[0] - 0.5, 2)); }
         double rx = 0.45, ry = 0.25;
         const double cx = 0., cy = -0.2, w = 10.;
         if (dim == 3)
         {
            const double s = (1. + 0.25 * cos(2 * M_PI * X[2]));
            rx *= s; ry *= s;
         }
         return (erfc(w * (X[0] - cx - rx)) * erfc(-w * (X[0] - cx + rx)) *
                 erfc(w * (X[1] - cy - ry)) * erfc(-w * (X[1] - cy + ry))) / 16;
      }
      case 2:
      {
         const double rho = hypot(X[0], X[1]), phi = atan2(X[1], X[0]);
         return pow(sin(M_PI * rho), 2) * sin(3 * phi);
      }
      case 3: return .5 * (sin(M_PI * X[0]) * sin(M_PI * X[1]) + 1.);
      case 4:
      {
         const double scale = 0.0225, coef = (0.5 / sqrt(scale));
         const bool slit = (X[0] <= -0.05) || (X[0] >= 0.05) || (X[1] >= 0.7);
         const double cone = coef * sqrt(pow(X[0], 2.) + pow(X[1] + 0.5, 2.));
         const double hump = coef * sqrt(pow(X[0] + 0.5, 2.) + pow(X[1], 2.));
         // operator precedence of the reference's ?: kept literally (remhos.cpp:2258-2261)
         return (slit && ((pow(X[0], 2.) + pow(X[1] - .5, 2.)) <= 4. * scale)) ? 1. : 0.
                + (1. - cone) * (pow(X[0], 2.) + pow(X[1] + .5, 2.) <= 4. * scale)
                + .25 * (1. + cos(M_PI * hump))
                * ((pow(X[0] + .5, 2.) + pow(X[1], 2.)) <= 4. * scale);
      }
      case 5:
      {
         double y[3] = {0, 0, 0};
         for (int i = 0; i < dim; i++) { y[i] = 50. * (x[i] + 1.); }
         if (dim == 2)
         {
            const double r1 = box2(14., 3., 17., 26., -45., 15.5, 11.5, y[0], y[1]);
            const double r2 = box2(7., 10., 32., 13., -45., 15.5, 11.5, y[0], y[1]);
            const double c1[2] = {40., 40.}, c2[2] = {40., 20.};
            return cross_u(r1, r2) + ring(2, 7., 10., c1, y) + ring(2, 3., 7., c2, y);
         }
         if (dim == 3)
         {
            double r1 = box3(7., 32., 10., 13., 10., 13., -45., 15.5, 11.5, y[0], y[1], y[2]);
            double r2 = box3(14., 17., 3., 26., 10., 13., -45., 15.5, 11.5, y[0], y[1], y[2]);
            double r3 = box3(14., 17., 10., 13., 3., 26., -45., 15.5, 11.5, y[0], y[1], y[2]);
            double cr = cross_u(cross_u(r1, r2), r3);
            const double c1[3] = {40., 40., 40.}, c2[3] = {40., 20., 20.};
            const double dom2 = cr + ring(3, 7., 10., c1, y) + ring(3, 3., 7., c2, y);
            r1 = box3(2., 27., 30., 33., 30., 33., 0., 0., 0., y[0], y[1], y[2]);
            r2 = box3(9., 12., 23., 46., 30., 33., 0., 0., 0., y[0], y[1], y[2]);
            r3 = box3(9., 12., 30., 33., 23., 46., 0., 0., 0., y[0], y[1], y[2]);
            cr = cross_u(cross_u(r1, r2), r3);
            const double dom3 = cr + ring(3, 0., 7., c1, y) + ring(3, 0., 3., c2, y) +
                                ring(3, 7., 10., c2, y);
            const double dom1 = 1. - cross_u(dom2, dom3);
            return dom1 + 2. * dom2 + 3. * dom3;
         }
         return 0.0;
      }
      case 6:
      {
         double r = 0.0;
         for (int i = 0; i < dim; i++) { r += x[i] * x[i]; }
         r = sqrt(r);
         if (r >= 0.15 && r < 0.45) { return 1.; }
         if (r >= 0.55 && r < 0.85) { return pow(cos(10. * M_PI * (r - 0.7) / 3.), 2.); }
         return 0.;
      }
      case 7:
      {
         double r = 0.0;
         for (int i = 0; i < dim; i++) { r += x[i] * x[i]; }
         r = sqrt(r);
         const double a = 0.5, b = 3.e-2, c = 0.1;
         return 0.25 * (1. + tanh((r + c - a) / b)) * (1. - tanh((r - c - a) / b));
      }
   }
   return 0.0;
}

static double inflow_pt(int problem, int dim, const double *x)
{
   double r = 0.0;
   for (int i = 0; i < dim; i++) { r += x[i] * x[i]; }
   r = sqrt(r);
   if ((problem % 10) == 6 && dim == 2)
   {
      if (r >= 0.15 && r < 0.45) { return 1.; }
      if (r >= 0.55 && r < 0.85) { return pow(cos(10. * M_PI * (r - 0.7) / 3.), 2.); }
      return 0.;
   }
   if ((problem % 10) == 7)
   {
      const double a = 0.5, b = 3.e-2, c = 0.1;
      return 0.25 * (1. + tanh((r + c - a) / b)) * (1. - tanh((r - c - a) / b));
   }
   return 0.0;
}

extern "C" int rmh_gauss_legendre_01(int n, double *x, double *w)
{
   if (n < 1) { rmh::set_error("rmh_gauss_legendre_01: n >= 1"); return 1; }
   std::vector<double> xv, wv;
   rmh::gauss_legendre_01(n, xv, wv);
   for (int i = 0; i < n; i++) { if (x) { x[i] = xv[i]; } if (w) { w[i] = wv[i]; } }
   return 0;
}

extern "C" int rmh_velocity(int problem, int dim, int64_t n, const double *x, const double *bmin,
                            const double *bmax, double *v)
{
   int bad = 0;
#pragma omp parallel for reduction(| : bad)
   for (int64_t i = 0; i < n; i++) { bad |= velocity_pt(problem, dim, x + i * dim, bmin, bmax, v + i * dim); }
   if (bad) { set_error("velocity: unknown problem number"); return 1; }
   return 0;
}

extern "C" int rmh_u0(int problem, int dim, int64_t n, const double *x, const double *bmin,
                      const double *bmax, double *u)
{
#pragma omp parallel for
   for (int64_t i = 0; i < n; i++) { u[i] = u0_pt(problem, dim, x + i * dim, bmin, bmax); }
   return 0;
}

extern "C" int rmh_inflow(int problem, int dim, int64_t n, const double *x, double *u)
{
#pragma omp parallel for
   for (int64_t i = 0; i < n; i++) { u[i] = inflow_pt(problem, dim, x + i * dim); }
   return 0;
}

// physical coordinates of a tensor lattice of reference points (pts1d[npts] per axis; on a
// face: the fixed axis takes the value `side`), out [ne][npts^d' ][dim], first axis fastest
extern "C" int rmh_mesh_eval(const rmh_mesh *m, int npts, const double *pts1d, int face,
                             double *out)
{
   const int dim = rmh_mesh_dim(m), g = rmh_mesh_geom_order(m), n1 = g + 1;
   const int64_t ne = rmh_mesh_ne(m);
   const double *X = rmh_mesh_nodes(m);
   const std::vector<double> gll = gauss_lobatto_01(n1);
   const std::vector<double> P(pts1d, pts1d + npts);
   const std::vector<double> L = lagrange(gll, P);
   int axis = -1, side = 0;
   if (face >= 0) { face_axis(dim, face, axis, side); }
   const std::vector<double> ends = {(double)side};
   const std::vector<double> Ls = lagrange(gll, ends);
   int nn = 1, np = 1;
   for (int a = 0; a < dim; a++) { nn *= n1; if (a != axis) { np *= npts; } }
#pragma omp parallel for
   for (int64_t e = 0; e < ne; e++)
   {
      const double *Xe = X + (size_t)e * nn * dim;
      for (int q = 0; q < np; q++)
      {
         const double *l[3];
         int mq = q;
         for (int a = 0; a < dim; a++)
         {
            if (a == axis) { l[a] = Ls.data(); }
            else { l[a] = L.data() + (size_t)(mq % npts) * n1; mq /= npts; }
         }
         double x[3] = {0, 0, 0};
         for (int n = 0; n < nn; n++)
         {
            int mn = n;
            double w = 1.0;
            for (int a = 0; a < dim; a++) { w *= l[a][mn % n1]; mn /= n1; }
            for (int i = 0; i < dim; i++) { x[i] += w * Xe[n * dim + i]; }
         }
         for (int i = 0; i < dim; i++) { out[((size_t)e * np + q) * dim + i] = x[i]; }
      }
   }
   return 0;
}

// CFL time step estimate (remhos.cpp:538-553): min_e 0.25 * |det J(center)|^(1/dim) / |v(center)|
// inflow_gf (remhos.cpp:625-636): the inflow function in the order-p Bernstein DG space,
// [ne][(p+1)^dim].  ProjectCoefficient on a positive basis samples at the lattice points i/p;
// problem 7 ("Convergence test: use high order projection") interpolates at the tensor
// Gauss-Legendre points instead and takes the Bernstein coefficients of that polynomial.
extern "C" int rmh_inflow_project(const rmh_mesh *m, int problem, int order, double *infl_out)
{
   const int dim = rmh_mesh_dim(m), n1 = order + 1;
   const int64_t ne = rmh_mesh_ne(m);
   int nd = 1;
   for (int a = 0; a < dim; a++) { nd *= n1; }
   std::vector<double> pts1(n1), w1(n1);
   const bool ho = (problem == 7);
   if (ho) { gauss_legendre_01(n1, pts1, w1); }
   else { for (int i = 0; i < n1; i++) { pts1[i] = order > 0 ? (double)i / order : 0.5; } }
   std::vector<double> x((size_t)ne * nd * dim), v((size_t)ne * nd);
   if (rmh_mesh_eval(m, n1, pts1.data(), -1, x.data())) { return 1; }
   rmh_inflow(problem, dim, ne * nd, x.data(), v.data());
   if (!ho) { std::copy(v.begin(), v.end(), infl_out); return 0; }
   // coefficients = (Bernstein values at the Gauss points)^-1 applied along every axis
   const std::vector<double> V = bernstein(order, pts1);          // [point][basis]
   const std::vector<double> Vinv = invert_small(V, n1);          // [basis][point]
#pragma omp parallel for
   for (int64_t e = 0; e < ne; e++)
   {
      std::vector<double> a(v.begin() + e * nd, v.begin() + (e + 1) * nd), b(nd);
      int stride = 1;
      for (int ax = 0; ax < dim; ax++)
      {
         for (int i = 0; i < nd; i++)
         {
            const int k = (i / stride) % n1, base = i - k * stride;
            double s = 0.0;
            for (int g = 0; g < n1; g++) { s += Vinv[k * n1 + g] * a[base + g * stride]; }
            b[i] = s;
         }
         a.swap(b);
         stride *= n1;
      }
      std::copy(a.begin(), a.end(), infl_out + e * nd);
   }
   return 0;
}

// Mesh::GetElementSize(e) (type 0): |det J at the element centre|^(1/dim), as used by the CFL
// estimate (remhos.cpp:544) and by MonoRDSolver's scale (remhos_mono.cpp:55)
extern "C" int rmh_mesh_elem_sizes(const rmh_mesh *m, double *h_out)
{
   const int dim = rmh_mesh_dim(m), g = rmh_mesh_geom_order(m), n1 = g + 1;
   const int64_t ne = rmh_mesh_ne(m);
   const double *X = rmh_mesh_nodes(m);
   const std::vector<double> gll = gauss_lobatto_01(n1);
   const std::vector<double> c = {0.5};
   const std::vector<double> L = lagrange(gll, c), dL = lagrange_deriv(gll, c);
   int nn = 1;
   for (int a = 0; a < dim; a++) { nn *= n1; }
#pragma omp parallel for
   for (int64_t e = 0; e < ne; e++)
   {
      const double *Xe = X + (size_t)e * nn * dim;
      double J[3][3] = {{0}};
      for (int n = 0; n < nn; n++)
      {
         int idx[3] = {0, 0, 0}, mn = n;
         for (int a = 0; a < dim; a++) { idx[a] = mn % n1; mn /= n1; }
         for (int j = 0; j < dim; j++)
         {
            double d = 1.0;
            for (int a = 0; a < dim; a++) { d *= (a == j) ? dL[idx[a]] : L[idx[a]]; }
            for (int i = 0; i < dim; i++) { J[i][j] += d * Xe[n * dim + i]; }
         }
      }
      double det;
      if (dim == 2) { det = J[0][0] * J[1][1] - J[0][1] * J[1][0]; }
      else
      {
         det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
               J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
               J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
      }
      h_out[e] = pow(fabs(det), 1.0 / dim);
   }
   return 0;
}

extern "C" int rmh_cfl_dt(const rmh_mesh *m, int problem, const double *bmin, const double *bmax,
                          double *dt_out)
{
   const int dim = rmh_mesh_dim(m), g = rmh_mesh_geom_order(m), n1 = g + 1;
   const int64_t ne = rmh_mesh_ne(m);
   const double *X = rmh_mesh_nodes(m);
   const std::vector<double> gll = gauss_lobatto_01(n1);
   const std::vector<double> c = {0.5};
   const std::vector<double> L = lagrange(gll, c), dL = lagrange_deriv(gll, c);
   int nn = 1;
   for (int a = 0; a < dim; a++) { nn *= n1; }
   double dt = INFINITY;
   int bad = 0;
#pragma omp parallel for reduction(min : dt) reduction(| : bad)
   for (int64_t e = 0; e < ne; e++)
   {
      const double *Xe = X + (size_t)e * nn * dim;
      double J[3][3] = {{0}}, xc[3] = {0, 0, 0};
      for (int n = 0; n < nn; n++)
      {
         int idx[3] = {0, 0, 0}, mn = n;
         for (int a = 0; a < dim; a++) { idx[a] = mn % n1; mn /= n1; }
         double lv = 1.0;
         for (int a = 0; a < dim; a++) { lv *= L[idx[a]]; }
         for (int i = 0; i < dim; i++) { xc[i] += lv * Xe[n * dim + i]; }
         for (int j = 0; j < dim; j++)
         {
            double d = 1.0;
            for (int a = 0; a < dim; a++) { d *= (a == j) ? dL[idx[a]] : L[idx[a]]; }
            for (int i = 0; i < dim; i++) { J[i][j] += d * Xe[n * dim + i]; }
         }
      }
      double det;
      if (dim == 2) { det = J[0][0] * J[1][1] - J[0][1] * J[1][0]; }
      else
      {
         det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
               J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
               J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
      }
      const double length = pow(fabs(det), 1.0 / dim);
      double v[3] = {0, 0, 0};
      bad |= velocity_pt(problem, dim, xc, bmin, bmax, v);
      double s2 = 1e-14;
      for (int i = 0; i < dim; i++) { s2 += v[i] * v[i]; }
      dt = fmin(dt, 0.25 * length / sqrt(s2));
   }
   if (bad) { set_error("cfl_dt: unknown problem number"); return 1; }
   *dt_out = dt;
   return 0;
}

// Remap mesh velocity (remhos.cpp:562-584): forward-Euler integration of the node positions,
// including the literal min(dt, t_final - t) evaluated after t += dt; v_nodes = x_final - x0.
extern "C" int rmh_remap_mesh_velocity(const rmh_mesh *m, int problem, const double *bmin,
                                       const double *bmax, double dt, double t_final,
                                       double *v_nodes)
{
   const int dim = rmh_mesh_dim(m), g = rmh_mesh_geom_order(m);
   const int64_t ne = rmh_mesh_ne(m);
   int nn = 1;
   for (int a = 0; a < dim; a++) { nn *= (g + 1); }
   const int64_t np = ne * nn;
   const double *X0 = rmh_mesh_nodes(m);
   std::vector<double> x(X0, X0 + (size_t)np * dim), v((size_t)np * dim);
   if (rmh_velocity(problem, dim, np, x.data(), bmin, bmax, v.data())) { return 1; }
   double t = 0.0;
   while (t < t_final)
   {
      t += dt;
      const double h = std::fmin(dt, t_final - t);
#pragma omp parallel for
      for (int64_t i = 0; i < np * dim; i++) { x[i] = x[i] + h * v[i]; }
      if (rmh_velocity(problem, dim, np, x.data(), bmin, bmax, v.data())) { return 1; }
   }
   for (int64_t i = 0; i < np * dim; i++) { v_nodes[i] = x[i] - X0[i]; }
   return 0;
}
