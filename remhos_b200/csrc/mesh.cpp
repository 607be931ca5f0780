// Host mesh module of remhos_b200: mesh reader (MFEM mesh v1.0 / INLINE v1.0, quads + hexes),
// uniform refinement, nodal geometry, face-neighbour topology and the DofInfo index maps.
//
// Replaces what Remhos obtains from MFEM's Mesh/ParMesh at remhos.cpp:448-463 (load, refine),
// :457 (bounding box), :513 (SetCurvature) and what DofInfo builds at
// remhos_tools.cpp:356-379,525-734,1356-1431.  Own design: geometry is always an element-wise
// (L2-style) Gauss-Lobatto nodal field, which represents H1 and periodic meshes alike, and all
// topology is derived from sorted vertex tuples (no orientation tables).
#include "../../include/remhos_b200.h"
#include "common.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <sstream>
#include <string>
#include <map>
#include <vector>

namespace rmh
{

// local face -> (fixed axis, side); quad: S E N W (remhos_tools.cpp:1367-1376),
// hex: bottom south east north west top (remhos_tools.cpp:1086-1286)
static const int FACE_AXIS2[4][2] = {{1, 0}, {0, 1}, {1, 1}, {0, 0}};
static const int FACE_AXIS3[6][2] = {{2, 0}, {1, 0}, {0, 1}, {1, 1}, {0, 0}, {2, 1}};

void face_axis(int dim, int f, int &axis, int &side)
{
   if (dim == 2) { axis = FACE_AXIS2[f][0]; side = FACE_AXIS2[f][1]; }
   else { axis = FACE_AXIS3[f][0]; side = FACE_AXIS3[f][1]; }
}

// lexicographic corner ids of local face f, ordered in the face's natural parametrisation
// (remaining axes ascending, first fastest)
static void face_corners(int dim, int f, int *c)
{
   int axis, side;
   face_axis(dim, f, axis, side);
   const int nfc = 1 << (dim - 1);
   for (int t = 0; t < nfc; t++)
   {
      int cc[3] = {0, 0, 0};
      cc[axis] = side;
      int m = 0;
      for (int a = 0; a < dim; a++)
      {
         if (a == axis) { continue; }
         cc[a] = (t >> m) & 1;
         m++;
      }
      int id = 0;
      for (int a = 0; a < dim; a++) { id |= cc[a] << a; }
      c[t] = id;
   }
}

std::vector<double> gauss_lobatto_01(int n)
{
   std::vector<double> x(n);
   if (n == 1) { x[0] = 0.5; return x; }
   x[0] = 0.0; x[n - 1] = 1.0;
   // interior: roots of P'_{n-1} on [-1,1] by Newton from Chebyshev-Lobatto guesses
   const int N = n - 1;
   for (int i = 1; i < N; i++)
   {
      double z = -std::cos(M_PI * i / N);
      for (int it = 0; it < 100; it++)
      {
         // Legendre P_N(z) and derivatives
         double p0 = 1.0, p1 = z;
         for (int k = 2; k <= N; k++)
         {
            const double pk = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k;
            p0 = p1; p1 = pk;
         }
         const double dp = N * (z * p1 - p0) / (z * z - 1.0);          // P_N'
         const double d2p = (2 * z * dp - N * (N + 1) * p1) / (1.0 - z * z);   // P_N''
         const double dz = dp / d2p;
         z -= dz;
         if (std::fabs(dz) < 1e-16) { break; }
      }
      x[i] = 0.5 * (z + 1.0);
   }
   for (int i = 0; i < n / 2; i++)   // symmetrise
   {
      const double a = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
      x[i] = a; x[n - 1 - i] = 1.0 - a;
   }
   if (n % 2) { x[n / 2] = 0.5; }
   return x;
}

// L[q*n + i] = l_i(x_q)
std::vector<double> lagrange(const std::vector<double> &nodes, const std::vector<double> &x)
{
   const int n = (int)nodes.size(), nq = (int)x.size();
   std::vector<double> L((size_t)nq * n, 1.0);
   for (int q = 0; q < nq; q++)
      for (int i = 0; i < n; i++)
      {
         double v = 1.0;
         for (int j = 0; j < n; j++)
            if (j != i) { v *= (x[q] - nodes[j]) / (nodes[i] - nodes[j]); }
         L[(size_t)q * n + i] = v;
      }
   return L;
}

struct Mesh
{
   int dim = 0, g = 1;
   int64_t ne = 0, nv = 0;
   std::vector<int64_t> ev;   // [ne][2^dim] lexicographic corners
   std::vector<double> X;     // [ne][(g+1)^dim][dim]
   int nvert() const { return 1 << dim; }
   int npe() const { int n = 1; for (int a = 0; a < dim; a++) { n *= (g + 1); } return n; }
};

// ---------------------------------------------------------------------------- unique keys
struct Key4 { int64_t a, b, c, d; int64_t idx; };
static inline bool key_less(const Key4 &x, const Key4 &y)
{
   if (x.a != y.a) { return x.a < y.a; }
   if (x.b != y.b) { return x.b < y.b; }
   if (x.c != y.c) { return x.c < y.c; }
   return x.d < y.d;
}
static inline bool key_eq(const Key4 &x, const Key4 &y)
{ return x.a == y.a && x.b == y.b && x.c == y.c && x.d == y.d; }

// keys: row-sorted tuples (width <= 4, padded with -1). Returns ids per original index.
static int64_t unique_ids(std::vector<Key4> &keys, std::vector<int64_t> &ids)
{
   std::sort(keys.begin(), keys.end(), key_less);
   ids.assign(keys.size(), 0);
   int64_t cnt = 0;
   for (size_t i = 0; i < keys.size(); i++)
   {
      if (i > 0 && !key_eq(keys[i], keys[i - 1])) { cnt++; }
      ids[keys[i].idx] = cnt;
   }
   return keys.empty() ? 0 : cnt + 1;
}

static Key4 make_key(const int64_t *v, int w, int64_t idx)
{
   int64_t t[4] = {-1, -1, -1, -1};
   for (int i = 0; i < w; i++) { t[i] = v[i]; }
   std::sort(t, t + w);
   return Key4{t[0], t[1], t[2], t[3], idx};
}

// Entity ids on the 3^dim lattice of each element (H1 order-2 style numbering).
static int64_t macro_lattice(const Mesh &m, std::vector<int64_t> &lat)
{
   const int dim = m.dim, nvx = m.nvert();
   int n3 = 1;
   for (int a = 0; a < dim; a++) { n3 *= 3; }
   lat.assign((size_t)m.ne * n3, -1);
   int64_t offset = m.nv;
   // classify lattice positions by the number of parent corners
   std::vector<std::vector<int>> corners(n3);
   for (int t = 0; t < n3; t++)
   {
      int tt[3], q = t;
      for (int a = 0; a < dim; a++) { tt[a] = q % 3; q /= 3; }
      for (int c = 0; c < nvx; c++)
      {
         bool ok = true;
         for (int a = 0; a < dim; a++)
         {
            const int ca = (c >> a) & 1;
            if (!(tt[a] == 1 || tt[a] == 2 * ca)) { ok = false; }
         }
         if (ok) { corners[t].push_back(c); }
      }
   }
   for (int width = 1; width <= nvx; width *= 2)
   {
      std::vector<int> pos;
      for (int t = 0; t < n3; t++)
         if ((int)corners[t].size() == width) { pos.push_back(t); }
      if (pos.empty()) { continue; }
      if (width == 1)
      {
         for (int t : pos)
            for (int64_t e = 0; e < m.ne; e++)
            { lat[e * n3 + t] = m.ev[e * nvx + corners[t][0]]; }
      }
      else if (width == nvx)
      {
         for (int t : pos)
            for (int64_t e = 0; e < m.ne; e++) { lat[e * n3 + t] = offset + e; }
         offset += m.ne;
      }
      else
      {
         std::vector<Key4> keys;
         keys.reserve((size_t)m.ne * pos.size());
         int64_t v[4];
         for (size_t k = 0; k < pos.size(); k++)
            for (int64_t e = 0; e < m.ne; e++)
            {
               for (int i = 0; i < width; i++) { v[i] = m.ev[e * nvx + corners[pos[k]][i]]; }
               keys.push_back(make_key(v, width, (int64_t)(k * m.ne + e)));
            }
         std::vector<int64_t> ids;
         const int64_t cnt = unique_ids(keys, ids);
         for (size_t k = 0; k < pos.size(); k++)
            for (int64_t e = 0; e < m.ne; e++)
            { lat[e * n3 + pos[k]] = offset + ids[k * m.ne + e]; }
         offset += cnt;
      }
   }
   return offset;
}

// apply a 1-D matrix R[no][ni] along axis `a` of an array [n_after][ni][n_before][ncomp]
static void apply_axis(const std::vector<double> &in, std::vector<double> &out,
                       const double *R, int no, int ni, size_t n_before, size_t n_after)
{
   out.assign(n_after * no * n_before, 0.0);
   for (size_t z = 0; z < n_after; z++)
      for (int o = 0; o < no; o++)
         for (int i = 0; i < ni; i++)
         {
            const double r = R[o * ni + i];
            if (r == 0.0) { continue; }
            const double *src = &in[(z * ni + i) * n_before];
            double *dst = &out[(z * no + o) * n_before];
            for (size_t b = 0; b < n_before; b++) { dst[b] += r * src[b]; }
         }
}

// interpolate one element's nodal block [n1^dim][dim] with per-axis matrices R[a] (no x n1)
static void interp_block(const double *Xe, int dim, int n1, const double *const *R, int no,
                         std::vector<double> &tmpa, std::vector<double> &tmpb, double *out)
{
   size_t nin = 1;
   for (int a = 0; a < dim; a++) { nin *= n1; }
   tmpa.assign(Xe, Xe + nin * dim);
   // array layout: [z][y][x][comp]; axis a has n_before = dim * n^(a) (current sizes)
   size_t before = dim;
   for (int a = 0; a < dim; a++)
   {
      size_t after = 1;
      for (int b = a + 1; b < dim; b++) { after *= n1; }
      apply_axis(tmpa, tmpb, R[a], no, n1, before, after);
      tmpa.swap(tmpb);
      before *= no;
   }
   std::copy(tmpa.begin(), tmpa.end(), out);
}

static void refine_once(Mesh &m)
{
   const int dim = m.dim, nvx = m.nvert(), nch = nvx, g = m.g, n1 = g + 1;
   int n3 = 1;
   for (int a = 0; a < dim; a++) { n3 *= 3; }
   std::vector<int64_t> lat;
   const int64_t nvnew = macro_lattice(m, lat);
   Mesh r;
   r.dim = dim; r.g = g; r.ne = m.ne * nch; r.nv = nvnew;
   r.ev.resize((size_t)r.ne * nvx);
   const int npe = m.npe();
   r.X.resize((size_t)r.ne * npe * dim);
   const std::vector<double> gll = gauss_lobatto_01(n1);
   std::vector<double> h0(n1), h1(n1);
   for (int i = 0; i < n1; i++) { h0[i] = 0.5 * gll[i]; h1[i] = 0.5 + 0.5 * gll[i]; }
   const std::vector<double> R0 = lagrange(gll, h0), R1 = lagrange(gll, h1);
   std::vector<double> ta, tb;
   for (int64_t e = 0; e < m.ne; e++)
   {
      for (int ch = 0; ch < nch; ch++)
      {
         const int64_t ce = e * nch + ch;
         for (int c = 0; c < nvx; c++)
         {
            int t = 0, mul = 1;
            for (int a = 0; a < dim; a++)
            {
               t += (((ch >> a) & 1) + ((c >> a) & 1)) * mul;
               mul *= 3;
            }
            r.ev[ce * nvx + c] = lat[e * n3 + t];
         }
         const double *R[3];
         for (int a = 0; a < dim; a++) { R[a] = ((ch >> a) & 1) ? R1.data() : R0.data(); }
         interp_block(&m.X[(size_t)e * npe * dim], dim, n1, R, n1, ta, tb,
                      &r.X[(size_t)ce * npe * dim]);
      }
   }
   m = std::move(r);
}

static void set_curvature(Mesh &m, int order)
{
   if (order == m.g) { return; }
   const int dim = m.dim, n1 = m.g + 1, no = order + 1;
   const std::vector<double> src = gauss_lobatto_01(n1), dst = gauss_lobatto_01(no);
   const std::vector<double> I1 = lagrange(src, dst);
   const int npe = m.npe();
   int npo = 1;
   for (int a = 0; a < dim; a++) { npo *= no; }
   std::vector<double> Xn((size_t)m.ne * npo * dim), ta, tb;
   const double *R[3] = {I1.data(), I1.data(), I1.data()};
   for (int64_t e = 0; e < m.ne; e++)
   {
      interp_block(&m.X[(size_t)e * npe * dim], dim, n1, R, no, ta, tb,
                   &Xn[(size_t)e * npo * dim]);
   }
   m.X.swap(Xn);
   m.g = order;
}

// ------------------------------------------------------------------------------ topology
struct Topology
{
   std::vector<int64_t> nbr_elem, nbr_face;   // [ne][nf]
   std::vector<int8_t> fmap;                  // [ne][nf][nfc]
};

static int build_topology(const Mesh &m, Topology &T)
{
   const int dim = m.dim, nvx = m.nvert(), nf = 2 * dim, nfc = 1 << (dim - 1);
   int fc[6][4];
   for (int f = 0; f < nf; f++) { face_corners(dim, f, fc[f]); }
   std::vector<Key4> keys((size_t)m.ne * nf);
   int64_t v[4];
   for (int64_t e = 0; e < m.ne; e++)
      for (int f = 0; f < nf; f++)
      {
         for (int t = 0; t < nfc; t++) { v[t] = m.ev[e * nvx + fc[f][t]]; }
         keys[e * nf + f] = make_key(v, nfc, e * nf + f);
      }
   std::sort(keys.begin(), keys.end(), key_less);
   T.nbr_elem.assign((size_t)m.ne * nf, -1);
   T.nbr_face.assign((size_t)m.ne * nf, -1);
   T.fmap.assign((size_t)m.ne * nf * nfc, -1);
   for (size_t i = 0; i < keys.size();)
   {
      size_t j = i + 1;
      while (j < keys.size() && key_eq(keys[j], keys[i])) { j++; }
      if (j - i > 2)
      {
         set_error("mesh: a face is shared by more than two elements "
                   "(non-manifold or too-coarse periodic mesh)");
         return 1;
      }
      if (j - i == 2)
      {
         const int64_t a = keys[i].idx, b = keys[i + 1].idx;
         T.nbr_elem[a] = b / nf; T.nbr_face[a] = b % nf;
         T.nbr_elem[b] = a / nf; T.nbr_face[b] = a % nf;
      }
      i = j;
   }
   for (int64_t e = 0; e < m.ne; e++)
      for (int f = 0; f < nf; f++)
      {
         const int64_t e2 = T.nbr_elem[e * nf + f];
         if (e2 < 0) { continue; }
         const int f2 = (int)T.nbr_face[e * nf + f];
         for (int t = 0; t < nfc; t++)
         {
            const int64_t vid = m.ev[e * nvx + fc[f][t]];
            int found = -1, cnt = 0;
            for (int s = 0; s < nfc; s++)
               if (m.ev[e2 * nvx + fc[f2][s]] == vid) { found = s; cnt++; }
            if (cnt != 1)
            {
               set_error("mesh: degenerate face (repeated vertex ids); periodic mesh too coarse");
               return 1;
            }
            T.fmap[(e * nf + f) * nfc + t] = (int8_t)found;
         }
      }
   return 0;
}

// BdrDofs[nfd][nf] (ExtractBdrDofs, remhos_tools.cpp:1356-1431)
void bdr_dofs(int p, int dim, std::vector<int> &bd)
{
   const int n = p + 1;
   if (dim == 2)
   {
      bd.assign((size_t)n * 4, 0);
      for (int i = 0; i <= p; i++)
      {
         bd[i * 4 + 0] = i;
         bd[i * 4 + 1] = i * n + p;
         bd[i * 4 + 2] = n * n - 1 - i;
         bd[i * 4 + 3] = (p - i) * n;
      }
      return;
   }
   const int nfd = n * n;
   bd.assign((size_t)nfd * 6, 0);
   for (int b = 0; b < n; b++)
      for (int a = 0; a < n; a++)
      {
         const int j = a + n * b;
         bd[j * 6 + 0] = a + n * b;                    // z = 0 : (x, y)
         bd[j * 6 + 1] = a + n * n * b;                // y = 0 : (x, z)
         bd[j * 6 + 2] = p + n * a + n * n * b;        // x = p : (y, z)
         bd[j * 6 + 3] = a + n * p + n * n * b;        // y = p : (x, z)
         bd[j * 6 + 4] = n * a + n * n * b;            // x = 0 : (y, z)
         bd[j * 6 + 5] = a + n * b + n * n * p;        // z = p : (x, y)
      }
}

void nat2ref_table(int p, int dim, std::vector<int> &n2r)
{
   const int n = p + 1, nf = 2 * dim;
   int nfd = 1;
   for (int a = 0; a < dim - 1; a++) { nfd *= n; }
   std::vector<int> bd;
   bdr_dofs(p, dim, bd);
   n2r.assign((size_t)nf * nfd, -1);
   for (int f = 0; f < nf; f++)
   {
      int axis, side;
      face_axis(dim, f, axis, side);
      for (int j = 0; j < nfd; j++)
      {
         int l[3] = {0, 0, 0}, m = j;
         for (int a = 0; a < dim; a++)
         {
            if (a == axis) { l[a] = side * p; }
            else { l[a] = m % n; m /= n; }
         }
         const int dof = l[0] + n * (l[1] + n * l[2]);
         for (int r = 0; r < nfd; r++) { if (bd[r * nf + f] == dof) { n2r[f * nfd + j] = r; } }
      }
   }
}

void sub2ind(int p, int dim, std::vector<int> &s)
{
   const int n = p + 1;
   int ns = 1;
   for (int a = 0; a < dim; a++) { ns *= p; }
   const int nc = 1 << dim;
   s.assign((size_t)ns * nc, 0);
   for (int m = 0; m < ns; m++)
   {
      if (dim == 2)
      {
         const int aux = m + m / p;
         const int v[4] = {aux, aux + 1, aux + p + 1, aux + p + 2};
         for (int j = 0; j < 4; j++) { s[m * nc + j] = v[j]; }
      }
      else
      {
         const int aux = m + m / p + (p + 1) * (m / (p * p));
         const int v[8] = {aux, aux + 1, aux + p + 1, aux + p + 2, aux + n * n, aux + n * n + 1,
                           aux + n * n + p + 1, aux + n * n + p + 2};
         for (int j = 0; j < 8; j++) { s[m * nc + j] = v[j]; }
      }
   }
}

} // namespace rmh

using namespace rmh;

struct rmh_mesh { Mesh m; };

// ------------------------------------------------------------------------------- reading
static int read_tokens(const char *path, std::string &first, std::vector<std::string> &toks)
{
   std::ifstream f(path);
   if (!f) { set_error(std::string("cannot open mesh file ") + path); return 1; }
   std::getline(f, first);
   std::string line;
   while (std::getline(f, line))
   {
      const size_t h = line.find('#');
      if (h != std::string::npos) { line.erase(h); }
      std::istringstream is(line);
      std::string t;
      while (is >> t) { toks.push_back(t); }
   }
   return 0;
}

static const int MFEM2LEX2[4] = {0, 1, 3, 2};
static const int MFEM2LEX3[8] = {0, 1, 3, 2, 4, 5, 7, 6};

static int make_cartesian(int dim, const int *n, const double *origin, const double *size,
                          int periodic, Mesh &m)
{
   m.dim = dim; m.g = 1;
   int64_t nvd[3] = {1, 1, 1}, ne = 1, nv = 1;
   for (int a = 0; a < dim; a++)
   {
      if (periodic && n[a] < 3) { set_error("periodic Cartesian mesh needs n >= 3"); return 1; }
      nvd[a] = periodic ? n[a] : n[a] + 1;
      ne *= n[a]; nv *= nvd[a];
   }
   m.ne = ne; m.nv = nv;
   const int nvx = 1 << dim;
   m.ev.resize((size_t)ne * nvx);
   m.X.resize((size_t)ne * nvx * dim);
   for (int64_t e = 0; e < ne; e++)
   {
      int64_t idx[3], q = e;
      for (int a = 0; a < dim; a++) { idx[a] = q % n[a]; q /= n[a]; }
      for (int c = 0; c < nvx; c++)
      {
         int64_t vid = 0;
         for (int a = dim - 1; a >= 0; a--)
         {
            int64_t vi = idx[a] + ((c >> a) & 1);
            if (periodic) { vi %= nvd[a]; }
            vid = vid * nvd[a] + vi;
         }
         m.ev[e * nvx + c] = vid;
         for (int a = 0; a < dim; a++)
         {
            m.X[((size_t)e * nvx + c) * dim + a] =
               (origin ? origin[a] : 0.0) + (idx[a] + ((c >> a) & 1)) * (size[a] / n[a]);
         }
      }
   }
   return 0;
}

extern "C" int rmh_mesh_cartesian(int dim, const int *n, const double *origin,
                                  const double *size, int periodic, rmh_mesh **out)
{
   if (dim != 2 && dim != 3) { set_error("dim must be 2 or 3"); return 1; }
   rmh_mesh *r = new rmh_mesh;
   if (make_cartesian(dim, n, origin, size, periodic, r->m)) { delete r; return 1; }
   *out = r;
   return 0;
}

extern "C" int rmh_mesh_load(const char *path, rmh_mesh **out)
{
   std::string first;
   std::vector<std::string> tk;
   if (read_tokens(path, first, tk)) { return 1; }
   rmh_mesh *r = new rmh_mesh;
   Mesh &m = r->m;
   if (first.rfind("MFEM INLINE mesh", 0) == 0)
   {
      std::string type;
      int n[3] = {1, 1, 1};
      double s[3] = {1, 1, 1};
      for (size_t i = 0; i + 2 < tk.size() + 0; i++)
      {
         if (tk[i + 1] != "=") { continue; }
         const std::string &k = tk[i], &v = tk[i + 2];
         if (k == "type") { type = v; }
         else if (k == "nx") { n[0] = atoi(v.c_str()); }
         else if (k == "ny") { n[1] = atoi(v.c_str()); }
         else if (k == "nz") { n[2] = atoi(v.c_str()); }
         else if (k == "sx") { s[0] = atof(v.c_str()); }
         else if (k == "sy") { s[1] = atof(v.c_str()); }
         else if (k == "sz") { s[2] = atof(v.c_str()); }
      }
      int dim = 0;
      if (type == "quad") { dim = 2; }
      else if (type == "hex") { dim = 3; }
      else { set_error("inline mesh: only type = quad | hex"); delete r; return 1; }
      if (make_cartesian(dim, n, nullptr, s, 0, m)) { delete r; return 1; }
      *out = r;
      return 0;
   }
   if (first.rfind("MFEM mesh v1.0", 0) != 0)
   {
      set_error("unsupported mesh format: " + first);
      delete r;
      return 1;
   }
   size_t pos = 0;
   auto fail = [&](const std::string &msg) { set_error("mesh parse: " + msg); delete r; return 1; };
   auto expect = [&](const char *w) { return pos < tk.size() && tk[pos++] == w; };
   if (!expect("dimension")) { return fail("expected dimension"); }
   const int dim = atoi(tk[pos++].c_str());
   if (dim != 2 && dim != 3) { return fail("only 2D/3D meshes"); }
   if (!expect("elements")) { return fail("expected elements"); }
   const int64_t ne = atoll(tk[pos++].c_str());
   const int nvx = 1 << dim;
   m.dim = dim; m.ne = ne;
   m.ev.resize((size_t)ne * nvx);
   std::vector<int64_t> ev_file((size_t)ne * nvx);        // the file's (MFEM's) vertex order
   for (int64_t e = 0; e < ne; e++)
   {
      const int geom = atoi(tk[pos + 1].c_str());
      if (!((dim == 2 && geom == 3) || (dim == 3 && geom == 5)))
      { return fail("only quadrilateral / hexahedral elements"); }
      for (int c = 0; c < nvx; c++)
      {
         const int src = (dim == 2) ? MFEM2LEX2[c] : MFEM2LEX3[c];
         m.ev[e * nvx + c] = atoll(tk[pos + 2 + src].c_str());
         ev_file[e * nvx + c] = atoll(tk[pos + 2 + c].c_str());
      }
      pos += 2 + nvx;
   }
   if (!expect("boundary")) { return fail("expected boundary"); }
   const int64_t nb = atoll(tk[pos++].c_str());
   for (int64_t b = 0; b < nb; b++)
   {
      const int geom = atoi(tk[pos + 1].c_str());
      const int nbv = (geom == 1) ? 2 : (geom == 3) ? 4 : 1;
      pos += 2 + nbv;
   }
   if (!expect("vertices")) { return fail("expected vertices"); }
   const int64_t nv = atoll(tk[pos++].c_str());
   m.nv = nv;
   std::vector<double> coords;
   if (pos < tk.size() && tk[pos] == "nodes")
   {
      pos++;
      if (!expect("FiniteElementSpace")) { return fail("expected FiniteElementSpace"); }
      if (!expect("FiniteElementCollection:")) { return fail("expected FiniteElementCollection"); }
      const std::string fec = tk[pos++];
      if (!expect("VDim:")) { return fail("expected VDim"); }
      const int vdim = atoi(tk[pos++].c_str());
      if (!expect("Ordering:")) { return fail("expected Ordering"); }
      const int ordering = atoi(tk[pos++].c_str());
      if (vdim != dim) { return fail("VDim != dimension"); }
      const size_t nvals = tk.size() - pos;
      if (fec.rfind("L2_T1_", 0) == 0)
      {
         const size_t pp = fec.find("_P");
         const int g = atoi(fec.c_str() + pp + 2);
         m.g = g;
         const int npe = m.npe();
         const size_t nd = (size_t)ne * npe;
         if (nvals != nd * dim) { return fail("wrong number of nodal values"); }
         m.X.resize(nd * dim);
         for (size_t i = 0; i < nd; i++)
            for (int c = 0; c < dim; c++)
            {
               const size_t src = ordering == 1 ? i * dim + c : c * nd + i;
               m.X[i * dim + c] = atof(tk[pos + src].c_str());
            }
         *out = r;
         return 0;
      }
      if ((fec == "Quadratic" || fec == "H1_2D_P2") && dim == 2)
      {
         // H1 order-2 nodal field on quadrilaterals (legacy `Quadratic` collection): global dofs
         // are [vertices | edges | elements]; edges are numbered in the order of their first
         // appearance over the elements and their local edges (0,1), (1,2), (2,3), (3,0) -- MFEM's
         // vertex-to-vertex table (validated by remhos_tests.cpp:88-91 through the star-q2 run)
         std::map<std::pair<int64_t, int64_t>, int64_t> edge_id;
         std::vector<int64_t> el_edges((size_t)ne * 4);
         static const int EV[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
         for (int64_t e = 0; e < ne; e++)
            for (int j = 0; j < 4; j++)
            {
               const int64_t a = ev_file[e * 4 + EV[j][0]], b = ev_file[e * 4 + EV[j][1]];
               const std::pair<int64_t, int64_t> key(std::min(a, b), std::max(a, b));
               auto it = edge_id.find(key);
               if (it == edge_id.end()) { it = edge_id.emplace(key, (int64_t)edge_id.size()).first; }
               el_edges[e * 4 + j] = it->second;
            }
         const int64_t nedge = (int64_t)edge_id.size(), ndof = nv + nedge + ne;
         if (nvals != (size_t)ndof * dim) { return fail("wrong number of nodal values"); }
         auto val = [&](int64_t i, int c)
         { return atof(tk[pos + (ordering == 1 ? (size_t)i * dim + c : (size_t)c * ndof + i)].c_str()); };
         m.g = 2;
         m.X.resize((size_t)ne * 9 * dim);
         for (int64_t e = 0; e < ne; e++)
         {
            const int64_t *v = &ev_file[e * 4], *ed = &el_edges[e * 4];
            // lexicographic 3 x 3 nodes: corners, edge midpoints, centre
            const int64_t dof[9] = {v[0], nv + ed[0], v[1], nv + ed[3], nv + nedge + e, nv + ed[1],
                                    v[3], nv + ed[2], v[2]};
            for (int n = 0; n < 9; n++)
               for (int c = 0; c < dim; c++) { m.X[((size_t)e * 9 + n) * dim + c] = val(dof[n], c); }
         }
         *out = r;
         return 0;
      }
      if ((fec == "Cubic" || fec == "H1_2D_P3") && dim == 2)
      {
         // legacy `Cubic` collection (equispaced nodes): [vertices | 2 per edge | 4 per element]; the
         // two edge dofs run from the edge's lower-numbered vertex to the higher one, the element
         // dofs are (1/3,1/3), (2/3,1/3), (1/3,2/3), (2/3,2/3); stored as Gauss-Lobatto nodes of the
         // same cubic map (DESIGN.md section 3 says how these conventions were checked)
         std::map<std::pair<int64_t, int64_t>, int64_t> edge_id;
         std::vector<int64_t> el_edges((size_t)ne * 4);
         std::vector<char> el_fwd((size_t)ne * 4);
         static const int EV[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
         for (int64_t e = 0; e < ne; e++)
            for (int j = 0; j < 4; j++)
            {
               const int64_t a = ev_file[e * 4 + EV[j][0]], b = ev_file[e * 4 + EV[j][1]];
               const std::pair<int64_t, int64_t> key(std::min(a, b), std::max(a, b));
               auto it = edge_id.find(key);
               if (it == edge_id.end()) { it = edge_id.emplace(key, (int64_t)edge_id.size()).first; }
               el_edges[e * 4 + j] = it->second;
               el_fwd[e * 4 + j] = a < b;
            }
         const int64_t nedge = (int64_t)edge_id.size(), ndof = nv + 2 * nedge + 4 * ne;
         if (nvals != (size_t)ndof * dim) { return fail("wrong number of nodal values"); }
         auto val = [&](int64_t i, int c)
         { return atof(tk[pos + (ordering == 1 ? (size_t)i * dim + c : (size_t)c * ndof + i)].c_str()); };
         const std::vector<double> equi = {0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0};
         const std::vector<double> T = lagrange(equi, gauss_lobatto_01(4));     // [gll point][equi node]
         m.g = 3;
         m.X.resize((size_t)ne * 16 * dim);
         for (int64_t e = 0; e < ne; e++)
         {
            const int64_t *v = &ev_file[e * 4];
            auto edof = [&](int j, int k)
            { return nv + 2 * el_edges[e * 4 + j] + (el_fwd[e * 4 + j] ? k : 1 - k); };
            const int64_t base = nv + 2 * nedge + 4 * e;
            int64_t dof[4][4];                                    // [iy][ix]
            dof[0][0] = v[0]; dof[0][3] = v[1]; dof[3][3] = v[2]; dof[3][0] = v[3];
            dof[0][1] = edof(0, 0); dof[0][2] = edof(0, 1);
            dof[1][3] = edof(1, 0); dof[2][3] = edof(1, 1);
            dof[3][2] = edof(2, 0); dof[3][1] = edof(2, 1);
            dof[2][0] = edof(3, 0); dof[1][0] = edof(3, 1);
            dof[1][1] = base; dof[1][2] = base + 1; dof[2][1] = base + 2; dof[2][2] = base + 3;
            for (int c = 0; c < dim; c++)
            {
               double E[4][4], tmp[4][4];
               for (int iy = 0; iy < 4; iy++) for (int ix = 0; ix < 4; ix++) { E[iy][ix] = val(dof[iy][ix], c); }
               for (int iy = 0; iy < 4; iy++)
                  for (int q = 0; q < 4; q++)
                  {
                     double sacc = 0.0;
                     for (int i = 0; i < 4; i++) { sacc += T[q * 4 + i] * E[iy][i]; }
                     tmp[iy][q] = sacc;
                  }
               for (int p = 0; p < 4; p++)
                  for (int q = 0; q < 4; q++)
                  {
                     double sacc = 0.0;
                     for (int j = 0; j < 4; j++) { sacc += T[p * 4 + j] * tmp[j][q]; }
                     m.X[((size_t)e * 16 + p * 4 + q) * dim + c] = sacc;
                  }
            }
         }
         *out = r;
         return 0;
      }
      if (fec == "Linear" || (fec.rfind("H1_", 0) == 0 && fec.size() >= 3 &&
                              fec.compare(fec.size() - 3, 3, "_P1") == 0))
      {
         if (nvals != (size_t)nv * dim) { return fail("wrong number of nodal values"); }
         coords.resize((size_t)nv * dim);
         for (int64_t i = 0; i < nv; i++)
            for (int c = 0; c < dim; c++)
            {
               const size_t src = ordering == 1 ? (size_t)i * dim + c : (size_t)c * nv + i;
               coords[i * dim + c] = atof(tk[pos + src].c_str());
            }
      }
      else { return fail("unsupported nodal collection " + fec); }
   }
   else
   {
      const int sdim = atoi(tk[pos++].c_str());
      if (sdim != dim) { return fail("space dimension != dimension"); }
      coords.resize((size_t)nv * dim);
      for (size_t i = 0; i < (size_t)nv * dim; i++) { coords[i] = atof(tk[pos + i].c_str()); }
   }
   m.g = 1;
   m.X.resize((size_t)ne * nvx * dim);
   for (int64_t e = 0; e < ne; e++)
      for (int c = 0; c < nvx; c++)
         for (int a = 0; a < dim; a++)
         { m.X[((size_t)e * nvx + c) * dim + a] = coords[m.ev[e * nvx + c] * dim + a]; }
   *out = r;
   return 0;
}

// Neighbourhood lattice of every element, derived from the lattice-entity map `lat`
// ([ne][3^dim] entity ids, as returned by rmh_mesh_dof_maps): nbr[e][d], d = sum (1 + d_a) 3^a with
// d_a in {-1, 0, 1}, is the element reached from e by crossing the entity in direction d (a face,
// an edge, a vertex), e itself for d = 0, -1 if there is none (domain boundary).
// *structured = 1 iff, for every element and every one of its 3^dim entities, the set of elements
// sharing the entity equals { nbr[e][d] : d_a in S(c_a) }, S(0) = {-1, 0}, S(1) = {0}, S(2) = {0, 1}
// (c = the entity's position in the element).  Then the overlap bounds of
// DofInfo::ComputeOverlapBounds (remhos_tools.cpp:432-495) -- min/max over the elements sharing a
// lattice entity -- can be formed from the 3^dim neighbourhood values without the entity table.
// Unstructured vertex valences (e.g. data/periodic-hexagon.mesh) give *structured = 0.
extern "C" int rmh_nbr_lattice(int dim, int64_t ne, int32_t n_ent, const int32_t *lat, int32_t *nbr,
                               int *structured)
{
   return rmh::nbr_lattice_rows(dim, ne, ne, n_ent, lat, nbr, structured);
}

namespace rmh
{
int nbr_lattice_rows(int dim, int64_t ne, int64_t ne_rows, int32_t n_ent, const int32_t *lat,
                     int32_t *nbr, int *structured)
{
   if (dim != 2 && dim != 3) { set_error("rmh_nbr_lattice: dim must be 2 or 3"); return 1; }
   int n3 = 1;
   for (int a = 0; a < dim; a++) { n3 *= 3; }
   // entity -> elements (CSR)
   std::vector<int64_t> off((size_t)n_ent + 1, 0);
   for (int64_t i = 0; i < ne * n3; i++)
   {
      if (lat[i] < 0 || lat[i] >= n_ent) { set_error("rmh_nbr_lattice: lat id out of range"); return 1; }
      off[lat[i] + 1]++;
   }
   for (int32_t i = 0; i < n_ent; i++) { off[i + 1] += off[i]; }
   std::vector<int32_t> el((size_t)ne * n3);
   {
      std::vector<int64_t> cur(off.begin(), off.end() - 1);
      for (int64_t e = 0; e < ne; e++)
         for (int t = 0; t < n3; t++) { el[cur[lat[e * n3 + t]]++] = (int32_t)e; }
   }
   auto members = [&](int64_t e, int t, std::vector<int32_t> &out)
   {
      out.clear();
      const int32_t id = lat[e * n3 + t];
      for (int64_t k = off[id]; k < off[id + 1]; k++) { out.push_back(el[k]); }
      std::sort(out.begin(), out.end());
      out.erase(std::unique(out.begin(), out.end()), out.end());
   };
   int ok = 1;
#pragma omp parallel
   {
   std::vector<int32_t> S, T, rest;
   // pass 1: neighbours, by increasing number of non-zero offset components
#pragma omp for schedule(static) reduction(min : ok)
   for (int64_t e = 0; e < ne_rows; e++)
   {
      int32_t *nb = nbr + e * n3;
      for (int t = 0; t < n3; t++) { nb[t] = -2; }
      for (int nz = 0; nz <= dim; nz++)
         for (int t = 0; t < n3; t++)
         {
            int c[3] = {1, 1, 1}, r = t, cnt = 0;
            for (int a = 0; a < dim; a++) { c[a] = r % 3; r /= 3; cnt += (c[a] != 1); }
            if (cnt != nz) { continue; }
            if (nz == 0) { nb[t] = (int32_t)e; continue; }
            // elements sharing the entity in direction d = c - 1, minus those reached by the proper
            // sub-offsets (some components of d zeroed)
            members(e, t, S);
            T.clear();
            for (int m = 0; m < (1 << dim); m++)
            {
               int tt = 0, mul = 1;
               bool proper = false, valid = true;
               for (int a = 0; a < dim; a++)
               {
                  int ca = c[a];
                  if ((m >> a) & 1) { if (c[a] == 1) { valid = false; } ca = 1; proper = true; }
                  tt += ca * mul; mul *= 3;
               }
               if (!valid || !proper) { continue; }
               if (nb[tt] >= 0) { T.push_back(nb[tt]); }
            }
            std::sort(T.begin(), T.end());
            T.erase(std::unique(T.begin(), T.end()), T.end());
            rest.clear();
            std::set_difference(S.begin(), S.end(), T.begin(), T.end(), std::back_inserter(rest));
            if (rest.empty()) { nb[t] = -1; }
            else if (rest.size() == 1) { nb[t] = rest[0]; }
            else { nb[t] = rest[0]; ok = 0; }
         }
   }
   // pass 2: the neighbourhood must reproduce every entity's element set
#pragma omp for schedule(static) reduction(min : ok)
   for (int64_t e = 0; e < ne_rows; e++)
   {
      const int32_t *nb = nbr + e * n3;
      for (int t = 0; t < n3 && ok; t++)
      {
         int c[3] = {1, 1, 1}, r = t;
         for (int a = 0; a < dim; a++) { c[a] = r % 3; r /= 3; }
         members(e, t, S);
         T.clear();
         for (int d = 0; d < n3; d++)
         {
            int q = d;
            bool in = true;
            for (int a = 0; a < dim; a++)
            {
               const int da = q % 3 - 1; q /= 3;
               if (!((c[a] == 0 && da <= 0) || (c[a] == 1 && da == 0) || (c[a] == 2 && da >= 0))) { in = false; }
            }
            if (in && nb[d] >= 0) { T.push_back(nb[d]); }
         }
         std::sort(T.begin(), T.end());
         T.erase(std::unique(T.begin(), T.end()), T.end());
         if (T != S) { ok = 0; }
      }
   }
   }
   if (structured) { *structured = ok; }
   return 0;
}
} // namespace rmh

extern "C" int rmh_mesh_free(rmh_mesh *m) { delete m; return 0; }

extern "C" int rmh_mesh_refine(rmh_mesh *m, int levels)
{
   for (int l = 0; l < levels; l++) { refine_once(m->m); }
   return 0;
}

extern "C" int rmh_mesh_set_curvature(rmh_mesh *m, int order)
{
   if (order < 1) { set_error("mesh order must be >= 1"); return 1; }
   set_curvature(m->m, order);
   return 0;
}

extern "C" int rmh_mesh_bounding_box(const rmh_mesh *m, double *bb_min, double *bb_max)
{
   const Mesh &M = m->m;
   for (int a = 0; a < M.dim; a++) { bb_min[a] = INFINITY; bb_max[a] = -INFINITY; }
   const size_t np = M.X.size() / M.dim;
   for (size_t i = 0; i < np; i++)
      for (int a = 0; a < M.dim; a++)
      {
         bb_min[a] = std::min(bb_min[a], M.X[i * M.dim + a]);
         bb_max[a] = std::max(bb_max[a], M.X[i * M.dim + a]);
      }
   return 0;
}

extern "C" int rmh_mesh_dim(const rmh_mesh *m) { return m->m.dim; }
extern "C" int rmh_mesh_ne(const rmh_mesh *m) { return (int)m->m.ne; }
extern "C" int rmh_mesh_nv(const rmh_mesh *m) { return (int)m->m.nv; }
extern "C" int rmh_mesh_geom_order(const rmh_mesh *m) { return m->m.g; }
extern "C" const double *rmh_mesh_nodes(const rmh_mesh *m) { return m->m.X.data(); }
extern "C" const int64_t *rmh_mesh_elem_vertices(const rmh_mesh *m) { return m->m.ev.data(); }

extern "C" int rmh_mesh_extract(const rmh_mesh *m, int64_t n, const int64_t *ids, rmh_mesh **out)
{
   const Mesh &M = m->m;
   rmh_mesh *r = new rmh_mesh;
   Mesh &S = r->m;
   S.dim = M.dim; S.g = M.g; S.nv = M.nv; S.ne = n;
   const int nvx = M.nvert(), npe = M.npe();
   S.ev.resize((size_t)n * nvx);
   S.X.resize((size_t)n * npe * M.dim);
   for (int64_t i = 0; i < n; i++)
   {
      if (ids[i] < 0 || ids[i] >= M.ne) { set_error("extract: bad element id"); delete r; return 1; }
      std::copy(&M.ev[ids[i] * nvx], &M.ev[ids[i] * nvx] + nvx, &S.ev[i * nvx]);
      std::copy(&M.X[(size_t)ids[i] * npe * M.dim], &M.X[(size_t)ids[i] * npe * M.dim] + npe * M.dim,
                &S.X[(size_t)i * npe * M.dim]);
   }
   // compact the vertex ids (topology only depends on their identity)
   {
      std::vector<int64_t> v(S.ev);
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
      for (auto &x : S.ev) { x = std::lower_bound(v.begin(), v.end(), x) - v.begin(); }
      S.nv = (int64_t)v.size();
   }
   *out = r;
   return 0;
}

// Mesh::MakeRefined(mesh, factor, BasisType::ClosedUniform) (remhos.cpp:801): every element split into
// factor^dim linear sub-elements on the uniform lattice i / factor -- the subcell mesh of the subcell
// residual-distribution schemes, written as meshLO_*.mesh by -save (remhos.cpp:1021-1026, 1371-1376).
// Vertices = distinct lattice points: the order-`factor` DG lattice points united through the face
// coincidences of the neighbour-DOF map (the same identification the H1 numbering of the smoothness
// indicator uses), so periodic meshes stay periodic.  nodes == NULL: the mesh's own nodes; else
// [ne][(g+1)^dim][dim] (the moved mesh of a remap run).  The result has geometry order 1 with
// element-wise (discontinuous) nodes, as SetCurvature(1, disc_nodes) gives for periodic meshes (:822).
extern "C" int rmh_mesh_make_refined(const rmh_mesh *mm, int factor, const double *nodes, rmh_mesh **out)
{
   const Mesh &M = mm->m;
   if (factor < 1) { set_error("make_refined: factor must be >= 1"); return 1; }
   const int dim = M.dim, p = factor, n = p + 1, nf = 2 * dim, nvx = M.nvert();
   int nd = 1, nfd = 1, ns = 1;
   for (int a = 0; a < dim; a++) { nd *= n; ns *= p; }
   for (int a = 0; a < dim - 1; a++) { nfd *= n; }
   if ((double)M.ne * nd >= 2147483647.0) { set_error("make_refined: int32 overflow"); return 1; }
   // ---- distinct lattice points
   std::vector<int32_t> bd((size_t)nf * nfd), nbr((size_t)M.ne * nf * nfd);
   if (rmh_mesh_dof_maps(mm, p, bd.data(), nbr.data(), nullptr, nullptr, nullptr, nullptr)) { return 1; }
   const int64_t N = M.ne * nd;
   std::vector<int64_t> par((size_t)N);
   for (int64_t i = 0; i < N; i++) { par[i] = i; }
   auto find = [&](int64_t i)
   {
      while (par[i] != i) { par[i] = par[par[i]]; i = par[i]; }
      return i;
   };
   for (int64_t e = 0; e < M.ne; e++)
      for (int f = 0; f < nf; f++)
         for (int j = 0; j < nfd; j++)
         {
            const int64_t g = nbr[((size_t)e * nf + f) * nfd + j];
            if (g < 0) { continue; }
            const int64_t a = find(e * nd + bd[j * nf + f]), b = find(g);       // bd is [nfd][nf]
            if (a != b) { par[std::max(a, b)] = std::min(a, b); }
         }
   std::vector<int64_t> vid((size_t)N, -1);
   int64_t nv = 0;
   for (int64_t i = 0; i < N; i++) { const int64_t r = find(i); if (vid[r] < 0) { vid[r] = nv++; } vid[i] = vid[r]; }
   // ---- positions of the lattice points
   const int g = M.g, n1 = g + 1;
   int nn = 1;
   for (int a = 0; a < dim; a++) { nn *= n1; }
   std::vector<double> pts((size_t)n);
   for (int i = 0; i < n; i++) { pts[i] = (double)i / p; }
   const std::vector<double> L = lagrange(gauss_lobatto_01(n1), pts);      // [n][n1]
   const double *X = nodes ? nodes : M.X.data();
   rmh_mesh *r = new rmh_mesh;
   Mesh &S = r->m;
   S.dim = dim; S.g = 1; S.nv = nv; S.ne = M.ne * ns;
   S.ev.resize((size_t)S.ne * nvx);
   S.X.resize((size_t)S.ne * nvx * dim);
   std::vector<double> xl((size_t)nd * dim);
   for (int64_t e = 0; e < M.ne; e++)
   {
      const double *Xe = X + (size_t)e * nn * dim;
      for (int q = 0; q < nd; q++)
      {
         int l[3] = {0, 0, 0}, mq = q;
         for (int a = 0; a < dim; a++) { l[a] = mq % n; mq /= n; }
         double x[3] = {0, 0, 0};
         for (int k = 0; k < nn; k++)
         {
            int mk = k;
            double w = 1.0;
            for (int a = 0; a < dim; a++) { w *= L[(size_t)l[a] * n1 + mk % n1]; mk /= n1; }
            for (int i = 0; i < dim; i++) { x[i] += w * Xe[k * dim + i]; }
         }
         for (int i = 0; i < dim; i++) { xl[(size_t)q * dim + i] = x[i]; }
      }
      for (int sc = 0; sc < ns; sc++)
      {
         int c[3] = {0, 0, 0}, ms = sc;
         for (int a = 0; a < dim; a++) { c[a] = ms % p; ms /= p; }
         const int64_t se = e * ns + sc;
         for (int k = 0; k < nvx; k++)
         {
            int q = 0, mul = 1;
            for (int a = 0; a < dim; a++) { q += (c[a] + ((k >> a) & 1)) * mul; mul *= n; }
            S.ev[se * nvx + k] = vid[e * nd + q];
            for (int i = 0; i < dim; i++) { S.X[((size_t)se * nvx + k) * dim + i] = xl[(size_t)q * dim + i]; }
         }
      }
   }
   *out = r;
   return 0;
}

// ------------------------------------------------------------------- on-disk formats
// Mesh::Print in "MFEM mesh v1.0" (what pmesh.PrintAsOne writes for -save, remhos.cpp:1016-1030,
// 1366-1380, and VisItDataCollection for -visit, :1034-1043): elements in MFEM's vertex order,
// boundary = the faces without a neighbour (attribute 1), nodes as the element-wise Gauss-Lobatto
// field L2_T1_<dim>D_P<g> (byVDIM) -- the representation MFEM itself uses for periodic meshes, valid
// for every mesh, and read back by rmh_mesh_load.  nodes == NULL: the mesh's own nodes; else
// [ne][(g+1)^dim][dim] (e.g. the moved mesh of a remap run).
extern "C" int rmh_mesh_save(const rmh_mesh *mm, const char *path, const double *nodes, int precision)
{
   const Mesh &m = mm->m;
   const int dim = m.dim, nvx = m.nvert(), nf = 2 * dim, nfc = 1 << (dim - 1);
   Topology T;
   if (build_topology(m, T)) { return 1; }
   FILE *fp = std::fopen(path, "w");
   if (!fp) { set_error(std::string("cannot write ") + path); return 1; }
   std::fprintf(fp, "MFEM mesh v1.0\n\ndimension\n%d\n\nelements\n%lld\n", dim, (long long)m.ne);
   const int *perm = (dim == 2) ? MFEM2LEX2 : MFEM2LEX3;
   for (int64_t e = 0; e < m.ne; e++)
   {
      std::fprintf(fp, "1 %d", dim == 2 ? 3 : 5);
      for (int k = 0; k < nvx; k++) { std::fprintf(fp, " %lld", (long long)m.ev[e * nvx + perm[k]]); }
      std::fprintf(fp, "\n");
   }
   int64_t nb = 0;
   for (int64_t i = 0; i < m.ne * nf; i++) { if (T.nbr_elem[i] < 0) { nb++; } }
   std::fprintf(fp, "\nboundary\n%lld\n", (long long)nb);
   for (int64_t e = 0; e < m.ne; e++)
      for (int f = 0; f < nf; f++)
      {
         if (T.nbr_elem[e * nf + f] >= 0) { continue; }
         int fc[4], axis, side;
         face_corners(dim, f, fc);
         face_axis(dim, f, axis, side);
         // counter-clockwise seen from outside (natural parametrisation: remaining axes ascending)
         int order[4] = {0, 1, 3, 2};
         if (dim == 2) { order[0] = 0; order[1] = 1; }
         const bool flip = (dim == 3) ? ((axis == 2 && side == 0) || (axis == 1 && side == 1) || (axis == 0 && side == 0))
                                      : ((axis == 1 && side == 1) || (axis == 0 && side == 0));
         std::fprintf(fp, "1 %d", dim == 2 ? 1 : 3);
         for (int k = 0; k < nfc; k++)
         {
            const int kk = flip ? (nfc - k) % nfc : k;       // reversed cycle, same starting corner
            std::fprintf(fp, " %lld", (long long)m.ev[e * nvx + fc[order[kk]]]);
         }
         std::fprintf(fp, "\n");
      }
   std::fprintf(fp, "\nvertices\n%lld\n\nnodes\nFiniteElementSpace\nFiniteElementCollection: L2_T1_%dD_P%d\n"
                    "VDim: %d\nOrdering: 1\n\n", (long long)m.nv, dim, m.g, dim);
   const double *X = nodes ? nodes : m.X.data();
   const size_t np = (size_t)m.ne * m.npe();
   for (size_t i = 0; i < np; i++)
   {
      for (int a = 0; a < dim; a++) { std::fprintf(fp, a ? " %.*g" : "%.*g", precision, X[i * dim + a]); }
      std::fprintf(fp, "\n");
   }
   std::fclose(fp);
   return 0;
}

// GridFunction::Save of a scalar field in the DG space of the run (u.SaveAsOne, remhos.cpp:1027-1029):
// L2_T<basis>_<dim>D_P<order>, basis 2 = Positive (Bernstein, remhos.cpp:588-590), element-major values
extern "C" int rmh_gf_save(const char *path, int dim, int order, int basis_type, int64_t n, const double *vals,
                           int precision)
{
   FILE *fp = std::fopen(path, "w");
   if (!fp) { set_error(std::string("cannot write ") + path); return 1; }
   std::fprintf(fp, "FiniteElementSpace\nFiniteElementCollection: L2_T%d_%dD_P%d\nVDim: 1\nOrdering: 0\n\n", basis_type,
                dim, order);
   for (int64_t i = 0; i < n; i++) { std::fprintf(fp, "%.*g\n", precision, vals[i]); }
   std::fclose(fp);
   return 0;
}

// ------------------------------------------------------------------- domain decomposition
// Recursive coordinate bisection of the element centroids into nparts (any count; splits are
// proportional).  Replaces the METIS / Cartesian partitioning ParMesh does (remhos.cpp:451-461).
static void rcb(const std::vector<double> &cen, int dim, std::vector<int64_t> &idx, int64_t lo,
                int64_t hi, int p0, int np, int32_t *part)
{
   if (np == 1) { for (int64_t i = lo; i < hi; i++) { part[idx[i]] = p0; } return; }
   double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
   for (int64_t i = lo; i < hi; i++)
      for (int a = 0; a < dim; a++)
      {
         mn[a] = std::min(mn[a], cen[idx[i] * dim + a]);
         mx[a] = std::max(mx[a], cen[idx[i] * dim + a]);
      }
   int ax = 0;
   for (int a = 1; a < dim; a++) { if (mx[a] - mn[a] > (mx[ax] - mn[ax]) * (1.0 + 1e-12)) { ax = a; } }
   const int npl = np / 2;
   const int64_t mid = lo + (hi - lo) * npl / np;
   std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                    [&](int64_t a, int64_t b)
                    {
                       const double ca = cen[a * dim + ax], cb = cen[b * dim + ax];
                       return ca < cb || (ca == cb && a < b);
                    });
   rcb(cen, dim, idx, lo, mid, p0, npl, part);
   rcb(cen, dim, idx, mid, hi, p0 + npl, np - npl, part);
}

extern "C" int rmh_mesh_partition(const rmh_mesh *m, int nparts, int32_t *part)
{
   const Mesh &M = m->m;
   if (nparts < 1) { set_error("partition: nparts < 1"); return 1; }
   const int dim = M.dim, npe = M.npe();
   std::vector<double> cen((size_t)M.ne * dim, 0.0);
   for (int64_t e = 0; e < M.ne; e++)
      for (int n = 0; n < npe; n++)
         for (int a = 0; a < dim; a++) { cen[e * dim + a] += M.X[((size_t)e * npe + n) * dim + a] / npe; }
   std::vector<int64_t> idx(M.ne);
   for (int64_t e = 0; e < M.ne; e++) { idx[e] = e; }
   rcb(cen, dim, idx, 0, M.ne, 0, nparts, part);
   return 0;
}

// Halo plan of `rank`: owned elements (ascending global id), ghost ring = elements of other
// ranks sharing at least a vertex with an owned element, ordered by (owner, global id), and per
// peer the owned elements that are in the peer's ghost ring (ascending global id).  Vertex
// adjacency is symmetric, so every rank derives matching send/receive lists without talking.
int32_t rmh_halo::local_of(int64_t g) const
{
   const auto it = std::lower_bound(owned_sorted.begin(), owned_sorted.end(), g);
   if (it == owned_sorted.end() || *it != g) { return -1; }
   return owned_pos[it - owned_sorted.begin()];
}

static void halo_index(rmh_halo *h)
{
   const size_t n = h->owned.size();
   std::vector<int32_t> idx(n);
   for (size_t i = 0; i < n; i++) { idx[i] = (int32_t)i; }
   std::sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return h->owned[a] < h->owned[b]; });
   h->owned_sorted.resize(n); h->owned_pos.resize(n);
   for (size_t i = 0; i < n; i++) { h->owned_sorted[i] = h->owned[idx[i]]; h->owned_pos[i] = idx[i]; }
}

extern "C" int rmh_halo_create(const rmh_mesh *m, const int32_t *part, int rank, rmh_halo **out)
{
   const Mesh &M = m->m;
   const int nvx = M.nvert();
   rmh_halo *h = new rmh_halo;
   // vertex -> elements CSR
   std::vector<int64_t> off((size_t)M.nv + 1, 0);
   for (size_t i = 0; i < M.ev.size(); i++) { off[M.ev[i] + 1]++; }
   for (int64_t v = 0; v < M.nv; v++) { off[v + 1] += off[v]; }
   std::vector<int64_t> cur(off.begin(), off.end() - 1), v2e(M.ev.size());
   for (int64_t e = 0; e < M.ne; e++)
      for (int c = 0; c < nvx; c++) { v2e[cur[M.ev[e * nvx + c]]++] = e; }
   std::vector<std::pair<int32_t, int64_t>> gh;      // (owner, ghost element)
   std::vector<std::pair<int32_t, int64_t>> sd;      // (peer, owned element)
   for (int64_t e = 0; e < M.ne; e++)
   {
      if (part[e] != rank) { continue; }
      h->owned.push_back(e);
      for (int c = 0; c < nvx; c++)
      {
         const int64_t v = M.ev[e * nvx + c];
         for (int64_t k = off[v]; k < off[v + 1]; k++)
         {
            const int64_t e2 = v2e[k];
            if (part[e2] != rank)
            {
               gh.emplace_back(part[e2], e2);
               sd.emplace_back(part[e2], e);
            }
         }
      }
   }
   std::sort(gh.begin(), gh.end());
   gh.erase(std::unique(gh.begin(), gh.end()), gh.end());
   std::sort(sd.begin(), sd.end());
   sd.erase(std::unique(sd.begin(), sd.end()), sd.end());
   for (auto &g : gh) { h->ghost.push_back(g.second); h->ghost_owner.push_back(g.first); }
   for (auto &g : gh) { if (h->peers.empty() || h->peers.back() != g.first) { h->peers.push_back(g.first); } }
   // peers from the send side must coincide (symmetry); build offsets per peer
   h->send_off.assign(h->peers.size() + 1, 0);
   h->recv_off.assign(h->peers.size() + 1, 0);
   for (size_t p = 0; p < h->peers.size(); p++)
   {
      int64_t ns = 0, nr = 0;
      for (auto &x : sd) { if (x.first == h->peers[p]) { ns++; } }
      for (auto &x : gh) { if (x.first == h->peers[p]) { nr++; } }
      h->send_off[p + 1] = h->send_off[p] + (int32_t)ns;
      h->recv_off[p + 1] = h->recv_off[p] + (int32_t)nr;
   }
   if ((size_t)h->send_off.back() != sd.size())
   { set_error("halo: asymmetric adjacency"); delete h; return 1; }
   for (auto &x : sd) { h->send.push_back(x.second); }
   halo_index(h);
   *out = h;
   return 0;
}

extern "C" int rmh_halo_free(rmh_halo *h) { delete h; return 0; }
extern "C" int rmh_halo_sizes(const rmh_halo *h, int64_t *n_owned, int64_t *n_ghost, int32_t *n_peers,
                              int64_t *n_send)
{
   *n_owned = (int64_t)h->owned.size(); *n_ghost = (int64_t)h->ghost.size();
   *n_peers = (int32_t)h->peers.size(); *n_send = (int64_t)h->send.size();
   return 0;
}
// owned / ghost: global element ids; send_local: LOCAL owned index (position in `owned`) of every
// element to send, concatenated by peer; offsets have n_peers + 1 entries
extern "C" int rmh_halo_get(const rmh_halo *h, int64_t *owned, int64_t *ghost, int32_t *peers,
                            int32_t *send_off, int32_t *recv_off, int32_t *send_local)
{
   std::copy(h->owned.begin(), h->owned.end(), owned);
   std::copy(h->ghost.begin(), h->ghost.end(), ghost);
   std::copy(h->peers.begin(), h->peers.end(), peers);
   std::copy(h->send_off.begin(), h->send_off.end(), send_off);
   std::copy(h->recv_off.begin(), h->recv_off.end(), recv_off);
   for (size_t i = 0; i < h->send.size(); i++) { send_local[i] = h->local_of(h->send[i]); }
   return 0;
}

// Reorder the owned elements: those no peer needs (= those sharing no vertex with a ghost element;
// vertex adjacency is symmetric) first, in their previous relative order; the others ("shell")
// last.  The stage kernel runs the leading elements while the halo is still in flight.
extern "C" int rmh_halo_interior_first(rmh_halo *h, int64_t *n_interior)
{
   const size_t n = h->owned.size();
   std::vector<char> bnd(n, 0);
   for (int64_t g : h->send) { bnd[h->local_of(g)] = 1; }
   std::vector<int64_t> o2;
   o2.reserve(n);
   for (size_t i = 0; i < n; i++) { if (!bnd[i]) { o2.push_back(h->owned[i]); } }
   const int64_t ni = (int64_t)o2.size();
   for (size_t i = 0; i < n; i++) { if (bnd[i]) { o2.push_back(h->owned[i]); } }
   h->owned.swap(o2);
   halo_index(h);
   h->n_interior = ni;
   if (n_interior) { *n_interior = ni; }
   return 0;
}

extern "C" int rmh_mesh_dof_maps(const rmh_mesh *mm, int p, int32_t *bdr_out, int32_t *nbr_dof,
                                 int32_t *s2i_out, int32_t *lat_out, int32_t *n_ent,
                                 int32_t *nbr_elem_out)
{
   const Mesh &m = mm->m;
   const int dim = m.dim, nf = 2 * dim, n = p + 1, nfc = 1 << (dim - 1);
   int nd = 1, nfd = 1;
   for (int a = 0; a < dim; a++) { nd *= n; }
   for (int a = 0; a < dim - 1; a++) { nfd *= n; }
   if ((double)m.ne * nd >= 2147483647.0) { set_error("dof_maps: int32 overflow"); return 1; }
   std::vector<int> bd;
   bdr_dofs(p, dim, bd);
   if (bdr_out) { for (size_t i = 0; i < bd.size(); i++) { bdr_out[i] = bd[i]; } }
   if (s2i_out && p >= 1)
   {
      std::vector<int> s;
      sub2ind(p, dim, s);
      for (size_t i = 0; i < s.size(); i++) { s2i_out[i] = s[i]; }
   }
   if (lat_out || n_ent)
   {
      std::vector<int64_t> lat;
      const int64_t cnt = macro_lattice(m, lat);
      if (cnt >= 2147483647LL) { set_error("dof_maps: entity count overflows int32"); return 1; }
      if (lat_out) { for (size_t i = 0; i < lat.size(); i++) { lat_out[i] = (int32_t)lat[i]; } }
      if (n_ent) { *n_ent = (int32_t)cnt; }
   }
   if (!nbr_dof && !nbr_elem_out) { return 0; }
   Topology T;
   if (build_topology(m, T)) { return 1; }
   if (nbr_elem_out)
   { for (size_t i = 0; i < T.nbr_elem.size(); i++) { nbr_elem_out[i] = (int32_t)T.nbr_elem[i]; } }
   if (!nbr_dof) { return 0; }
   int fc[6][4];
   for (int f = 0; f < nf; f++) { face_corners(dim, f, fc[f]); }
   for (int64_t e = 0; e < m.ne; e++)
      for (int f = 0; f < nf; f++)
      {
         int32_t *dst = &nbr_dof[((size_t)e * nf + f) * nfd];
         const int64_t e2 = T.nbr_elem[e * nf + f];
         if (e2 < 0) { for (int j = 0; j < nfd; j++) { dst[j] = -1; } continue; }
         const int f2 = (int)T.nbr_face[e * nf + f];
         const int8_t *fm = &T.fmap[((size_t)e * nf + f) * nfc];
         int axis, side;
         face_axis(dim, f, axis, side);
         int rem[2], nr = 0;
         for (int a = 0; a < dim; a++) { if (a != axis) { rem[nr++] = a; } }
         // image of own face corner 0 and of the own face axes in the neighbour's lattice
         int o[3], d[2][3];
         const int c0 = fc[f2][fm[0]];
         for (int a = 0; a < dim; a++) { o[a] = ((c0 >> a) & 1) * p; }
         for (int mx = 0; mx < dim - 1; mx++)
         {
            const int c1 = fc[f2][fm[1 << mx]];
            for (int a = 0; a < dim; a++) { d[mx][a] = ((c1 >> a) & 1) - ((c0 >> a) & 1); }
         }
         for (int j = 0; j < nfd; j++)
         {
            const int own = bd[j * nf + f];
            int l[3], q = own;
            for (int a = 0; a < dim; a++) { l[a] = q % n; q /= n; }
            int loc = 0;
            for (int a = dim - 1; a >= 0; a--)
            {
               int v = o[a];
               for (int mx = 0; mx < dim - 1; mx++) { v += l[rem[mx]] * d[mx][a]; }
               loc = loc * n + v;
            }
            dst[j] = (int32_t)(e2 * nd + loc);
         }
      }
   return 0;
}
