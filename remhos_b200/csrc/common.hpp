// Shared host-side helpers of remhos_b200 (error reporting, 1-D finite-element tables).
#ifndef RMH_COMMON_HPP
#define RMH_COMMON_HPP

#include <cstdint>
#include <string>
#include <vector>

namespace rmh
{

void set_error(const std::string &msg);

// n-point Gauss-Legendre rule on [0,1] (IntRules.Get(Segment, 2n-1))
void gauss_legendre_01(int n, std::vector<double> &x, std::vector<double> &w);
// n Gauss-Lobatto points on [0,1]
std::vector<double> gauss_lobatto_01(int n);
// L[q*n + i] = l_i(x_q), Lagrange basis on `nodes`
std::vector<double> lagrange(const std::vector<double> &nodes, const std::vector<double> &x);
std::vector<double> lagrange_deriv(const std::vector<double> &nodes, const std::vector<double> &x);
// B[q*(p+1) + i] = B_i^p(x_q) and its derivative (Bernstein / "Positive" basis,
// DG_FECollection(p, dim, BasisType::Positive), remhos.cpp:588-590)
std::vector<double> bernstein(int p, const std::vector<double> &x);
std::vector<double> bernstein_deriv(int p, const std::vector<double> &x);
// dense inverse of a small n x n matrix (Gauss-Jordan with partial pivoting)
std::vector<double> invert_small(const std::vector<double> &A, int n);

void face_axis(int dim, int f, int &axis, int &side);
void bdr_dofs(int p, int dim, std::vector<int> &bd);
// Sub2Ind [p^dim][2^dim]: DG dofs of the lexicographic corners of every subcell (remhos_tools.cpp:678-734)
void sub2ind(int p, int dim, std::vector<int> &s);
// natural face order (remaining axes ascending, first fastest) -> row of BdrDofs: [nf][nfd]
void nat2ref_table(int p, int dim, std::vector<int> &n2r);
// rmh_nbr_lattice restricted to the first ne_rows elements (the owned ones of a decomposed mesh;
// lat still lists owned + ghost rows, ne_all of them)
int nbr_lattice_rows(int dim, int64_t ne_all, int64_t ne_rows, int32_t n_ent, const int32_t *lat,
                     int32_t *nbr, int *structured);

} // namespace rmh

// Halo plan of one rank (rmh_halo_create): global element ids.  `owned` may be reordered by
// rmh_halo_interior_first; owned_sorted / owned_pos locate a global id among the owned elements.
struct rmh_halo
{
   std::vector<int64_t> owned, ghost, send;         // global element ids
   std::vector<int32_t> ghost_owner, peers, send_off, recv_off;
   std::vector<int64_t> owned_sorted;               // ascending global ids
   std::vector<int32_t> owned_pos;                  // position in `owned` of owned_sorted[i]
   int64_t n_interior = -1;
   int32_t local_of(int64_t g) const;               // -1 if not owned
};

#endif
