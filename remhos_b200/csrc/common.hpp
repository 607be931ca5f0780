// Shared host-side helpers of remhos_b200 (error reporting, 1-D finite-element tables).
#ifndef RMH_COMMON_HPP
#define RMH_COMMON_HPP

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

} // namespace rmh

#endif
