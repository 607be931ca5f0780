// remhos_b200 device context, kernels' entry points and the C ABI of the RK-stage path.
// See include/remhos_b200.h for the reference method each entry point replaces.
#include "../../include/remhos_b200.h"
#include "common.hpp"
#include "kernels.cuh"
#include "stage3d.cuh"
#include "stage3p.cuh"
#include "stage3w.cuh"
#include "stage3c.cuh"
#include "fa.cuh"
#include "product.cuh"
#include "geom3.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace rmh;

#define RMH_MAX_PEERS 32      // epoch flags in a rank's window (one per peer)
#define RMH_MAX_DEVICES 16

static std::atomic<int64_t> g_launches{0};

#define CUDA_OK(call)                                                                  \
   do {                                                                                \
      cudaError_t err__ = (call);                                                      \
      if (err__ != cudaSuccess)                                                        \
      {                                                                                \
         set_error(std::string(#call) + ": " + cudaGetErrorString(err__));             \
         return 1;                                                                     \
      }                                                                                \
   } while (0)

#define LAUNCH_OK()                                                                    \
   do {                                                                                \
      g_launches++;                                                                    \
      cudaError_t err__ = cudaGetLastError();                                          \
      if (err__ != cudaSuccess)                                                        \
      {                                                                                \
         set_error(std::string("kernel launch: ") + cudaGetErrorString(err__));        \
         return 1;                                                                     \
      }                                                                                \
   } while (0)

// ============================================================================ context
struct HostPipe;
// remap mode: the time-dependent operator data of ONE other time (see rmh_set_time)
struct GeomSet
{
   double *Dvol = nullptr, *detJw = nullptr, *Dface = nullptr, *ml = nullptr, *einv = nullptr, *BL = nullptr;
   double t = 0.0;
   bool valid = false;
};
struct rmh_ctx
{
   int dim, p, mo, exec_mode, bounds_type, device;
   int D1, Q, ND, NQ, NF, NFD, NQF, NG1, NGN, N3;
   int64_t ne, ne_ghost, N;
   double t_cur;
   // host tables
   std::vector<double> hB, hG, hMinv, hw, hxq;
   // device tables (generic kernels)
   double *dB = nullptr, *dw = nullptr, *dL = nullptr, *ddL = nullptr, *dLs = nullptr,
          *ddLs = nullptr;
   // geometry inputs
   double *X0 = nullptr, *V = nullptr, *velq = nullptr, *velf = nullptr;
   // operator (PA) data
   double *Dvol = nullptr, *detJw = nullptr, *Dface = nullptr, *ml = nullptr, *inflow = nullptr;
   double *einv = nullptr;   // [ne] 1/volume for elements with constant Jacobian determinant, else 0
   // face neighbours
   int32_t *nbr_elem = nullptr;
   uint8_t *nbr_pat = nullptr;
   int32_t *nbr_pat32 = nullptr;   // same ids, cp.async granularity (pipelined stage kernel)
   int16_t *pat = nullptr;
   int npat = 0;
   // bounds
   int32_t *lat = nullptr, *ent_off = nullptr, *ent_el = nullptr, *bnbr = nullptr;
   int32_t n_ent = 0;
   double *ent_mm = nullptr;   // [n_ent][2] (min, max) over the elements sharing the entity
   double *xe_min = nullptr, *xe_max = nullptr;
   double2 *xe_mm = nullptr;  // interleaved copy for the entity pass
   int num_sms = 0;
   bool pipelined = true;      // RMH_NO_PIPELINE=1 selects the one-batch-per-block stage kernel
   bool tensor = true;         // RMH_NO_TENSOR=1 selects the DFMA pipelined kernel
   bool frag = false;          // Dvol/Dface stored in the fragment order of stage3w.cuh
   bool xe_valid = false;
   bool all_affine = false;   // every element has constant det J (transport meshes only)
   bool op_lin = false;       // ... and adj(J) v is linear over every element: opc is valid
   bool op_const = false;     // ... and constant over every element: opa is valid
   // decomposed meshes: owned elements [0, n_split) share no vertex with a ghost element
   int64_t n_split = -1;
   // smoothness indicator (rmh_si_setup): H1 order-1 operators in CSR
   int si_type = 0, si_n = 0;
   double si_param = 0.0;
   int32_t *si_MI = nullptr, *si_MJ = nullptr, *si_LI = nullptr, *si_LJ = nullptr, *si_XI = nullptr,
           *si_XJ = nullptr, *si_d2c = nullptr;
   double *si_MA = nullptr, *si_LA = nullptr, *si_XA = nullptr, *si_ml = nullptr, *si_y = nullptr,
          *si_z = nullptr, *si_r = nullptr, *si_val = nullptr, *si_tmp = nullptr, *si_nrm = nullptr;
   int dt_control = 0;        // rmh_dt_control: 1 = LO bounds error (remhos.cpp:312-316)
   double *dt_ratio = nullptr;   // device scalar, min over the LimitMult calls since the last reset
   int mono_type = 0;         // rmh_mono_setup: 1 MonoRDSolver, 2 with subcells; 0 none
   int mono_mass_lim = 1;
   double *mono_scale = nullptr;   // [ne] (remhos_mono.cpp:40-57)
   bool trust_state = false;  // rmh_ctx_trust_state: the caller leaves the state alone between steps
   const double *xe_ptr = nullptr;   // state vector whose element min/max the context currently holds
   double *opc = nullptr;     // [ne][12] (k_op_linear)
   double *opa = nullptr;     // [ne][4]  (k_op_linear)
   double *dxq = nullptr;     // device copy of the 1-D quadrature points
   // matrix-based ("FA") solver data: lumped face matrices always; dense blocks after rmh_fa_setup
   double *dG = nullptr, *BL = nullptr;
   double *faK = nullptr, *faKH = nullptr, *faM = nullptr, *faBI = nullptr;
   // decomposed FluxBasedFCT: neighbour-side blocks of the ghost faces, private copies of ghost traces (u, R+, R-)
   double *faBIg = nullptr, *gtr_u = nullptr, *gtr_cp = nullptr, *gtr_cn = nullptr;
   double *faKP = nullptr;     // M_L M^-1 K (PrecondConvectionIntegrator), valid at time kp_t
   double kp_t = -1.0e300;
   int16_t *pat_idx = nullptr;
   uint8_t *pat_face = nullptr;
   bool fa_on = false;
   // subcell residual distribution (-lo 4): lattice points, velocity samples, weights
   double *sub_x = nullptr, *sub_v = nullptr, *sub_w = nullptr;
   bool sub_on = false;
   // product remap (-ps): the state of rmh_mult* / rmh_limit_mult / rmh_ode_step is the block (u, us)
   bool product = false, idp_mask = false;
   double *pw[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   uint8_t *pf_el[2] = {nullptr, nullptr}, *pf_dof[2] = {nullptr, nullptr}, *pmask = nullptr;
   // work vectors of the unfused solver path (allocated on first use)
   double *wk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   double *rk[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   // halo (multi-GPU, dist.cuh).  Faces with a ghost neighbour get one slot each: nbr_elem = ne + slot,
   // the neighbour's face trace (this element's natural face order) is ughost[slot][NFD].
   const double *ughost = nullptr;
   const double *halo_ptr = nullptr;    // state whose ghost traces / (min,max) the window holds (unfused path)
   struct HostPipe *pipe = nullptr;      // rmh_rk_step_host_async
   GeomSet gspare;                       // remap: second set of operator data (rmh_set_time)
   size_t n_dvol = 0, n_dface = 0;
   bool geom_ok = false;                 // the active set holds the operators of t_cur
   // fused halo send (dist.cuh): the stage kernel that wrote sent_ptr has already stored its halo for epoch sent_epoch
   bool send_next = false;
   const double *sent_ptr = nullptr;
   unsigned long long sent_epoch = 0;
   int64_t n_gslots = 0;
   std::vector<int32_t> gs_ghost, gs_pid;   // per slot: ghost element (0 .. ne_ghost-1), pattern id
   std::vector<int16_t> pat_h;              // host copy of the pattern table [npat][NFD]
   // k_stage3c<FOLD>: 3^dim neighbourhood of every owned element (rmh_nbr_lattice; boundary -> the
   // sentinel pair); element (min,max) pairs ping-pong between xe_mm2[0|1], each [ne + ne_ghost + 1]
   int32_t *nb27 = nullptr;
   bool fold = false;
   double2 *xe_mm2[2] = {nullptr, nullptr};
   double *zeros = nullptr;              // 64 zeros
   unsigned long long epoch = 0;   // stage counter: stage k reads pairs / ghost traces [k & 1], writes pairs [(k+1) & 1]
   // the window peers write into (one allocation = one IPC handle):
   // flags[RMH_MAX_PEERS] | xe_mm2[0] | xe_mm2[1] | gtr[0] | gtr[1]
   void *win = nullptr;
   size_t win_bytes = 0, win_off_mm[2] = {0, 0}, win_off_tr[2] = {0, 0};
   unsigned long long *flags = nullptr;
   double *gtr[2] = {nullptr, nullptr};
   struct rmh_dist *dist = nullptr;
   // scratch
   double *w1 = nullptr, *w2 = nullptr, *w3 = nullptr, *red = nullptr;
   double *pin = nullptr;   // pinned host staging (e2e entry point)
   std::vector<double> x0_e0, v_e0;   // nodes / node velocities of element 0 (Mesh::GetElementSize(0, 0))
   // optional per-launch timing of the fused stage kernel (bench.py roofline)
   bool prof = false;
   std::vector<cudaEvent_t> prof_ev;
   size_t prof_used = 0;
   double pcg_tol2 = 1e-28;
   int pcg_maxit = 60;
   std::vector<void *> allocs;
};

namespace rmh { struct StagePArgs; }
// multi-GPU hook (dist.cuh): flags / epoch / shell range of the in-kernel halo wait
static void dist_stage_args(rmh_ctx *c, rmh::StagePArgs &pa, bool in_kernel_wait);
// decomposed meshes, unfused solver path: element min/max of u, then face traces + (min,max) pairs of u
// exchanged with the peers (one put kernel + k_halo_wait); a no-op on a single rank
static int dist_halo(rmh_ctx *c, const double *u, cudaStream_t s);
// face traces of any DOF vector exchanged with the peers and copied into dst [n_gslots][NFD]; the window no
// longer holds the halo of the state afterwards
static int dist_traces(rmh_ctx *c, const double *vec, double *dst, cudaStream_t s);

template <typename Tp>
static int dev_alloc(rmh_ctx *c, Tp **p, size_t n)
{
   void *q = nullptr;
   if (n == 0) { n = 1; }
   cudaError_t e = cudaMalloc(&q, n * sizeof(Tp));
   if (e != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return 1; }
   c->allocs.push_back(q);
   *p = (Tp *)q;
   return 0;
}
template <typename Tp>
static int dev_upload(rmh_ctx *c, Tp **p, const Tp *h, size_t n)
{
   if (dev_alloc(c, p, n)) { return 1; }
   CUDA_OK(cudaMemcpy(*p, h, n * sizeof(Tp), cudaMemcpyHostToDevice));
   return 0;
}

// ============================================================================ kernels
struct GeomArgs
{
   int dim, Q, NG1, exec_mode;
   int frag;
   int64_t ne;
   double t;
   const double *X0, *V, *velq, *velf, *L, *dL, *Ls, *dLs, *w;
   double *Dvol, *detJw, *Dface;
};

template <int DIM>
__device__ __forceinline__ void det_adj(const double (&J)[3][3], double &det, double (&adj)[3][3])
{
   if (DIM == 2)
   {
      det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      adj[0][0] = J[1][1]; adj[0][1] = -J[0][1];
      adj[1][0] = -J[1][0]; adj[1][1] = J[0][0];
   }
   else
   {
      adj[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      adj[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
      adj[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
      adj[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      adj[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
      adj[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
      adj[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      adj[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
      adj[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      det = J[0][0] * adj[0][0] + J[0][1] * adj[1][0] + J[0][2] * adj[2][0];
   }
}

// Jacobian (and optionally the interpolated nodal velocity) at a reference point given by its
// per-axis 1-D basis rows l[a][.], dl[a][.]
template <int DIM>
__device__ __forceinline__ void eval_geom(const GeomArgs &g, int64_t e, const double *const *l,
                                          const double *const *dl, double (&J)[3][3],
                                          double (&v)[3], bool need_v)
{
   const int n1 = g.NG1;
   int nn = 1;
   for (int a = 0; a < DIM; a++) { nn *= n1; }
   for (int i = 0; i < 3; i++) { v[i] = 0.0; for (int j = 0; j < 3; j++) { J[i][j] = 0.0; } }
   const double *X = g.X0 + (size_t)e * nn * DIM;
   const double *V = g.V ? g.V + (size_t)e * nn * DIM : nullptr;
   for (int n = 0; n < nn; n++)
   {
      int idx[3] = {0, 0, 0}, m = n;
      for (int a = 0; a < DIM; a++) { idx[a] = m % n1; m /= n1; }
      double x[3];
      for (int i = 0; i < DIM; i++)
      {
         x[i] = X[n * DIM + i];
         if (g.exec_mode == 1) { x[i] += g.t * V[n * DIM + i]; }
      }
      double lv = 1.0;
      for (int a = 0; a < DIM; a++) { lv *= l[a][idx[a]]; }
      for (int j = 0; j < DIM; j++)
      {
         double d = 1.0;
         for (int a = 0; a < DIM; a++) { d *= (a == j) ? dl[a][idx[a]] : l[a][idx[a]]; }
         for (int i = 0; i < DIM; i++) { J[i][j] += d * x[i]; }
      }
      if (need_v) { for (int i = 0; i < DIM; i++) { v[i] += lv * V[n * DIM + i]; } }
   }
}

template <int DIM>
__global__ void k_geom_vol(GeomArgs g)
{
   const int Q = g.Q;
   int NQ = 1;
   for (int a = 0; a < DIM; a++) { NQ *= Q; }
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= g.ne * NQ) { return; }
   const int64_t e = idx / NQ;
   const int q = (int)(idx - e * NQ);
   int qa[3] = {0, 0, 0}, m = q;
   double w = 1.0;
   const double *l[3], *dl[3];
   for (int a = 0; a < DIM; a++)
   {
      qa[a] = m % Q; m /= Q;
      w *= g.w[qa[a]];
      l[a] = g.L + qa[a] * g.NG1;
      dl[a] = g.dL + qa[a] * g.NG1;
   }
   double J[3][3], v[3], det, adj[3][3];
   const bool nodal_v = (g.velq == nullptr);
   eval_geom<DIM>(g, e, l, dl, J, v, nodal_v);
   det_adj<DIM>(J, det, adj);
   if (!nodal_v) { for (int i = 0; i < DIM; i++) { v[i] = g.velq[((size_t)e * NQ + q) * DIM + i]; } }
   const double alpha = (g.exec_mode == 1) ? 1.0 : -1.0;   // remhos.cpp:648-657
   for (int c = 0; c < DIM; c++)
   {
      double s = 0.0;
      for (int j = 0; j < DIM; j++) { s += adj[c][j] * v[j]; }
      if (g.frag)
      {
         // [e][col = qy*Q+qx][qz (RQ)][3]
         const int RQ = (Q + 1) & ~1, QQ = Q * Q;
         const int col = q % QQ, qz = q / QQ;
         g.Dvol[(size_t)e * QQ * RQ * 3 + ((size_t)col * RQ + qz) * 3 + c] = alpha * w * s;
      }
      else { g.Dvol[((size_t)e * DIM + c) * NQ + q] = alpha * w * s; }
   }
   g.detJw[(size_t)e * NQ + q] = w * det;
}

template <int DIM>
__global__ void k_geom_face(GeomArgs g)
{
   const int Q = g.Q, NF = 2 * DIM;
   int NQF = 1;
   for (int a = 0; a < DIM - 1; a++) { NQF *= Q; }
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= g.ne * NF * NQF) { return; }
   const int64_t e = idx / (NF * NQF);
   const int r = (int)(idx - e * NF * NQF), f = r / NQF, qf = r - f * NQF;
   int axis, side;
   face_axis_side(DIM, f, axis, side);
   const double *l[3], *dl[3];
   double w = 1.0;
   int m = qf;
   for (int a = 0; a < DIM; a++)
   {
      if (a == axis) { l[a] = g.Ls + side * g.NG1; dl[a] = g.dLs + side * g.NG1; }
      else
      {
         const int q = m % Q; m /= Q;
         w *= g.w[q];
         l[a] = g.L + q * g.NG1; dl[a] = g.dL + q * g.NG1;
      }
   }
   double J[3][3], v[3], det, adj[3][3];
   const bool nodal_v = (g.velf == nullptr);
   eval_geom<DIM>(g, e, l, dl, J, v, nodal_v);
   det_adj<DIM>(J, det, adj);
   if (!nodal_v) { for (int i = 0; i < DIM; i++) { v[i] = g.velf[(size_t)idx * DIM + i]; } }
   const double sgn = side ? 1.0 : -1.0;
   double vn = 0.0;
   for (int i = 0; i < DIM; i++) { vn += v[i] * sgn * adj[axis][i]; }
   // upwinded normal velocity (remhos_tools.cpp:833-845): transport min(0, v.n), remap -max(0, v.n)
   const double vs = (g.exec_mode == 1) ? -fmax(0.0, vn) : fmin(0.0, vn);
   if (DIM == 3)
   {
      // 3D layout [e][qb][f][qa] (coalesced for the (qa, f) thread mapping of stage3d.cuh)
      const int qa = qf % Q, qb = qf / Q;
      if (g.frag)
      {
         const int RQ = (Q + 1) & ~1;     // [e][f][qa][qb (RQ)]
         g.Dface[(size_t)e * NF * Q * RQ + ((size_t)f * Q + qa) * RQ + qb] = w * vs;
      }
      else { g.Dface[(size_t)e * NF * NQF + (size_t)qb * NF * Q + f * Q + qa] = w * vs; }
   }
   else { g.Dface[idx] = w * vs; }
}

// per element: einv = 1/volume if det J is constant over the element (then M = vol * M_ref
// exactly and the Kronecker inverse is the exact mass inverse), else 0
__global__ void k_elem_affine(int dim, int Q, int ngn, int exec_mode, double t, int64_t ne,
                              const double *w, const double *detJw, const double *X0,
                              const double *V, double *einv)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= ne) { return; }
   int NQ = 1;
   for (int a = 0; a < dim; a++) { NQ *= Q; }
   const double *d = detJw + (size_t)e * NQ;
   double vol = 0.0;
   for (int q = lane; q < NQ; q += 32) { vol += d[q]; }
   vol = warp_sum(vol);
   double dev = 0.0;
   for (int q = lane; q < NQ; q += 32)
   {
      double wq = 1.0;
      int m = q;
      for (int a = 0; a < dim; a++) { wq *= w[m % Q]; m /= Q; }
      dev = fmax(dev, fabs(d[q] / wq - vol));
   }
   dev = warp_max(dev);
   // the Jacobian is a difference of coordinates of size |x| over a distance h, so det J carries
   // a relative round-off of ~eps*|x|/h: deviations below that are not geometry, they are noise
   double xmax = 0.0;
   for (int n = lane; n < ngn * dim; n += 32)
   {
      double x = X0[(size_t)e * ngn * dim + n];
      if (exec_mode == 1) { x += t * V[(size_t)e * ngn * dim + n]; }
      xmax = fmax(xmax, fabs(x));
   }
   xmax = warp_max(xmax);
   const double h = pow(fabs(vol), 1.0 / dim);
   const double tol = 100.0 * 2.220446049250313e-16 * fmax(1.0, xmax / h);
   if (lane == 0) { einv[e] = (vol > 0.0 && dev <= tol * vol) ? 1.0 / vol : 0.0; }
}

// Linear-operator detection for the tensor-core stage kernel (fragment-ordered 3D transport data).
// One warp per element: least-squares fit (quadrature-weighted, so the basis 1, x-1/2, y-1/2, z-1/2
// is orthogonal) of adj(J) v = W0 + W1 x + W2 y + W3 z in reference coordinates to the stored
// Dvol = -w_q adj(J) v, then check that the fit reproduces every stored volume value and every
// stored face value w min(0, +-(adj(J) v)_axis).  Tolerance: the stored data carries the round-off
// of a Jacobian formed from coordinate differences, eps*|x|/h relative (see k_elem_affine).
// opc[e][k*3 + c] = W_k[c]; nfail[0] counts the elements that do not fit.  Elements whose slopes
// W1..W3 vanish to the same tolerance carry a constant coefficient a = W at the centroid:
// opa[e] = (a_x, a_y, a_z, 1/vol); nfail[1] counts the elements that are not of that kind.
__global__ void k_op_linear(int Q, int ngn, int64_t ne, const double *xq, const double *w,
                            const double *Dvol, const double *Dface, const double *X0,
                            const double *einv, double *opc, double *opa, unsigned int *nfail)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= ne) { return; }
   const int RQ = (Q + 1) & ~1, QQ = Q * Q;
   const double *dv = Dvol + (size_t)e * QQ * RQ * 3;
   const double *df = Dface + (size_t)e * 6 * Q * RQ;
   double acc[12], s2 = 0.0;
   for (int i = 0; i < 12; i++) { acc[i] = 0.0; }
   for (int q = 0; q < Q; q++) { s2 += w[q] * (xq[q] - 0.5) * (xq[q] - 0.5); }
   for (int q = lane; q < QQ * Q; q += 32)
   {
      const int qx = q % Q, qy = (q / Q) % Q, qz = q / QQ;
      for (int c = 0; c < 3; c++)
      {
         // stored value = -w_q W_c(q): the quadrature weight is already on it
         const double wv = -dv[((size_t)(qy * Q + qx) * RQ + qz) * 3 + c];
         acc[c] += wv;
         acc[3 + c] += wv * (xq[qx] - 0.5);
         acc[6 + c] += wv * (xq[qy] - 0.5);
         acc[9 + c] += wv * (xq[qz] - 0.5);
      }
   }
   double cf[12];
   for (int i = 0; i < 12; i++) { acc[i] = warp_sum(acc[i]); }
   for (int c = 0; c < 3; c++)
   {
      cf[3 + c] = acc[3 + c] / s2; cf[6 + c] = acc[6 + c] / s2; cf[9 + c] = acc[9 + c] / s2;
      cf[c] = acc[c] - 0.5 * (cf[3 + c] + cf[6 + c] + cf[9 + c]);
   }
   double scale = 0.0, dev = 0.0;
   for (int q = lane; q < QQ * Q; q += 32)
   {
      const int qx = q % Q, qy = (q / Q) % Q, qz = q / QQ;
      for (int c = 0; c < 3; c++)
      {
         const double v = -dv[((size_t)(qy * Q + qx) * RQ + qz) * 3 + c] / (w[qx] * w[qy] * w[qz]);
         const double fit = cf[c] + cf[3 + c] * xq[qx] + cf[6 + c] * xq[qy] + cf[9 + c] * xq[qz];
         scale = fmax(scale, fabs(v));
         dev = fmax(dev, fabs(v - fit));
      }
   }
   for (int i = lane; i < 6 * QQ; i += 32)
   {
      const int f = i / QQ, qa = i % Q, qb = (i / Q) % Q;
      const int axis = (f == 0 || f == 5) ? 2 : ((f == 1 || f == 3) ? 1 : 0);
      const int side = (f == 2 || f == 3 || f == 5) ? 1 : 0;
      const int ia = (axis == 0) ? 1 : 0, ib = (axis == 2) ? 1 : 2;
      double vn = cf[axis] + side * cf[3 * (1 + axis) + axis] + cf[3 * (1 + ia) + axis] * xq[qa] +
                  cf[3 * (1 + ib) + axis] * xq[qb];
      if (!side) { vn = -vn; }
      const double stored = df[((size_t)f * Q + qa) * RQ + qb] / (w[qa] * w[qb]);
      dev = fmax(dev, fabs(stored - fmin(0.0, vn)));
   }
   double xmax = 0.0;
   for (int n = lane; n < ngn * 3; n += 32) { xmax = fmax(xmax, fabs(X0[(size_t)e * ngn * 3 + n])); }
   scale = warp_max(scale);
   dev = warp_max(dev);
   xmax = warp_max(xmax);
   const double h = cbrt(1.0 / einv[e]);
   const double tol = 100.0 * 2.220446049250313e-16 * fmax(1.0, xmax / h);
   if (lane < 12) { opc[e * 12 + lane] = cf[lane]; }
   if (lane == 0)
   {
      const bool lin = (dev <= tol * scale);
      double slope = 0.0;
      for (int i = 3; i < 12; i++) { slope = fmax(slope, fabs(cf[i])); }
      if (!lin) { atomicAdd(nfail, 1u); }
      if (!(lin && slope <= tol * scale)) { atomicAdd(nfail + 1, 1u); }
      opa[e * 4 + 0] = acc[0]; opa[e * 4 + 1] = acc[1]; opa[e * 4 + 2] = acc[2]; opa[e * 4 + 3] = einv[e];
   }
}

// lumped mass m_i = sum_q B_qi detJw_q  (M_HO * 1, remhos.cpp:721-727)
__global__ void k_lumped_mass(int dim, int D1, int Q, int64_t ne, const double *B,
                              const double *detJw, double *ml)
{
   int ND = 1, NQ = 1;
   for (int a = 0; a < dim; a++) { ND *= D1; NQ *= Q; }
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= ne * ND) { return; }
   const int64_t e = idx / ND;
   const int i = (int)(idx - e * ND);
   int ia[3] = {0, 0, 0}, m = i;
   for (int a = 0; a < dim; a++) { ia[a] = m % D1; m /= D1; }
   const double *d = detJw + (size_t)e * NQ;
   double s = 0.0;
   if (dim == 2)
   {
      for (int qy = 0; qy < Q; qy++)
      {
         double sx = 0.0;
         for (int qx = 0; qx < Q; qx++) { sx += B[qx * D1 + ia[0]] * d[qy * Q + qx]; }
         s += B[qy * D1 + ia[1]] * sx;
      }
   }
   else
   {
      for (int qz = 0; qz < Q; qz++)
      {
         double sy = 0.0;
         for (int qy = 0; qy < Q; qy++)
         {
            double sx = 0.0;
            for (int qx = 0; qx < Q; qx++) { sx += B[qx * D1 + ia[0]] * d[(qz * Q + qy) * Q + qx]; }
            sy += B[qy * D1 + ia[1]] * sx;
         }
         s += B[qz * D1 + ia[2]] * sy;
      }
   }
   ml[idx] = s;
}

// ---------------------------------------------------------------- HO kernel (a) + (b)
struct HoArgs
{
   int64_t ne;
   const double *u;       // input: solution (mode has MULT) or rhs (mode == SOLVE only)
   double *out;
   const double *Dvol, *detJw, *Dface, *einv;
   FaceNbr fn;
   int mode;              // bit 0: apply K_HO, bit 1: apply M^-1, bit 2: volume terms only
   int frag;              // stored quadrature data in fragment order
   double tol2;
   int maxit;
};

// block shape / shared-memory plan: 3D uses the (a,b)-mapped fast path of stage3d.cuh
template <int DIM, int D1, int Q, int E>
struct KCfg
{
   using S2 = Smem<DIM, D1, Q, E>;
   using S3 = Smem3<D1, Q, E>;
   static constexpr int T = (DIM == 3) ? S3::T : 32 * E;
   static constexpr size_t BYTES = (DIM == 3) ? S3::BYTES : S2::BYTES;
   static constexpr int OFF_U = (DIM == 3) ? S3::OFF_U : S2::OFF_U;
   static constexpr int OFF_R = (DIM == 3) ? S3::OFF_R : S2::OFF_R;
   static constexpr int OFF_X = (DIM == 3) ? S3::OFF_X : S2::OFF_X;
   // resident blocks per SM the register allocation must allow (shared memory permitting)
#ifndef RMH_MINB
#define RMH_MINB 2
#endif
   static constexpr int MINB = (DIM == 3 && BYTES * RMH_MINB <= 200 * 1024) ? RMH_MINB : 1;
};

// (a) + (b) on the block's element batch: U (smem) -> R = K_HO u -> X = M^-1 R; returns result ptr
template <int DIM, int D1, int Q, int E, bool AFF = false>
__device__ __forceinline__ double *ho_phases(const HoArgs &a, double *sm, int64_t e0, int ne,
                                             const Tab<D1, Q> &tab);

template <int DIM, int D1, int Q, int E>
__global__ void __launch_bounds__(KCfg<DIM, D1, Q, E>::T) k_ho(HoArgs a, const Tab<D1, Q> tab)
{
   using K = KCfg<DIM, D1, Q, E>;
   constexpr int T = K::T, ND = ipow(D1, DIM);
   extern __shared__ double sm[];
   const int64_t e0 = (int64_t)blockIdx.x * E;
   const int ne = (int)min((int64_t)E, a.ne - e0);
   double *U = sm + K::OFF_U;
   for (int t = threadIdx.x; t < E * ND; t += T) { U[t] = (t < ne * ND) ? a.u[e0 * ND + t] : 0.0; }
   const double *res = ho_phases<DIM, D1, Q, E>(a, sm, e0, ne, tab);
   for (int t = threadIdx.x; t < ne * ND; t += T) { a.out[e0 * ND + t] = res[t]; }
}

template <int DIM, int D1, int Q, int E, bool AFF>
__device__ __forceinline__ double *ho_phases(const HoArgs &a, double *sm, int64_t e0, int ne,
                                             const Tab<D1, Q> &tab)
{
   using K = KCfg<DIM, D1, Q, E>;
   constexpr int T = K::T, ND = ipow(D1, DIM), NQ = ipow(Q, DIM), NF = 2 * DIM,
                 NQF = ipow(Q, DIM - 1);
   double *U = sm + K::OFF_U, *R = sm + K::OFF_R, *X = sm + K::OFF_X;
   double *res = R;
   if constexpr (DIM == 3)
   {
      Pre3<D1, Q, E> pre;
      if constexpr (AFF)
      {
         // every element of the mesh has constant det J (checked on the host at set-up)
         pre.load(a.Dvol, a.Dface, e0, ne, a.frag != 0);
         face3_gather<D1, Q, E>(sm, a.u, a.fn, e0, ne);
         __syncthreads();
         ho3_affine<D1, Q, E>(U, X, sm, pre, a.einv + e0, ne, tab);
         res = X;
      }
      else
      {
         if (a.mode & 1)
         {
            pre.load(a.Dvol, a.Dface, e0, ne, a.frag != 0);
            if (!(a.mode & 4)) { face3_gather<D1, Q, E>(sm, a.u, a.fn, e0, ne); }
         }
         __syncthreads();
         if ((a.mode & 1) && (a.mode & 4)) { vol3_apply<D1, Q, E, false>(U, R, sm, pre, tab); }
         else if (a.mode & 1)
         {
            face3_apply<D1, Q, E>(sm, pre, tab);
            vol3_apply<D1, Q, E, true>(U, R, sm, pre, tab);
         }
         else
         {
            for (int i = threadIdx.x; i < E * ND; i += T) { R[i] = U[i]; }
            __syncthreads();
         }
         if (a.mode & 2)
         {
            mass3_solve<D1, Q, E>(R, X, sm, a.detJw + (size_t)e0 * NQ, a.einv + e0, ne, a.tol2,
                                  a.maxit, tab);
            res = X;
         }
      }
   }
   else
   {
      __syncthreads();
      if (a.mode & 1)
      {
         vol_apply<DIM, D1, Q, E>(U, R, sm, a.Dvol + (size_t)e0 * DIM * NQ, ne, tab);
         if (!(a.mode & 4))
         {
            face_apply<DIM, D1, Q, E>(U, R, sm, a.u, a.Dface + (size_t)e0 * NF * NQF, a.fn, e0, ne, tab);
         }
      }
      else
      {
         for (int i = threadIdx.x; i < E * ND; i += T) { R[i] = U[i]; }
         __syncthreads();
      }
      if (a.mode & 2)
      {
         mass_solve<DIM, D1, Q, E>(R, X, sm, a.detJw + (size_t)e0 * NQ, ne, a.tol2, a.maxit, tab);
         res = X;
      }
   }
   return res;
}

__device__ __forceinline__ int lattice_class(int dim, int D1, int i)
{
   int t = 0, mul = 1, m = i;
   for (int a = 0; a < dim; a++)
   {
      const int l = m % D1; m /= D1;
      const int c = (l == 0) ? 0 : ((l == D1 - 1) ? 2 : 1);
      t += c * mul; mul *= 3;
   }
   return t;
}

// ---------------------------------------------------- fused RK-stage kernel (a)-(e)
// One launch = HO advection action + element mass solve + MassBasedAvg LO + per-DOF bounds
// gather + ClipScale + RK combination (+ element min/max of the output for the next stage).
struct StageArgs
{
   HoArgs ho;                 // ho.u = stage input y; ho.out unused
   const double *ml;          // lumped mass
   const double *x0;          // RK base state (may alias ho.u)
   double *out;               // out_mode 0: k = F(y); 1: a*x0 + b*(y + dt*k)
   double a, b, dt;
   int out_mode;
   int bounds_type, dim_n3;
   const int32_t *lat;        // [NE][3^dim]
   const double *ent_mm;      // [n_ent][2]
   const int32_t *bnbr;       // [NE][NF] (bounds_type 1)
   const double *xe_min, *xe_max;
   double *xe_min_out, *xe_max_out;   // may be NULL
};

template <int DIM, int D1, int Q, int E, bool AFF>
__global__ void __launch_bounds__(KCfg<DIM, D1, Q, E>::T, KCfg<DIM, D1, Q, E>::MINB) k_stage(StageArgs a, const Tab<D1, Q> tab)
{
   using K = KCfg<DIM, D1, Q, E>;
   constexpr int T = K::T, ND = ipow(D1, DIM), NF = 2 * DIM;
   constexpr int NK = (ND + 31) / 32, N3 = ipow(3, DIM);
   extern __shared__ double sm[];
   const int64_t e0 = (int64_t)blockIdx.x * E;
   const int ne = (int)min((int64_t)E, a.ho.ne - e0);
   double *U = sm + K::OFF_U;
   for (int t = threadIdx.x; t < E * ND; t += T) { U[t] = (t < ne * ND) ? a.ho.u[e0 * ND + t] : 0.0; }
   HoArgs ha = a.ho;
   ha.mode = 3;
   const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const bool mine = (w < ne);
   const int64_t ge = e0 + (mine ? w : 0);
   double m[NK], x0v[NK], bmn[NK], bmx[NK];
   // element-wise inputs of this warp's element (lumped mass, RK base state, bounds)
   auto load_ew = [&]()
   {
#pragma unroll
      for (int k = 0; k < NK; k++)
      {
         const int j = lane + 32 * k;
         m[k] = 1.0; x0v[k] = 0.0; bmn[k] = 0.0; bmx[k] = 0.0;
         if (mine && j < ND)
         {
            m[k] = a.ml[ge * ND + j];
            if (a.out_mode == 1 && a.a != 0.0) { x0v[k] = a.x0[ge * ND + j]; }
            if (a.bounds_type == 0)
            {
               const int ent = a.lat[ge * N3 + lattice_class(DIM, D1, j)];
               bmn[k] = a.ent_mm[2 * ent]; bmx[k] = a.ent_mm[2 * ent + 1];
            }
         }
      }
      if (mine && a.bounds_type == 1)
      {
         double bmin = a.xe_min[ge], bmax = a.xe_max[ge];
         for (int fc = 0; fc < NF; fc++)
         {
            const int nb = a.bnbr[ge * NF + fc];
            if (nb >= 0) { bmin = fmin(bmin, a.xe_min[nb]); bmax = fmax(bmax, a.xe_max[nb]); }
         }
#pragma unroll
         for (int k = 0; k < NK; k++) { bmn[k] = bmin; bmx[k] = bmax; }
      }
   };
#ifndef RMH_LATE_EW
#define RMH_LATE_EW 1
#endif
   // general path: issue them before the HO phases so their HBM latency overlaps the
   // contractions; affine path: register pressure is the tighter constraint, load afterwards
   constexpr bool LATE = AFF && (RMH_LATE_EW != 0);
   if (!LATE) { load_ew(); }
   const double *X = ho_phases<DIM, D1, Q, E, AFF>(ha, sm, e0, ne, tab);
   if (LATE) { load_ew(); }
   __syncthreads();
   // ---- element-wise part: one warp per element
   if (!mine) { return; }
   const double dt = a.dt;
   double u[NK], f[NK], lo[NK];
   double s1 = 0.0, s0 = 0.0;
#pragma unroll
   for (int k = 0; k < NK; k++)
   {
      const int j = lane + 32 * k;
      if (j < ND)
      {
         u[k] = U[w * ND + j];
         s1 += m[k] * (u[k] + dt * X[w * ND + j]);
         s0 += m[k];
      }
   }
   warp_sum2(s1, s0);
   const double ubar = s1 / s0;                        // MassBasedAvg, remhos_lo.cpp:278-285
   double sumPos = 0.0, sumNeg = 0.0;
#pragma unroll
   for (int k = 0; k < NK; k++)
   {
      const int j = lane + 32 * k;
      if (j < ND)
      {
         const double umin = bmn[k], umax = bmx[k];
         lo[k] = (ubar - u[k]) / dt;
         const double u_new_lo = u[k] + dt * lo[k];
         const double fmn = m[k] / dt * (umin - u_new_lo);
         const double fmx = m[k] / dt * (umax - u_new_lo);
         double fcl = m[k] * (X[w * ND + j] - lo[k]);
         fcl = fmin(fmx, fmax(fmn, fcl));               // ClipScale, remhos_fct.cpp:490-515
         f[k] = fcl;
         sumNeg += fmin(fcl, 0.0);
         sumPos += fmax(fcl, 0.0);
      }
   }
   warp_sum2(sumNeg, sumPos);
   const double new_mass = sumNeg + sumPos;
   constexpr double eps = 1.0e-15;
   double omin = INFINITY, omax = -INFINITY;
#pragma unroll
   for (int k = 0; k < NK; k++)
   {
      const int j = lane + 32 * k;
      if (j < ND)
      {
         double fcl = f[k];
         if (new_mass > eps) { fcl = fmin(0.0, fcl) - fmax(0.0, fcl) * sumNeg / sumPos; }
         if (new_mass < -eps) { fcl = fmax(0.0, fcl) - fmin(0.0, fcl) * sumPos / sumNeg; }
         const double du = lo[k] + fcl / m[k];
         double o = du;
         if (a.out_mode == 1) { o = a.a * x0v[k] + a.b * (u[k] + dt * du); }
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

// ------------------------------------------------------------- element-wise kernels
// one warp per element; generic in nd

// MassBasedAvg (remhos_lo.cpp:247-324): du_lo = (ubar - u)/dt, ubar = sum m (u + dt du_ho) / sum m
__global__ void k_mass_avg(int64_t ne, int nd, double dt, const double *u, const double *du_ho,
                           const double *ml, double *du_lo)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= ne) { return; }
   double s1 = 0.0, s0 = 0.0;
   for (int j = lane; j < nd; j += 32)
   {
      const int64_t i = e * nd + j;
      const double m = ml[i];
      s1 += m * (u[i] + dt * du_ho[i]);
      s0 += m;
   }
   s1 = warp_sum(s1); s0 = warp_sum(s0);
   const double ubar = s1 / s0;
   for (int j = lane; j < nd; j += 32)
   {
      const int64_t i = e * nd + j;
      du_lo[i] = (ubar - u[i]) / dt;
   }
}

// ComputeElementsMinMax (remhos_tools.cpp:497-523)
__global__ void k_elem_min_max(int64_t ne, int nd, const double *u, double *xe_min, double *xe_max,
                               double2 *xe_mm = nullptr)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= ne) { return; }
   double mn = INFINITY, mx = -INFINITY;
   for (int j = lane; j < nd; j += 32)
   {
      const double v = u[e * nd + j];
      mn = fmin(mn, v); mx = fmax(mx, v);
   }
   mn = warp_min(mn); mx = warp_max(mx);
   if (lane == 0)
   {
      if (xe_mm) { xe_mm[e] = make_double2(mn, mx); }
      else { xe_min[e] = mn; xe_max[e] = mx; }
   }
}

// same, streaming variant for even nd and 16-byte aligned u: a warp owns EW consecutive elements
// and issues all of its 16-byte loads before the first reduction (the one-element-per-warp form
// keeps two 8-byte loads in flight per lane and reaches a third of the HBM bandwidth)
template <int EW>
__global__ void __launch_bounds__(256) k_elem_min_max_v(int64_t ne, int nd, const double *__restrict__ u,
                                                        double *xe_min, double *xe_max, double2 *xe_mm)
{
   const int64_t e0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * EW;
   const int lane = threadIdx.x & 31;
   if (e0 >= ne) { return; }
   const int nd2 = nd >> 1;                 // double2 words per element
   double mn[EW], mx[EW];
#pragma unroll
   for (int k = 0; k < EW; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
   for (int j0 = 0; j0 < nd2; j0 += 32)
   {
      double2 v[EW];
#pragma unroll
      for (int k = 0; k < EW; k++)
      {
         const int j = j0 + lane;
         v[k] = make_double2(INFINITY, -INFINITY);
         if (j < nd2 && e0 + k < ne)
         {
            const double2 t = __ldcs(reinterpret_cast<const double2 *>(u + (e0 + k) * nd) + j);
            v[k] = make_double2(t.x < t.y ? t.x : t.y, t.x < t.y ? t.y : t.x);
         }
      }
#pragma unroll
      for (int k = 0; k < EW; k++)
      {
         mn[k] = v[k].x < mn[k] ? v[k].x : mn[k];
         mx[k] = v[k].y > mx[k] ? v[k].y : mx[k];
      }
   }
#pragma unroll
   for (int k = 0; k < EW; k++)
   {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
         const double a = __shfl_xor_sync(0xffffffffu, mn[k], o), b = __shfl_xor_sync(0xffffffffu, mx[k], o);
         mn[k] = a < mn[k] ? a : mn[k];
         mx[k] = b > mx[k] ? b : mx[k];
      }
   }
   if (lane < EW && e0 + lane < ne)
   {
      double a = mn[0], b = mx[0];
#pragma unroll
      for (int k = 1; k < EW; k++) { if (lane == k) { a = mn[k]; b = mx[k]; } }
      if (xe_mm) { xe_mm[e0 + lane] = make_double2(a, b); }
      else { xe_min[e0 + lane] = a; xe_max[e0 + lane] = b; }
   }
}

// (min, max) of every element: into two arrays, or -- xe_mm != nullptr -- as one pair per element
static void launch_elem_min_max(int64_t ne, int nd, const double *u, double *xe_min, double *xe_max, cudaStream_t s,
                                double2 *xe_mm = nullptr)
{
   const int bs = 256;
   if ((nd & 1) == 0 && (((uintptr_t)u) & 15) == 0)
   {
      constexpr int EW = 4;
      const int64_t nw = (ne + EW - 1) / EW;
      k_elem_min_max_v<EW><<<(unsigned)((nw * 32 + bs - 1) / bs), bs, 0, s>>>(ne, nd, u, xe_min, xe_max, xe_mm);
   }
   else
   {
      k_elem_min_max<<<(unsigned)((ne * 32 + bs - 1) / bs), bs, 0, s>>>(ne, nd, u, xe_min, xe_max, xe_mm);
   }
}

// entity min/max over the elements sharing the entity (CG-dof overlap, remhos_tools.cpp:449-458)
// (min, max) pairs of the elements as one 16-byte word each: the entity pass below is bound by the
// number of gathered words (l1tex 86 %, profiles/r01/ncu_ent_min_max_v12_rs5.txt), not by bytes
__global__ void k_xe_interleave(int64_t n, const double *xe_min, const double *xe_max, double2 *xe_mm)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { xe_mm[i] = make_double2(xe_min[i], xe_max[i]); }
}
__global__ void k_ent_min_max(int32_t n_ent, const int32_t *list, const int32_t *off, const int32_t *el,
                              const double2 *__restrict__ xe_mm, double *ent_mm)
{
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= n_ent) { return; }
   const int i = list ? list[t] : t;
   double mn = INFINITY, mx = -INFINITY;
   const int k1 = off[i + 1];
   // batches of four: all index loads, then all value loads, are in flight together
   for (int k = off[i]; k < k1; k += 4)
   {
      int id[4];
      double2 v[4];
#pragma unroll
      for (int q = 0; q < 4; q++) { id[q] = (k + q < k1) ? el[k + q] : -1; }
#pragma unroll
      for (int q = 0; q < 4; q++) { v[q] = (id[q] >= 0) ? xe_mm[id[q]] : make_double2(INFINITY, -INFINITY); }
#pragma unroll
      for (int q = 0; q < 4; q++) { mn = v[q].x < mn ? v[q].x : mn; mx = v[q].y > mx ? v[q].y : mx; }
   }
   *reinterpret_cast<double2 *>(ent_mm + 2 * (size_t)i) = make_double2(mn, mx);
}

// per-DOF gather (remhos_tools.cpp:468-494)
__global__ void k_bounds_overlap(int64_t ne, int dim, int D1, int nd, int n3, const int32_t *lat,
                                 const double *ent_mm, double *xi_min, double *xi_max)
{
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= ne * nd) { return; }
   const int64_t e = idx / nd;
   const int i = (int)(idx - e * nd);
   const int ent = lat[e * n3 + lattice_class(dim, D1, i)];
   xi_min[idx] = ent_mm[2 * ent]; xi_max[idx] = ent_mm[2 * ent + 1];
}

// ComputeMatrixSparsityBounds (remhos_tools.cpp:381-430)
__global__ void k_bounds_sparsity(int64_t ne, int nf, int nd, const int32_t *nbr,
                                  const double *xe_min, const double *xe_max, double *xi_min,
                                  double *xi_max)
{
   const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= ne * nd) { return; }
   const int64_t e = idx / nd;
   double mn = xe_min[e], mx = xe_max[e];
   for (int f = 0; f < nf; f++)
   {
      const int nb = nbr[e * nf + f];
      if (nb >= 0) { mn = fmin(mn, xe_min[nb]); mx = fmax(mx, xe_max[nb]); }
   }
   xi_min[idx] = mn; xi_max[idx] = mx;
}

// ClipScaleSolver::CalcFCTSolution (remhos_fct.cpp:484-539)
__global__ void k_clip_scale(int64_t ne, int nd, double dt, const double *u, const double *m,
                             const double *du_ho, const double *du_lo, const double *umin,
                             const double *umax, double *du)
{
   const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (e >= ne) { return; }
   constexpr int MAXK = 8;   // nd <= 256
   double f[MAXK];
   double sumPos = 0.0, sumNeg = 0.0;
   int k = 0;
   for (int j = lane; j < nd; j += 32, k++)
   {
      const int64_t i = e * nd + j;
      const double mi = m[i], lo = du_lo[i];
      const double u_new_lo = u[i] + dt * lo;
      const double fmn = mi / dt * (umin[i] - u_new_lo);
      const double fmx = mi / dt * (umax[i] - u_new_lo);
      double fc = mi * (du_ho[i] - lo);
      fc = fmin(fmx, fmax(fmn, fc));
      f[k] = fc;
      sumNeg += fmin(fc, 0.0);
      sumPos += fmax(fc, 0.0);
   }
   sumNeg = warp_sum(sumNeg); sumPos = warp_sum(sumPos);
   const double new_mass = sumNeg + sumPos;
   constexpr double eps = 1.0e-15;
   k = 0;
   for (int j = lane; j < nd; j += 32, k++)
   {
      const int64_t i = e * nd + j;
      double fc = f[k];
      if (new_mass > eps) { fc = fmin(0.0, fc) - fmax(0.0, fc) * sumNeg / sumPos; }
      if (new_mass < -eps) { fc = fmax(0.0, fc) - fmin(0.0, fc) * sumPos / sumNeg; }
      du[i] = du_lo[i] + fc / m[i];
   }
}

// out = a*x0 + b*(y + dt*k)
__global__ void k_rk_combine(int64_t n, double a, double b, double dt, const double *x0,
                             const double *y, const double *k, double *out)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { out[i] = a * x0[i] + b * (y[i] + dt * k[i]); }
}

// block-level reductions -> partial results, finished on the host (deterministic)
__global__ void k_reduce(int64_t n, int op, const double *a, const double *b, double *part)
{
   __shared__ double sh[32];
   double v = (op == 0) ? 0.0 : (op == 1 ? INFINITY : -INFINITY);
   for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
        i += (int64_t)gridDim.x * blockDim.x)
   {
      if (op == 0) { v += b ? a[i] * b[i] : a[i]; }
      else if (op == 1) { v = fmin(v, a[i]); }
      else { v = fmax(v, a[i]); }
   }
   v = (op == 0) ? warp_sum(v) : (op == 1 ? warp_min(v) : warp_max(v));
   const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
   if (lane == 0) { sh[w] = v; }
   __syncthreads();
   if (w == 0)
   {
      const int nw = blockDim.x >> 5;
      v = (lane < nw) ? sh[lane] : ((op == 0) ? 0.0 : (op == 1 ? INFINITY : -INFINITY));
      v = (op == 0) ? warp_sum(v) : (op == 1 ? warp_min(v) : warp_max(v));
      if (lane == 0) { part[blockIdx.x] = v; }
   }
}

// ============================================================================ dispatch
template <int DIM, int D1, int Q, int E>
static int launch_ho_E(rmh_ctx *c, const HoArgs &a, cudaStream_t s)
{
   using S = KCfg<DIM, D1, Q, E>;
   Tab<D1, Q> tab;
   for (int q = 0; q < Q; q++)
      for (int i = 0; i < D1; i++) { tab.B[q][i] = c->hB[q * D1 + i]; tab.G[q][i] = c->hG[q * D1 + i]; }
   for (int i = 0; i < D1; i++)
      for (int j = 0; j < D1; j++) { tab.Minv[i][j] = c->hMinv[i * D1 + j]; }
   for (int i = 0; i < D1; i++)
      for (int q = 0; q < Q; q++)
      {
         double v = 0.0;
         for (int j = 0; j < D1; j++) { v += c->hMinv[i * D1 + j] * c->hB[q * D1 + j]; }
         tab.C[i][q] = v;
      }
   static bool attr_set_dev[RMH_MAX_DEVICES] = {false};     // function attributes are per device
   bool &attr_set = attr_set_dev[c->device % RMH_MAX_DEVICES];
   if (!attr_set)
   {
      CUDA_OK(cudaFuncSetAttribute(k_ho<DIM, D1, Q, E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)S::BYTES));
      attr_set = true;
   }
   const int64_t nb = (a.ne + E - 1) / E;
   k_ho<DIM, D1, Q, E><<<(unsigned)nb, S::T, S::BYTES, s>>>(a, tab);
   LAUNCH_OK();
   return 0;
}

template <int DIM, int D1, int Q>
static int launch_ho(rmh_ctx *c, const HoArgs &a, cudaStream_t s)
{
   constexpr size_t LIM = 80 * 1024;
   if constexpr (KCfg<DIM, D1, Q, 8>::BYTES <= LIM) { return launch_ho_E<DIM, D1, Q, 8>(c, a, s); }
   else if constexpr (KCfg<DIM, D1, Q, 4>::BYTES <= LIM) { return launch_ho_E<DIM, D1, Q, 4>(c, a, s); }
   else if constexpr (KCfg<DIM, D1, Q, 2>::BYTES <= LIM) { return launch_ho_E<DIM, D1, Q, 2>(c, a, s); }
   else { return launch_ho_E<DIM, D1, Q, 1>(c, a, s); }
}

#define RMH_DISPATCH(FN, c, ...)                                                          \
   do {                                                                                   \
      const int key__ = (c)->dim * 10000 + (c)->D1 * 100 + (c)->Q;                        \
      switch (key__)                                                                      \
      {                                                                                   \
         case 20203: return FN<2, 2, 3>(c, __VA_ARGS__);                                  \
         case 20304: return FN<2, 3, 4>(c, __VA_ARGS__);                                  \
         case 20405: return FN<2, 4, 5>(c, __VA_ARGS__);                                  \
         case 20506: return FN<2, 5, 6>(c, __VA_ARGS__);                                  \
         case 30204: return FN<3, 2, 4>(c, __VA_ARGS__);                                  \
         case 30305: return FN<3, 3, 5>(c, __VA_ARGS__);                                  \
         case 30406: return FN<3, 4, 6>(c, __VA_ARGS__);                                  \
         case 30507: return FN<3, 5, 7>(c, __VA_ARGS__);                                  \
         default:                                                                         \
            set_error("no kernel instantiated for (dim, order+1, nq1d) = (" +            \
                      std::to_string((c)->dim) + "," + std::to_string((c)->D1) + "," +    \
                      std::to_string((c)->Q) + ")");                                      \
            return 1;                                                                     \
      }                                                                                   \
   } while (0)

static int dispatch_ho(rmh_ctx *c, const HoArgs &a, cudaStream_t s)
{
   RMH_DISPATCH(launch_ho, c, a, s);
}

template <int DIM, int D1, int Q, int E, bool AFF>
static int launch_stage_EA(rmh_ctx *c, const StageArgs &a, cudaStream_t s)
{
   using S = KCfg<DIM, D1, Q, E>;
   Tab<D1, Q> tab;
   for (int q = 0; q < Q; q++)
      for (int i = 0; i < D1; i++) { tab.B[q][i] = c->hB[q * D1 + i]; tab.G[q][i] = c->hG[q * D1 + i]; }
   for (int i = 0; i < D1; i++)
      for (int j = 0; j < D1; j++) { tab.Minv[i][j] = c->hMinv[i * D1 + j]; }
   for (int i = 0; i < D1; i++)
      for (int q = 0; q < Q; q++)
      {
         double v = 0.0;
         for (int j = 0; j < D1; j++) { v += c->hMinv[i * D1 + j] * c->hB[q * D1 + j]; }
         tab.C[i][q] = v;
      }
   static bool attr_set_dev[RMH_MAX_DEVICES] = {false};
   bool &attr_set = attr_set_dev[c->device % RMH_MAX_DEVICES];
   if (!attr_set)
   {
      CUDA_OK(cudaFuncSetAttribute(k_stage<DIM, D1, Q, E, AFF>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES));
      attr_set = true;
   }
   const int64_t nb = (a.ho.ne + E - 1) / E;
   k_stage<DIM, D1, Q, E, AFF><<<(unsigned)nb, S::T, S::BYTES, s>>>(a, tab);
   LAUNCH_OK();
   return 0;
}

template <int DIM, int D1, int Q, int E>
static int launch_stage_E(rmh_ctx *c, const StageArgs &a, cudaStream_t s)
{
   if constexpr (DIM == 3)
   {
      if (c->all_affine) { return launch_stage_EA<DIM, D1, Q, E, true>(c, a, s); }
   }
   return launch_stage_EA<DIM, D1, Q, E, false>(c, a, s);
}

template <int DIM, int D1, int Q>
static int launch_stage(rmh_ctx *c, const StageArgs &a, cudaStream_t s)
{
   constexpr size_t LIM = 80 * 1024;
   if constexpr (KCfg<DIM, D1, Q, 8>::BYTES <= LIM) { return launch_stage_E<DIM, D1, Q, 8>(c, a, s); }
   else if constexpr (KCfg<DIM, D1, Q, 4>::BYTES <= LIM) { return launch_stage_E<DIM, D1, Q, 4>(c, a, s); }
   else if constexpr (KCfg<DIM, D1, Q, 2>::BYTES <= LIM) { return launch_stage_E<DIM, D1, Q, 2>(c, a, s); }
   else { return launch_stage_E<DIM, D1, Q, 1>(c, a, s); }
}

static int dispatch_stage(rmh_ctx *c, const StageArgs &a, cudaStream_t s)
{
   RMH_DISPATCH(launch_stage, c, a, s);
}

// ---- persistent pipelined stage kernel (3D, every element affine): stage3p.cuh
template <int D1, int Q>
static Tab<D1, Q> make_tab(const rmh_ctx *c)
{
   Tab<D1, Q> tab;
   for (int q = 0; q < Q; q++)
      for (int i = 0; i < D1; i++) { tab.B[q][i] = c->hB[q * D1 + i]; tab.G[q][i] = c->hG[q * D1 + i]; }
   for (int i = 0; i < D1; i++)
      for (int j = 0; j < D1; j++) { tab.Minv[i][j] = c->hMinv[i * D1 + j]; }
   for (int i = 0; i < D1; i++)
      for (int q = 0; q < Q; q++)
      {
         double v = 0.0;
         for (int j = 0; j < D1; j++) { v += c->hMinv[i * D1 + j] * c->hB[q * D1 + j]; }
         tab.C[i][q] = v;
      }
   for (int q = 0; q < Q; q++) { tab.xq[q] = c->hxq[q]; tab.wq[q] = c->hw[q]; }
   return tab;
}

template <int D1, int Q, int E>
static int launch_stagep_E(rmh_ctx *c, const StagePArgs &a, cudaStream_t s)
{
   using S = SmemP<D1, Q, E>;
   constexpr int MINB0 = (int)((227 * 1024) / (S::BYTES + 1024));
   constexpr int MINB = MINB0 < 1 ? 1 : (MINB0 > 3 ? 3 : MINB0);
   static int blocks_per_sm_dev[RMH_MAX_DEVICES] = {0};
   int &blocks_per_sm = blocks_per_sm_dev[c->device % RMH_MAX_DEVICES];
   if (blocks_per_sm == 0)
   {
      CUDA_OK(cudaFuncSetAttribute(k_stage3p<D1, Q, E, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)S::BYTES));
      int nb = 0;
      CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_stage3p<D1, Q, E, MINB>, S::T, S::BYTES));
      if (nb < 1) { set_error("k_stage3p does not fit on an SM"); return 1; }
      blocks_per_sm = nb;
   }
   const int64_t nbatch = (a.ne + E - 1) / E;
   const int64_t grid = std::min<int64_t>(nbatch, (int64_t)blocks_per_sm * c->num_sms);
   k_stage3p<D1, Q, E, MINB><<<(unsigned)grid, S::T, S::BYTES, s>>>(a, make_tab<D1, Q>(c));
   LAUNCH_OK();
   return 0;
}

template <int DIM, int D1, int Q>
static int launch_stagep(rmh_ctx *c, const StagePArgs &a, cudaStream_t s)
{
   if constexpr (DIM == 3)
   {
      constexpr size_t LIM = 112 * 1024;
      static int force_e = -1;
      if (force_e < 0) { const char *ev = getenv("RMH_PIPE_E"); force_e = ev ? atoi(ev) : 0; }
      // 4 elements per block: three resident blocks per SM at order 3 and no register spills
      // (15 warps per SM leave 128 registers per thread); RMH_PIPE_E=8 selects the larger batch
      if (force_e != 8) { return launch_stagep_E<D1, Q, 4>(c, a, s); }
      if constexpr (SmemP<D1, Q, 8>::BYTES <= LIM) { return launch_stagep_E<D1, Q, 8>(c, a, s); }
      else if constexpr (SmemP<D1, Q, 4>::BYTES <= LIM) { return launch_stagep_E<D1, Q, 4>(c, a, s); }
      else { return launch_stagep_E<D1, Q, 2>(c, a, s); }
   }
   else
   {
      set_error("pipelined stage kernel is 3D only");
      return 1;
   }
}

static int dispatch_stagep(rmh_ctx *c, const StagePArgs &a, cudaStream_t s)
{
   RMH_DISPATCH(launch_stagep, c, a, s);
}


// ---- warp-per-element FP64 tensor-core kernel (stage3w.cuh)
// NW warps per block, MINB resident blocks per SM the register allocation must allow.  The kernel
// is latency-bound: 16 warps per SM at 128 registers (operator-data loads hoisted two phases
// ahead) beat 20 warps at 96 (profiles/r01/README.md).
template <int D1, int Q, int NW, int MINB, bool LIN>
static int launch_stagew_L(rmh_ctx *c, const StagePArgs &a, cudaStream_t s)
{
   using S = SmemW<D1, Q>;
   constexpr size_t BYTES = S::bytes(NW);
   static int blocks_per_sm_dev[RMH_MAX_DEVICES] = {0};
   int &blocks_per_sm = blocks_per_sm_dev[c->device % RMH_MAX_DEVICES];
   if (blocks_per_sm == 0)
   {
      CUDA_OK(cudaFuncSetAttribute(k_stage3w<D1, Q, NW, MINB, LIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)BYTES));
      int nb = 0;
      CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_stage3w<D1, Q, NW, MINB, LIN>, NW * 32,
                                                            BYTES));
      if (nb < 1) { set_error("k_stage3w does not fit on an SM"); return 1; }
      blocks_per_sm = std::min(nb, MINB);
      if (getenv("RMH_VERBOSE"))
      {
         fprintf(stderr, "k_stage3w<%d,%d,%d,%d,%s>: %d blocks/SM (occupancy %d), %zu B shared\n", D1, Q, NW, MINB,
                 LIN ? "lin" : "stored", blocks_per_sm, nb, BYTES);
      }
   }
   const int64_t nblk = (a.ne + NW - 1) / NW;
   const int64_t grid = std::min<int64_t>(nblk, (int64_t)blocks_per_sm * c->num_sms);
   k_stage3w<D1, Q, NW, MINB, LIN><<<(unsigned)grid, NW * 32, BYTES, s>>>(a, make_tab<D1, Q>(c));
   LAUNCH_OK();
   return 0;
}

template <int D1, int Q, int NW, int MINB>
static int launch_stagew_N(rmh_ctx *c, const StagePArgs &a, cudaStream_t s)
{
   return a.opc ? launch_stagew_L<D1, Q, NW, MINB, true>(c, a, s)
                : launch_stagew_L<D1, Q, NW, MINB, false>(c, a, s);
}

template <int DIM, int D1, int Q>
static int launch_stagew(rmh_ctx *c, const StagePArgs &a, cudaStream_t s)
{
   if constexpr (DIM == 3 && D1 <= 5)
   {
      static int cfg = -1;
      if (cfg < 0)
      {
         const char *ev = getenv("RMH_W_NW"), *em = getenv("RMH_W_MINB");
         // order 4 needs 18.7 KB of shared memory per warp: 2 blocks x 6 warps (168 registers)
         const int nw = ev ? atoi(ev) : (D1 >= 5 ? 6 : 8), mb = em ? atoi(em) : 0;
         cfg = nw * 10 + mb;
      }
      switch (cfg)
      {
         case 44: case 40: return launch_stagew_N<D1, Q, 4, 4>(c, a, s);    // 16 warps, 128 regs
         case 43: return launch_stagew_N<D1, Q, 4, 3>(c, a, s);             // 12 warps, 168 regs
         case 54: return launch_stagew_N<D1, Q, 5, 4>(c, a, s);             // 20 warps,  96 regs
         case 53: case 50: return launch_stagew_N<D1, Q, 5, 3>(c, a, s);    // 15 warps, 136 regs
         case 62: case 60: return launch_stagew_N<D1, Q, 6, 2>(c, a, s);    // 12 warps, 168 regs
         case 72: case 70: return launch_stagew_N<D1, Q, 7, 2>(c, a, s);    // 14 warps, 144 regs
         case 102: case 100: return launch_stagew_N<D1, Q, 10, 2>(c, a, s); // 20 warps,  96 regs
         case 101: return launch_stagew_N<D1, Q, 10, 1>(c, a, s);           // 10 warps, 200 regs
         default: return launch_stagew_N<D1, Q, 8, 2>(c, a, s);             // 16 warps, 128 regs
      }
   }
   else
   {
      set_error("tensor-core stage kernel: 3D, order <= 4 only");
      return 1;
   }
}

static int dispatch_stagew(rmh_ctx *c, const StagePArgs &a, cudaStream_t s)
{
   RMH_DISPATCH(launch_stagew, c, a, s);
}

// ---- constant-coefficient kernel (stage3c.cuh)
// 1-D operators of the constant-coefficient kernel.  The integrals are rational numbers (Bernstein
// products), so they are formed in extended precision from their closed forms -- the same matrices the
// Q-point Gauss rule gives exactly, M1 = B^T W B and S = B^T W G -- and rounded once:
//    M1_ij = C(p,i) C(p,j) / (C(2p,i+j) (2p+1)),     S_ik = int B_i^p (B_k^p)' = p (I_{i,k-1} - I_{i,k}),
//    I_ij = int B_i^p B_j^{p-1} = C(p,i) C(p-1,j) / (C(2p-1,i+j) 2p),          T = M1^-1 S.
// Then the column sums are made exact in double: sum_i T[i][k] = D1 (delta_kp - delta_k0) and
// sum_i Minv[i][0] = sum_i Minv[i][p] = D1 (partition of unity: 1^T M1 = 1^T / D1).  These sums are what
// makes the scheme conservative -- the mass an element loses through a face is D1 Minv-weighted on one
// side and on the other -- and with tables rounded entry by entry they were off by cond(M1) eps ~ 1e-14,
// a SYSTEMATIC source (u > 0 everywhere) that showed as a mass drift of 2e-14 per stage.
static long double binom_ld(int n, int k)
{
   if (k < 0 || k > n) { return 0.0L; }
   long double r = 1.0L;
   for (int i = 1; i <= k; i++) { r = r * (long double)(n - k + i) / (long double)i; }
   return r;
}

template <int D1, int Q>
static TabC<D1> make_tabc(const rmh_ctx *c)
{
   (void)c;
   constexpr int p = D1 - 1;
   long double M[D1][2 * D1], S[D1][D1];
   for (int i = 0; i < D1; i++)
   {
      for (int j = 0; j < D1; j++)
      {
         M[i][j] = binom_ld(p, i) * binom_ld(p, j) / (binom_ld(2 * p, i + j) * (long double)(2 * p + 1));
         M[i][D1 + j] = (i == j) ? 1.0L : 0.0L;
      }
      for (int k = 0; k < D1; k++)
      {
         auto I = [&](int jj) -> long double
         {
            if (p == 0 || jj < 0 || jj > p - 1) { return 0.0L; }
            return binom_ld(p, i) * binom_ld(p - 1, jj) / (binom_ld(2 * p - 1, i + jj) * (long double)(2 * p));
         };
         S[i][k] = (long double)p * (I(k - 1) - I(k));
      }
   }
   for (int col = 0; col < D1; col++)           // Gauss-Jordan with partial pivoting
   {
      int piv = col;
      for (int r = col + 1; r < D1; r++) { if (fabsl(M[r][col]) > fabsl(M[piv][col])) { piv = r; } }
      for (int j = 0; j < 2 * D1; j++) { std::swap(M[col][j], M[piv][j]); }
      const long double d = M[col][col];
      for (int j = 0; j < 2 * D1; j++) { M[col][j] /= d; }
      for (int r = 0; r < D1; r++)
      {
         if (r == col) { continue; }
         const long double f = M[r][col];
         for (int j = 0; j < 2 * D1; j++) { M[r][j] -= f * M[col][j]; }
      }
   }
   TabC<D1> o;
   for (int i = 0; i < D1; i++)
   {
      for (int k = 0; k < D1; k++)
      {
         long double v = 0.0L;
         for (int j = 0; j < D1; j++) { v += M[i][D1 + j] * S[j][k]; }
         o.T[i][k] = (double)v;
      }
      o.M0[i] = (double)M[i][D1]; o.Mp[i] = (double)M[i][D1 + D1 - 1];
   }
   // exact column sums in double (the entry of largest magnitude absorbs the rounding)
   auto fix = [&](auto get, auto put, double target)
   {
      for (int it = 0; it < 3; it++)
      {
         double sum = 0.0;
         int big = 0;
         for (int i = 0; i < D1; i++) { sum += get(i); if (std::fabs(get(i)) > std::fabs(get(big))) { big = i; } }
         if (sum == target) { break; }
         put(big, get(big) + (target - sum));
      }
   };
   for (int k = 0; k < D1; k++)
   {
      const double target = (double)D1 * ((k == p ? 1.0 : 0.0) - (k == 0 ? 1.0 : 0.0));
      fix([&](int i) { return o.T[i][k]; }, [&](int i, double v) { o.T[i][k] = v; }, target);
   }
   fix([&](int i) { return o.M0[i]; }, [&](int i, double v) { o.M0[i] = v; }, (double)D1);
   fix([&](int i) { return o.Mp[i]; }, [&](int i, double v) { o.Mp[i] = v; }, (double)D1);
   return o;
}

template <int D1, int Q, int NW, int MINB, int NST, bool GH, bool FOLD, bool SEND = false>
static int launch_stagec_G(rmh_ctx *c, const StagePArgs &a, cudaStream_t s)
{
   using S = SmemC<D1, NST>;
   constexpr size_t BYTES = S::bytes(NW);
   static int blocks_per_sm[RMH_MAX_DEVICES] = {0};     // function attributes are per device
   int &bps = blocks_per_sm[c->device % RMH_MAX_DEVICES];
   if (bps == 0)
   {
      CUDA_OK(cudaFuncSetAttribute(k_stage3c<D1, NW, MINB, NST, GH, FOLD, SEND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)BYTES));
      int nb = 0;
      CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_stage3c<D1, NW, MINB, NST, GH, FOLD, SEND>, NW * 32, BYTES));
      if (nb < 1) { set_error("k_stage3c does not fit on an SM"); return 1; }
      bps = std::min(nb, MINB);
      if (getenv("RMH_VERBOSE"))
      {
         fprintf(stderr, "k_stage3c<%d,%d,%d,%d,%s,%s%s>: %d blocks/SM (occupancy %d), %zu B shared\n", D1, NW, MINB, NST,
                 GH ? "ghosts" : "local", FOLD ? "fold" : "entities", SEND ? ",send" : "", bps, nb, BYTES);
      }
   }
   const int64_t ngrp = (a.ne + S::E - 1) / S::E - a.e_begin / S::E;
   const int64_t nblk = (ngrp + NW - 1) / NW;
   const int64_t grid = std::min<int64_t>(nblk, (int64_t)bps * c->num_sms);
   k_stage3c<D1, NW, MINB, NST, GH, FOLD, SEND><<<(unsigned)grid, NW * 32, BYTES, s>>>(a, make_tabc<D1, Q>(c));
   LAUNCH_OK();
   return 0;
}

template <int D1, int Q, int NW, int MINB, int NST>
static int launch_stagec_N(rmh_ctx *c, const StagePArgs &a, cudaStream_t s)
{
   const bool fold = (a.xe_mm_out != nullptr);
   static int force_gh = -1;        // RMH_FORCE_GH=1: the ghost-aware instantiation on a mesh without ghosts (profiling)
   if (force_gh < 0) { const char *ev = getenv("RMH_FORCE_GH"); force_gh = (ev && ev[0] == '1') ? 1 : 0; }
   if (c->ne_ghost > 0 || force_gh)
   {
      if (fold && a.send != nullptr) { return launch_stagec_G<D1, Q, NW, MINB, NST, true, true, true>(c, a, s); }
      return fold ? launch_stagec_G<D1, Q, NW, MINB, NST, true, true>(c, a, s)
                  : launch_stagec_G<D1, Q, NW, MINB, NST, true, false>(c, a, s);
   }
   return fold ? launch_stagec_G<D1, Q, NW, MINB, NST, false, true>(c, a, s)
               : launch_stagec_G<D1, Q, NW, MINB, NST, false, false>(c, a, s);
}

template <int DIM, int D1, int Q>
static int launch_stagec(rmh_ctx *c, const StagePArgs &a, cudaStream_t s)
{
   if constexpr (DIM == 3 && D1 <= 5)
   {
      static int cfg = -1;
      if (cfg < 0)
      {
         const char *ev = getenv("RMH_C_CFG");     // NW * 100 + MINB * 10 + NST
         cfg = ev ? atoi(ev) : 0;
      }
      switch (cfg)
      {
         case 822: return launch_stagec_N<D1, Q, 8, 2, 2>(c, a, s);     // 16 warps in two blocks, ring of 2
         default:                                                        // 16 warps (one block), ring of 2
            if constexpr (SmemC<D1, 2>::bytes(16) <= 227 * 1024) { return launch_stagec_N<D1, Q, 16, 1, 2>(c, a, s); }
            else { return launch_stagec_N<D1, Q, 8, 1, 2>(c, a, s); }   // order 1: eight elements per warp
      }
   }
   else
   {
      set_error("constant-coefficient stage kernel: 3D, order <= 4 only");
      return 1;
   }
}

static int dispatch_stagec(rmh_ctx *c, const StagePArgs &a, cudaStream_t s)
{
   RMH_DISPATCH(launch_stagec, c, a, s);
}

// Decomposed meshes, FluxBasedFCT: the face block the NEIGHBOUR assembles for a face on the rank boundary,
// k_ji = BI_nbr[b][a] = -sum_q w S(-v.n) phi_a phi_b with S the upwind switch of k_geom_face, formed from
// this side's geometry at the same quadrature points (conforming face: same points, same weights, opposite
// normal).  One block per (owned element, face); BIg[slot][a][b] in this element's face numbering.
template <int DIM>
__global__ void k_fa_ghost_blocks(GeomArgs g, OpData o, const int32_t *nbr_elem, int64_t ne_owned, double *BIg)
{
   __shared__ double dn[64];
   const int Q = g.Q, NF = 2 * DIM, NFD = o.NFD, NQF = o.NQF;
   const int64_t e = blockIdx.x / NF;
   const int f = (int)(blockIdx.x - e * NF);
   const int64_t nb = nbr_elem[e * NF + f];
   if (nb < ne_owned) { return; }
   const int64_t slot = nb - ne_owned;
   int axis, side;
   face_axis_side(DIM, f, axis, side);
   for (int qf = threadIdx.x; qf < NQF; qf += blockDim.x)
   {
      const double *l[3], *dl[3];
      double w = 1.0;
      int m = qf;
      for (int a = 0; a < DIM; a++)
      {
         if (a == axis) { l[a] = g.Ls + side * g.NG1; dl[a] = g.dLs + side * g.NG1; }
         else
         {
            const int q = m % Q; m /= Q;
            w *= g.w[q];
            l[a] = g.L + q * g.NG1; dl[a] = g.dL + q * g.NG1;
         }
      }
      double J[3][3], v[3], det, adj[3][3];
      const bool nodal_v = (g.velf == nullptr);
      eval_geom<DIM>(g, e, l, dl, J, v, nodal_v);
      det_adj<DIM>(J, det, adj);
      if (!nodal_v) { for (int i = 0; i < DIM; i++) { v[i] = g.velf[((size_t)(e * NF + f) * NQF + qf) * DIM + i]; } }
      const double sgn = side ? 1.0 : -1.0;
      double vn = 0.0;
      for (int i = 0; i < DIM; i++) { vn += v[i] * sgn * adj[axis][i]; }
      vn = -vn;                                    // seen from the neighbour
      dn[qf] = w * ((g.exec_mode == 1) ? -fmax(0.0, vn) : fmin(0.0, vn));
   }
   __syncthreads();
   for (int t = threadIdx.x; t < NFD * NFD; t += blockDim.x)
   {
      const int a = t / NFD, b = t - a * NFD;
      double s = 0.0;
      for (int qf = 0; qf < NQF; qf++) { s += dn[qf] * (face_phi(o, a, qf) * face_phi(o, b, qf)); }
      BIg[(size_t)slot * NFD * NFD + t] = -s;
   }
}

static OpData op_data(const rmh_ctx *c)
{
   OpData o;
   o.dim = c->dim; o.D1 = c->D1; o.Q = c->Q; o.frag = c->frag ? 1 : 0;
   o.ND = c->ND; o.NQ = c->NQ; o.NF = c->NF; o.NFD = c->NFD; o.NQF = c->NQF;
   o.Dvol = c->Dvol; o.Dface = c->Dface; o.detJw = c->detJw; o.B = c->dB; o.G = c->dG;
   return o;
}

static int run_geom(rmh_ctx *c, double t, cudaStream_t s)
{
   GeomArgs g;
   g.dim = c->dim; g.Q = c->Q; g.NG1 = c->NG1; g.exec_mode = c->exec_mode; g.ne = c->ne; g.t = t;
   g.frag = c->frag ? 1 : 0;
   g.X0 = c->X0; g.V = c->V; g.velq = c->velq; g.velf = c->velf;
   g.L = c->dL; g.dL = c->ddL; g.Ls = c->dLs; g.dLs = c->ddLs; g.w = c->dw;
   g.Dvol = c->Dvol; g.detJw = c->detJw; g.Dface = c->Dface;
   const int bs = 128;
   const int64_t nv = c->ne * c->NQ, nfq = c->ne * c->NF * c->NQF;
   static int use_g3 = -1;   // RMH_NO_GEOM3=1 selects the point-wise set-up kernels in 3D too
   if (use_g3 < 0) { const char *ev = getenv("RMH_NO_GEOM3"); use_g3 = (ev && ev[0] == '1') ? 0 : 1; }
   const bool g3 = (c->dim == 3) && use_g3 && c->NG1 == 3 &&
                   (c->D1 * 100 + c->Q == 204 || c->D1 * 100 + c->Q == 305 || c->D1 * 100 + c->Q == 406 ||
                    c->D1 * 100 + c->Q == 507);
   if (g3)
   {
      Geom3Args a;
      a.Q = c->Q; a.NG1 = c->NG1; a.D1 = c->D1; a.exec_mode = c->exec_mode; a.frag = c->frag ? 1 : 0;
      a.ne = c->ne; a.t = t;
      a.X0 = c->X0; a.V = c->V; a.velq = c->velq; a.velf = c->velf;
      a.L = c->dL; a.dL = c->ddL; a.Ls = c->dLs; a.dLs = c->ddLs; a.w = c->dw; a.B = c->dB;
      a.Dvol = c->Dvol; a.detJw = c->detJw; a.Dface = c->Dface; a.ml = c->ml; a.einv = c->einv; a.BL = c->BL;
      const size_t shb = geom3_smem_doubles(c->Q, c->NG1, c->D1) * sizeof(double);
      constexpr int GT = 128;
      switch (c->D1 * 100 + c->Q)
      {
         case 204: k_geom3<3, 4, 2, GT><<<(unsigned)c->ne, GT, shb, s>>>(a); break;
         case 305: k_geom3<3, 5, 3, GT><<<(unsigned)c->ne, GT, shb, s>>>(a); break;
         case 406: k_geom3<3, 6, 4, GT><<<(unsigned)c->ne, GT, shb, s>>>(a); break;
         case 507: k_geom3<3, 7, 5, GT><<<(unsigned)c->ne, GT, shb, s>>>(a); break;
         default: set_error("k_geom3: no instantiation"); return 1;
      }
      LAUNCH_OK();
   }
   else
   {
   if (c->dim == 2)
   {
      k_geom_vol<2><<<(unsigned)((nv + bs - 1) / bs), bs, 0, s>>>(g); LAUNCH_OK();
      k_geom_face<2><<<(unsigned)((nfq + bs - 1) / bs), bs, 0, s>>>(g); LAUNCH_OK();
   }
   else
   {
      k_geom_vol<3><<<(unsigned)((nv + bs - 1) / bs), bs, 0, s>>>(g); LAUNCH_OK();
      k_geom_face<3><<<(unsigned)((nfq + bs - 1) / bs), bs, 0, s>>>(g); LAUNCH_OK();
   }
   k_lumped_mass<<<(unsigned)((c->N + bs - 1) / bs), bs, 0, s>>>(c->dim, c->D1, c->Q, c->ne, c->dB,
                                                                 c->detJw, c->ml);
   LAUNCH_OK();
   k_elem_affine<<<(unsigned)((c->ne * 32 + bs - 1) / bs), bs, 0, s>>>(
      c->dim, c->Q, c->NGN, c->exec_mode, t, c->ne, c->dw, c->detJw, c->X0, c->V, c->einv);
   LAUNCH_OK();
   }
   {
      const OpData o = op_data(c);
      const int64_t n = c->ne * c->NF * c->NFD;
      if (!g3)
      {
         k_face_lump<<<(unsigned)((n + bs - 1) / bs), bs, 0, s>>>(o, c->ne, c->BL);
         LAUNCH_OK();
      }
      if (c->sub_on)
      {
         int ns = 1;
         for (int a = 0; a < c->dim; a++) { ns *= c->p; }
         const int64_t nsub = c->ne * ns;
         k_subcell_weights<<<(unsigned)((nsub + bs - 1) / bs), bs, 0, s>>>(
            c->dim, c->p, c->exec_mode, t, c->ne, c->sub_x, c->sub_v, c->sub_w);
         LAUNCH_OK();
      }
      if (c->fa_on)
      {
         const size_t shb = (size_t)((c->dim + 1) * c->NQ + 2 * c->Q * c->D1) * sizeof(double);
         k_fa_dense<<<(unsigned)c->ne, 256, shb, s>>>(o, c->ne, c->faK, c->faKH, c->faM, c->faBI);
         LAUNCH_OK();
         if (c->ne_ghost > 0 && c->n_gslots > 0)
         {
            if (c->NQF > 64) { set_error("k_fa_ghost_blocks: more than 64 face quadrature points"); return 1; }
            if (c->dim == 2) { k_fa_ghost_blocks<2><<<(unsigned)(c->ne * c->NF), 64, 0, s>>>(g, o, c->nbr_elem, c->ne, c->faBIg); }
            else { k_fa_ghost_blocks<3><<<(unsigned)(c->ne * c->NF), 64, 0, s>>>(g, o, c->nbr_elem, c->ne, c->faBIg); }
            LAUNCH_OK();
         }
      }
   }
   c->t_cur = t;
   return 0;
}

// ============================================================================ C ABI
extern "C" int64_t rmh_launch_count(int reset)
{
   const int64_t v = g_launches.load();
   if (reset) { g_launches = 0; }
   return v;
}

static void host_pipe_free(rmh_ctx *c);

extern "C" int rmh_ctx_destroy(rmh_ctx *c)
{
   if (!c) { return 0; }
   cudaSetDevice(c->device);
   host_pipe_free(c);
   for (void *p : c->allocs) { cudaFree(p); }
   for (cudaEvent_t ev : c->prof_ev) { cudaEventDestroy(ev); }
   if (c->pin) { cudaFreeHost(c->pin); }
   delete c;
   return 0;
}

extern "C" int rmh_ctx_create(const rmh_desc *d, rmh_ctx **out)
{
   if (!d || !out) { set_error("null argument"); return 1; }
   if (d->dim != 2 && d->dim != 3) { set_error("dim must be 2 or 3"); return 1; }
   if (d->order < 1) { set_error("order must be >= 1 (order 0 disables limiting, remhos.cpp:600-608)"); return 1; }
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
   {
      set_error("no CUDA device: remhos_b200 has no CPU fallback");
      return 1;
   }
   CUDA_OK(cudaSetDevice(d->device));
   rmh_ctx *c = new rmh_ctx;
   c->dim = d->dim; c->p = d->order; c->mo = d->mesh_order; c->exec_mode = d->exec_mode;
   c->bounds_type = d->bounds_type; c->device = d->device;
   c->D1 = c->p + 1;
   c->Q = (2 * c->p + c->dim * c->mo - 1) / 2 + 1;
   c->ND = ipow(c->D1, c->dim); c->NQ = ipow(c->Q, c->dim);
   c->NF = 2 * c->dim; c->NFD = ipow(c->D1, c->dim - 1); c->NQF = ipow(c->Q, c->dim - 1);
   c->NG1 = c->mo + 1; c->NGN = ipow(c->NG1, c->dim); c->N3 = ipow(3, c->dim);
   c->ne = d->ne; c->ne_ghost = d->ne_ghost; c->N = c->ne * c->ND;
   c->t_cur = 0.0;
   {
      cudaDeviceProp prop;
      CUDA_OK(cudaGetDeviceProperties(&prop, d->device));
      c->num_sms = prop.multiProcessorCount;
      const char *np = getenv("RMH_NO_PIPELINE");
      c->pipelined = !(np && np[0] == '1');
      const char *nt = getenv("RMH_NO_TENSOR");
      c->tensor = !(nt && nt[0] == '1');
   }
   auto fail = [&]() { rmh_ctx_destroy(c); return 1; };
   if ((double)(c->ne + c->ne_ghost) * c->ND >= 2147483647.0)
   { set_error("too many DOFs for int32 index maps"); return fail(); }
   // ---- 1-D tables
   gauss_legendre_01(c->Q, c->hxq, c->hw);
   c->hB = bernstein(c->p, c->hxq);
   c->hG = bernstein_deriv(c->p, c->hxq);
   {
      std::vector<double> M1((size_t)c->D1 * c->D1, 0.0);
      for (int q = 0; q < c->Q; q++)
         for (int i = 0; i < c->D1; i++)
            for (int j = 0; j < c->D1; j++)
            { M1[i * c->D1 + j] += c->hw[q] * c->hB[q * c->D1 + i] * c->hB[q * c->D1 + j]; }
      c->hMinv = invert_small(M1, c->D1);
   }
   const std::vector<double> gll = gauss_lobatto_01(c->NG1);
   const std::vector<double> L = lagrange(gll, c->hxq), dL = lagrange_deriv(gll, c->hxq);
   const std::vector<double> ends = {0.0, 1.0};
   const std::vector<double> Ls = lagrange(gll, ends), dLs = lagrange_deriv(gll, ends);
   if (dev_upload(c, &c->dB, c->hB.data(), c->hB.size())) { return fail(); }
   if (dev_upload(c, &c->dG, c->hG.data(), c->hG.size())) { return fail(); }
   if (dev_upload(c, &c->dw, c->hw.data(), c->hw.size())) { return fail(); }
   if (dev_upload(c, &c->dL, L.data(), L.size())) { return fail(); }
   if (dev_upload(c, &c->ddL, dL.data(), dL.size())) { return fail(); }
   if (dev_upload(c, &c->dLs, Ls.data(), Ls.size())) { return fail(); }
   if (dev_upload(c, &c->ddLs, dLs.data(), dLs.size())) { return fail(); }
   // ---- geometry / velocity inputs
   const size_t nnod = (size_t)c->ne * c->NGN * c->dim;
   if (!d->nodes) { set_error("desc.nodes is required"); return fail(); }
   if (dev_upload(c, &c->X0, d->nodes, nnod)) { return fail(); }
   if (d->vel_nodes) { if (dev_upload(c, &c->V, d->vel_nodes, nnod)) { return fail(); } }
   if (c->ne > 0)
   {
      c->x0_e0.assign(d->nodes, d->nodes + (size_t)c->NGN * c->dim);
      if (d->vel_nodes && c->exec_mode == 1) { c->v_e0.assign(d->vel_nodes, d->vel_nodes + (size_t)c->NGN * c->dim); }
   }
   if (c->exec_mode == 1 && !d->vel_nodes)
   { set_error("remap mode needs desc.vel_nodes (mesh velocity)"); return fail(); }
   if (c->exec_mode == 0)
   {
      if (d->vel_quad && d->vel_face)
      {
         if (dev_upload(c, &c->velq, d->vel_quad, (size_t)c->ne * c->NQ * c->dim)) { return fail(); }
         if (dev_upload(c, &c->velf, d->vel_face, (size_t)c->ne * c->NF * c->NQF * c->dim)) { return fail(); }
      }
      else if (!d->vel_nodes)
      { set_error("transport mode needs vel_quad+vel_face or vel_nodes"); return fail(); }
   }
   // ---- face neighbour map: compress NbrDof into (element, pattern)
   if (!d->nbr_dof) { set_error("desc.nbr_dof is required"); return fail(); }
   {
      const int NF = c->NF, NFD = c->NFD, ND = c->ND;
      std::vector<int> bd;
      bdr_dofs(c->p, c->dim, bd);   // [NFD][NF] reference ordering
      // natural face order j -> reference BdrDofs position
      std::vector<int> nat2ref((size_t)NF * NFD);
      for (int f = 0; f < NF; f++)
      {
         int axis, side;
         face_axis(c->dim, f, axis, side);
         for (int j = 0; j < NFD; j++)
         {
            int l[3] = {0, 0, 0}, m = j;
            for (int a = 0; a < c->dim; a++)
            {
               if (a == axis) { l[a] = side * c->p; }
               else { l[a] = m % c->D1; m /= c->D1; }
            }
            const int dof = l[0] + c->D1 * (l[1] + c->D1 * l[2]);
            int pos = -1;
            for (int r = 0; r < NFD; r++) { if (bd[r * NF + f] == dof) { pos = r; } }
            nat2ref[f * NFD + j] = pos;
         }
      }
      std::vector<int32_t> ne_h((size_t)c->ne * NF);
      std::vector<uint8_t> pid_h((size_t)c->ne * NF, 0);
      std::map<std::vector<int16_t>, int> pats;
      std::vector<int16_t> patv, key(NFD);
      for (int64_t e = 0; e < c->ne; e++)
         for (int f = 0; f < NF; f++)
         {
            const int32_t *nd = &d->nbr_dof[((size_t)e * NF + f) * NFD];
            if (nd[0] < 0) { ne_h[e * NF + f] = -1; continue; }
            const int32_t nb = nd[0] / ND;
            for (int j = 0; j < NFD; j++)
            {
               const int32_t g = nd[nat2ref[f * NFD + j]];
               if (g < 0 || g / ND != nb)
               { set_error("nbr_dof: face DOFs of one face must map into one neighbour element"); return fail(); }
               key[j] = (int16_t)(g - nb * ND);
            }
            auto it = pats.find(key);
            int id;
            if (it == pats.end())
            {
               id = (int)pats.size();
               if (id >= 255) { set_error("nbr_dof: too many distinct face patterns"); return fail(); }
               pats[key] = id;
               patv.insert(patv.end(), key.begin(), key.end());
            }
            else { id = it->second; }
            ne_h[e * NF + f] = nb;
            pid_h[e * NF + f] = (uint8_t)id;
            if (nb >= c->ne + c->ne_ghost) { set_error("nbr_dof: neighbour beyond ghost range"); return fail(); }
            if (nb >= c->ne)
            {
               // ghost neighbour: one trace slot per (element, face), in scan order
               c->gs_ghost.push_back((int32_t)(nb - c->ne));
               c->gs_pid.push_back(id);
               if (c->ne + c->n_gslots >= 2147483647LL) { set_error("too many ghost faces"); return fail(); }
               ne_h[e * NF + f] = (int32_t)(c->ne + c->n_gslots);
               c->n_gslots++;
            }
         }
      c->npat = (int)pats.size();
      if (patv.empty()) { patv.assign(NFD, 0); }
      c->pat_h = patv;
      if (dev_upload(c, &c->nbr_elem, ne_h.data(), ne_h.size())) { return fail(); }
      if (dev_upload(c, &c->nbr_pat, pid_h.data(), pid_h.size())) { return fail(); }
      {
         std::vector<int32_t> pid32(pid_h.begin(), pid_h.end());
         if (dev_upload(c, &c->nbr_pat32, pid32.data(), pid32.size())) { return fail(); }
      }
      if (dev_upload(c, &c->pat, patv.data(), patv.size())) { return fail(); }
      // per pattern: the neighbour's local face and the natural index on it of every matched DOF
      // (the reverse view the flux-based FCT needs for k_ji across a face)
      const int np = (int)(patv.size() / NFD);
      std::vector<uint8_t> pface(np, 0);
      std::vector<int16_t> pidx(patv.size(), 0);
      for (int id = 0; id < np; id++)
      {
         int found = -1;
         for (int f2 = 0; f2 < NF && found < 0; f2++)
         {
            int axis, side;
            face_axis(c->dim, f2, axis, side);
            bool all = true;
            for (int j = 0; j < NFD && all; j++)
            {
               int m = patv[(size_t)id * NFD + j], l[3] = {0, 0, 0};
               for (int a = 0; a < c->dim; a++) { l[a] = m % c->D1; m /= c->D1; }
               all = (l[axis] == side * c->p);
            }
            if (all) { found = f2; }
         }
         if (found < 0) { found = 0; }
         pface[id] = (uint8_t)found;
         int axis, side;
         face_axis(c->dim, found, axis, side);
         for (int j = 0; j < NFD; j++)
         {
            int m = patv[(size_t)id * NFD + j], l[3] = {0, 0, 0};
            for (int a = 0; a < c->dim; a++) { l[a] = m % c->D1; m /= c->D1; }
            int nat = 0, mul = 1;
            for (int a = 0; a < c->dim; a++) { if (a != axis) { nat += l[a] * mul; mul *= c->D1; } }
            pidx[(size_t)id * NFD + j] = (int16_t)nat;
         }
      }
      if (dev_upload(c, &c->pat_face, pface.data(), pface.size())) { return fail(); }
      if (dev_upload(c, &c->pat_idx, pidx.data(), pidx.size())) { return fail(); }
   }
   // ---- bounds structures
   const int64_t ne_all = c->ne + c->ne_ghost;
   if (dev_alloc(c, &c->xe_min, (size_t)ne_all)) { return fail(); }
   if (dev_alloc(c, &c->xe_max, (size_t)ne_all)) { return fail(); }
   if (dev_alloc(c, &c->xe_mm, (size_t)ne_all)) { return fail(); }
   {
      // window: epoch flags, the two (min,max) pair arrays (+1: sentinel), the two ghost trace arrays
      auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
      size_t off = up(RMH_MAX_PEERS * sizeof(unsigned long long));
      for (int k = 0; k < 2; k++) { c->win_off_mm[k] = off; off += up((size_t)(ne_all + 1) * sizeof(double2)); }
      for (int k = 0; k < 2; k++) { c->win_off_tr[k] = off; off += up((size_t)std::max<int64_t>(c->n_gslots, 1) * c->NFD * sizeof(double)); }
      c->win_bytes = off;
      CUDA_OK(cudaMalloc(&c->win, off));
      c->allocs.push_back(c->win);
      CUDA_OK(cudaMemset(c->win, 0, off));
      c->flags = (unsigned long long *)c->win;
      if (dev_alloc(c, &c->zeros, 64)) { return fail(); }
      CUDA_OK(cudaMemset(c->zeros, 0, 64 * sizeof(double)));
      const double2 sentinel = make_double2(INFINITY, -INFINITY);
      for (int k = 0; k < 2; k++)
      {
         c->xe_mm2[k] = (double2 *)((char *)c->win + c->win_off_mm[k]);
         c->gtr[k] = (double *)((char *)c->win + c->win_off_tr[k]);
         CUDA_OK(cudaMemcpy(c->xe_mm2[k] + ne_all, &sentinel, sizeof(sentinel), cudaMemcpyHostToDevice));
      }
      c->ughost = c->gtr[0];
   }
   if (c->bounds_type == 0)
   {
      if (!d->lat || d->n_ent <= 0) { set_error("bounds_type 0 needs desc.lat / n_ent"); return fail(); }
      c->n_ent = d->n_ent;
      const size_t nl = (size_t)(c->ne + c->ne_ghost) * c->N3;   // owned + ghost rows
      std::vector<int32_t> off((size_t)c->n_ent + 1, 0), el(nl);
      for (size_t i = 0; i < nl; i++)
      {
         if (d->lat[i] < 0 || d->lat[i] >= c->n_ent) { set_error("lat id out of range"); return fail(); }
         off[d->lat[i] + 1]++;
      }
      for (int i = 0; i < c->n_ent; i++) { off[i + 1] += off[i]; }
      std::vector<int32_t> cur(off.begin(), off.end() - 1);
      for (int64_t e = 0; e < c->ne + c->ne_ghost; e++)
         for (int t = 0; t < c->N3; t++) { el[cur[d->lat[e * c->N3 + t]]++] = (int32_t)e; }
      if (dev_upload(c, &c->lat, d->lat, nl)) { return fail(); }
      if (dev_upload(c, &c->ent_off, off.data(), off.size())) { return fail(); }
      if (dev_upload(c, &c->ent_el, el.data(), el.size())) { return fail(); }
      if (dev_alloc(c, &c->ent_mm, 2 * (size_t)c->n_ent)) { return fail(); }
   }
   else
   {
      if (!d->nbr_elem) { set_error("bounds_type 1 needs desc.nbr_elem"); return fail(); }
      if (dev_upload(c, &c->bnbr, d->nbr_elem, (size_t)c->ne * c->NF)) { return fail(); }
   }
   // ---- operator data + scratch
   // 3D: room for the fragment-ordered layout (rows of Q padded to even length)
   const size_t rq = (size_t)((c->Q + 1) & ~1);
   const size_t n_dvol = (c->dim == 3) ? (size_t)c->ne * c->Q * c->Q * rq * 3 : (size_t)c->ne * c->dim * c->NQ;
   const size_t n_dface = (c->dim == 3) ? (size_t)c->ne * c->NF * c->Q * rq : (size_t)c->ne * c->NF * c->NQF;
   c->n_dvol = n_dvol; c->n_dface = n_dface;
   if (dev_alloc(c, &c->Dvol, n_dvol)) { return fail(); }
   if (dev_alloc(c, &c->detJw, (size_t)c->ne * c->NQ)) { return fail(); }
   if (dev_alloc(c, &c->Dface, n_dface)) { return fail(); }
   CUDA_OK(cudaMemset(c->Dvol, 0, n_dvol * sizeof(double)));
   CUDA_OK(cudaMemset(c->Dface, 0, n_dface * sizeof(double)));
   if (dev_alloc(c, &c->ml, (size_t)c->N)) { return fail(); }
   if (dev_alloc(c, &c->einv, (size_t)c->ne)) { return fail(); }
   if (dev_alloc(c, &c->BL, (size_t)c->ne * c->NF * c->NFD)) { return fail(); }
   if (dev_alloc(c, &c->w1, (size_t)c->N)) { return fail(); }
   if (dev_alloc(c, &c->w2, (size_t)c->N)) { return fail(); }
   if (dev_alloc(c, &c->w3, (size_t)c->N)) { return fail(); }
   if (dev_alloc(c, &c->red, 4096)) { return fail(); }
   if (d->inflow) { if (dev_upload(c, &c->inflow, d->inflow, (size_t)c->N)) { return fail(); } }
   if (run_geom(c, 0.0, 0)) { return fail(); }
   CUDA_OK(cudaDeviceSynchronize());
   c->geom_ok = true;
   if (c->exec_mode == 0)
   {
      std::vector<double> ei((size_t)c->ne);
      CUDA_OK(cudaMemcpy(ei.data(), c->einv, ei.size() * sizeof(double), cudaMemcpyDeviceToHost));
      c->all_affine = true;
      for (double v : ei) { if (!(v > 0.0)) { c->all_affine = false; break; } }
      // affine 3D meshes run the tensor-core stage kernel: store the quadrature data in its
      // fragment order (every other 3D kernel reads it through the layout flag)
      if (c->all_affine && c->dim == 3 && c->pipelined && c->tensor && c->D1 <= 5)
      {
         c->frag = true;
         CUDA_OK(cudaMemset(c->Dvol, 0, n_dvol * sizeof(double)));
         CUDA_OK(cudaMemset(c->Dface, 0, n_dface * sizeof(double)));
         if (run_geom(c, 0.0, 0)) { return fail(); }
         CUDA_OK(cudaDeviceSynchronize());
         // velocity linear over every element (constant, rotation, ...): the stage kernel rebuilds
         // the quadrature data from 12 doubles per element instead of streaming it
         const char *nl = getenv("RMH_NO_LINEAR_OP");
         if (!(nl && nl[0] == '1'))
         {
            unsigned int *nfail = nullptr;
            if (dev_alloc(c, &c->opc, (size_t)c->ne * 12)) { return fail(); }
            if (dev_alloc(c, &c->opa, (size_t)c->ne * 4)) { return fail(); }
            if (dev_alloc(c, &nfail, 2)) { return fail(); }
            if (dev_upload(c, &c->dxq, c->hxq.data(), c->hxq.size())) { return fail(); }
            CUDA_OK(cudaMemset(nfail, 0, 2 * sizeof(unsigned int)));
            const int bs = 128;
            k_op_linear<<<(unsigned)((c->ne * 32 + bs - 1) / bs), bs>>>(
               c->Q, c->NGN, c->ne, c->dxq, c->dw, c->Dvol, c->Dface, c->X0, c->einv, c->opc, c->opa, nfail);
            LAUNCH_OK();
            unsigned int nf[2] = {1, 1};
            CUDA_OK(cudaMemcpy(nf, nfail, sizeof(nf), cudaMemcpyDeviceToHost));
            // the rebuilt-data DMMA variant is correct but slower than streaming the stored data
            // (the kernel is not HBM-bound, profiles/r01/README.md): opt-in with RMH_LINEAR_OP=1
            const char *el = getenv("RMH_LINEAR_OP");
            c->op_lin = (nf[0] == 0) && el && el[0] == '1';
            const char *nc = getenv("RMH_NO_CONST_OP");
            // (the constant-coefficient kernel keeps the whole orientation-pattern table in shared memory)
            c->op_const = (nf[1] == 0) && !(nc && nc[0] == '1') && c->npat <= 16;
            if (getenv("RMH_VERBOSE"))
            {
               fprintf(stderr, "k_op_linear: %u of %lld elements not linear, %u not constant\n", nf[0],
                       (long long)c->ne, nf[1]);
            }
         }
      }
   }
   // overlap bounds inside the constant-coefficient stage kernel: needs the 3x3x3 element
   // neighbourhoods to reproduce every lattice entity's element set (structured vertex topology)
   {
      const char *nfold = getenv("RMH_NO_FOLD");
      if (c->op_const && c->bounds_type == 0 && c->dim == 3 && !(nfold && nfold[0] == '1'))
      {
         std::vector<int32_t> nb((size_t)c->ne * c->N3);
         int structured = 0;
         if (nbr_lattice_rows(c->dim, ne_all, c->ne, c->n_ent, d->lat, nb.data(), &structured)) { return fail(); }
         if (structured)
         {
            for (auto &v : nb) { if (v < 0) { v = (int32_t)ne_all; } }     // -> the (inf,-inf) sentinel pair
            if (dev_upload(c, &c->nb27, nb.data(), nb.size())) { return fail(); }
            c->fold = true;
         }
         if (getenv("RMH_VERBOSE")) { fprintf(stderr, "neighbourhood lattice: structured = %d\n", structured); }
      }
   }
   *out = c;
   return 0;
}

extern "C" int64_t rmh_ctx_ndofs(const rmh_ctx *c) { return c->N; }
extern "C" int rmh_ctx_nd(const rmh_ctx *c) { return c->ND; }
extern "C" int rmh_ctx_nq1d(const rmh_ctx *c) { return c->Q; }
extern "C" int rmh_ctx_trust_state(rmh_ctx *c, int on)
{
   c->trust_state = (on != 0); c->xe_ptr = nullptr;
   return 0;
}
// bit 4: overlap bounds formed inside the stage kernel (k_stage3c<FOLD>)
extern "C" int rmh_ctx_path_flags(const rmh_ctx *c)
{ return (c->all_affine ? 1 : 0) | (c->frag ? 2 : 0) | (c->op_lin ? 4 : 0) | (c->op_const ? 8 : 0) | (c->fold ? 16 : 0); }
extern "C" int rmh_ctx_quad_points_1d(const rmh_ctx *c, double *q1d, double *w1d)
{
   for (int q = 0; q < c->Q; q++) { if (q1d) { q1d[q] = c->hxq[q]; } if (w1d) { w1d[q] = c->hw[q]; } }
   return 0;
}

static void swap_geom_sets(rmh_ctx *c)
{
   GeomSet &g = c->gspare;
   std::swap(c->Dvol, g.Dvol); std::swap(c->detJw, g.detJw); std::swap(c->Dface, g.Dface);
   std::swap(c->ml, g.ml); std::swap(c->einv, g.einv); std::swap(c->BL, g.BL);
   const double t = c->t_cur; const bool ok = c->geom_ok;
   c->t_cur = g.t; c->geom_ok = g.valid;
   g.t = t; g.valid = ok;
}

// Remap: the operators are functions of t only (mesh position x0 + t v, remhos.cpp:1598-1677).  The context
// keeps the data of the last TWO distinct times: an RK3-SSP step asks for t, t + dt, t + dt/2 and the next
// step starts at t + dt again, so one rebuild in three is a pointer swap; asking for the current time again
// (rmh_mult_unlimited followed by rmh_limit_mult) costs nothing.  Matrix-based solvers and subcell weights
// keep their data outside the sets: with them every call rebuilds, as before.  RMH_NO_GEOM_CACHE=1: ditto.
extern "C" int rmh_set_time(rmh_ctx *c, double t, void *stream)
{
   if (c->exec_mode != 1) { c->t_cur = t; return 0; }
   static int no_cache = -1;
   if (no_cache < 0) { const char *ev = getenv("RMH_NO_GEOM_CACHE"); no_cache = (ev && ev[0] == '1') ? 1 : 0; }
   if (no_cache || c->fa_on || c->sub_on)
   {
      c->gspare.valid = false;
      c->geom_ok = false;
      const int rc = run_geom(c, t, (cudaStream_t)stream);
      c->geom_ok = (rc == 0);
      return rc;
   }
   if (c->geom_ok && c->t_cur == t) { return 0; }
   GeomSet &g = c->gspare;
   if (!g.Dvol)
   {
      if (dev_alloc(c, &g.Dvol, c->n_dvol) || dev_alloc(c, &g.detJw, (size_t)c->ne * c->NQ) ||
          dev_alloc(c, &g.Dface, c->n_dface) || dev_alloc(c, &g.ml, (size_t)c->N) ||
          dev_alloc(c, &g.einv, (size_t)c->ne) || dev_alloc(c, &g.BL, (size_t)c->ne * c->NF * c->NFD)) { return 1; }
      CUDA_OK(cudaMemsetAsync(g.Dvol, 0, c->n_dvol * sizeof(double), (cudaStream_t)stream));     // padding rows
      CUDA_OK(cudaMemsetAsync(g.Dface, 0, c->n_dface * sizeof(double), (cudaStream_t)stream));
      g.valid = false;
   }
   swap_geom_sets(c);                                  // the set of the previous time stays around
   if (c->geom_ok && c->t_cur == t) { return 0; }      // ... and this one was built for t
   c->geom_ok = false;
   const int rc = run_geom(c, t, (cudaStream_t)stream);
   c->geom_ok = (rc == 0);
   return rc;
}

extern "C" int rmh_lumped_mass(rmh_ctx *c, double *m_dev, void *stream)
{
   CUDA_OK(cudaMemcpyAsync(m_dev, c->ml, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
   return 0;
}

static HoArgs ho_args(rmh_ctx *c, const double *in, double *out, int mode)
{
   HoArgs a;
   a.ne = c->ne; a.u = in; a.out = out;
   a.Dvol = c->Dvol; a.detJw = c->detJw; a.Dface = c->Dface; a.einv = c->einv;
   a.fn.nbr_elem = c->nbr_elem; a.fn.nbr_pat = c->nbr_pat; a.fn.pat = c->pat;
   a.fn.ughost = c->ughost; a.fn.ne_owned = c->ne;
   a.mode = mode; a.tol2 = c->pcg_tol2; a.maxit = c->pcg_maxit; a.frag = c->frag ? 1 : 0;
   return a;
}

extern "C" int rmh_ho_mult(rmh_ctx *c, const double *u, double *rhs, void *stream)
{
   return dispatch_ho(c, ho_args(c, u, rhs, 1), (cudaStream_t)stream);
}
extern "C" int rmh_mass_inv(rmh_ctx *c, const double *rhs, double *du, void *stream)
{
   return dispatch_ho(c, ho_args(c, rhs, du, 2), (cudaStream_t)stream);
}
extern "C" int rmh_ho_local_inverse(rmh_ctx *c, const double *u, double *du, void *stream)
{
   return dispatch_ho(c, ho_args(c, u, du, 3), (cudaStream_t)stream);
}

extern "C" int rmh_lo_mass_avg(rmh_ctx *c, double dt, const double *u, const double *du_ho,
                               double *du_lo, void *stream)
{
   const int bs = 256;
   const int64_t nb = (c->ne * 32 + bs - 1) / bs;
   k_mass_avg<<<(unsigned)nb, bs, 0, (cudaStream_t)stream>>>(c->ne, c->ND, dt, u, du_ho, c->ml, du_lo);
   LAUNCH_OK();
   return 0;
}

extern "C" int rmh_elem_min_max(rmh_ctx *c, const double *u, double *xe_min, double *xe_max,
                                void *stream)
{
   if (xe_min == c->xe_min || xe_max == c->xe_max) { c->xe_ptr = nullptr; }   // the cached min/max change owner
   launch_elem_min_max(c->ne, c->ND, u, xe_min, xe_max, (cudaStream_t)stream);
   LAUNCH_OK();
   return 0;
}

extern "C" int rmh_bounds(rmh_ctx *c, const double *xe_min, const double *xe_max, double *xi_min,
                          double *xi_max, void *stream)
{
   cudaStream_t s = (cudaStream_t)stream;
   const int bs = 256;
   if (c->bounds_type == 0)
   {
      const int64_t na = c->ne + c->ne_ghost;
      k_xe_interleave<<<(unsigned)((na + bs - 1) / bs), bs, 0, s>>>(na, xe_min, xe_max, c->xe_mm);
      k_ent_min_max<<<(c->n_ent + bs - 1) / bs, bs, 0, s>>>(c->n_ent, nullptr, c->ent_off, c->ent_el, c->xe_mm,
                                                            c->ent_mm);
      LAUNCH_OK();
      k_bounds_overlap<<<(unsigned)((c->N + bs - 1) / bs), bs, 0, s>>>(
         c->ne, c->dim, c->D1, c->ND, c->N3, c->lat, c->ent_mm, xi_min, xi_max);
      LAUNCH_OK();
   }
   else
   {
      k_bounds_sparsity<<<(unsigned)((c->N + bs - 1) / bs), bs, 0, s>>>(c->ne, c->NF, c->ND, c->bnbr,
                                                                        xe_min, xe_max, xi_min, xi_max);
      LAUNCH_OK();
   }
   return 0;
}

extern "C" int rmh_fct_clip_scale(rmh_ctx *c, double dt, const double *u, const double *m,
                                  const double *du_ho, const double *du_lo, const double *xi_min,
                                  const double *xi_max, double *du, void *stream)
{
   if (c->ND > 256) { set_error("clip_scale: nd > 256 unsupported"); return 1; }
   const int bs = 256;
   const int64_t nb = (c->ne * 32 + bs - 1) / bs;
   k_clip_scale<<<(unsigned)nb, bs, 0, (cudaStream_t)stream>>>(c->ne, c->ND, dt, u, m, du_ho, du_lo,
                                                               xi_min, xi_max, du);
   LAUNCH_OK();
   return 0;
}

extern "C" int rmh_reduce(rmh_ctx *c, int op, const double *a, const double *b, double *out,
                          void *stream)
{
   cudaStream_t s = (cudaStream_t)stream;
   const int nb = 1024, bs = 256;
   k_reduce<<<nb, bs, 0, s>>>(c->N, op, a, b, c->red);
   LAUNCH_OK();
   std::vector<double> part(nb);
   CUDA_OK(cudaMemcpyAsync(part.data(), c->red, nb * sizeof(double), cudaMemcpyDeviceToHost, s));
   CUDA_OK(cudaStreamSynchronize(s));
   double v = (op == 0) ? 0.0 : (op == 1 ? INFINITY : -INFINITY);
   for (int i = 0; i < nb; i++)
   {
      if (op == 0) { v += part[i]; }
      else if (op == 1) { v = std::min(v, part[i]); }
      else { v = std::max(v, part[i]); }
   }
   *out = v;
   return 0;
}

// ---------------------------------------------------------------- fused stage entry points
// Element (min,max) of a stage input live in the context: as pairs in xe_mm2[(epoch + 1) & 1] when
// the bounds are folded into the stage kernel (c->fold), else in xe_min / xe_max.
static int stage_minmax(rmh_ctx *c, const double *y, cudaStream_t s)
{
   c->xe_ptr = nullptr;
   if (c->fold) { launch_elem_min_max(c->ne, c->ND, y, nullptr, nullptr, s, c->xe_mm2[(c->epoch + 1) & 1]); }
   else { launch_elem_min_max(c->ne, c->ND, y, c->xe_min, c->xe_max, s); }
   LAUNCH_OK();
   return 0;
}

// xe_valid: the context already holds the element min/max of y (and, on a decomposed mesh, the
// peers' k_halo_put for stage epoch + 1 has been issued: dist.cuh).  Every call is one "stage
// epoch": ghost traces and (min,max) pairs of parity epoch & 1 are read, pairs of the output are
// written with the other parity.
static int stage_impl(rmh_ctx *c, int lo_type, double dt, int out_mode, double a, double b,
                      const double *x0, const double *y, double *out, bool xe_valid,
                      bool write_xe, cudaStream_t s)
{
   if (lo_type != 5)
   {
      set_error("fused stage: only -lo 5 (MassBasedAvg) is fused; use the separate LO entry points");
      return 1;
   }
   if (out == y) { set_error("stage: output must not alias the stage input"); return 1; }
   c->xe_ptr = nullptr;     // the context's element min/max are about to change owner
   const int bs = 256;
   if (!xe_valid) { if (stage_minmax(c, y, s)) { return 1; } }
   c->epoch++;
   const int par = (int)(c->epoch & 1);
   c->ughost = c->gtr[par];
   c->halo_ptr = nullptr;
   // 3D meshes with constant-Jacobian elements: persistent pipelined kernels
   const bool use_p = c->pipelined && c->dim == 3 && c->all_affine &&
                      ((((uintptr_t)y | (uintptr_t)x0 | (uintptr_t)out) & 15) == 0);
   const bool use_c = use_p && c->frag && c->op_const && c->npat <= 16;
   const bool fold = use_c && c->fold;
   if (c->fold && !fold) { set_error("stage: state vectors must be 16-byte aligned"); return 1; }
   if (c->bounds_type == 0 && !fold)
   {
      const int64_t na = c->ne + c->ne_ghost;
      k_xe_interleave<<<(unsigned)((na + bs - 1) / bs), bs, 0, s>>>(na, c->xe_min, c->xe_max, c->xe_mm);
      k_ent_min_max<<<(c->n_ent + bs - 1) / bs, bs, 0, s>>>(c->n_ent, nullptr, c->ent_off, c->ent_el, c->xe_mm,
                                                            c->ent_mm);
      LAUNCH_OK();
   }
   StageArgs sa;
   sa.ho = ho_args(c, y, nullptr, 3);
   sa.ml = c->ml; sa.x0 = x0; sa.out = out; sa.a = a; sa.b = b; sa.dt = dt; sa.out_mode = out_mode;
   sa.bounds_type = c->bounds_type; sa.dim_n3 = c->N3; sa.lat = c->lat;
   sa.ent_mm = c->ent_mm; sa.bnbr = c->bnbr;
   sa.xe_min = c->xe_min; sa.xe_max = c->xe_max;
   // with overlap bounds the stage kernel leaves the element min/max of its output for the next
   // stage (bounds_type 1 reads the neighbours' values during the kernel: no in-place update)
   sa.xe_min_out = nullptr; sa.xe_max_out = nullptr;
   if (write_xe && c->bounds_type == 0) { sa.xe_min_out = c->xe_min; sa.xe_max_out = c->xe_max; }
   StagePArgs pa;
   if (use_p)
   {
      pa.ne = c->ne; pa.y = y; pa.x0 = x0; pa.out = out;
      pa.e_begin = 0;
      pa.Dvol = c->Dvol; pa.Dface = c->Dface; pa.einv = c->einv;
      pa.opc = c->op_lin ? c->opc : nullptr;
      pa.opa = c->op_const ? c->opa : nullptr;
      pa.fn = sa.ho.fn; pa.npat = c->npat; pa.nbr_pat32 = c->nbr_pat32;
      pa.a = a; pa.b = b; pa.dt = dt; pa.out_mode = out_mode;
      pa.has_x0 = (out_mode == 1 && a != 0.0) ? 1 : 0;
      pa.frag = c->frag ? 1 : 0;
      pa.bounds_type = c->bounds_type;
      pa.bidx = (c->bounds_type == 0) ? c->lat : c->bnbr;
      pa.ent_mm = c->ent_mm; pa.xe_min = c->xe_min; pa.xe_max = c->xe_max;
      pa.xe_min_out = sa.xe_min_out; pa.xe_max_out = sa.xe_max_out;
      if (fold)
      {
         pa.bidx = c->nb27;
         pa.ent_mm = reinterpret_cast<const double *>(c->xe_mm2[par]);
         pa.xe_mm_out = c->xe_mm2[par ^ 1];
      }
      pa.zeros = c->zeros;
      dist_stage_args(c, pa, fold);
   }
   auto run = [&]()
   {
      if (use_p && c->frag) { return use_c ? dispatch_stagec(c, pa, s) : dispatch_stagew(c, pa, s); }
      return use_p ? dispatch_stagep(c, pa, s) : dispatch_stage(c, sa, s);
   };
   if (c->prof)
   {
      if (c->prof_used + 2 > c->prof_ev.size())
      {
         for (int i = 0; i < 64; i++)
         {
            cudaEvent_t ev;
            CUDA_OK(cudaEventCreate(&ev));
            c->prof_ev.push_back(ev);
         }
      }
      CUDA_OK(cudaEventRecord(c->prof_ev[c->prof_used], s));
      const int rc = run();
      CUDA_OK(cudaEventRecord(c->prof_ev[c->prof_used + 1], s));
      c->prof_used += 2;
      return rc;
   }
   return run();
}

extern "C" int rmh_profile(rmh_ctx *c, int enable, double *total_ms, int64_t *launches)
{
   // enable: 1 start/continue recording, 0 stop; always reports (and clears) what was recorded
   double tot = 0.0;
   int64_t n = 0;
   for (size_t i = 0; i + 1 < c->prof_used; i += 2)
   {
      CUDA_OK(cudaEventSynchronize(c->prof_ev[i + 1]));
      float ms = 0.f;
      CUDA_OK(cudaEventElapsedTime(&ms, c->prof_ev[i], c->prof_ev[i + 1]));
      tot += ms; n++;
   }
   c->prof_used = 0;
   c->prof = enable != 0;
   if (total_ms) { *total_ms = tot; }
   if (launches) { *launches = n; }
   return 0;
}

extern "C" int rmh_stage(rmh_ctx *c, int lo_type, double dt, const double *u, double *k, void *stream)
{
   return stage_impl(c, lo_type, dt, 0, 0.0, 0.0, u, u, k, false, false, (cudaStream_t)stream);
}

extern "C" int rmh_rk_stage(rmh_ctx *c, int lo_type, double dt, double a, double b,
                            const double *x0, const double *y, double *out, void *stream)
{
   return stage_impl(c, lo_type, dt, 1, a, b, x0, y, out, false, false, (cudaStream_t)stream);
}

extern "C" int rmh_rk_step(rmh_ctx *c, int ode, int lo_type, double *t, double dt, double *u,
                           void *stream)
{
   cudaStream_t s = (cudaStream_t)stream;
   const double t0 = *t;
   // with overlap bounds the stage kernel leaves the element min/max of its output in the
   // context, so only the first stage needs the stand-alone min/max pass
   const bool chain = (c->bounds_type == 0);
   // rmh_ctx_trust_state: the last stage of the previous step left the element min/max of this
   // very vector in the context
   const bool have_xe = chain && c->trust_state && c->xe_ptr == u;
   const bool keep_xe = chain && c->trust_state;
   c->xe_ptr = nullptr;
   if (ode == 1)          // ForwardEulerSolver
   {
      if (rmh_set_time(c, t0, stream)) { return 1; }
      if (stage_impl(c, lo_type, dt, 1, 0.0, 1.0, u, u, c->w1, false, false, s)) { return 1; }
      CUDA_OK(cudaMemcpyAsync(u, c->w1, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToDevice, s));
   }
   else if (ode == 2)     // RK2Solver(1.0): u_new = 1/2 u + 1/2 (y + dt F(y)), y = u + dt F(u)
   {
      if (rmh_set_time(c, t0, stream)) { return 1; }
      if (stage_impl(c, lo_type, dt, 1, 0.0, 1.0, u, u, c->w1, have_xe, chain, s)) { return 1; }
      if (rmh_set_time(c, t0 + dt, stream)) { return 1; }
      if (stage_impl(c, lo_type, dt, 1, 0.5, 0.5, u, c->w1, u, chain, keep_xe, s)) { return 1; }
      if (keep_xe) { c->xe_ptr = u; }
   }
   else if (ode == 3)     // RK3SSPSolver
   {
      if (rmh_set_time(c, t0, stream)) { return 1; }
      if (stage_impl(c, lo_type, dt, 1, 0.0, 1.0, u, u, c->w1, have_xe, chain, s)) { return 1; }
      if (rmh_set_time(c, t0 + dt, stream)) { return 1; }
      if (stage_impl(c, lo_type, dt, 1, 0.75, 0.25, u, c->w1, c->w2, chain, chain, s)) { return 1; }
      if (rmh_set_time(c, t0 + dt / 2, stream)) { return 1; }
      if (stage_impl(c, lo_type, dt, 1, 1.0 / 3.0, 2.0 / 3.0, u, c->w2, u, chain, keep_xe, s)) { return 1; }
      if (keep_xe) { c->xe_ptr = u; }
   }
   else
   {
      set_error("rmh_rk_step: ode solver type must be 1, 2 or 3 (use rmh_stage for general RK)");
      return 1;
   }
   *t = t0 + dt;
   return 0;
}

extern "C" int rmh_rk_step_host(rmh_ctx *c, int ode, int lo_type, double *t, double dt, double *u_host)
{
   // end-to-end entry point: state lives in host memory (as the reference's ODESolver vectors do,
   // remhos.cpp:1680); H2D, one step, D2H.  u_host should be pinned for full PCIe bandwidth.
   const size_t bytes = (size_t)c->N * sizeof(double);
   if (rmh_host_sync(c)) { return 1; }      // queued steps share the stage intermediates with this call
   CUDA_OK(cudaMemcpyAsync(c->w3, u_host, bytes, cudaMemcpyHostToDevice, 0));
   c->xe_ptr = nullptr;     // fresh state from the host
   if (rmh_rk_step(c, ode, lo_type, t, dt, c->w3, nullptr)) { return 1; }
   CUDA_OK(cudaMemcpyAsync(u_host, c->w3, bytes, cudaMemcpyDeviceToHost, 0));
   CUDA_OK(cudaStreamSynchronize(0));
   return 0;
}

// ---- pipelined host-state stepping.  A step of a host-resident state is three transfers over two different
// resources: H2D (PCIe down), the stages (HBM), D2H (PCIe up).  rmh_rk_step_host runs them back to back;
// here they are queued on three streams with a ring of device buffers, so that consecutive calls overlap:
//  * independent states (several fields advected by the same velocity, one call each): H2D of call n+1,
//    the stages of call n and D2H of call n-1 run at the same time;
//  * the SAME host buffer stepped repeatedly: the copies are cut into slabs and slab k of the next H2D is
//    ordered behind slab k of the previous D2H only, so both directions of the link stay busy.
// Host buffers of different calls must be identical or disjoint.
struct HostPipe
{
   static constexpr int NBUF = 3, NSLAB = 16;
   cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
   double *buf[NBUF] = {nullptr, nullptr, nullptr};
   cudaEvent_t in_done[NBUF] = {}, cmp_done[NBUF] = {}, out_done[NBUF] = {}, slab_out[NBUF][NSLAB] = {};
   const double *in_host[NBUF] = {nullptr, nullptr, nullptr};
   double *out_host[NBUF] = {nullptr, nullptr, nullptr};
   uint64_t n = 0;
};

static void host_pipe_free(rmh_ctx *c)
{
   HostPipe *P = c->pipe;
   if (!P) { return; }
   // queued work may still read the buffers that are about to be freed
   for (cudaStream_t st : {P->s_in, P->s_cmp, P->s_out}) { if (st) { cudaStreamSynchronize(st); } }
   auto drop = [](cudaEvent_t e) { if (e) { cudaEventDestroy(e); } };
   for (int i = 0; i < HostPipe::NBUF; i++)
   {
      drop(P->in_done[i]); drop(P->cmp_done[i]); drop(P->out_done[i]);
      for (int k = 0; k < HostPipe::NSLAB; k++) { drop(P->slab_out[i][k]); }
   }
   for (cudaStream_t st : {P->s_in, P->s_cmp, P->s_out}) { if (st) { cudaStreamDestroy(st); } }
   delete P;
   c->pipe = nullptr;
}

static int host_pipe_get(rmh_ctx *c, HostPipe **out)
{
   if (c->pipe) { *out = c->pipe; return 0; }
   HostPipe *P = new HostPipe;
   c->pipe = P;
   CUDA_OK(cudaStreamCreateWithFlags(&P->s_in, cudaStreamNonBlocking));
   CUDA_OK(cudaStreamCreateWithFlags(&P->s_cmp, cudaStreamNonBlocking));
   CUDA_OK(cudaStreamCreateWithFlags(&P->s_out, cudaStreamNonBlocking));
   for (int i = 0; i < HostPipe::NBUF; i++)
   {
      if (dev_alloc(c, &P->buf[i], (size_t)c->N)) { return 1; }
      CUDA_OK(cudaEventCreateWithFlags(&P->in_done[i], cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&P->cmp_done[i], cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&P->out_done[i], cudaEventDisableTiming));
      for (int k = 0; k < HostPipe::NSLAB; k++) { CUDA_OK(cudaEventCreateWithFlags(&P->slab_out[i][k], cudaEventDisableTiming)); }
   }
   *out = P;
   return 0;
}

// step(u_dev, stream) advances the device copy in place
template <typename StepFn>
static int host_pipe_step(rmh_ctx *c, const double *u_in_host, double *u_out_host, StepFn step)
{
   if (!u_in_host || !u_out_host) { set_error("rmh_rk_step_host_async: null host buffer"); return 1; }
   HostPipe *P = nullptr;
   if (host_pipe_get(c, &P)) { return 1; }
   constexpr int NB = HostPipe::NBUF, NS = HostPipe::NSLAB;
   const int i = (int)(P->n % NB);
   const int64_t N = c->N, per = ((N + NS - 1) / NS + 1) & ~(int64_t)1;
   // the device buffer is free once the D2H of the call that used it last is through
   if (P->n >= (uint64_t)NB) { CUDA_OK(cudaStreamWaitEvent(P->s_in, P->out_done[i], 0)); }
   // a call in flight that writes the buffer we are about to read: follow it slab by slab
   int dep = -1;
   for (int r = 0; r < NB; r++) { if (r != i && P->out_host[r] == u_in_host && P->n > 0) { dep = r; } }
   // ... and the newest such call only (older ones are ordered before it on s_out)
   if (dep >= 0)
   {
      const int last = (int)((P->n - 1) % NB);
      if (P->out_host[last] == u_in_host) { dep = last; }
   }
   for (int k = 0; k < NS; k++)
   {
      const int64_t o = k * per, len = std::min<int64_t>(per, N - o);
      if (len <= 0) { break; }
      if (dep >= 0) { CUDA_OK(cudaStreamWaitEvent(P->s_in, P->slab_out[dep][k], 0)); }
      CUDA_OK(cudaMemcpyAsync(P->buf[i] + o, u_in_host + o, (size_t)len * sizeof(double), cudaMemcpyHostToDevice, P->s_in));
   }
   CUDA_OK(cudaEventRecord(P->in_done[i], P->s_in));
   CUDA_OK(cudaStreamWaitEvent(P->s_cmp, P->in_done[i], 0));
   c->xe_ptr = nullptr; c->sent_ptr = nullptr;          // fresh state from the host
   if (step(P->buf[i], P->s_cmp)) { return 1; }
   c->xe_ptr = nullptr;
   CUDA_OK(cudaEventRecord(P->cmp_done[i], P->s_cmp));
   CUDA_OK(cudaStreamWaitEvent(P->s_out, P->cmp_done[i], 0));
   // a call in flight that still reads the buffer we are about to write (other than this one)
   for (int r = 0; r < NB; r++)
   { if (r != i && P->in_host[r] == u_out_host && P->n > 0) { CUDA_OK(cudaStreamWaitEvent(P->s_out, P->in_done[r], 0)); } }
   for (int k = 0; k < NS; k++)
   {
      const int64_t o = k * per, len = std::min<int64_t>(per, N - o);
      if (len > 0)
      { CUDA_OK(cudaMemcpyAsync(u_out_host + o, P->buf[i] + o, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, P->s_out)); }
      CUDA_OK(cudaEventRecord(P->slab_out[i][k], P->s_out));
   }
   CUDA_OK(cudaEventRecord(P->out_done[i], P->s_out));
   P->in_host[i] = u_in_host; P->out_host[i] = u_out_host;
   P->n++;
   return 0;
}

extern "C" int rmh_rk_step_host_async(rmh_ctx *c, int ode, int lo_type, double t, double dt, const double *u_in_host,
                                      double *u_out_host)
{
   return host_pipe_step(c, u_in_host, u_out_host, [&](double *u, cudaStream_t s)
   {
      double tt = t;
      return rmh_rk_step(c, ode, lo_type, &tt, dt, u, (void *)s);
   });
}

extern "C" int rmh_host_sync(rmh_ctx *c)
{
   HostPipe *P = c->pipe;
   if (!P) { return 0; }
   CUDA_OK(cudaStreamSynchronize(P->s_in));
   CUDA_OK(cudaStreamSynchronize(P->s_cmp));
   CUDA_OK(cudaStreamSynchronize(P->s_out));
   for (int r = 0; r < HostPipe::NBUF; r++) { P->in_host[r] = nullptr; P->out_host[r] = nullptr; }
   return 0;
}

// ---------------------------------------------------------------- matrix-based solver entry points
static FaArgs fa_args(rmh_ctx *c)
{
   FaArgs A;
   A.ne = c->ne; A.dim = c->dim; A.D1 = c->D1; A.ND = c->ND; A.NF = c->NF; A.NFD = c->NFD;
   A.fn.nbr_elem = c->nbr_elem; A.fn.nbr_pat = c->nbr_pat; A.fn.pat = c->pat;
   A.fn.ughost = c->ughost; A.fn.ne_owned = c->ne;
   A.pat_idx = c->pat_idx; A.pat_face = c->pat_face;
   A.K = c->faK; A.KH = c->faKH; A.M = c->faM; A.BI = c->faBI; A.BL = c->BL; A.ml = c->ml;
   A.inflow = c->inflow;
   return A;
}

static int64_t state_len(const rmh_ctx *c) { return c->product ? 2 * c->N : c->N; }
static int work_vec(rmh_ctx *c, double **slot)
{
   if (*slot) { return 0; }
   return dev_alloc(c, slot, (size_t)state_len(c));
}

extern "C" int rmh_fa_setup(rmh_ctx *c, void *stream)
{
   if (c->fa_on) { return 0; }
   const size_t nn = (size_t)c->ne * c->ND * c->ND;
   if (dev_alloc(c, &c->faK, nn) || dev_alloc(c, &c->faKH, nn) || dev_alloc(c, &c->faM, nn) ||
       dev_alloc(c, &c->faBI, (size_t)c->ne * c->NF * c->NFD * c->NFD)) { return 1; }
   if (c->ne_ghost > 0 && c->n_gslots > 0)
   {
      const size_t ng = (size_t)c->n_gslots * c->NFD;
      if (dev_alloc(c, &c->faBIg, ng * c->NFD) || dev_alloc(c, &c->gtr_u, ng) || dev_alloc(c, &c->gtr_cp, ng) ||
          dev_alloc(c, &c->gtr_cn, ng)) { return 1; }
   }
   c->fa_on = true;
   return run_geom(c, c->t_cur, (cudaStream_t)stream);
}

extern "C" int rmh_fa_get(rmh_ctx *c, int which, double *host_out)
{
   if (!c->fa_on && which != 4) { set_error("rmh_fa_get: call rmh_fa_setup first"); return 1; }
   const size_t nn = (size_t)c->ne * c->ND * c->ND;
   const double *src = nullptr;
   size_t n = nn;
   switch (which)
   {
      case 0: src = c->faK; break;
      case 1: src = c->faKH; break;
      case 2: src = c->faM; break;
      case 3: src = c->faBI; n = (size_t)c->ne * c->NF * c->NFD * c->NFD; break;
      case 4: src = c->BL; n = (size_t)c->ne * c->NF * c->NFD; break;
      default: set_error("rmh_fa_get: which must be 0..4"); return 1;
   }
   CUDA_OK(cudaDeviceSynchronize());
   CUDA_OK(cudaMemcpy(host_out, src, n * sizeof(double), cudaMemcpyDeviceToHost));
   return 0;
}

extern "C" int rmh_lo_discrete_upwind(rmh_ctx *c, const double *u, double *du_lo, void *stream)
{
   if (!c->fa_on) { set_error("rmh_lo_discrete_upwind: call rmh_fa_setup first (assembled K)"); return 1; }
   const int bs = std::min(256, ((c->ND + 31) / 32) * 32);
   k_lo_du<<<(unsigned)c->ne, bs, c->ND * sizeof(double), (cudaStream_t)stream>>>(fa_args(c), u, du_lo);
   LAUNCH_OK();
   return 0;
}

// PrecondConvectionIntegrator blocks M_L M^-1 K (remhos_tools.cpp:975-1031), column by column
// through the exact element mass inverse; rebuilt when the mesh has moved
static int build_precond_conv(rmh_ctx *c, cudaStream_t s)
{
   if (c->faKP && (c->exec_mode == 0 || c->kp_t == c->t_cur)) { return 0; }
   const size_t nn = (size_t)c->ne * c->ND * c->ND;
   if (!c->faKP) { if (dev_alloc(c, &c->faKP, nn)) { return 1; } }
   if (work_vec(c, &c->wk[0]) || work_vec(c, &c->wk[1])) { return 1; }
   const int bs = 256;
   const unsigned nb = (unsigned)((c->N + bs - 1) / bs);
   for (int j = 0; j < c->ND; j++)
   {
      k_col_get<<<nb, bs, 0, s>>>(c->ne, c->ND, j, c->faK, c->wk[0]);
      LAUNCH_OK();
      if (dispatch_ho(c, ho_args(c, c->wk[0], c->wk[1], 2), s)) { return 1; }
      k_col_put<<<nb, bs, 0, s>>>(c->ne, c->ND, j, c->wk[1], c->ml, c->faKP);
      LAUNCH_OK();
   }
   c->kp_t = c->t_cur;
   return 0;
}

extern "C" int rmh_lo_discrete_upwind_prec(rmh_ctx *c, const double *u, double *du_lo, void *stream)
{
   if (!c->fa_on) { set_error("rmh_lo_discrete_upwind_prec: call rmh_fa_setup first (assembled K, M)"); return 1; }
   if (build_precond_conv(c, (cudaStream_t)stream)) { return 1; }
   FaArgs A = fa_args(c);
   A.K = c->faKP;
   const int bs = std::min(256, ((c->ND + 31) / 32) * 32);
   k_lo_du<<<(unsigned)c->ne, bs, c->ND * sizeof(double), (cudaStream_t)stream>>>(A, u, du_lo);
   LAUNCH_OK();
   return 0;
}

extern "C" int rmh_ho_neumann(rmh_ctx *c, const double *u, double *du, void *stream)
{
   if (!c->fa_on) { set_error("rmh_ho_neumann: call rmh_fa_setup first (the Neumann solver is FA only, remhos_ho.cpp:138-139)"); return 1; }
   cudaStream_t s = (cudaStream_t)stream;
   if (work_vec(c, &c->wk[0]) || work_vec(c, &c->wk[1])) { return 1; }
   double *rhs = c->wk[0], *res = c->wk[1];
   if (dispatch_ho(c, ho_args(c, u, rhs, 1 | 4), s)) { return 1; }
   const int bs = 256;
   const unsigned nb = (unsigned)((c->N + bs - 1) / bs);
   k_face_galerkin<<<nb, bs, 0, s>>>(fa_args(c), u, rhs);
   LAUNCH_OK();
   CUDA_OK(cudaMemsetAsync(du, 0, (size_t)c->N * sizeof(double), s));
   for (int iter = 1; iter <= 20; iter++)
   {
      k_mass_residual<<<nb, bs, 0, s>>>(c->ne, c->ND, c->faM, du, rhs, res);
      LAUNCH_OK();
      double r2 = 0.0;
      if (rmh_reduce(c, 0, res, res, &r2, stream)) { return 1; }
      // the stopping test is on the GLOBAL residual (MPI_Allreduce, remhos_ho.cpp:174-177): every rank of a
      // decomposed run does the same number of sweeps as the single-GPU run
      if (c->dist) { if (rmh_dist_allreduce(c->dist, 0, &r2, 1, stream)) { return 1; } }
      if (std::sqrt(r2) <= 1.0e-4) { return 0; }
      k_neumann_update<<<nb, bs, 0, s>>>(c->N, res, c->ml, du);
      LAUNCH_OK();
   }
   return 0;
}

extern "C" int rmh_lo_res_dist(rmh_ctx *c, const double *u, double *du_lo, void *stream)
{
   // z = K u with the volume-only convection operator (sum-factorised), then the element-local
   // redistribution
   if (work_vec(c, &c->wk[0])) { return 1; }
   if (dispatch_ho(c, ho_args(c, u, c->wk[0], 1 | 4), (cudaStream_t)stream)) { return 1; }
   const int bs = 256;
   const int64_t nb = (c->ne * 32 + bs - 1) / bs;
   k_lo_rd<<<(unsigned)nb, bs, 0, (cudaStream_t)stream>>>(fa_args(c), u, c->wk[0], du_lo);
   LAUNCH_OK();
   return 0;
}

extern "C" int rmh_subcell_setup(rmh_ctx *c, const double *xlat_host, const double *vel_host,
                                 void *stream)
{
   if (c->p < 2) { set_error("Subcell schemes require FE order > 1."); return 1; }   // remhos.cpp:613-616
   int ns = 1, nc = 1;
   for (int a = 0; a < c->dim; a++) { ns *= c->p; nc *= 2; }
   const size_t nx = (size_t)c->ne * c->ND * c->dim;
   const size_t nv = (c->exec_mode == 1) ? nx : (size_t)c->ne * ns * c->dim;
   if (!c->sub_on)
   {
      if (dev_alloc(c, &c->sub_x, nx) || dev_alloc(c, &c->sub_v, nv) ||
          dev_alloc(c, &c->sub_w, (size_t)c->ne * ns * nc)) { return 1; }
   }
   CUDA_OK(cudaMemcpy(c->sub_x, xlat_host, nx * sizeof(double), cudaMemcpyHostToDevice));
   CUDA_OK(cudaMemcpy(c->sub_v, vel_host, nv * sizeof(double), cudaMemcpyHostToDevice));
   c->sub_on = true;
   return run_geom(c, c->t_cur, (cudaStream_t)stream);
}

extern "C" int rmh_lo_res_dist_subcell(rmh_ctx *c, const double *u, double *du_lo, void *stream)
{
   if (!c->sub_on) { set_error("rmh_lo_res_dist_subcell: call rmh_subcell_setup first"); return 1; }
   if (work_vec(c, &c->wk[0])) { return 1; }
   if (dispatch_ho(c, ho_args(c, u, c->wk[0], 1 | 4), (cudaStream_t)stream)) { return 1; }
   int ns = 1;
   for (int a = 0; a < c->dim; a++) { ns *= c->p; }
   const int wpb = 4;
   const int64_t nb = (c->ne + wpb - 1) / wpb;
   k_lo_rd_sub<<<(unsigned)nb, wpb * 32, (size_t)wpb * ns * 6 * sizeof(double), (cudaStream_t)stream>>>(
      fa_args(c), c->p, c->sub_w, u, c->wk[0], du_lo);
   LAUNCH_OK();
   return 0;
}

// ---- SmoothnessIndicator (remhos_tools.cpp:24-354) for order-1 spaces: the H1 space of positive
// order-1 elements on the mesh itself (remhos.cpp:870) has one DOF per vertex.  Element matrices by
// a 3-point Gauss rule on the multilinear map through the element corners (exact on parallelograms).
namespace
{
struct Trip { int32_t r, c; double v; };
void to_csr(int nrows, std::vector<Trip> &t, std::vector<int32_t> &I, std::vector<int32_t> &J,
            std::vector<double> &A)
{
   std::sort(t.begin(), t.end(), [](const Trip &a, const Trip &b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
   I.assign((size_t)nrows + 1, 0); J.clear(); A.clear();
   for (size_t k = 0; k < t.size(); k++)
   {
      if (k > 0 && t[k].r == t[k - 1].r && t[k].c == t[k - 1].c) { A.back() += t[k].v; continue; }
      J.push_back(t[k].c); A.push_back(t[k].v); I[t[k].r + 1]++;
   }
   for (int i = 0; i < nrows; i++) { I[i + 1] += I[i]; }
}
// Q1 mass and (negative stiffness + boundary normal-derivative) matrices of one element
void q1_element(int dim, const double *Xc, const bool *bnd, double *Me, double *Ke)
{
   const int nv = 1 << dim;
   std::vector<double> xq, wq;
   gauss_legendre_01(3, xq, wq);
   for (int i = 0; i < nv * nv; i++) { Me[i] = 0.0; Ke[i] = 0.0; }
   auto eval = [&](const double *xi, double *N, double (*G)[3], double &det, double (*Jinv)[3])
   {
      double dN[8][3], J[3][3] = {{0}};
      for (int n = 0; n < nv; n++)
      {
         N[n] = 1.0;
         for (int a = 0; a < dim; a++) { N[n] *= ((n >> a) & 1) ? xi[a] : 1.0 - xi[a]; }
         for (int a = 0; a < dim; a++)
         {
            double d = 1.0;
            for (int b = 0; b < dim; b++)
            {
               const int bit = (n >> b) & 1;
               d *= (a == b) ? (bit ? 1.0 : -1.0) : (bit ? xi[b] : 1.0 - xi[b]);
            }
            dN[n][a] = d;
            for (int i = 0; i < dim; i++) { J[i][a] += d * Xc[n * dim + i]; }
         }
      }
      if (dim == 2)
      {
         det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
         Jinv[0][0] = J[1][1] / det; Jinv[0][1] = -J[0][1] / det;
         Jinv[1][0] = -J[1][0] / det; Jinv[1][1] = J[0][0] / det;
      }
      else
      {
         double c[3][3];
         c[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1]; c[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
         c[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1]; c[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
         c[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0]; c[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
         c[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0]; c[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
         c[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
         det = J[0][0] * c[0][0] + J[0][1] * c[1][0] + J[0][2] * c[2][0];
         for (int a = 0; a < 3; a++) for (int i = 0; i < 3; i++) { Jinv[a][i] = c[a][i] / det; }
      }
      for (int n = 0; n < nv; n++)
         for (int i = 0; i < dim; i++)
         {
            double g = 0.0;
            for (int a = 0; a < dim; a++) { g += dN[n][a] * Jinv[a][i]; }
            G[n][i] = g;
         }
   };
   int nq = 1;
   for (int a = 0; a < dim; a++) { nq *= 3; }
   for (int q = 0; q < nq; q++)
   {
      double xi[3] = {0, 0, 0}, w = 1.0, N[8], G[8][3], det, Jinv[3][3];
      int r = q;
      for (int a = 0; a < dim; a++) { xi[a] = xq[r % 3]; w *= wq[r % 3]; r /= 3; }
      eval(xi, N, G, det, Jinv);
      w *= std::fabs(det);
      for (int i = 0; i < nv; i++)
         for (int j = 0; j < nv; j++)
         {
            double gg = 0.0;
            for (int d = 0; d < dim; d++) { gg += G[i][d] * G[j][d]; }
            Me[i * nv + j] += w * N[i] * N[j];
            Ke[i * nv + j] -= w * gg;
         }
   }
   for (int f = 0; f < 2 * dim; f++)
   {
      if (!bnd[f]) { continue; }
      int axis, side;
      face_axis(dim, f, axis, side);
      int nqf = 1;
      for (int a = 0; a < dim - 1; a++) { nqf *= 3; }
      for (int q = 0; q < nqf; q++)
      {
         double xi[3] = {0, 0, 0}, w = 1.0, N[8], G[8][3], det, Jinv[3][3];
         int r = q;
         for (int a = 0; a < dim; a++)
         {
            if (a == axis) { xi[a] = side; }
            else { xi[a] = xq[r % 3]; w *= wq[r % 3]; r /= 3; }
         }
         eval(xi, N, G, det, Jinv);
         for (int j = 0; j < nv; j++)
         {
            double dn = 0.0;      // grad phi_j . (outward normal * surface element)
            for (int i = 0; i < dim; i++) { dn += G[j][i] * (side ? 1.0 : -1.0) * det * Jinv[axis][i]; }
            for (int i = 0; i < nv; i++) { Ke[i * nv + j] += w * N[i] * dn; }
         }
      }
   }
}
} // namespace

extern "C" int rmh_si_setup(rmh_ctx *c, int si_type, void *stream)
{
   if (si_type == 0) { c->si_type = 0; return 0; }
   if (si_type != 1 && si_type != 2) { set_error("Bad smoothness indicator id!"); return 1; }   // remhos_tools.cpp:36
   if (c->ne_ghost > 0) { set_error("rmh_si_setup: decomposed meshes are not supported"); return 1; }
   // Any order: the H1 space is the positive order-1 space on the SUBCELL mesh (p^dim subcells per element,
   // vertices = the lattice points i/p; for p = 1 the mesh itself), one dof per distinct lattice point.
   const int dim = c->dim, nv = 1 << dim, NF = c->NF, ND = c->ND, NFD = c->NFD, p = c->p, D1 = c->D1;
   const int64_t ne = c->ne;
   std::vector<double> X((size_t)ne * c->NGN * dim);
   std::vector<int32_t> nbe((size_t)ne * NF);
   std::vector<uint8_t> npat((size_t)ne * NF);
   CUDA_OK(cudaMemcpy(X.data(), c->X0, X.size() * sizeof(double), cudaMemcpyDeviceToHost));
   CUDA_OK(cudaMemcpy(nbe.data(), c->nbr_elem, nbe.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
   CUDA_OK(cudaMemcpy(npat.data(), c->nbr_pat, npat.size(), cudaMemcpyDeviceToHost));
   // ---- H1 dofs: classes of coincident lattice points.  DG dof (e, face dof j of face f) coincides with
   // the neighbour's dof pat[pid][j] (NbrDof, remhos_tools.cpp:525-676): union-find over the face pairs
   std::vector<int32_t> parent((size_t)ne * ND);
   for (size_t i = 0; i < parent.size(); i++) { parent[i] = (int32_t)i; }
   auto find = [&](int32_t x)
   {
      while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
      return x;
   };
   for (int64_t e = 0; e < ne; e++)
      for (int f = 0; f < NF; f++)
      {
         const int32_t nb = nbe[e * NF + f];
         if (nb < 0) { continue; }
         int axis, side;
         face_axis(dim, f, axis, side);
         for (int j = 0; j < NFD; j++)
         {
            int l[3] = {0, 0, 0}, m = j;
            for (int a = 0; a < dim; a++)
            {
               if (a == axis) { l[a] = side * p; }
               else { l[a] = m % D1; m /= D1; }
            }
            const int own = l[0] + D1 * (l[1] + D1 * l[2]);
            const int32_t x = find((int32_t)(e * ND + own));
            const int32_t y = find((int32_t)(nb * ND + c->pat_h[(size_t)npat[e * NF + f] * NFD + j]));
            if (x != y) { parent[std::max(x, y)] = std::min(x, y); }
         }
      }
   std::vector<int32_t> cgdof((size_t)ne * ND);
   int N = 0;
   {
      std::vector<int32_t> root_id((size_t)ne * ND, -1);
      for (size_t i = 0; i < cgdof.size(); i++)
      {
         const int32_t r = find((int32_t)i);
         if (root_id[r] < 0) { root_id[r] = N++; }
         cgdof[i] = root_id[r];
      }
   }
   // ---- lattice point coordinates (subcell vertices) and ShapeEval (Bernstein -> lattice values, :107-125)
   std::vector<double> lat1((size_t)D1);
   for (int i = 0; i < D1; i++) { lat1[i] = (double)i / std::max(p, 1); }
   const std::vector<double> gll = gauss_lobatto_01(c->NG1);
   const std::vector<double> Lg = lagrange(gll, lat1);            // [D1][NG1]
   const std::vector<double> Bl = bernstein(p, lat1);             // [D1 points][D1 coefficients]
   std::vector<double> xlat((size_t)ne * ND * dim, 0.0);
   for (int64_t e = 0; e < ne; e++)
      for (int j = 0; j < ND; j++)
      {
         int lj[3] = {0, 0, 0}, r = j;
         for (int a = 0; a < dim; a++) { lj[a] = r % D1; r /= D1; }
         for (int n = 0; n < c->NGN; n++)
         {
            int ln[3] = {0, 0, 0}, q = n;
            double w = 1.0;
            for (int a = 0; a < dim; a++) { ln[a] = q % c->NG1; q /= c->NG1; w *= Lg[lj[a] * c->NG1 + ln[a]]; }
            if (w == 0.0) { continue; }
            for (int i = 0; i < dim; i++) { xlat[((size_t)e * ND + j) * dim + i] += w * X[((size_t)e * c->NGN + n) * dim + i]; }
         }
      }
   std::vector<double> V((size_t)ND * ND);                        // [lattice point][coefficient]
   for (int l = 0; l < ND; l++)
      for (int k = 0; k < ND; k++)
      {
         int ll[3] = {0, 0, 0}, kk[3] = {0, 0, 0}, r = l, q = k;
         double w = 1.0;
         for (int a = 0; a < dim; a++) { ll[a] = r % D1; r /= D1; kk[a] = q % D1; q /= D1; w *= Bl[ll[a] * D1 + kk[a]]; }
         V[(size_t)l * ND + k] = w;
      }
   std::vector<int> s2i;
   sub2ind(p, dim, s2i);
   int nsub = 1;
   for (int a = 0; a < dim; a++) { nsub *= p; }
   std::vector<Trip> tm, tl, tx;
   std::vector<int32_t> d2c(cgdof);
   std::vector<int> bd;
   bdr_dofs(p, dim, bd);     // [NFD][NF]
   for (int64_t e = 0; e < ne; e++)
   {
      for (int m = 0; m < nsub; m++)
      {
         double Xc[8 * 3], Me[64], Ke[64];
         bool bnd[6];
         int sl[3] = {0, 0, 0}, r = m;
         for (int a = 0; a < dim; a++) { sl[a] = r % p; r /= p; }
         const int *sc = &s2i[(size_t)m * nv];
         for (int j = 0; j < nv; j++)
            for (int i = 0; i < dim; i++) { Xc[j * dim + i] = xlat[((size_t)e * ND + sc[j]) * dim + i]; }
         for (int f = 0; f < NF; f++)
         {
            int axis, side;
            face_axis(dim, f, axis, side);
            bnd[f] = (nbe[e * NF + f] < 0) && (sl[axis] == (side ? p - 1 : 0));
         }
         q1_element(dim, Xc, bnd, Me, Ke);
         for (int i = 0; i < nv; i++)
         {
            const int32_t ri = cgdof[e * ND + sc[i]];
            for (int j = 0; j < nv; j++)
            {
               const int32_t cj = cgdof[e * ND + sc[j]];
               tm.push_back({ri, cj, Me[i * nv + j]});
               tl.push_back({ri, cj, Ke[i * nv + j]});
               // MassMixed x ShapeEval: the lattice value sc[j] is sum_k V[sc[j]][k] u_k
               for (int k = 0; k < ND; k++)
               {
                  const double v = Me[i * nv + j] * V[(size_t)sc[j] * ND + k];
                  if (v != 0.0) { tx.push_back({ri, (int32_t)(e * ND + k), v}); }
               }
            }
         }
      }
      for (int f = 0; f < NF; f++)
      {
         if (nbe[e * NF + f] >= 0) { continue; }
         for (int j = 0; j < NFD; j++) { d2c[e * ND + bd[j * NF + f]] = -1; }     // remhos_tools.cpp:94-105
      }
   }
   std::vector<int32_t> MI, MJ, LI, LJ, XI, XJ;
   std::vector<double> MA, LA, XA;
   to_csr(N, tm, MI, MJ, MA); to_csr(N, tl, LI, LJ, LA); to_csr(N, tx, XI, XJ, XA);
   std::vector<double> ml((size_t)N, 0.0);
   for (int i = 0; i < N; i++) { for (int k = MI[i]; k < MI[i + 1]; k++) { ml[i] += MA[k]; } }   // LumpedIntegrator
   if (dev_upload(c, &c->si_MI, MI.data(), MI.size()) || dev_upload(c, &c->si_MJ, MJ.data(), MJ.size()) ||
       dev_upload(c, &c->si_MA, MA.data(), MA.size()) || dev_upload(c, &c->si_LI, LI.data(), LI.size()) ||
       dev_upload(c, &c->si_LJ, LJ.data(), LJ.size()) || dev_upload(c, &c->si_LA, LA.data(), LA.size()) ||
       dev_upload(c, &c->si_XI, XI.data(), XI.size()) || dev_upload(c, &c->si_XJ, XJ.data(), XJ.size()) ||
       dev_upload(c, &c->si_XA, XA.data(), XA.size()) || dev_upload(c, &c->si_ml, ml.data(), ml.size()) ||
       dev_upload(c, &c->si_d2c, d2c.data(), d2c.size())) { return 1; }
   if (dev_alloc(c, &c->si_y, (size_t)N) || dev_alloc(c, &c->si_z, (size_t)N) || dev_alloc(c, &c->si_r, (size_t)N) ||
       dev_alloc(c, &c->si_val, (size_t)N) || dev_alloc(c, &c->si_tmp, (size_t)c->N) ||
       dev_alloc(c, &c->si_nrm, 1)) { return 1; }
   c->si_n = N; c->si_type = si_type; c->si_param = (si_type == 1) ? 5.0 : 3.0;
   (void)stream;
   return 0;
}

// ComputeSmoothnessIndicator (remhos_tools.cpp:153-184) -> per DG dof: the indicator at the dof's
// vertex, 1 on the domain boundary (the `tmp` of remhos_mono.cpp:134-135).  out may be null
// (the values stay in the context for rmh_mono_rd).
extern "C" int rmh_si_values(rmh_ctx *c, const double *u, double *out, void *stream)
{
   if (!c->si_type) { set_error("rmh_si_values: call rmh_si_setup first"); return 1; }
   cudaStream_t s = (cudaStream_t)stream;
   const int N = c->si_n, bs = 128, nb = (N + bs - 1) / bs;
   auto solve2 = [&](const double *rhs, double *y)      // ApproximateLaplacian: two sweeps from y = 0
   {
      cudaMemsetAsync(y, 0, (size_t)N * sizeof(double), s);
      for (int it = 0; it < 2; it++)
      {
         k_csr_spmv<<<nb, bs, 0, s>>>(N, c->si_MI, c->si_MJ, c->si_MA, y, rhs, c->si_z);
         k_sum_sq<<<1, 1024, 0, s>>>(N, c->si_z, c->si_nrm);
         k_si_sweep<<<nb, bs, 0, s>>>(N, c->si_z, c->si_ml, c->si_nrm, y);
      }
   };
   // u at the lattice points = u (ShapeEval is the identity for order 1), rhs = MassMixed u
   k_csr_spmv<<<nb, bs, 0, s>>>(N, c->si_XI, c->si_XJ, c->si_XA, u, nullptr, c->si_r);
   solve2(c->si_r, c->si_y);
   k_csr_spmv<<<nb, bs, 0, s>>>(N, c->si_LI, c->si_LJ, c->si_LA, c->si_y, nullptr, c->si_r);
   solve2(c->si_r, c->si_val);                          // g (reuses si_val as scratch ...)
   CUDA_OK(cudaMemcpyAsync(c->si_y, c->si_val, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s));
   k_si_value<<<nb, bs, 0, s>>>(N, c->si_MI, c->si_MJ, c->si_y, c->si_type, c->si_param, c->si_val);
   k_si_gather<<<(unsigned)((c->N + bs - 1) / bs), bs, 0, s>>>(c->N, c->si_d2c, c->si_val, c->si_tmp);
   LAUNCH_OK();
   if (out) { CUDA_OK(cudaMemcpyAsync(out, c->si_tmp, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToDevice, s)); }
   return 0;
}

// ElementFCTProjection::CalcFCTSolution (remhos_fct.cpp:613-733); dense element mass blocks
extern "C" int rmh_fct_project(rmh_ctx *c, double dt, const double *u, const double *du_ho,
                               const double *du_lo, const double *xi_min, const double *xi_max,
                               double *du, void *stream)
{
   if (!c->fa_on) { set_error("rmh_fct_project: call rmh_fa_setup first (assembled element mass)"); return 1; }
   const int wpb = 2;
   const size_t shb = (size_t)wpb * 6 * c->ND * sizeof(double);
   k_fct_project<<<(unsigned)((c->ne + wpb - 1) / wpb), wpb * 32, shb, (cudaStream_t)stream>>>(
      fa_args(c), dt, u, du_ho, du_lo, xi_min, xi_max, du);
   LAUNCH_OK();
   return 0;
}

// automatic time step control (remhos.cpp:312-316, 1153-1154, 1178-1197).  mode 1: every
// rmh_limit_mult lowers the ratio dt_estimate / dt (UpdateTimeStepEstimate on the LO rate).
// rmh_dt_ratio returns the minimum since the last reset (GetTimeStepRatio) and, with reset != 0,
// starts over (ResetTimeStepRatio); it synchronises the device.
extern "C" int rmh_dt_control(rmh_ctx *c, int mode)
{
   if (mode != 0 && mode != 1) { set_error("rmh_dt_control: mode must be 0 (fixed) or 1 (LO bounds error)"); return 1; }
   if (mode && !c->dt_ratio)
   {
      if (dev_alloc(c, &c->dt_ratio, 1)) { return 1; }
      const double inf = INFINITY;
      CUDA_OK(cudaMemcpy(c->dt_ratio, &inf, sizeof(double), cudaMemcpyHostToDevice));
   }
   c->dt_control = mode;
   return 0;
}
extern "C" int rmh_dt_ratio(rmh_ctx *c, int reset, double *ratio)
{
   if (!c->dt_ratio) { set_error("rmh_dt_ratio: call rmh_dt_control first"); return 1; }
   CUDA_OK(cudaDeviceSynchronize());
   if (ratio) { CUDA_OK(cudaMemcpy(ratio, c->dt_ratio, sizeof(double), cudaMemcpyDeviceToHost)); }
   if (reset)
   {
      const double inf = INFINITY;
      CUDA_OK(cudaMemcpy(c->dt_ratio, &inf, sizeof(double), cudaMemcpyHostToDevice));
   }
   return 0;
}

// MonolithicSolver set-up (remhos.cpp:997-1011): the operator owns at most one monolithic solver,
// which then takes precedence over HO/LO/FCT in rmh_mult / rmh_mult_unlimited (remhos.cpp:1687).
extern "C" int rmh_mono_setup(rmh_ctx *c, int mono_type, int mass_lim, const double *scale_host,
                              void *stream)
{
   if (mono_type < 0 || mono_type > 2) { set_error("rmh_mono_setup: mono type must be 0, 1 or 2"); return 1; }
   if (mono_type == 0) { c->mono_type = 0; return 0; }
   if (!scale_host) { set_error("rmh_mono_setup: scale (one value per element) is required"); return 1; }
   if (mono_type == 2 && !c->sub_on)
   { set_error("rmh_mono_setup: the subcell variant needs rmh_subcell_setup first"); return 1; }
   if (c->ne_ghost > 0)
   { set_error("rmh_mono_setup: decomposed meshes are not supported (serial only in the reference too, remhos_mono.cpp:283)"); return 1; }
   if (c->ND > 256) { set_error("rmh_mono_setup: at most 256 DOFs per element"); return 1; }
   if (rmh_fa_setup(c, stream)) { return 1; }
   if (!c->mono_scale) { if (dev_alloc(c, &c->mono_scale, (size_t)c->ne)) { return 1; } }
   CUDA_OK(cudaMemcpy(c->mono_scale, scale_host, (size_t)c->ne * sizeof(double), cudaMemcpyHostToDevice));
   c->mono_type = mono_type; c->mono_mass_lim = mass_lim ? 1 : 0;
   return 0;
}

// MonoRDSolver::CalcSolution (remhos_mono.cpp:60-356)
extern "C" int rmh_mono_rd(rmh_ctx *c, const double *u, double *du, void *stream)
{
   if (!c->mono_type) { set_error("rmh_mono_rd: call rmh_mono_setup first"); return 1; }
   if (du == u) { set_error("rmh_mono_rd: output must not alias the input"); return 1; }
   cudaStream_t s = (cudaStream_t)stream;
   if (work_vec(c, &c->wk[0]) || work_vec(c, &c->wk[5]) || work_vec(c, &c->wk[6])) { return 1; }
   double *z = c->wk[0], *xmn = c->wk[5], *xmx = c->wk[6];
   if (dispatch_ho(c, ho_args(c, u, z, 1 | 4), s)) { return 1; }
   c->xe_ptr = nullptr;
   if (rmh_elem_min_max(c, u, c->xe_min, c->xe_max, stream)) { return 1; }
   if (rmh_bounds(c, c->xe_min, c->xe_max, xmn, xmx, stream)) { return 1; }
   int ns = 1;
   for (int a = 0; a < c->dim; a++) { ns *= c->p; }
   const int wpb = 2;
   const size_t shb = (size_t)wpb * (8 * c->ND + 3 * c->NFD + ns * 6) * sizeof(double);
   static bool attr_dev[RMH_MAX_DEVICES] = {false};
   bool &attr = attr_dev[c->device % RMH_MAX_DEVICES];
   if (!attr)
   {
      CUDA_OK(cudaFuncSetAttribute(k_mono_rd, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr = true;
   }
   if (shb > 96 * 1024) { set_error("rmh_mono_rd: element too large for the shared-memory scratch"); return 1; }
   if (c->si_type) { if (rmh_si_values(c, u, nullptr, stream)) { return 1; } }
   k_mono_rd<<<(unsigned)((c->ne + wpb - 1) / wpb), wpb * 32, shb, s>>>(
      fa_args(c), c->p, c->mono_type == 2 ? 1 : 0, c->mono_mass_lim, c->sub_w, c->mono_scale, u, z, xmn,
      xmx, c->si_type ? c->si_tmp : nullptr, du);
   LAUNCH_OK();
   return 0;
}

extern "C" int rmh_fct_flux_based(rmh_ctx *c, double dt, const double *u, const double *m,
                                  const double *du_ho, const double *du_lo, const double *xi_min,
                                  const double *xi_max, double *du, void *stream)
{
   if (!c->fa_on) { set_error("rmh_fct_flux_based: call rmh_fa_setup first (assembled K_HO, M)"); return 1; }
   if (work_vec(c, &c->wk[1]) || work_vec(c, &c->wk[2])) { return 1; }
   FaArgs A = fa_args(c);
   A.ml = m;
   cudaStream_t s = (cudaStream_t)stream;
   const bool ghosts = (c->ne_ghost > 0 && c->n_gslots > 0);
   if (ghosts)
   {
      // the window holds the halo of u (rmh_limit_mult / the caller's rmh_dist_halo): keep a private copy,
      // the two coefficient exchanges below reuse the window
      if (!c->dist) { set_error("rmh_fct_flux_based: ghost elements without a connected rmh_dist"); return 1; }
      if (c->halo_ptr != u) { if (dist_halo(c, u, s)) { return 1; } }
      CUDA_OK(cudaMemcpyAsync(c->gtr_u, c->ughost, (size_t)c->n_gslots * c->NFD * sizeof(double), cudaMemcpyDeviceToDevice, s));
      A.fn.ughost = c->gtr_u;
      A.BIg = c->faBIg; A.gcp = c->gtr_cp; A.gcn = c->gtr_cn;
   }
   const int bs = std::min(256, ((c->ND + 31) / 32) * 32);
   const size_t shb = 2 * c->ND * sizeof(double);
   k_flux_coeff<<<(unsigned)c->ne, bs, shb, s>>>(A, dt, u, du_ho, du_lo, xi_min, xi_max, c->wk[1], c->wk[2]);
   LAUNCH_OK();
   if (ghosts)
   {
      // coeff_pos / coeff_neg of the face neighbours (ExchangeFaceNbrData, remhos_fct.cpp:406-409)
      if (dist_traces(c, c->wk[1], c->gtr_cp, s) || dist_traces(c, c->wk[2], c->gtr_cn, s)) { return 1; }
   }
   k_flux_apply<<<(unsigned)c->ne, bs, shb, s>>>(A, dt, u, du_ho, du_lo, c->wk[1], c->wk[2], du);
   LAUNCH_OK();
   return 0;
}

static int check_combo(int ho_type, int lo_type, int fct_type)
{
   if (ho_type != 0 && ho_type != 1 && ho_type != 3)
   { set_error("stage operator: HO solver must be 0, 1 (Neumann) or 3 (LocalInverse)"); return 1; }
   if (lo_type < 0 || lo_type > 5)
   { set_error("stage operator: LO solver must be 0 .. 5"); return 1; }
   if (fct_type < 0 || fct_type > 4)
   { set_error("stage operator: FCT solver must be 0, 1 (FluxBased), 2 (ClipScale), 3 (NonlinearPenalty) or 4 (FCTProject)"); return 1; }
   if (fct_type && (ho_type == 0 || lo_type == 0))
   { set_error("FCT requires HO and LO solvers."); return 1; }        // remhos.cpp:1690
   if (!fct_type && lo_type == 5 && ho_type == 0)
   { set_error("Mass-Based LO solver requires a choice of a HO solver."); return 1; }   // remhos.cpp:991
   if (!ho_type && !lo_type) { set_error("No solver was chosen."); return 1; }          // remhos.cpp:1711
   return 0;
}

// AdvectionOperator::MultUnlimited (remhos.cpp:1596-1739): remap re-assembly at time t, then the
// HO rate when an FCT solver will limit it later, else the LO or HO rate.
static int mult_unlimited_1(rmh_ctx *c, int ho_type, int lo_type, int fct_type, double t,
                            double dt, const double *u, double *k, void *stream)
{
   if (c->mono_type)      // the monolithic solver takes precedence (remhos.cpp:1687)
   {
      if (rmh_set_time(c, t, stream)) { return 1; }
      return rmh_mono_rd(c, u, k, stream);
   }
   if (check_combo(ho_type, lo_type, fct_type)) { return 1; }
   if (k == u) { set_error("rmh_mult_unlimited: output must not alias the input"); return 1; }
   if (rmh_set_time(c, t, stream)) { return 1; }
   if (dist_halo(c, u, (cudaStream_t)stream)) { return 1; }      // x_gf.ExchangeFaceNbrData() (remhos.cpp:1813)
   auto HO = [&](double *out)
   { return ho_type == 1 ? rmh_ho_neumann(c, u, out, stream) : rmh_ho_local_inverse(c, u, out, stream); };
   if (fct_type) { return HO(k); }
   if (lo_type == 1) { return rmh_lo_discrete_upwind(c, u, k, stream); }
   if (lo_type == 2) { return rmh_lo_discrete_upwind_prec(c, u, k, stream); }
   if (lo_type == 3) { return rmh_lo_res_dist(c, u, k, stream); }
   if (lo_type == 4) { return rmh_lo_res_dist_subcell(c, u, k, stream); }
   if (lo_type == 5)
   {
      if (work_vec(c, &c->wk[3])) { return 1; }
      if (HO(c->wk[3])) { return 1; }
      return rmh_lo_mass_avg(c, dt, u, c->wk[3], k, stream);
   }
   return HO(k);
}

// AdvectionOperator::LimitMult (remhos.cpp:1798-1916): k holds the (possibly combined) HO rate
// on entry and the limited rate on exit; a no-op without an FCT solver.
// Mesh::GetElementSize(0, 0) at the current mesh position: |det J(centre of element 0)|^(1/dim)
static double elem0_size(const rmh_ctx *c)
{
   const int dim = c->dim, n1 = c->NG1;
   const std::vector<double> gll = gauss_lobatto_01(n1), half = {0.5};
   const std::vector<double> L = lagrange(gll, half), dL = lagrange_deriv(gll, half);
   double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
   for (int node = 0; node < c->NGN; node++)
   {
      int id[3] = {0, 0, 0}, r = node;
      for (int a = 0; a < dim; a++) { id[a] = r % n1; r /= n1; }
      for (int a = 0; a < dim; a++)          // derivative direction
      {
         double g = 1.0;
         for (int b = 0; b < dim; b++) { g *= (a == b) ? dL[id[b]] : L[id[b]]; }
         for (int i = 0; i < dim; i++)
         {
            double x = c->x0_e0[(size_t)node * dim + i];
            if (!c->v_e0.empty()) { x += c->t_cur * c->v_e0[(size_t)node * dim + i]; }
            J[i][a] += g * x;
         }
      }
   }
   double det;
   if (dim == 2) { det = J[0][0] * J[1][1] - J[0][1] * J[1][0]; }
   else
   {
      det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
            J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
   }
   return std::pow(std::fabs(det), 1.0 / dim);
}

static int limit_mult_1(rmh_ctx *c, int lo_type, int fct_type, double dt, const double *u,
                        double *k, void *stream)
{
   if (!fct_type || c->mono_type) { return 0; }            // remhos.cpp:1803: no FCT solver, nothing to limit
   if (check_combo(3, lo_type, fct_type)) { return 1; }
   if (k == u) { set_error("rmh_limit_mult: the rate must not alias the state"); return 1; }
   if (c->halo_ptr != u) { if (dist_halo(c, u, (cudaStream_t)stream)) { return 1; } }
   for (int i = 3; i < 7; i++) { if (work_vec(c, &c->wk[i])) { return 1; } }
   double *du_ho = c->wk[3], *du_lo = c->wk[4], *xmn = c->wk[5], *xmx = c->wk[6];
   CUDA_OK(cudaMemcpyAsync(du_ho, k, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
   if (lo_type == 5) { if (rmh_lo_mass_avg(c, dt, u, du_ho, du_lo, stream)) { return 1; } }
   else if (lo_type == 1) { if (rmh_lo_discrete_upwind(c, u, du_lo, stream)) { return 1; } }
   else if (lo_type == 2) { if (rmh_lo_discrete_upwind_prec(c, u, du_lo, stream)) { return 1; } }
   else if (lo_type == 4) { if (rmh_lo_res_dist_subcell(c, u, du_lo, stream)) { return 1; } }
   else { if (rmh_lo_res_dist(c, u, du_lo, stream)) { return 1; } }
   c->xe_ptr = nullptr;       // the context's element min/max change owner (rmh_ctx_trust_state)
   if (rmh_elem_min_max(c, u, c->xe_min, c->xe_max, stream)) { return 1; }
   if (rmh_bounds(c, c->xe_min, c->xe_max, xmn, xmx, stream)) { return 1; }
   int rc;
   if (c->si_type && (fct_type == 2 || fct_type == 3))
   {
      // bounds relaxed where the solution is smooth (SmoothnessIndicator::UpdateBounds: NonlinearPenaltySolver
      // remhos_fct.cpp:780-795; ClipScaleSolver :498-504, which the reference's device kernel aborts on)
      if (work_vec(c, &c->wk[7])) { return 1; }
      if (rmh_si_values(c, u, c->wk[7], stream)) { return 1; }
      if (rmh_si_update_bounds(c, dt, u, du_ho, c->wk[7], xmn, xmx, stream)) { return 1; }
   }
   if (fct_type == 2) { rc = rmh_fct_clip_scale(c, dt, u, c->ml, du_ho, du_lo, xmn, xmx, k, stream); }
   else if (fct_type == 3)
   { rc = rmh_fct_nonlinear_penalty(c, dt, elem0_size(c) / c->p, u, c->ml, du_ho, du_lo, xmn, xmx, k, stream); }
   else if (fct_type == 4) { rc = rmh_fct_project(c, dt, u, du_ho, du_lo, xmn, xmx, k, stream); }
   else { rc = rmh_fct_flux_based(c, dt, u, c->ml, du_ho, du_lo, xmn, xmx, k, stream); }
   if (rc) { return rc; }
   if (c->dt_control)                                         // remhos.cpp:1839-1842
   {
      const int bs = 256;
      k_dt_estimate<<<(unsigned)((c->N + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(c->N, dt, u, du_lo, xmn,
                                                                                       xmx, c->dt_ratio);
      LAUNCH_OK();
   }
   return 0;
}

// SmoothnessIndicator::UpdateBounds (remhos_tools.cpp:183-190) on all dofs; si_dev from rmh_si_values
extern "C" int rmh_si_update_bounds(rmh_ctx *c, double dt, const double *u, const double *du_ho, const double *si,
                                    double *xi_min, double *xi_max, void *stream)
{
   const int bs = 256;
   k_si_update_bounds<<<(unsigned)((c->N + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(c->N, dt, u, du_ho, si,
                                                                                        xi_min, xi_max);
   LAUNCH_OK();
   return 0;
}

// NonlinearPenaltySolver::CalcFCTSolution (remhos_fct.cpp:760-996); eps_w = GetElementSize(0, 0) /
// GetOrder(0) (:961).  The bound relaxation by a smoothness indicator (:780-795) is rmh_si_update_bounds.
extern "C" int rmh_fct_nonlinear_penalty(rmh_ctx *c, double dt, double eps_w, const double *u, const double *m,
                                         const double *du_ho, const double *du_lo, const double *xi_min,
                                         const double *xi_max, double *du, void *stream)
{
   if (work_vec(c, &c->wk[0]) || work_vec(c, &c->wk[1]) || work_vec(c, &c->wk[2])) { return 1; }
   cudaStream_t s = (cudaStream_t)stream;
   const int bs = 256;
   k_penalty_flux<<<(unsigned)((c->N + bs - 1) / bs), bs, 0, s>>>(c->N, dt, u, m, du_ho, du_lo, xi_min, xi_max,
                                                                 c->wk[0], c->wk[1]);
   LAUNCH_OK();
   k_penalty_correct<<<(unsigned)((c->ne + 63) / 64), 64, 0, s>>>(c->ne, c->ND, eps_w, m, du_lo, c->wk[0], c->wk[1],
                                                                 c->wk[2], du);
   LAUNCH_OK();
   return 0;
}

// ---------------------------------------------------------------- product-field remap (-ps)
static int prod_buffers(rmh_ctx *c)
{
   for (int i = 0; i < 10; i++) { if (!c->pw[i]) { if (dev_alloc(c, &c->pw[i], (size_t)c->N)) { return 1; } } }
   for (int i = 0; i < 2; i++)
   {
      if (!c->pf_el[i]) { if (dev_alloc(c, &c->pf_el[i], (size_t)c->ne)) { return 1; } }
      if (!c->pf_dof[i]) { if (dev_alloc(c, &c->pf_dof[i], (size_t)c->N)) { return 1; } }
   }
   if (!c->pmask) { if (dev_alloc(c, &c->pmask, (size_t)2 * c->N)) { return 1; } }
   return 0;
}

static unsigned warp_grid(int64_t ne, int bs) { return (unsigned)((ne * 32 + bs - 1) / bs); }

// on != 0: the state vectors of rmh_mult_unlimited / rmh_limit_mult / rmh_mult / rmh_ode_step are the
// block (u, us) of 2 N doubles (BlockVector S, remhos.cpp:594-598,886-903); remap mode only
extern "C" int rmh_product_enable(rmh_ctx *c, int on)
{
   if (on && c->exec_mode != 1) { set_error("Products are processed only in remap mode."); return 1; }   // remhos.cpp:1850
   if ((on != 0) != c->product)
   {
      // work vectors are sized by the state: start over
      for (int i = 0; i < 8; i++) { c->wk[i] = nullptr; }
      for (int i = 0; i < 9; i++) { c->rk[i] = nullptr; }
   }
   c->product = (on != 0);
   c->xe_ptr = nullptr;
   return on ? prod_buffers(c) : 0;
}

// use_masks of RKIDPSolver (remhos_solvers.hpp; the driver switches them off, remhos.cpp:502-507)
extern "C" int rmh_idp_use_mask(rmh_ctx *c, int on)
{
   c->idp_mask = (on != 0);
   return 0;
}

extern "C" int rmh_prod_bool_indicators(rmh_ctx *c, const double *u, uint8_t *el, uint8_t *dof, void *stream)
{
   const int bs = 128;
   k_bool_indicators<<<warp_grid(c->ne, bs), bs, 0, (cudaStream_t)stream>>>(c->ne, c->ND, u, el, dof);
   LAUNCH_OK();
   return 0;
}

extern "C" int rmh_prod_compute_ratio(rmh_ctx *c, const double *us, const double *u, double *s, uint8_t *el,
                                      uint8_t *dof, void *stream)
{
   const int bs = 128;
   k_prod_ratio<<<warp_grid(c->ne, bs), bs, 0, (cudaStream_t)stream>>>(c->ne, c->ND, us, u, s, el, dof, nullptr,
                                                                      nullptr);
   LAUNCH_OK();
   return 0;
}

extern "C" int rmh_elem_min_max_masked(rmh_ctx *c, const double *u, const uint8_t *el, const uint8_t *dof,
                                       double *xe_min, double *xe_max, void *stream)
{
   if (xe_min == c->xe_min || xe_max == c->xe_max) { c->xe_ptr = nullptr; }
   const int bs = 128;
   k_elem_min_max_masked<<<warp_grid(c->ne, bs), bs, 0, (cudaStream_t)stream>>>(c->ne, c->ND, u, el, dof, xe_min,
                                                                               xe_max);
   LAUNCH_OK();
   return 0;
}

extern "C" int rmh_prod_compatible_lo(rmh_ctx *c, double dt, const double *us, const double *m,
                                      const double *d_us_ho, double *s_min, double *s_max, const double *u_new,
                                      const uint8_t *el, const uint8_t *dof, double *d_us_lo_new, void *stream)
{
   const int bs = 128;
   k_prod_compatible_lo<<<warp_grid(c->ne, bs), bs, 0, (cudaStream_t)stream>>>(
      c->ne, c->ND, dt, us, m, d_us_ho, s_min, s_max, u_new, el, dof, d_us_lo_new, nullptr, nullptr);
   LAUNCH_OK();
   return 0;
}

extern "C" int rmh_prod_zero_empty(rmh_ctx *c, const uint8_t *el, const uint8_t *dof, double *d_us, void *stream)
{
   const int bs = 256;
   k_zero_empty<<<(unsigned)((c->N + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(c->N, c->ND, el, dof, d_us);
   LAUNCH_OK();
   return 0;
}

// FCTSolver::CalcFCTProduct of FluxBasedFCT (fct_type 1, remhos_fct.cpp:183-294), ClipScaleSolver (2,
// :543-563) and ElementFCTProjection (4, :735-758): compatible LO product, scaled bounds, the solver's
// own limiter on us, empty dofs zeroed.  s_min / s_max are adjusted in place as in the reference;
// d_us_lo is read by the flux-based solver only (NeedsLOProductInput).
extern "C" int rmh_fct_product(rmh_ctx *c, int fct_type, double dt, const double *us, const double *m,
                               const double *d_us_ho, const double *d_us_lo, double *s_min, double *s_max,
                               const double *u_new, const uint8_t *el, const uint8_t *dof, double *d_us,
                               void *stream)
{
   if (fct_type != 1 && fct_type != 2 && fct_type != 4) { set_error("rmh_fct_product: FCT solver must be 1, 2 or 4"); return 1; }
   if (prod_buffers(c)) { return 1; }
   cudaStream_t s = (cudaStream_t)stream;
   double *d_lo_c = c->pw[5], *us_min = c->pw[6], *us_max = c->pw[7];
   const int bs = 128;
   k_prod_compatible_lo<<<warp_grid(c->ne, bs), bs, 0, s>>>(c->ne, c->ND, dt, us, m, d_us_ho, s_min, s_max, u_new,
                                                            el, dof, d_lo_c, us_min, us_max);
   LAUNCH_OK();
   if (fct_type == 2) { if (rmh_fct_clip_scale(c, dt, us, m, d_us_ho, d_lo_c, us_min, us_max, d_us, stream)) { return 1; } }
   else if (fct_type == 4) { if (rmh_fct_project(c, dt, us, d_us_ho, d_lo_c, us_min, us_max, d_us, stream)) { return 1; } }
   else
   {
      if (!c->fa_on) { set_error("rmh_fct_product: call rmh_fa_setup first (assembled K_HO, M)"); return 1; }
      if (!d_us_lo) { set_error("rmh_fct_product: the flux-based solver needs the LO product rate"); return 1; }
      double *fel = c->pw[8], *beta = c->pw[9];
      k_prod_flux_el<<<warp_grid(c->ne, bs), bs, 0, s>>>(c->ne, c->ND, dt, m, d_us_lo, d_lo_c, u_new, el, fel, beta);
      LAUNCH_OK();
      if (work_vec(c, &c->wk[1]) || work_vec(c, &c->wk[2])) { return 1; }
      FaArgs A = fa_args(c);
      A.ml = m; A.pbeta = beta; A.pfel = fel;
      const int fb = std::min(256, ((c->ND + 31) / 32) * 32);
      const size_t shb = 2 * c->ND * sizeof(double);
      k_flux_coeff<<<(unsigned)c->ne, fb, shb, s>>>(A, dt, us, d_us_ho, d_lo_c, us_min, us_max, c->wk[1], c->wk[2]);
      LAUNCH_OK();
      k_flux_apply<<<(unsigned)c->ne, fb, shb, s>>>(A, dt, us, d_us_ho, d_lo_c, c->wk[1], c->wk[2], d_us);
      LAUNCH_OK();
   }
   return rmh_prod_zero_empty(c, el, dof, d_us, stream);
}

static int calc_lo_1(rmh_ctx *c, int lo_type, double dt, const double *u, const double *du_ho, double *du_lo,
                     void *stream)
{
   if (lo_type == 5) { return rmh_lo_mass_avg(c, dt, u, du_ho, du_lo, stream); }
   if (lo_type == 1) { return rmh_lo_discrete_upwind(c, u, du_lo, stream); }
   if (lo_type == 2) { return rmh_lo_discrete_upwind_prec(c, u, du_lo, stream); }
   if (lo_type == 4) { return rmh_lo_res_dist_subcell(c, u, du_lo, stream); }
   return rmh_lo_res_dist(c, u, du_lo, stream);
}

// second pass of AdvectionOperator::LimitMult (remhos.cpp:1848-1915): d_us holds the HO product rate
// on entry and the limited one on exit; d_u is the limited rate of u
static int limit_product(rmh_ctx *c, int lo_type, int fct_type, double dt, const double *u, const double *d_u,
                         const double *us, double *d_us, void *stream)
{
   if (!fct_type) { return 0; }
   if (c->dt_control) { set_error("Automatic time step is not implemented for product remap."); return 1; }   // :1851
   if (prod_buffers(c)) { return 1; }
   cudaStream_t s = (cudaStream_t)stream;
   double *d_us_ho = c->pw[0], *d_us_lo = c->pw[1], *smin = c->pw[2], *smax = c->pw[3], *u_new = c->pw[4];
   CUDA_OK(cudaMemcpyAsync(d_us_ho, d_us, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToDevice, s));
   if (fct_type == 1) { if (calc_lo_1(c, lo_type, dt, us, d_us_ho, d_us_lo, stream)) { return 1; } }   // NeedsLOProductInput
   // s = us_old / u_old on the old active dofs, and its bounds from those dofs only
   const int bs = 128;
   c->xe_ptr = nullptr;
   k_prod_ratio<<<warp_grid(c->ne, bs), bs, 0, s>>>(c->ne, c->ND, us, u, nullptr, c->pf_el[0], c->pf_dof[0], c->xe_min,
                                                    c->xe_max);
   LAUNCH_OK();
   if (rmh_bounds(c, c->xe_min, c->xe_max, smin, smax, stream)) { return 1; }
   // evolve u, new active dofs
   k_axpy_out<<<(unsigned)((c->N + 255) / 256), 256, 0, s>>>(c->N, dt, u, d_u, u_new);
   LAUNCH_OK();
   if (rmh_prod_bool_indicators(c, u_new, c->pf_el[1], c->pf_dof[1], stream)) { return 1; }
   return rmh_fct_product(c, fct_type, dt, us, c->ml, d_us_ho, d_us_lo, smin, smax, u_new, c->pf_el[1], c->pf_dof[1],
                          d_us, stream);
}

// AdvectionOperator::ComputeMask (remhos.cpp:1741-1796) for the product state (u, us) [2 N bytes]
extern "C" int rmh_compute_mask(rmh_ctx *c, const double *state, uint8_t *mask, void *stream)
{
   const int bs = 128;
   if (!c->product)
   {
      CUDA_OK(cudaMemsetAsync(mask, 1, (size_t)c->N, (cudaStream_t)stream));     // only product fields are masked
      return 0;
   }
   k_compute_mask<<<warp_grid(c->ne, bs), bs, 0, (cudaStream_t)stream>>>(c->ne, c->ND, state, mask, 2);
   LAUNCH_OK();
   return 0;
}

// AdvectionOperator::MultUnlimited (remhos.cpp:1596-1739): see rmh_mult_unlimited in the header.  With a
// product field the second block is remapped with the same HO operator (:1714-1738).
extern "C" int rmh_mult_unlimited(rmh_ctx *c, int ho_type, int lo_type, int fct_type, double t,
                                  double dt, const double *u, double *k, void *stream)
{
   if (mult_unlimited_1(c, ho_type, lo_type, fct_type, t, dt, u, k, stream)) { return 1; }
   if (!c->product) { return 0; }
   if (!fct_type || c->mono_type) { set_error("product remap needs an FCT solver (remhos.cpp:1858)"); return 1; }
   const double *us = u + c->N;
   double *kus = k + c->N;
   return ho_type == 1 ? rmh_ho_neumann(c, us, kus, stream) : rmh_ho_local_inverse(c, us, kus, stream);
}

// AdvectionOperator::LimitMult (remhos.cpp:1798-1916): see rmh_limit_mult in the header
extern "C" int rmh_limit_mult(rmh_ctx *c, int lo_type, int fct_type, double dt, const double *u,
                              double *k, void *stream)
{
   if (limit_mult_1(c, lo_type, fct_type, dt, u, k, stream)) { return 1; }
   if (!c->product) { return 0; }
   return limit_product(c, lo_type, fct_type, dt, u, k, u + c->N, k + c->N, stream);
}

// LimitedTimeDependentOperator::Mult = MultUnlimited + LimitMult (remhos_solvers.hpp:46-50);
// the combination -ho 3 -lo 5 -fct 2 goes to the fused stage kernel.
extern "C" int rmh_mult(rmh_ctx *c, int ho_type, int lo_type, int fct_type, double t, double dt,
                        const double *u, double *k, void *stream)
{
   if (c->mono_type) { return rmh_mult_unlimited(c, ho_type, lo_type, fct_type, t, dt, u, k, stream); }
   if (ho_type == 3 && lo_type == 5 && fct_type == 2 && !c->dt_control && !c->product && !c->si_type && !c->dist)
   {
      if (k == u) { set_error("rmh_mult: output must not alias the input"); return 1; }
      if (rmh_set_time(c, t, stream)) { return 1; }
      return stage_impl(c, 5, dt, 0, 0.0, 0.0, u, u, k, false, false, (cudaStream_t)stream);
   }
   if (rmh_mult_unlimited(c, ho_type, lo_type, fct_type, t, dt, u, k, stream)) { return 1; }
   return rmh_limit_mult(c, lo_type, fct_type, dt, u, k, stream);
}

// over the state length (N, or 2 N for a product state) unless len is given
static int lincomb(rmh_ctx *c, int n, const double *coef, const double *const *x, double *out,
                   cudaStream_t s, int64_t len = -1)
{
   LinComb L;
   L.n = 0;
   for (int i = 0; i < n; i++)
   {
      if (coef[i] == 0.0) { continue; }
      L.c[L.n] = coef[i]; L.x[L.n] = x[i]; L.n++;
   }
   const int bs = 256;
   const int64_t n_state = (len >= 0) ? len : state_len(c);
   k_lincomb<<<(unsigned)((n_state + bs - 1) / bs), bs, 0, s>>>(n_state, L, out);
   LAUNCH_OK();
   return 0;
}

extern "C" int rmh_lincomb(rmh_ctx *c, int n, const double *coef, const double *const *x_dev,
                           double *out_dev, void *stream)
{
   if (n < 1 || n > 9) { set_error("rmh_lincomb: 1..9 terms"); return 1; }
   return lincomb(c, n, coef, x_dev, out_dev, (cudaStream_t)stream, c->N);
}

// ODESolver::Step for -s 1/2/3/4/6 (remhos.cpp:488-492) over rmh_mult: explicit RK in Butcher
// form.  Stage times t + c_i dt are pushed through rmh_set_time (remap); dt is the full step in
// every stage, as the reference's operator keeps it (remhos.cpp:176-182).
extern "C" int rmh_ode_step(rmh_ctx *c, int ode, int ho_type, int lo_type, int fct_type, double *t,
                            double dt, double *u, void *stream)
{
   cudaStream_t s = (cudaStream_t)stream;
   // (with automatic time step control the LO rate must be visible to the dt estimate: unfused path)
   if (!c->mono_type && !c->dt_control && !c->product && !c->si_type && !c->dist && ho_type == 3 && lo_type == 5 &&
       fct_type == 2 && ode >= 1 && ode <= 3)
   { return rmh_rk_step(c, ode, lo_type, t, dt, u, stream); }
   const double t0 = *t;
   const size_t bytes = (size_t)state_len(c) * sizeof(double);
   auto F = [&](const double *x, double tt, double *k) { return rmh_mult(c, ho_type, lo_type, fct_type, tt, dt, x, k, stream); };
   for (int i = 0; i < 2; i++) { if (work_vec(c, &c->rk[i])) { return 1; } }
   double *k0 = c->rk[0], *y = c->rk[1];
   if (ode == 1)
   {
      if (F(u, t0, k0)) { return 1; }
      const double cf[2] = {1.0, dt}; const double *xs[2] = {u, k0};
      if (lincomb(c, 2, cf, xs, u, s)) { return 1; }
   }
   else if (ode == 2)       // RK2Solver(1.0): Heun
   {
      if (F(u, t0, k0)) { return 1; }
      { const double cf[2] = {1.0, dt}; const double *xs[2] = {u, k0}; if (lincomb(c, 2, cf, xs, y, s)) { return 1; } }
      { const double cf[2] = {1.0, 0.5 * dt}; const double *xs[2] = {u, k0}; if (lincomb(c, 2, cf, xs, u, s)) { return 1; } }
      if (F(y, t0 + dt, k0)) { return 1; }
      { const double cf[2] = {1.0, 0.5 * dt}; const double *xs[2] = {u, k0}; if (lincomb(c, 2, cf, xs, u, s)) { return 1; } }
   }
   else if (ode == 3)       // RK3SSPSolver
   {
      if (F(u, t0, k0)) { return 1; }
      { const double cf[2] = {1.0, dt}; const double *xs[2] = {u, k0}; if (lincomb(c, 2, cf, xs, y, s)) { return 1; } }
      if (F(y, t0 + dt, k0)) { return 1; }
      { const double cf[3] = {0.75, 0.25, 0.25 * dt}; const double *xs[3] = {u, y, k0}; if (lincomb(c, 3, cf, xs, y, s)) { return 1; } }
      if (F(y, t0 + dt / 2, k0)) { return 1; }
      { const double cf[3] = {1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0 * dt}; const double *xs[3] = {u, y, k0}; if (lincomb(c, 3, cf, xs, u, s)) { return 1; } }
   }
   else if (ode == 4 || ode == 6)
   {
      // ExplicitRKSolver tables: classical RK4; RK6Solver = Verner's 8-stage 6th-order method
      static const double a4[] = {0.5, 0.0, 0.5, 0.0, 0.0, 1.0};
      static const double b4[] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
      static const double c4[] = {0.5, 0.5, 1.0};
      static const double a6[] = {
         .6e-1,
         .1923996296296296296296296296296296296296e-1, .7669337037037037037037037037037037037037e-1,
         .35975e-1, 0., .107925,
         1.318683415233148260919747276431735612861, 0., -5.042058063628562225427761634715637693344,
         4.220674648395413964508014358283902080483,
         -41.87259166432751461803757780644346812905, 0., 159.4325621631374917700365669070346830453,
         -122.1192135650100309202516203389242140663, 5.531743066200053768252631238332999150076,
         -54.43015693531650433250642051294142461271, 0., 207.0672513650184644273657173866509835987,
         -158.6108137845899991828742424365058599469, 6.991816585950242321992597280791793907096,
         -.1859723106220323397765171799549294623692e-1,
         -54.66374178728197680241215648050386959351, 0., 207.9528062553893734515824816699834244238,
         -159.2889574744995071508959805871426654216, 7.018743740796944434698170760964252490817,
         -.1833878590504572306472782005141738268361e-1, -.5119484997882099077875432497245168395840e-3};
      static const double b6[] = {
         .3438957868357036009278820124728322386520e-1, 0., 0.,
         .2582624555633503404659558098586120858767, .4209371189673537150642551514069801967032,
         4.405396469669310170148836816197095664891, -176.4831190242986576151740942499002125029,
         172.3641334014150730294022582711902413315};
      static const double c6[] = {.6e-1, .9593333333333333333333333333333333333333e-1, .1439, .4973,
                                  .9725, .9995, 1.};
      const int ns = (ode == 4) ? 4 : 8;
      const double *a = (ode == 4) ? a4 : a6, *b = (ode == 4) ? b4 : b6, *cc = (ode == 4) ? c4 : c6;
      for (int i = 0; i < ns + 1 && i < 9; i++) { if (work_vec(c, &c->rk[i])) { return 1; } }
      // k_i in rk[0..ns-1]; y reuses wk[7]
      if (work_vec(c, &c->wk[7])) { return 1; }
      double *yy = c->wk[7];
      if (F(u, t0, c->rk[0])) { return 1; }
      for (int i = 1; i < ns; i++)
      {
         const double *ai = a + i * (i - 1) / 2;
         double cf[9]; const double *xs[9];
         cf[0] = 1.0; xs[0] = u;
         for (int j = 0; j < i; j++) { cf[j + 1] = dt * ai[j]; xs[j + 1] = c->rk[j]; }
         if (lincomb(c, i + 1, cf, xs, yy, s)) { return 1; }
         if (F(yy, t0 + cc[i - 1] * dt, c->rk[i])) { return 1; }
      }
      double cf[9]; const double *xs[9];
      cf[0] = 1.0; xs[0] = u;
      for (int j = 0; j < ns; j++) { cf[j + 1] = dt * b[j]; xs[j + 1] = c->rk[j]; }
      if (lincomb(c, ns + 1, cf, xs, u, s)) { return 1; }
   }
   else if (ode == 11)      // ForwardEulerIDPSolver (remhos_solvers.cpp:29-38)
   {
      if (rmh_mult_unlimited(c, ho_type, lo_type, fct_type, t0, dt, u, k0, stream)) { return 1; }
      if (rmh_limit_mult(c, lo_type, fct_type, dt, u, k0, stream)) { return 1; }
      const double cf[2] = {1.0, dt}; const double *xs[2] = {u, k0};
      if (lincomb(c, 2, cf, xs, u, s)) { return 1; }
   }
   else if (ode == 12 || ode == 13 || ode == 14 || ode == 16)
   {
      // RKIDPSolver::Step without masks (remhos_solvers.cpp:171-249; remhos.cpp:502-507 turns the
      // masks off); tables :252-279
      static const double a2[] = {.5}, b2[] = {0., 1.}, c2[] = {.5};
      static const double a3[] = {1. / 3., 0., 2. / 3.}, b3[] = {.25, 0., .75}, c3[] = {1. / 3., 2. / 3.};
      static const double a4i[] = {1. / 3., -1. / 3., 1., 1., -1., 1.};
      static const double b4i[] = {1. / 8., 3. / 8., 3. / 8., 1. / 8.}, c4i[] = {1. / 3., 2. / 3., 1.};
      static const double a6i[] = {.25, 1. / 8., 1. / 8., 0., -.5, 1., 3. / 16., 0., 0., 9. / 16.,
                                   -3. / 7., 2. / 7., 12. / 7., -12. / 7., 8. / 7.};
      static const double b6i[] = {7. / 90., 0., 32. / 90., 12. / 90., 32. / 90., 7. / 90.};
      static const double c6i[] = {.25, .25, .5, .75, 1.};
      const int ns = (ode == 12) ? 2 : (ode == 13) ? 3 : (ode == 14) ? 4 : 6;
      const double *a = (ode == 12) ? a2 : (ode == 13) ? a3 : (ode == 14) ? a4i : a6i;
      const double *b = (ode == 12) ? b2 : (ode == 13) ? b3 : (ode == 14) ? b4i : b6i;
      const double *cc = (ode == 12) ? c2 : (ode == 13) ? c3 : (ode == 14) ? c4i : c6i;
      // ConstructD (remhos_solvers.cpp:40-95)
      double d[21] = {0};
      {
         const double *a_n = a, *a_o = a;
         int i_o = -1;
         double c_o = 0.;
         for (int i = 0; i < ns; i++)
         {
            const double c_n = (i < ns - 1) ? cc[i] : 1.;
            const double dc = c_n - c_o;
            double *di = d + i * (i + 1) / 2;
            for (int j = 0; j < i; j++)
            {
               const double a_oj = (j <= i_o) ? a_o[j] : 0.;
               const double m = (a_n[j] - a_oj) / dc;
               if (m == 0.) { di[j] = 0.; continue; }
               const double *dj = d + j * (j + 1) / 2;
               const double dij = m / dj[j];
               for (int k = 0; k < j; k++) { di[k] -= dj[k] * dij; }
               di[j] = dij;
            }
            di[i] = a_n[i] / dc;
            const double c_next = (i < ns - 2) ? cc[i + 1] : 1.;
            if (c_next > c_n) { i_o = i; c_o = c_n; a_o = a_n; }
            if (i < ns - 2) { a_n += i + 1; } else { a_n = b; }
         }
      }
      for (int i = 0; i < ns; i++) { if (work_vec(c, &c->rk[i])) { return 1; } }
      // masks (RKIDPSolver::use_masks): only a product state is ever masked (remhos.cpp:1746-1752)
      const bool masks = c->idp_mask && c->product;
      const int64_t nstate = state_len(c);
      uint8_t *mask = c->pmask, *mask_new = nullptr;
      double *x_new = nullptr;
      if (masks)
      {
         if (work_vec(c, &c->wk[7])) { return 1; }
         x_new = c->wk[7];
         if (!c->pf_dof[0]) { if (prod_buffers(c)) { return 1; } }
         uint8_t *mn = nullptr;
         if (dev_alloc(c, &mn, (size_t)nstate)) { return 1; }     // (kept until the context goes)
         mask_new = mn;
      }
      const int mbs = 256;
      const unsigned mgrid = (unsigned)((nstate + mbs - 1) / mbs);
      double c_o = 0.;
      double tcur = t0;
      if (rmh_mult_unlimited(c, ho_type, lo_type, fct_type, tcur, cc[0] * dt, u, c->rk[0], stream)) { return 1; }
      if (rmh_limit_mult(c, lo_type, fct_type, cc[0] * dt, u, c->rk[0], stream)) { return 1; }
      {
         const double c_next = (ns > 2) ? cc[1] : 1.;
         if (c_next > cc[0])
         {
            const double cf[2] = {1.0, cc[0] * dt}; const double *xs[2] = {u, c->rk[0]};
            if (lincomb(c, 2, cf, xs, u, s)) { return 1; }
            if (masks) { if (rmh_compute_mask(c, u, mask, stream)) { return 1; } }
            tcur = t0 + cc[0] * dt;
            c_o = cc[0];
         }
         else if (masks)
         {
            const double cf[2] = {1.0, cc[0] * dt}; const double *xs[2] = {u, c->rk[0]};
            if (lincomb(c, 2, cf, xs, x_new, s)) { return 1; }
            if (rmh_compute_mask(c, x_new, mask, stream)) { return 1; }
         }
      }
      const double *d_i = d + 1;
      for (int i = 1; i < ns; i++)
      {
         const double c_n = (i < ns - 1) ? cc[i] : 1.;
         const double dc = c_n - c_o, dct = dc * dt;
         if (rmh_mult_unlimited(c, ho_type, lo_type, fct_type, tcur, dct, u, c->rk[i], stream)) { return 1; }
         if (masks)
         {
            // UpdateMask with the HO update, then the masked combination: where the mask is off the
            // stage stays the plain HO rate (forward Euler), remhos_solvers.cpp:213-231
            const double *xm = u;
            if (dct != 0.)
            {
               const double cf[2] = {1.0, dct}; const double *xs[2] = {u, c->rk[i]};
               if (lincomb(c, 2, cf, xs, x_new, s)) { return 1; }
               xm = x_new;
            }
            if (rmh_compute_mask(c, xm, mask_new, stream)) { return 1; }
            k_mask_and<<<mgrid, mbs, 0, s>>>(nstate, mask, mask_new); LAUNCH_OK();
            k_add_masked<<<mgrid, mbs, 0, s>>>(nstate, mask, d_i[i] - 1., c->rk[i], c->rk[i]); LAUNCH_OK();
            for (int j = 0; j < i; j++)
            { k_add_masked<<<mgrid, mbs, 0, s>>>(nstate, mask, d_i[j], c->rk[j], c->rk[i]); LAUNCH_OK(); }
         }
         else
         {
            double cf[9]; const double *xs[9];
            cf[0] = d_i[i]; xs[0] = c->rk[i];
            for (int j = 0; j < i; j++) { cf[j + 1] = d_i[j]; xs[j + 1] = c->rk[j]; }
            if (lincomb(c, i + 1, cf, xs, c->rk[i], s)) { return 1; }
         }
         if (rmh_limit_mult(c, lo_type, fct_type, dct, u, c->rk[i], stream)) { return 1; }
         const double c_next = (i < ns - 2) ? cc[i + 1] : 1.;
         if (i == ns - 1 || c_next > c_n)
         {
            tcur = t0 + c_n * dt;
            const double cf[2] = {1.0, dct}; const double *xs[2] = {u, c->rk[i]};
            if (lincomb(c, 2, cf, xs, u, s)) { return 1; }
            c_o = c_n;
         }
         d_i += i + 1;
      }
   }
   else
   {
      set_error("rmh_ode_step: unknown ODE solver type (remhos.cpp:499-500 returns 3)");
      return 3;
   }
   (void)bytes;
   *t = t0 + dt;
   return 0;
}

// ---------------------------------------------------------------- device memory helpers
extern "C" int rmh_dev_malloc(rmh_ctx *c, int64_t n, double **out)
{
   CUDA_OK(cudaSetDevice(c->device));
   void *q = nullptr;
   CUDA_OK(cudaMalloc(&q, (size_t)std::max<int64_t>(n, 1) * sizeof(double)));
   *out = (double *)q;
   return 0;
}
extern "C" int rmh_dev_free(rmh_ctx *c, double *p)
{
   CUDA_OK(cudaSetDevice(c->device));
   CUDA_OK(cudaFree(p));
   return 0;
}
extern "C" int rmh_copy_h2d(rmh_ctx *c, double *dst, const double *src, int64_t n)
{
   if (c && dst == c->xe_ptr) { c->xe_ptr = nullptr; }     // the cached element min/max go stale
   CUDA_OK(cudaMemcpy(dst, src, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
   return 0;
}
extern "C" int rmh_copy_d2h(rmh_ctx *, double *dst, const double *src, int64_t n)
{
   CUDA_OK(cudaMemcpy(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
   return 0;
}
extern "C" int rmh_copy_d2d(rmh_ctx *c, double *dst, const double *src, int64_t n)
{
   if (c && dst == c->xe_ptr) { c->xe_ptr = nullptr; }
   CUDA_OK(cudaMemcpy(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice));
   return 0;
}
extern "C" int rmh_sync(rmh_ctx *c)
{
   CUDA_OK(cudaSetDevice(c->device));
   CUDA_OK(cudaDeviceSynchronize());
   return 0;
}

// ---------------------------------------------------------------- multi-GPU layer
extern "C" int rmh_stage_minmax(rmh_ctx *c, const double *y, void *stream)
{
   return stage_minmax(c, y, (cudaStream_t)stream);
}

#include "dist.cuh"
