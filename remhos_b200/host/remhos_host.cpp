// Host-side mirror of the Remhos solver interfaces (see remhos_host.hpp) and the remhos() driver.
#include "remhos_host.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <fstream>
#include <map>
#include <sstream>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

namespace remhos
{

void Abort(const std::string &msg) { throw std::runtime_error(msg); }
void Verify(bool cond, const std::string &msg) { if (!cond) { Abort(msg); } }
void Check(int status) { if (status != 0) { Abort(rmh_last_error()); } }

// ------------------------------------------------------------------------------------ Communicator
Communicator::Communicator()
{
   auto env_int = [](const char *k, int dflt) { const char *v = std::getenv(k); return v ? std::atoi(v) : dflt; };
   rank = env_int("RANK", 0); world = env_int("WORLD_SIZE", 1); local_rank = env_int("LOCAL_RANK", rank);
   Verify(world >= 1 && rank >= 0 && rank < world, "bad RANK / WORLD_SIZE");
   if (world > 1)
   {
      const char *d = std::getenv("RMH_RDZV_DIR");
      if (d) { dir = d; }
      else
      {
         // workers of one launcher share the parent process and the rendezvous port
         std::ostringstream os;
         os << "/dev/shm/rmh_rdzv_" << (long)getppid() << "_" << env_int("MASTER_PORT", 0);
         dir = os.str();
      }
      mkdir(dir.c_str(), 0700);
   }
}

// every rank writes <dir>/<seq>.<rank> (temporary name, then rename: readers never see a partial
// file) and reads the others'; blobs come back ordered by rank
std::vector<std::string> Communicator::AllGather(const std::string &mine)
{
   std::vector<std::string> out(world);
   if (world == 1) { out[0] = mine; return out; }
   const int id = seq++;
   auto name = [&](int r) { std::ostringstream os; os << dir << "/" << id << "." << r; return os.str(); };
   {
      const std::string tmp = name(rank) + ".tmp";
      std::ofstream f(tmp, std::ios::binary);
      f.write(mine.data(), (std::streamsize)mine.size());
      f.close();
      Verify((bool)f, "rendezvous: cannot write " + tmp);
      Verify(std::rename(tmp.c_str(), name(rank).c_str()) == 0, "rendezvous: rename failed");
   }
   const auto t0 = std::chrono::steady_clock::now();
   for (int r = 0; r < world; r++)
   {
      if (r == rank) { out[r] = mine; continue; }
      while (true)
      {
         std::ifstream f(name(r), std::ios::binary);
         if (f)
         {
            std::ostringstream ss;
            ss << f.rdbuf();
            out[r] = ss.str();
            break;
         }
         const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
         Verify(waited < 600.0, "rendezvous: timed out waiting for rank " + std::to_string(r));
         std::this_thread::sleep_for(std::chrono::milliseconds(2));
      }
   }
   return out;
}

void Communicator::Finalize()
{
   if (world == 1) { return; }
   Barrier();
   // everybody has passed gather seq-1, hence read every earlier file: drop my own (the files of the
   // last barrier stay until the launcher removes the directory)
   for (int id = 0; id + 1 < seq; id++)
   {
      std::ostringstream os; os << dir << "/" << id << "." << rank;
      std::remove(os.str().c_str());
   }
}

// ------------------------------------------------------------------------------------ Vector
Vector::Vector(const ParFiniteElementSpace &space, int blocks_) { SetSpace(space, blocks_); }
Vector::Vector(const Vector &o)
{
   if (o.fes) { SetSpace(*o.fes, o.blocks); Check(rmh_copy_d2d(fes->ctx, d, o.d, n)); }
}
const double *Vector::Block(int b) const { return d + (int64_t)b * (n / blocks); }
double *Vector::Block(int b) { return d + (int64_t)b * (n / blocks); }
Vector &Vector::operator=(const Vector &o)
{
   if (this == &o) { return *this; }
   if (!fes && o.fes) { SetSpace(*o.fes, o.blocks); }
   Verify(n == o.n, "Vector::operator=: size mismatch");
   if (n) { Check(rmh_copy_d2d(fes->ctx, d, o.d, n)); }
   return *this;
}
Vector &Vector::operator=(double v)
{
   const double c = 0.0;
   const double *x = d;
   if (v == 0.0) { Check(rmh_lincomb(fes->ctx, 1, &c, &x, d, nullptr)); }
   else { SetFromHost(std::vector<double>((size_t)n, v)); }
   return *this;
}
Vector::~Vector() { if (d && fes && fes->ctx) { rmh_dev_free(fes->ctx, d); } }
void Vector::SetSpace(const ParFiniteElementSpace &space, int blocks_)
{
   Verify(d == nullptr, "Vector::SetSpace: already allocated");
   fes = &space;
   blocks = blocks_;
   n = space.GetVSize() * blocks;
   Check(rmh_dev_malloc(space.ctx, n, &d));
   const double c = 0.0;
   const double *x = d;
   // zero-fill through the library (0 * x): allocation returns uninitialised memory
   Check(rmh_copy_h2d(space.ctx, d, std::vector<double>((size_t)n, 0.0).data(), n));
   (void)c; (void)x;
}
void Vector::SetFromHost(const std::vector<double> &h)
{
   Verify((int64_t)h.size() == n, "Vector::SetFromHost: size mismatch");
   Check(rmh_copy_h2d(fes->ctx, d, h.data(), n));
}
std::vector<double> Vector::HostRead() const
{
   std::vector<double> h((size_t)n);
   if (n) { Check(rmh_copy_d2h(fes->ctx, h.data(), d, n)); }
   return h;
}
void Vector::Add(double a, const Vector &x)
{
   const double c[2] = {1.0, a};
   const double *xs[2] = {d, x.d};
   Check(rmh_lincomb(fes->ctx, 2, c, xs, d, nullptr));
}

// ------------------------------------------------------------------------------------ space
ParFiniteElementSpace::ParFiniteElementSpace(rmh_mesh *m, int problem_, int order_, int mesh_order_,
                                             int bounds_type_, double &dt, double &t_final_,
                                             int device, Communicator *comm_)
   : mesh(m), order(order_), mesh_order(mesh_order_), bounds_type(bounds_type_), problem(problem_),
     comm(comm_)
{
   dim = rmh_mesh_dim(m);
   exec_mode = (problem < 10) ? 0 : 1;                                   // remhos.cpp:438-440
   bb_min.assign(dim, 0.0); bb_max.assign(dim, 0.0);
   Check(rmh_mesh_bounding_box(m, bb_min.data(), bb_max.data()));        // :457
   if (comm && comm->world > 1) { SetupDistributed(dt, t_final_, device); return; }
   Check(rmh_mesh_set_curvature(m, mesh_order));                         // :513
   if (dt < 0.0) { Check(rmh_cfl_dt(m, problem, bb_min.data(), bb_max.data(), &dt)); }   // :538-553
   dt_cfl = dt;
   const int64_t ne = rmh_mesh_ne(m);
   int ngn = 1, nd = 1, nf = 2 * dim, nfd = 1, n3 = 1, nsub = 1, ncorner = 1;
   for (int a = 0; a < dim; a++) { ngn *= mesh_order + 1; nd *= order + 1; n3 *= 3; nsub *= std::max(order, 1); ncorner *= 2; }
   for (int a = 0; a < dim - 1; a++) { nfd *= order + 1; }
   const double *nodes = rmh_mesh_nodes(m);
   std::vector<double> vel_nodes, vel_quad, vel_face;
   const size_t nnod = (size_t)ne * ngn * dim;
   if (exec_mode == 1)                                                   // :562-584
   {
      vel_nodes.resize(nnod);
      Check(rmh_remap_mesh_velocity(m, problem, bb_min.data(), bb_max.data(), dt, t_final_,
                                    vel_nodes.data()));
      vel_nodes_host = vel_nodes;
      t_final_ = 1.0;                                                    // :1128-1134
   }
   else
   {
      // velocities that are polynomials of degree <= 1 are reproduced exactly by their nodal
      // interpolant; anything else is sampled at the quadrature points (VectorFunctionCoefficient)
      const int pv = problem % 20;
      const bool nodal_ok = (pv == 0 || pv == 1 || pv == 2 || pv == 4 || pv == 5 || pv == 6 || pv == 7);
      if (nodal_ok)
      {
         vel_nodes.resize(nnod);
         Check(rmh_velocity(problem, dim, (int64_t)ne * ngn, nodes, bb_min.data(), bb_max.data(),
                            vel_nodes.data()));
      }
      else
      {
         const int Q = (2 * order + dim * mesh_order - 1) / 2 + 1;
         std::vector<double> xq(Q), wq(Q);
         Check(rmh_gauss_legendre_01(Q, xq.data(), wq.data()));
         int nq = 1, nqf = 1;
         for (int a = 0; a < dim; a++) { nq *= Q; }
         for (int a = 0; a < dim - 1; a++) { nqf *= Q; }
         std::vector<double> pts((size_t)ne * nq * dim);
         vel_quad.resize(pts.size());
         Check(rmh_mesh_eval(m, Q, xq.data(), -1, pts.data()));
         Check(rmh_velocity(problem, dim, (int64_t)ne * nq, pts.data(), bb_min.data(), bb_max.data(),
                            vel_quad.data()));
         vel_face.resize((size_t)ne * nf * nqf * dim);
         std::vector<double> fp((size_t)ne * nqf * dim), fv(fp.size());
         for (int f = 0; f < nf; f++)
         {
            Check(rmh_mesh_eval(m, Q, xq.data(), f, fp.data()));
            Check(rmh_velocity(problem, dim, (int64_t)ne * nqf, fp.data(), bb_min.data(),
                               bb_max.data(), fv.data()));
            for (int64_t e = 0; e < ne; e++)
            {
               std::memcpy(&vel_face[((size_t)e * nf + f) * nqf * dim], &fv[(size_t)e * nqf * dim],
                           sizeof(double) * nqf * dim);
            }
         }
      }
   }
   t_final = t_final_;
   std::vector<int32_t> bd((size_t)nfd * nf), nbr((size_t)ne * nf * nfd), s2i((size_t)nsub * ncorner),
       lat((size_t)ne * n3), nbe((size_t)ne * nf);
   int32_t n_ent = 0;
   Check(rmh_mesh_dof_maps(m, order, bd.data(), nbr.data(), s2i.data(), lat.data(), &n_ent, nbe.data()));
   // DOF positions: uniform lattice i/p (ProjectCoefficient on the positive basis, remhos.cpp:883)
   std::vector<double> lp(order + 1);
   for (int i = 0; i <= order; i++) { lp[i] = (double)i / std::max(order, 1); }
   std::vector<double> xdof((size_t)ne * nd * dim), infl((size_t)ne * nd);
   Check(rmh_mesh_eval(m, order + 1, lp.data(), -1, xdof.data()));
   Check(rmh_inflow_project(m, problem, order, infl.data()));              // remhos.cpp:625-636
   rmh_desc d;
   std::memset(&d, 0, sizeof(d));
   d.dim = dim; d.order = order; d.mesh_order = mesh_order; d.exec_mode = exec_mode;
   d.bounds_type = bounds_type; d.device = device; d.ne = ne; d.ne_ghost = 0;
   d.nodes = nodes;
   d.vel_nodes = vel_nodes.empty() ? nullptr : vel_nodes.data();
   d.vel_quad = vel_quad.empty() ? nullptr : vel_quad.data();
   d.vel_face = vel_face.empty() ? nullptr : vel_face.data();
   d.nbr_dof = nbr.data(); d.lat = lat.data(); d.n_ent = n_ent; d.nbr_elem = nbe.data();
   d.inflow = infl.data();
   Check(rmh_ctx_create(&d, &ctx));
   u0.resize((size_t)ne * nd);
   Check(rmh_u0(problem, dim, (int64_t)ne * nd, xdof.data(), bb_min.data(), bb_max.data(), u0.data()));
   xlat = xdof;
   bdr_dofs = bd;
   nbr_elem = nbe;
}
// One rank's share of a decomposed run: what ParMesh(MPI_COMM_WORLD, mesh) and the parallel space do
// at remhos.cpp:459-463,588-623.  Partition (every rank computes the same recursive bisection), halo
// plan with the owned elements interior first, local mesh = owned + ghost ring, context with a ghost
// layer, exchange plan, device layer; the set-up blobs travel through the Communicator.
void ParFiniteElementSpace::SetupDistributed(double &dt, double &t_final_, int device)
{
   const int pv = problem % 20;
   Verify(exec_mode == 1 || pv == 0 || pv == 1 || pv == 2 || pv == 4 || pv == 5 || pv == 6 || pv == 7,
          "decomposed transport runs sample the velocity at the mesh nodes (problems 0, 1, 2, 4, 5, 6, 7)");
   global_mesh = mesh;
   const int rank = comm->rank, world = comm->world;
   std::vector<int32_t> part((size_t)rmh_mesh_ne(global_mesh));
   Check(rmh_mesh_partition(global_mesh, world, part.data()));
   Check(rmh_halo_create(global_mesh, part.data(), rank, &halo));
   int64_t n_interior = 0;
   Check(rmh_halo_interior_first(halo, &n_interior));
   int64_t no = 0, ng = 0, nsend = 0;
   int32_t npeer = 0;
   Check(rmh_halo_sizes(halo, &no, &ng, &npeer, &nsend));
   std::vector<int64_t> ids((size_t)(no + ng));
   {
      std::vector<int32_t> peers(npeer), soff(npeer + 1), roff(npeer + 1), sloc((size_t)nsend);
      Check(rmh_halo_get(halo, ids.data(), ids.data() + no, peers.data(), soff.data(), roff.data(), sloc.data()));
   }
   Check(rmh_mesh_extract(global_mesh, no + ng, ids.data(), &local_mesh));
   Check(rmh_mesh_set_curvature(local_mesh, mesh_order));                // :513
   std::vector<int64_t> own((size_t)no);
   for (int64_t i = 0; i < no; i++) { own[i] = i; }
   Check(rmh_mesh_extract(local_mesh, no, own.data(), &mesh));           // owned part: what GetNE() counts
   int ngn = 1, nd = 1, nf = 2 * dim, nfd = 1, n3 = 1, nsub = 1, ncorner = 1;
   for (int a = 0; a < dim; a++) { ngn *= mesh_order + 1; nd *= order + 1; n3 *= 3; nsub *= std::max(order, 1); ncorner *= 2; }
   for (int a = 0; a < dim - 1; a++) { nfd *= order + 1; }
   const double *nodes = rmh_mesh_nodes(mesh);
   std::vector<double> vel_nodes((size_t)no * ngn * dim);
   auto rank_min = [&](double v)          // MPI_MIN before the device layer exists: through the rendezvous
   {
      const std::vector<std::string> all = comm->AllGather(std::string(reinterpret_cast<const char *>(&v), sizeof(v)));
      for (const std::string &sv : all) { double w; std::memcpy(&w, sv.data(), sizeof(w)); v = std::min(v, w); }
      return v;
   };
   bool dt_done = false;
   if (exec_mode == 1)                                                   // remhos.cpp:538-584
   {
      if (dt < 0.0)
      {
         Check(rmh_cfl_dt(mesh, problem, bb_min.data(), bb_max.data(), &dt));
         dt = rank_min(dt);
      }
      dt_done = true;
      Check(rmh_remap_mesh_velocity(mesh, problem, bb_min.data(), bb_max.data(), dt, t_final_, vel_nodes.data()));
      vel_nodes_host = vel_nodes;
      t_final_ = 1.0;                                                    // :1128-1134
   }
   else { Check(rmh_velocity(problem, dim, no * ngn, nodes, bb_min.data(), bb_max.data(), vel_nodes.data())); }
   const int64_t na = no + ng;
   std::vector<int32_t> bd((size_t)nfd * nf), nbr((size_t)na * nf * nfd), s2i((size_t)nsub * ncorner),
       lat((size_t)na * n3), nbe((size_t)na * nf);
   int32_t n_ent = 0;
   Check(rmh_mesh_dof_maps(local_mesh, order, bd.data(), nbr.data(), s2i.data(), lat.data(), &n_ent, nbe.data()));
   std::vector<double> lp(order + 1);
   for (int i = 0; i <= order; i++) { lp[i] = (double)i / std::max(order, 1); }
   std::vector<double> xdof((size_t)no * nd * dim), infl((size_t)no * nd);
   Check(rmh_mesh_eval(mesh, order + 1, lp.data(), -1, xdof.data()));
   Check(rmh_inflow_project(mesh, problem, order, infl.data()));
   rmh_desc d;
   std::memset(&d, 0, sizeof(d));
   d.dim = dim; d.order = order; d.mesh_order = mesh_order; d.exec_mode = exec_mode;
   d.bounds_type = bounds_type; d.device = device; d.ne = no; d.ne_ghost = ng;
   d.nodes = nodes; d.vel_nodes = vel_nodes.data();
   d.nbr_dof = nbr.data(); d.lat = lat.data(); d.n_ent = n_ent; d.nbr_elem = nbe.data();
   d.inflow = infl.data();
   Check(rmh_ctx_create(&d, &ctx));
   Check(rmh_dplan_create(halo, rank, world, dim, order, nbr.data(), &dplan));
   Check(rmh_dist_create(ctx, dplan, rank, world, n_interior, &dist));
   {
      std::string blob((size_t)rmh_dist_blob_bytes(dist), '\0');
      Check(rmh_dist_export(dist, &blob[0]));
      const std::vector<std::string> all = comm->AllGather(blob);
      std::vector<const void *> ptrs(world);
      std::vector<int64_t> sizes(world);
      for (int r = 0; r < world; r++) { ptrs[r] = all[r].data(); sizes[r] = (int64_t)all[r].size(); }
      Check(rmh_dist_connect(dist, world, ptrs.data(), sizes.data()));
   }
   if (dt < 0.0 && !dt_done)                                             // :538-553 (MPI_MIN at :551)
   {
      Check(rmh_cfl_dt(mesh, problem, bb_min.data(), bb_max.data(), &dt));
      dt = Reduce(dt, 1);
   }
   dt_cfl = dt;
   t_final = t_final_;
   global_vsize = (int64_t)std::llround(Reduce((double)(no * nd), 0));
   u0.resize((size_t)no * nd);
   Check(rmh_u0(problem, dim, no * nd, xdof.data(), bb_min.data(), bb_max.data(), u0.data()));
   xlat = xdof;
   bdr_dofs = bd;
   nbr_elem.assign(nbe.begin(), nbe.begin() + (size_t)no * nf);
}

void ParFiniteElementSpace::SaveMesh(const std::string &path, double t, int precision, int refine_factor) const
{
   std::vector<double> x;
   if (exec_mode == 1 && !vel_nodes_host.empty())
   {
      const double *x0 = rmh_mesh_nodes(mesh);
      x.resize(vel_nodes_host.size());
      for (size_t i = 0; i < x.size(); i++) { x[i] = x0[i] + t * vel_nodes_host[i]; }   // remhos.cpp:1602
   }
   const double *nodes = x.empty() ? nullptr : x.data();
   if (refine_factor > 1)
   {
      // the subcell mesh: ParMesh::MakeRefined(pmesh, order, ClosedUniform), moved with the HO mesh
      // (remhos.cpp:801, 1021-1026, 1371-1376)
      rmh_mesh *sub = nullptr;
      Check(rmh_mesh_make_refined(mesh, refine_factor, nodes, &sub));
      const int rc = rmh_mesh_save(sub, path.c_str(), nullptr, precision);
      rmh_mesh_free(sub);
      Check(rc);
      return;
   }
   Check(rmh_mesh_save(mesh, path.c_str(), nodes, precision));
}

void ParFiniteElementSpace::SaveGridFunction(const std::string &path, const std::vector<double> &vals,
                                             int precision) const
{
   Check(rmh_gf_save(path.c_str(), dim, order, 2, (int64_t)vals.size(), vals.data(), precision));
}

void VisItDataCollection::Save() const
{
   char cyc[16], rk[16];
   std::snprintf(cyc, sizeof(cyc), "%06d", cycle);
   const int rank = pfes.comm ? pfes.comm->rank : 0, world = pfes.comm ? pfes.comm->world : 1;
   std::snprintf(rk, sizeof(rk), "%06d", rank);
   const std::string dir = name + "_" + cyc;
   mkdir(dir.c_str(), 0755);
   pfes.SaveMesh(dir + "/mesh." + rk, time, precision);
   for (const auto &f : fields)
   {
      std::vector<double> h = f.second->HostRead();
      h.resize((size_t)pfes.GetVSize());                     // block 0 of a product state
      pfes.SaveGridFunction(dir + "/" + f.first + "." + rk, h, precision);
   }
   if (rank != 0) { return; }
   std::ofstream root(dir + ".mfem_root");
   root << "{\n  \"dsets\": {\n    \"main\": {\n      \"cycle\": " << cycle << ",\n      \"domains\": " << world
        << ",\n      \"fields\": {\n";
   for (size_t i = 0; i < fields.size(); i++)
   {
      root << "        \"" << fields[i].first << "\": {\n          \"path\": \"" << dir << "/" << fields[i].first
           << ".%06d\",\n          \"tags\": {\n            \"assoc\": \"nodes\",\n            \"comps\": \"1\",\n"
           << "            \"lod\": \"" << std::max(pfes.order, 1) << "\"\n          }\n        }"
           << (i + 1 < fields.size() ? "," : "") << "\n";
   }
   root << "      },\n      \"mesh\": {\n        \"path\": \"" << dir << "/mesh.%06d\",\n        \"tags\": {\n"
        << "          \"max_lods\": \"32\"\n        }\n      },\n      \"time\": " << std::setprecision(16) << time
        << ",\n      \"time_step\": 0.0\n    }\n  }\n}\n";
}

double ParFiniteElementSpace::Reduce(double v, int op) const
{
   if (!dist) { return v; }
   Check(rmh_dist_allreduce(dist, op, &v, 1, nullptr));
   return v;
}

ParFiniteElementSpace::~ParFiniteElementSpace()
{
   if (dist)
   {
      // no rank may tear its window down while a peer can still write into it
      Check(rmh_sync(ctx));
      comm->Barrier();
      rmh_dist_destroy(dist); dist = nullptr;
   }
   if (ctx) { rmh_ctx_destroy(ctx); ctx = nullptr; }
   if (dplan) { rmh_dplan_free(dplan); }
   if (halo) { rmh_halo_free(halo); }
   if (local_mesh) { rmh_mesh_free(local_mesh); }
   if (global_mesh && mesh && mesh != global_mesh) { rmh_mesh_free(mesh); mesh = global_mesh; }
}
int64_t ParFiniteElementSpace::GetNE() const { return rmh_mesh_ne(mesh); }
int64_t ParFiniteElementSpace::GetVSize() const { return rmh_ctx_ndofs(ctx); }
int ParFiniteElementSpace::GetNDofs() const { return rmh_ctx_nd(ctx); }

// ------------------------------------------------------------------------------------ solvers
void LocalInverseHOSolver::CalcHOSolution(const Vector &u, Vector &du) const
{
   Verify(timer != nullptr, "Timer not set.");                           // remhos_ho.cpp:86
   Check(rmh_ho_local_inverse(pfes.ctx, u.Read(), du.Write(), nullptr));
}

NeumannHOSolver::NeumannHOSolver(ParFiniteElementSpace &space) : HOSolver(space)
{
   Check(rmh_fa_setup(space.ctx, nullptr));
}
void NeumannHOSolver::CalcHOSolution(const Vector &u, Vector &du) const
{
   Check(rmh_ho_neumann(pfes.ctx, u.Read(), du.Write(), nullptr));
}

DiscreteUpwind::DiscreteUpwind(ParFiniteElementSpace &space, bool preconditioned)
   : LOSolver(space), prec(preconditioned)
{
   Check(rmh_fa_setup(space.ctx, nullptr));
}
void DiscreteUpwind::CalcLOSolution(const Vector &u, Vector &du) const
{
   if (prec) { Check(rmh_lo_discrete_upwind_prec(pfes.ctx, u.Read(), du.Write(), nullptr)); }
   else { Check(rmh_lo_discrete_upwind(pfes.ctx, u.Read(), du.Write(), nullptr)); }
}
void ResidualDistribution::CalcLOSolution(const Vector &u, Vector &du) const
{
   Check(rmh_lo_res_dist(pfes.ctx, u.Read(), du.Write(), nullptr));
}
ResidualDistributionSubcell::ResidualDistributionSubcell(ParFiniteElementSpace &space)
   : LOSolver(space)
{
   SetupSubcells(space);
}
void SetupSubcells(ParFiniteElementSpace &space)
{
   if (space.subcells_ready) { return; }
   Verify(space.order > 1, "Subcell schemes require FE order > 1.");     // remhos.cpp:613-616
   const int dim = space.dim, p = space.order, nf = 2 * dim;
   int nd = 1, ns = 1, nc = 1, nfd = 1;
   for (int a = 0; a < dim; a++) { nd *= p + 1; ns *= p; nc *= 2; }
   for (int a = 0; a < dim - 1; a++) { nfd *= p + 1; }
   const int64_t ne = space.GetNE();
   std::vector<double> vel;
   if (space.exec_mode == 1)
   {
      // v_sub_gf: velocity_function at the subcell vertices, zero on the boundary (remhos.cpp:838-852)
      vel.resize((size_t)ne * nd * dim);
      Check(rmh_velocity(space.problem, dim, ne * nd, space.xlat.data(), space.bb_min.data(),
                         space.bb_max.data(), vel.data()));
      for (int64_t e = 0; e < ne; e++)
         for (int f = 0; f < nf; f++)
         {
            if (space.nbr_elem[(size_t)e * nf + f] >= 0) { continue; }
            for (int j = 0; j < nfd; j++)
            {
               const int loc = space.bdr_dofs[(size_t)j * nf + f];
               for (int c = 0; c < dim; c++) { vel[((size_t)e * nd + loc) * dim + c] = 0.0; }
            }
         }
   }
   else
   {
      // velocity at the subcell centres (midpoint rule of MixedConvectionIntegrator)
      std::vector<double> xc((size_t)ne * ns * dim, 0.0);
      for (int64_t e = 0; e < ne; e++)
         for (int m = 0; m < ns; m++)
         {
            int sc[3] = {0, 0, 0}, r = m;
            for (int a = 0; a < dim; a++) { sc[a] = r % p; r /= p; }
            for (int c = 0; c < nc; c++)
            {
               int loc = 0, mul = 1;
               for (int a = 0; a < dim; a++) { loc += (sc[a] + ((c >> a) & 1)) * mul; mul *= p + 1; }
               for (int i = 0; i < dim; i++)
               { xc[((size_t)e * ns + m) * dim + i] += space.xlat[((size_t)e * nd + loc) * dim + i] / nc; }
            }
         }
      vel.resize(xc.size());
      Check(rmh_velocity(space.problem, dim, ne * ns, xc.data(), space.bb_min.data(),
                         space.bb_max.data(), vel.data()));
   }
   Check(rmh_subcell_setup(space.ctx, space.xlat.data(), vel.data(), nullptr));
   space.subcells_ready = true;
}

// ---- smoothness indicator (remhos_tools.cpp:24-354)
SmoothnessIndicator::SmoothnessIndicator(int type_id, ParFiniteElementSpace &pfes_DG)
   : pfes(pfes_DG), type(type_id)
{
   Verify(type_id == 1 || type_id == 2, "Bad smoothness indicator id!");
   Check(rmh_si_setup(pfes.ctx, type_id, nullptr));
}
SmoothnessIndicator::~SmoothnessIndicator() { rmh_si_setup(pfes.ctx, 0, nullptr); }
void SmoothnessIndicator::ComputeSmoothnessIndicator(const Vector &u, Vector &si_vals_u) const
{
   Check(rmh_si_values(pfes.ctx, u.Read(), si_vals_u.Write(), nullptr));
}
void SmoothnessIndicator::UpdateBounds(double dt, const Vector &u, const Vector &du_ho, const Vector &si_vals_u,
                                       Vector &u_min, Vector &u_max) const
{
   Check(rmh_si_update_bounds(pfes.ctx, dt, u.Read(), du_ho.Read(), si_vals_u.Read(), u_min.Write(), u_max.Write(),
                              nullptr));
}

// ---- monolithic solver (remhos_mono.cpp)
MonoRDSolver::MonoRDSolver(ParFiniteElementSpace &space, SmoothnessIndicator *si, bool subcell,
                           bool timedep, bool masslim)
   : MonolithicSolver(space), subcell_scheme(subcell), time_dep(timedep), mass_lim(masslim)
{
   (void)si;      // the indicator, if any, is already registered with the device context
   // scale(e) = vmax / (2 sqrt(dim) h_e / order), vmax over the element quadrature rule of order
   // OrderW + 2 p + 2 max(OrderGrad, 0) (remhos_mono.cpp:40-57); tensor elements: OrderW = dim mo - 1,
   // OrderGrad = mo (dim - 1) + p - 1, Gauss-Legendre with order/2 + 1 points per direction
   const int dim = space.dim, p = space.order, mo = space.mesh_order;
   const int64_t ne = space.GetNE();
   const int q_ord = (dim * mo - 1) + 2 * p + 2 * std::max(mo * (dim - 1) + p - 1, 0);
   const int n = q_ord / 2 + 1;
   std::vector<double> xq(n), wq(n);
   Check(rmh_gauss_legendre_01(n, xq.data(), wq.data()));
   int nq = 1;
   for (int a = 0; a < dim; a++) { nq *= n; }
   std::vector<double> pts((size_t)ne * nq * dim), vel(pts.size()), h((size_t)ne);
   Check(rmh_mesh_eval(space.mesh, n, xq.data(), -1, pts.data()));
   Check(rmh_velocity(space.problem, dim, ne * nq, pts.data(), space.bb_min.data(), space.bb_max.data(),
                      vel.data()));
   Check(rmh_mesh_elem_sizes(space.mesh, h.data()));
   scale.resize((size_t)ne);
   for (int64_t e = 0; e < ne; e++)
   {
      double vmax = 0.0;
      for (int q = 0; q < nq; q++)
      {
         double v2 = 0.0;
         for (int c = 0; c < dim; c++) { const double v = vel[((size_t)e * nq + q) * dim + c]; v2 += v * v; }
         vmax = std::max(vmax, std::sqrt(v2));
      }
      scale[e] = vmax / (2. * (std::sqrt((double)dim) * h[e] / p));
   }
   if (subcell_scheme) { SetupSubcells(space); }
   Check(rmh_mono_setup(space.ctx, subcell_scheme ? 2 : 1, mass_lim ? 1 : 0, scale.data(), nullptr));
}
MonoRDSolver::~MonoRDSolver() { rmh_mono_setup(pfes.ctx, 0, 0, nullptr, nullptr); }
void MonoRDSolver::CalcSolution(const Vector &u, Vector &du) const
{
   Check(rmh_mono_rd(pfes.ctx, u.Read(), du.Write(), nullptr));
}
void ResidualDistributionSubcell::CalcLOSolution(const Vector &u, Vector &du) const
{
   Check(rmh_lo_res_dist_subcell(pfes.ctx, u.Read(), du.Write(), nullptr));
}
void MassBasedAvg::CalcLOSolution(const Vector &u, Vector &du) const
{
   // remhos_lo.cpp:253-262: use the HO solution handed over by LimitMult, or compute it
   Vector *own = nullptr;
   const Vector *ho = du_HO;
   if (!ho)
   {
      own = new Vector(pfes);
      ho_solver.CalcHOSolution(u, *own);
      ho = own;
   }
   Check(rmh_lo_mass_avg(pfes.ctx, dt, u.Read(), ho->Read(), du.Write(), nullptr));
   du_HO = nullptr;
   delete own;
}

FluxBasedFCT::FluxBasedFCT(ParFiniteElementSpace &space, double dt_) : FCTSolver(space, dt_)
{
   Check(rmh_fa_setup(space.ctx, nullptr));
}
void FluxBasedFCT::CalcFCTSolution(const Vector &u, const Vector &m, const Vector &du_ho,
                                   const Vector &du_lo, const Vector &u_min, const Vector &u_max,
                                   Vector &du) const
{
   Check(rmh_fct_flux_based(pfes.ctx, dt, u.Read(), m.Read(), du_ho.Read(), du_lo.Read(),
                            u_min.Read(), u_max.Read(), du.Write(), nullptr));
}
ElementFCTProjection::ElementFCTProjection(ParFiniteElementSpace &space, double dt_)
   : FCTSolver(space, dt_)
{
   Check(rmh_fa_setup(space.ctx, nullptr));      // dense element mass blocks
}
void ElementFCTProjection::CalcFCTSolution(const Vector &u, const Vector &m, const Vector &du_ho,
                                           const Vector &du_lo, const Vector &u_min,
                                           const Vector &u_max, Vector &du) const
{
   (void)m;                                       // the solver lumps its own element mass (remhos_fct.cpp:650)
   Check(rmh_fct_project(pfes.ctx, dt, u.Read(), du_ho.Read(), du_lo.Read(), u_min.Read(), u_max.Read(),
                         du.Write(), nullptr));
}
void ClipScaleSolver::CalcFCTSolution(const Vector &u, const Vector &m, const Vector &du_ho,
                                      const Vector &du_lo, const Vector &u_min, const Vector &u_max,
                                      Vector &du) const
{
   Check(rmh_fct_clip_scale(pfes.ctx, dt, u.Read(), m.Read(), du_ho.Read(), du_lo.Read(),
                            u_min.Read(), u_max.Read(), du.Write(), nullptr));
}

void NonlinearPenaltySolver::CalcFCTSolution(const Vector &u, const Vector &m, const Vector &du_ho,
                                             const Vector &du_lo, const Vector &u_min, const Vector &u_max,
                                             Vector &du) const
{
   // eps of CorrectFlux: GetElementSize(0, 0) / GetOrder(0) (remhos_fct.cpp:961)
   std::vector<double> h((size_t)rmh_mesh_ne(pfes.mesh));
   Check(rmh_mesh_elem_sizes(pfes.mesh, h.data()));
   const double eps_w = h.empty() ? 0.0 : h[0] / pfes.order;
   if (smth_indicator)
   {
      Vector si(pfes), mn(u_min), mx(u_max);
      smth_indicator->ComputeSmoothnessIndicator(u, si);
      smth_indicator->UpdateBounds(dt, u, du_ho, si, mn, mx);
      Check(rmh_fct_nonlinear_penalty(pfes.ctx, dt, eps_w, u.Read(), m.Read(), du_ho.Read(), du_lo.Read(),
                                      mn.Read(), mx.Read(), du.Write(), nullptr));
      return;
   }
   Check(rmh_fct_nonlinear_penalty(pfes.ctx, dt, eps_w, u.Read(), m.Read(), du_ho.Read(), du_lo.Read(),
                                   u_min.Read(), u_max.Read(), du.Write(), nullptr));
}

// ------------------------------------------------------------------------------------ DofInfo
DofInfo::DofInfo(ParFiniteElementSpace &space, int btype)
   : pfes(space), bounds_type(btype), xi_min(space), xi_max(space)
{
   const int dim = space.dim, p = space.order;
   int nd = 1, nfd = 1, nsub = 1, nc = 1;
   for (int a = 0; a < dim; a++) { nd *= p + 1; nsub *= std::max(p, 1); nc *= 2; }
   for (int a = 0; a < dim - 1; a++) { nfd *= p + 1; }
   numBdrs = 2 * dim; numFaceDofs = nfd; numSubcells = nsub; numDofsSubcell = nc;
   const int64_t ne = space.GetNE();
   BdrDofs.resize((size_t)nfd * numBdrs);
   NbrDof.resize((size_t)ne * numBdrs * nfd);
   Sub2Ind.resize((size_t)nsub * nc);
   Check(rmh_mesh_dof_maps(space.mesh, p, BdrDofs.data(), NbrDof.data(), Sub2Ind.data(), nullptr,
                           nullptr, nullptr));
   Check(rmh_dev_malloc(space.ctx, ne, &xe_min));
   Check(rmh_dev_malloc(space.ctx, ne, &xe_max));
}
DofInfo::~DofInfo()
{
   if (xe_min) { rmh_dev_free(pfes.ctx, xe_min); }
   if (xe_max) { rmh_dev_free(pfes.ctx, xe_max); }
}
void DofInfo::ComputeElementsMinMax(const Vector &u, double *u_min, double *u_max) const
{
   Check(rmh_elem_min_max(pfes.ctx, u.Read(), u_min, u_max, nullptr));
}
void DofInfo::ComputeBounds(const double *el_min, const double *el_max, Vector &dof_min,
                            Vector &dof_max) const
{
   Check(rmh_bounds(pfes.ctx, el_min, el_max, dof_min.Write(), dof_max.Write(), nullptr));
}

// ------------------------------------------------------------------------------------ operator
AdvectionOperator::AdvectionOperator(ParFiniteElementSpace &space, Vector &lumpedM_, DofInfo &dofs_,
                                     HOSolver *hos, LOSolver *los, FCTSolver *fct, MonolithicSolver *mos)
   : pfes(space), lumpedM(lumpedM_), dofs(dofs_), ho_solver(hos), lo_solver(los), fct_solver(fct),
     mono_solver(mos)
{
   if (ho_solver) { ho_solver->timer = &timer; }
   if (lo_solver) { lo_solver->timer = &timer; }
   if (fct_solver) { fct_solver->timer = &timer; }
}
void AdvectionOperator::SetDt(double dt_)                                // remhos.cpp:176-182
{
   dt = dt_;
   if (lo_solver) { lo_solver->UpdateTimeStep(dt_); }
   if (fct_solver) { fct_solver->UpdateTimeStep(dt_); }
}
void AdvectionOperator::SetTime(double t_) { t = t_; }
void AdvectionOperator::MultUnlimited(const Vector &x, Vector &y) const  // remhos.cpp:1596-1739
{
   // remap: move the mesh to x0 + t v and re-assemble (:1598-1677)
   Check(rmh_set_time(pfes.ctx, t, nullptr));
   if (pfes.exec_mode == 1) { Check(rmh_lumped_mass(pfes.ctx, lumpedM.Write(), nullptr)); }
   if (mono_solver) { mono_solver->CalcSolution(x, y); }                 // remhos.cpp:1687
   else if (fct_solver)
   {
      Verify(ho_solver && lo_solver, "FCT requires HO and LO solvers.");
      ho_solver->CalcHOSolution(x, y);
   }
   else if (lo_solver) { lo_solver->CalcLOSolution(x, y); }
   else if (ho_solver) { ho_solver->CalcHOSolution(x, y); }
   else { Abort("No solver was chosen."); }
}
void AdvectionOperator::LimitMult(const Vector &x, Vector &y) const     // remhos.cpp:1798-1916
{
   if (!fct_solver || mono_solver) { return; }
   Verify(ho_solver && lo_solver, "FCT requires HO and LO solvers.");
   Vector du_HO(y), du_LO(pfes);
   auto mba = dynamic_cast<MassBasedAvg *>(lo_solver);
   if (mba) { mba->SetHOSolution(du_HO); }
   lo_solver->CalcLOSolution(x, du_LO);
   dofs.ComputeElementsMinMax(x, dofs.xe_min, dofs.xe_max);
   dofs.ComputeBounds(dofs.xe_min, dofs.xe_max, dofs.xi_min, dofs.xi_max);
   fct_solver->CalcFCTSolution(x, lumpedM, du_HO, du_LO, dofs.xi_min, dofs.xi_max, y);
   if (verify_bounds)                                                    // check_violation, :1576-1594
   {
      const std::vector<double> u = x.HostRead(), du = y.HostRead(), mn = dofs.xi_min.HostRead(),
                                mx = dofs.xi_max.HostRead();
      for (size_t i = 0; i < u.size(); i++)
      {
         const double un = u[i] + dt * du[i];
         if (un + 1e-12 < mn[i] || un > mx[i] + 1e-12)
         {
            std::ostringstream os;
            os << std::setprecision(12) << "LimitMult FCT solution u bounds: " << mn[i] << " " << un
               << " " << mx[i];
            Abort(os.str());
         }
      }
   }
}
void AdvectionOperator::Mult(const Vector &x, Vector &y) const
{
   // the fused stage kernel covers -ho 3 -lo 5 -fct 2; anything else goes solver by solver
   if (HOType() == 3 && LOType() == 5 && FCTType() == 2 && !verify_bounds)
   {
      Check(rmh_mult(pfes.ctx, 3, 5, 2, t, dt, x.Read(), y.Write(), nullptr));
      return;
   }
   MultUnlimited(x, y);
   LimitMult(x, y);
}
void AdvectionOperator::PrintTimingData(int steps, double stage_seconds) const   // remhos.cpp:1918-1966
{
   if (!(HOType() == 3 && LOType() == 5 && FCTType() == 2)) { return; }
   const double dofs_steps = (double)pfes.GlobalVSize() * steps;
   std::cout << "---" << std::endl;
   std::cout << "RHS+L2inv+LO+FCT run as one fused stage kernel per RK stage" << std::endl
             << "Total kernel time: " << stage_seconds << std::endl;
   std::cout << "---" << std::endl;
   std::cout << "FOM:     " << 1e-6 * dofs_steps / stage_seconds << std::endl;
   std::cout << "(megadofs x time steps / second)\n---" << std::endl;
}

// ------------------------------------------------------------------------------------ ODE
bool ODESolver::Known(int t)
{
   return t == 1 || t == 2 || t == 3 || t == 4 || t == 6 || t == 11 || t == 12 || t == 13 ||
          t == 14 || t == 16;
}
int ODESolver::Stages() const
{
   switch (type)            // remhos.cpp:1340-1347 (the FOM counts 6 for RK6)
   {
      case 2: return 2;
      case 3: return 3;
      case 4: return 4;
      case 6: return 6;
      default: return 1;
   }
}
void ODESolver::Step(Vector &x, double &t, double &dt)
{
   Verify(f != nullptr, "ODESolver::Init was not called");
   if (f->verify_bounds && (type >= 1 && type <= 3))
   {
      // checked path: same Butcher forms through AdvectionOperator::Mult
      ParFiniteElementSpace &sp = f->Space();
      Vector k(sp), y(sp);
      auto F = [&](const Vector &in, double tt) { f->SetTime(tt); f->Mult(in, k); };
      if (type == 1) { F(x, t); x.Add(dt, k); }
      else if (type == 2)
      {
         F(x, t); y = x; y.Add(dt, k); x.Add(0.5 * dt, k);
         F(y, t + dt); x.Add(0.5 * dt, k);
      }
      else
      {
         F(x, t); y = x; y.Add(dt, k);
         F(y, t + dt); y.Add(dt, k);
         { const double c[2] = {0.75, 0.25}; const double *xs[2] = {x.Read(), y.Read()};
           Check(rmh_lincomb(sp.ctx, 2, c, xs, y.Write(), nullptr)); }
         F(y, t + dt / 2); y.Add(dt, k);
         { const double c[2] = {1.0 / 3.0, 2.0 / 3.0}; const double *xs[2] = {x.Read(), y.Read()};
           Check(rmh_lincomb(sp.ctx, 2, c, xs, x.Write(), nullptr)); }
      }
      t += dt;
      return;
   }
   if (f->Space().dist && type >= 1 && type <= 3 && f->HOType() == 3 && f->LOType() == 5 && f->FCTType() == 2 &&
       !f->Space().dt_control_on)
   {
      // decomposed mesh: fused stage path with the halo exchange behind the C ABI (remhos.cpp:1813)
      Check(rmh_dist_rk_step(f->Space().dist, type, f->LOType(), &t, dt, x.ReadWrite(), nullptr));
      return;
   }
   // (on a decomposed mesh rmh_ode_step exchanges the halo of every operator input itself)
   const int rc = rmh_ode_step(f->Space().ctx, type, f->HOType(), f->LOType(), f->FCTType(), &t, dt,
                               x.ReadWrite(), nullptr);
   Check(rc);
}

// ------------------------------------------------------------------------------------ driver
namespace
{
struct Opt
{
   std::string mesh_file = "default", device = "cpu";
   int dim = 3, epm = 1, problem = 0, rs = 2, rp = 0, order = 3, mesh_order = 2, ode = 3, ho = 3,
       lo = 0, fct = 0, mono = 0, bt = 0, si = 0, dtc = 0, max_steps = -1, vis_steps = 100, pool = 4;
   int gpus = 1;          // -gpus N: one process per GPU, started by main() (the reference: mpirun -np N)
   bool pa = false, full = false, gam = false, vis = true, save = false, visit = false, vb = false,
        ps = false;
   double t_final = 4.0, dt = 0.005;
};

void usage(std::ostream &os)
{
   os << "Usage: remhos [options]\n"
         "  -m <mesh>  -dim <d>  -epm <n>  -p <problem>  -rs <n>  -rp <n>  -o <order>  -mo <order>\n"
         "  -s <ode: 1,2,3,4,6,11,12,13,14,16>  -ho <0|1|2|3>  -lo <0..5>  -fct <0|1|2>  -mono <0|1|2>\n"
         "  -bt <0|1>  -pa/-no-pa  -full/-no-full  -d <device>  -gam/-no-gam  -si <0>  -tf <t>\n"
         "  -dtc <0>  -dt <dt>  -ms <steps>  -vis/-no-vis  -save/-no-save  -visit/-no-visit\n"
         "  -vb/-no-vb  -ps/-no-ps  -vs <steps>  -pool <GB>\n"
         "  -gpus <N>  decompose the mesh over N GPUs of this node (one process per GPU; -ho 3 -lo 5 -fct 2)\n";
}

// returns false on a bad command line (remhos.cpp:335-339 prints the usage and returns 1)
bool parse(int argc, char *argv[], Opt &o)
{
   std::map<std::string, int *> ints = {
      {"-dim", &o.dim}, {"--dimension", &o.dim}, {"-epm", &o.epm}, {"--elem-per-mpi", &o.epm},
      {"-p", &o.problem}, {"--problem", &o.problem}, {"-rs", &o.rs}, {"--refine-serial", &o.rs},
      {"-rp", &o.rp}, {"--refine-parallel", &o.rp}, {"-o", &o.order}, {"--order", &o.order},
      {"-mo", &o.mesh_order}, {"--mesh-order", &o.mesh_order}, {"-s", &o.ode}, {"--ode-solver", &o.ode},
      {"-ho", &o.ho}, {"--ho-type", &o.ho}, {"-lo", &o.lo}, {"--lo-type", &o.lo},
      {"-fct", &o.fct}, {"--fct-type", &o.fct}, {"-mono", &o.mono}, {"--mono-type", &o.mono},
      {"-bt", &o.bt}, {"--bounds-type", &o.bt}, {"-si", &o.si}, {"--smth_ind", &o.si},
      {"-dtc", &o.dtc}, {"--dt-control", &o.dtc}, {"-ms", &o.max_steps}, {"--max-steps", &o.max_steps},
      {"-vs", &o.vis_steps}, {"--visualization-steps", &o.vis_steps}, {"-pool", &o.pool},
      {"-gpus", &o.gpus}, {"--num-gpus", &o.gpus},
      {"--dev-pool-size", &o.pool}};
   std::map<std::string, double *> dbls = {{"-tf", &o.t_final}, {"--t-final", &o.t_final},
                                           {"-dt", &o.dt}, {"--time-step", &o.dt}};
   std::map<std::string, std::string *> strs = {{"-m", &o.mesh_file}, {"--mesh", &o.mesh_file},
                                                {"-d", &o.device}, {"--device", &o.device}};
   std::map<std::string, std::pair<bool *, bool>> flags = {
      {"-pa", {&o.pa, true}}, {"--partial-assembly", {&o.pa, true}}, {"-no-pa", {&o.pa, false}},
      {"--no-partial-assembly", {&o.pa, false}}, {"-full", {&o.full, true}}, {"-no-full", {&o.full, false}},
      {"--next-gen-full", {&o.full, true}}, {"--no-next-gen-full", {&o.full, false}},
      {"-gam", {&o.gam, true}}, {"-no-gam", {&o.gam, false}}, {"--gpu-aware-mpi", {&o.gam, true}},
      {"--no-gpu-aware-mpi", {&o.gam, false}}, {"-vis", {&o.vis, true}}, {"-no-vis", {&o.vis, false}},
      {"--visualization", {&o.vis, true}}, {"--no-visualization", {&o.vis, false}},
      {"-save", {&o.save, true}}, {"-no-save", {&o.save, false}}, {"-visit", {&o.visit, true}},
      {"-no-visit", {&o.visit, false}}, {"--visit-datafiles", {&o.visit, true}},
      {"--no-visit-datafiles", {&o.visit, false}}, {"-vb", {&o.vb, true}}, {"-no-vb", {&o.vb, false}},
      {"--verify-bounds", {&o.vb, true}}, {"--dont-verify-bounds", {&o.vb, false}},
      {"-ps", {&o.ps, true}}, {"-no-ps", {&o.ps, false}}, {"--product-sync", {&o.ps, true}},
      {"--no-product-sync", {&o.ps, false}}};
   for (int i = 1; i < argc; i++)
   {
      const std::string a = argv[i];
      if (a == "-h" || a == "--help") { return false; }
      auto fi = flags.find(a);
      if (fi != flags.end()) { *fi->second.first = fi->second.second; continue; }
      if (i + 1 >= argc) { std::cout << "Missing value for option " << a << std::endl; return false; }
      const char *v = argv[++i];
      char *end = nullptr;
      auto ii = ints.find(a);
      if (ii != ints.end())
      {
         *ii->second = (int)std::strtol(v, &end, 10);
         if (*end) { std::cout << "Wrong value for option " << a << std::endl; return false; }
         continue;
      }
      auto di = dbls.find(a);
      if (di != dbls.end())
      {
         *di->second = std::strtod(v, &end);
         if (*end) { std::cout << "Wrong value for option " << a << std::endl; return false; }
         continue;
      }
      auto si = strs.find(a);
      if (si != strs.end()) { *si->second = v; continue; }
      std::cout << "Unrecognized option: " << a << std::endl;
      return false;
   }
   return true;
}
} // namespace

int remhos(int argc, char *argv[], double &final_mass_u)
{
   Opt o;
   Communicator comm;                                                    // remhos.cpp:212-214 (Mpi::Init)
   // like the reference, only the root rank talks (remhos.cpp:335-339 and every "if (myid == 0)")
   std::ostringstream null_out;
   std::streambuf *cout_buf = comm.Root() ? nullptr : std::cout.rdbuf(null_out.rdbuf());
   struct Restore { std::streambuf *b; ~Restore() { if (b) { std::cout.rdbuf(b); } } } restore{cout_buf};
   if (!parse(argc, argv, o)) { usage(std::cout); return 1; }           // remhos.cpp:335-339
   if (comm.world > 1)
   {
      // fused stage path for -ho 3 -lo 5 -fct 2 -s 1/2/3; every other combination runs solver by solver with
      // one halo exchange per operator evaluation (the flux-based FCT: two more, for R+ and R-,
      // remhos_fct.cpp:406-409).  Not decomposed: the monolithic solver (serial in the reference too),
      // smoothness indicators, product remap, the per-stage -vb checks.
      Verify(o.mono == 0 && o.si == 0 && !o.ps && !o.vb,
             "decomposed runs (WORLD_SIZE > 1) do not cover -mono, -si, -ps and -vb");
      Verify(!(o.fct == 1 && o.problem >= 10), "decomposed remap runs do not cover -fct 1");
   }
   // ---- combinations the reference rejects (Appendix A of SURVEY.md) or this build lacks
   if (!ODESolver::Known(o.ode))
   {
      std::cout << "Unknown ODE solver type: " << o.ode << '\n';       // remhos.cpp:499-500
      return 3;
   }
   Verify(o.mono >= 0 && o.mono <= 2, "monolithic solver type must be 0, 1 (ResDistMono) or 2 (ResDistMonoSubcell)");
   if (o.mono == 2) { Verify(o.order > 1, "Subcell schemes require FE order > 1."); }
   // -ho 2 (CGHOSolver, remhos_ho.cpp:30-70) solves the block-diagonal system M du = K u by PCG to
   // a relative tolerance of 1e-12: on this path it is served by the exact element-local inverse
   Verify(o.ho >= 0 && o.ho <= 3, "HO solver type must be 0 .. 3");
   if (o.ho == 2) { o.ho = 3; }
   Verify(!(o.ho == 1 && o.pa), "PA for DG is not supported for Neummann Solver.");   // remhos_ho.cpp:138-139
   Verify(o.lo >= 0 && o.lo <= 5, "LO solver type must be 0 .. 5");
   if (o.lo == 4) { Verify(o.order > 1, "Subcell schemes require FE order > 1."); }
   Verify(o.fct >= 0 && o.fct <= 4, "FCT solver type must be 0 .. 4");
   Verify(!(o.fct == 3 && o.problem >= 10), "-fct 3 (NonlinearPenalty) is built for transport mode");
   Verify(!(o.fct == 4 && o.pa), "FCTProject needs the assembled element mass (no -pa).");
   if (o.ps)
   {
      Verify(o.problem >= 10, "Products are processed only in remap mode.");                 // remhos.cpp:1850
      Verify(!o.dtc, "Automatic time step is not implemented for product remap.");          // :1851
      Verify(o.fct != 0 && !o.mono && !o.vb, "product remap (-ps) needs an FCT solver (and no -vb / -mono) in this build");
   }
   Verify(o.si >= 0 && o.si <= 2, "Bad smoothness indicator id!");
   if (o.si)
   {
      Verify(o.mono != 0 || o.fct == 2 || o.fct == 3, "smoothness indicators (-si) act on -mono, -fct 2 and -fct 3");
   }
   Verify(o.dtc == 0 || o.dtc == 1, "time step control must be 0 (fixed) or 1 (LO bounds error)");
   if (o.dtc) { Verify(o.fct != 0 && !o.vb, "-dtc 1 needs an FCT solver (and no -vb) in this build"); }
   Verify(!(o.fct == 1 && o.pa), "Flux-based FCT is not compatible with partial assembly.");   // :1088
   Verify(o.order >= 1, "order 0 disables limiting; not part of this build");
   if (o.fct) { Verify(o.ho && o.lo, "FCT requires HO and LO solvers."); }     // :1690
   if (o.lo == 5) { Verify(o.ho != 0, "Mass-Based LO solver requires a choice of a HO solver."); }   // :991
   Verify(o.ho || o.lo || o.mono, "No solver was chosen.");
   // ---- mesh (remhos.cpp:448-463)
   rmh_mesh *mesh = nullptr;
   if (o.mesh_file == "default")
   {
      // PartitionMPI(dim, world, elem_per_mpi, ..., rp_levels) (remhos.cpp:451-455): a Cartesian mesh of
      // the unit box with world * epm elements AFTER the -rp refinements, so that every rank ends up
      // with exactly elem_per_mpi elements (verified at remhos.cpp:466-471); -rs applies to file meshes only
      Verify(o.dim == 2 || o.dim == 3, "-dim must be 2 or 3");
      Verify(o.epm >= 1 && o.rp >= 0, "Mesh generation error.");
      const long total = (long)comm.world * o.epm, per = 1L << (o.dim * o.rp);
      Verify(total % per == 0, "Mesh generation error.");
      int n[3] = {1, 1, 1};
      long left = total / per;                 // coarse elements, as cubic as possible
      for (int a = 0; a < o.dim; a++)
      {
         long k = std::lround(std::pow((double)left, 1.0 / (o.dim - a)));
         while (k > 1 && left % k) { k--; }
         n[a] = (int)std::max(k, 1L); left /= n[a];
      }
      const double org[3] = {0, 0, 0}, sz[3] = {1, 1, 1};
      Check(rmh_mesh_cartesian(o.dim, n, org, sz, 0, &mesh));
      Check(rmh_mesh_refine(mesh, o.rp));
      Verify((long)rmh_mesh_ne(mesh) == total, "Mesh generation error.");
   }
   else
   {
      Check(rmh_mesh_load(o.mesh_file.c_str(), &mesh));
      Check(rmh_mesh_refine(mesh, o.rs + o.rp));
   }
   double dt = o.dt, t_final = o.t_final;
   int rc = 0;
   {
      ParFiniteElementSpace pfes(mesh, o.problem, o.order, o.mesh_order, o.bt, dt, t_final, comm.local_rank, &comm);
      std::cout << "Number of unknowns: " << pfes.GlobalVSize() << std::endl;   // remhos.cpp:623
      // S = (u, us) with a product field (BlockVector, remhos.cpp:594-598); `u` names block 0 throughout
      Vector u(pfes, o.ps ? 2 : 1), lumpedM(pfes);
      const int64_t NV = pfes.GetVSize();
      if (!o.ps) { u.SetFromHost(pfes.u0); }
      else
      {
         // us = u * s, s = s0_function on the elements where u is active (BoolFunctionCoefficient over
         // ComputeBoolIndicators), sampled at the lattice points (remhos.cpp:886-903, 2357-2361)
         Verify(comm.world == 1, "product remap runs on one GPU");
         std::vector<double> S0(pfes.u0);
         S0.resize((size_t)2 * NV, 0.0);
         const int nd = pfes.GetNDofs(), dim = pfes.dim;
         const double two_pi = 2.0 * M_PI;
         for (int64_t e = 0; e < NV / nd; e++)
         {
            bool active = false;
            for (int j = 0; j < nd; j++) { if (pfes.u0[e * nd + j] > 1e-12) { active = true; } }   // EMPTY_ZONE_TOL
            for (int j = 0; j < nd; j++)
            {
               const double *x = &pfes.xlat[((size_t)e * nd + j) * dim];
               const double s0 = active ? 2.0 + std::sin(two_pi * x[0]) * std::sin(two_pi * x[1]) : 0.0;
               S0[NV + e * nd + j] = pfes.u0[e * nd + j] * s0;
            }
         }
         u.SetFromHost(S0);
         Check(rmh_product_enable(pfes.ctx, 1));
      }
      Check(rmh_lumped_mass(pfes.ctx, lumpedM.Write(), nullptr));
      DofInfo dofs(pfes, o.bt);
      HOSolver *ho_solver = nullptr;
      if (o.ho == 3) { ho_solver = new LocalInverseHOSolver(pfes); }
      else if (o.ho == 1) { ho_solver = new NeumannHOSolver(pfes); }
      LOSolver *lo_solver = nullptr;
      if (o.lo == 1) { lo_solver = new DiscreteUpwind(pfes); }
      else if (o.lo == 2) { lo_solver = new DiscreteUpwind(pfes, true); }
      else if (o.lo == 3) { lo_solver = new ResidualDistribution(pfes); }
      else if (o.lo == 4) { lo_solver = new ResidualDistributionSubcell(pfes); }
      else if (o.lo == 5) { lo_solver = new MassBasedAvg(pfes, *ho_solver); }
      SmoothnessIndicator *smth_indicator = nullptr;                     // remhos.cpp:905-911
      if (o.si) { smth_indicator = new SmoothnessIndicator(o.si, pfes); }
      FCTSolver *fct_solver = nullptr;
      if (o.fct == 1) { fct_solver = new FluxBasedFCT(pfes, dt); }
      else if (o.fct == 2) { fct_solver = new ClipScaleSolver(pfes, dt); }
      else if (o.fct == 3) { fct_solver = new NonlinearPenaltySolver(pfes, smth_indicator, dt); }
      else if (o.fct == 4) { fct_solver = new ElementFCTProjection(pfes, dt); }
      if (o.dtc) { Check(rmh_dt_control(pfes.ctx, 1)); pfes.dt_control_on = true; }
      // monolithic solver (remhos.cpp:997-1011)
      MonolithicSolver *mono_solver = nullptr;
      const bool mass_lim = (o.problem != 6 && o.problem != 7);
      if (o.mono)
      { mono_solver = new MonoRDSolver(pfes, smth_indicator, o.mono == 2, pfes.exec_mode == 1, mass_lim); }
      AdvectionOperator adv(pfes, lumpedM, dofs, ho_solver, lo_solver, fct_solver, mono_solver);
      adv.verify_bounds = o.vb;
      double mass0_u = 0.0, u_min = 0.0, u_max = 0.0;
      Check(rmh_reduce(pfes.ctx, 0, u.Read(), lumpedM.Read(), &mass0_u, nullptr));   // :1073-1076
      Check(rmh_reduce(pfes.ctx, 1, u.Read(), nullptr, &u_min, nullptr));
      Check(rmh_reduce(pfes.ctx, 2, u.Read(), nullptr, &u_max, nullptr));
      mass0_u = pfes.Reduce(mass0_u, 0); u_min = pfes.Reduce(u_min, 1); u_max = pfes.Reduce(u_max, 2);   // MPI_Allreduce
      double mass0_us = 0.0;
      if (o.ps) { Check(rmh_reduce(pfes.ctx, 0, u.Block(1), lumpedM.Read(), &mass0_us, nullptr)); }   // :1079-1083
      // Print the starting mesh and initial condition (remhos.cpp:1015-1030); VisIt collection (:1032-1043)
      const int precision = 8;
      const int lo_factor = (o.lo == 4 || o.mono == 2) ? o.order : 1;     // use_subcell_RD (remhos.cpp:610-612)
      if (o.save)
      {
         Verify(comm.world == 1, "-save writes one rank's files (PrintAsOne): run it on one GPU");
         pfes.SaveMesh("meshHO_init.mesh", 0.0, precision);
         pfes.SaveMesh("meshLO_init.mesh", 0.0, precision, lo_factor);   // the HO mesh itself without a subcell scheme (:870)
         std::vector<double> h = u.HostRead();
         h.resize((size_t)NV);
         pfes.SaveGridFunction("sltn_init.gf", h, precision);
      }
      VisItDataCollection *dc = nullptr;
      if (o.visit)
      {
         dc = new VisItDataCollection("Remhos", pfes);
         dc->SetPrecision(precision);
         dc->RegisterField("solution", &u);
         dc->SetCycle(0); dc->SetTime(0.0);
         dc->Save();
      }
      ODESolver ode_solver(o.ode);
      ode_solver.Init(adv);
      // the time loop below only reads the state between steps: the element min/max the last RK
      // stage leaves for its output are reused by the next step (fused RK1/2/3 path)
      Check(rmh_ctx_trust_state(pfes.ctx, 1));
      const bool steady = (o.problem == 6 || o.problem == 7 || o.problem == 8);
      std::vector<double> res_h, ml_h;
      if (steady) { res_h = u.HostRead(); ml_h = lumpedM.HostRead(); }
      double t = 0.0, residual = 0.0;
      bool done = false;
      int ti = 0, ti_total = 0;       // ti_total also counts the steps -dtc repeats (remhos.cpp:1142,1176)
      const bool forced_bounds = (o.lo != 0 || o.mono != 0);            // remhos.cpp:593-594
      Check(rmh_sync(pfes.ctx));
      const auto w0 = std::chrono::steady_clock::now();
      while (!done)                                                      // remhos.cpp:1146-1330
      {
         double dt_real = std::min(dt, t_final - t);
         adv.SetDt(dt_real);
         Vector *u_old = o.dtc ? new Vector(u) : nullptr;             // Sold = S (remhos.cpp:1173)
         if (o.dtc) { Check(rmh_dt_ratio(pfes.ctx, 1, nullptr)); }    // ResetTimeStepRatio
         ode_solver.Step(u, t, dt_real);
         ti++;
         ti_total++;
         if (o.dtc)                                                    // remhos.cpp:1178-1197
         {
            double dt_ratio = 0.0;
            Check(rmh_dt_ratio(pfes.ctx, 0, &dt_ratio));
            dt_ratio = pfes.Reduce(dt_ratio, 1);                        // MPI_MIN (remhos.cpp:1990-1996)
            if (dt_ratio < 1.)
            {
               std::cout << "Repeat / decrease dt: " << dt_real << " --> " << 0.85 * dt << std::endl;
               ti--; t -= dt_real; u = *u_old; dt = 0.85 * dt;
               delete u_old;
               Verify(dt >= 1e-12, "The time step crashed!");
               continue;
            }
            else if (dt_ratio > 1.25) { dt *= 1.02; }
            delete u_old;
         }
         // Monotonicity check (remhos.cpp:1218-1260): global extrema against the previous step's
         if (o.vb && forced_bounds && smth_indicator == nullptr)
         {
            const double eps = 1e-10;
            double mn = 0.0, mx = 0.0;
            Check(rmh_reduce(pfes.ctx, 1, u.Read(), nullptr, &mn, nullptr));
            Check(rmh_reduce(pfes.ctx, 2, u.Read(), nullptr, &mx, nullptr));
            mn = pfes.Reduce(mn, 1); mx = pfes.Reduce(mx, 2);
            auto msg = [](const char *what, double v) { std::ostringstream os; os << what << v; return os.str(); };
            if (o.problem % 10 != 6 && o.problem % 10 != 7)
            {
               Verify(mn > u_min - eps, msg("Undershoot of ", u_min - mn));
               Verify(mx < u_max + eps, msg("Overshoot of ", mx - u_max));
               u_min = mn; u_max = mx;
            }
            else
            {
               Verify(mn > 0.0 - eps, msg("Undershoot of ", 0.0 - mn));
               Verify(mx < 1.0 + eps, msg("Overshoot of ", mx - 1.0));
            }
         }
         if (!steady) { done = (t >= t_final - 1.e-8 * dt); }
         else
         {
            const std::vector<double> uh = u.HostRead();
            double r = 0.0;
            for (size_t i = 0; i < uh.size(); i++)
            {
               r += std::pow((ml_h[i] * uh[i] / dt) - (ml_h[i] * res_h[i] / dt), 2.);
            }
            residual = std::sqrt(pfes.Reduce(r, 0));                    // MPI_SUM (remhos.cpp:1288)
            if (residual < 1.e-12 && t >= 1.) { done = true; u.SetFromHost(res_h); }
            else { res_h = uh; }
         }
         if (ti_total == o.max_steps) { done = true; }                 // remhos.cpp:1296
         if (done || ti % o.vis_steps == 0)
         {
            std::cout << "time step: " << ti << ", time: " << t << ", dt: " << dt
                      << ", residual: " << residual << std::endl;
            if (dc) { dc->SetCycle(ti); dc->SetTime(t); dc->Save(); }   // remhos.cpp:1323-1328
         }
      }
      Check(rmh_sync(pfes.ctx));
      const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
      adv.PrintTimingData(ti_total * ode_solver.Stages(), wall);      // remhos.cpp:1340-1348
      // final mass on the final mesh (remhos.cpp:1382-1415)
      if (pfes.exec_mode == 1)
      {
         Check(rmh_set_time(pfes.ctx, t, nullptr));
         Check(rmh_lumped_mass(pfes.ctx, lumpedM.Write(), nullptr));
      }
      double mass_u = 0.0;
      Check(rmh_reduce(pfes.ctx, 0, u.Read(), lumpedM.Read(), &mass_u, nullptr));
      Check(rmh_reduce(pfes.ctx, 2, u.Read(), nullptr, &u_max, nullptr));
      mass_u = pfes.Reduce(mass_u, 0); u_max = pfes.Reduce(u_max, 2);     // remhos.cpp:1403-1415
      final_mass_u = mass_u;
      std::cout << std::setprecision(10) << "Final mass u:  " << mass_u << std::endl
                << "Max value u:   " << u_max << std::endl << std::setprecision(6)
                << "Mass loss u:   " << std::abs(mass0_u - mass_u) << std::endl;
      if (o.ps)                                                          // remhos.cpp:1416-1436
      {
         double mass_us = 0.0, s_max = 0.0;
         Check(rmh_reduce(pfes.ctx, 0, u.Block(1), lumpedM.Read(), &mass_us, nullptr));
         // ComputeRatio(us, u, s, ...); s.Max()
         Vector s(pfes), flags(pfes);      // flags: byte arrays (ne + N bytes fit in N doubles)
         uint8_t *el = reinterpret_cast<uint8_t *>(flags.Write());
         uint8_t *dof = el + ((pfes.GetNE() + 15) / 16) * 16;
         Check(rmh_prod_compute_ratio(pfes.ctx, u.Block(1), u.Block(0), s.Write(), el, dof, nullptr));
         Check(rmh_reduce(pfes.ctx, 2, s.Read(), nullptr, &s_max, nullptr));
         std::cout << std::setprecision(10) << "Final mass us: " << mass_us << std::endl
                   << "Max value s:   " << s_max << std::endl << std::setprecision(6)
                   << "Mass loss us:  " << std::abs(mass0_us - mass_us) << std::endl;
      }
      if (o.save)                                                        // remhos.cpp:1365-1380,1472-1482
      {
         pfes.SaveMesh("meshHO_final.mesh", t, precision);
         pfes.SaveMesh("meshLO_final.mesh", t, precision, lo_factor);
         std::vector<double> h = u.HostRead();
         h.resize((size_t)NV);
         pfes.SaveGridFunction("sltn_final.gf", h, precision);
         if (smth_indicator)
         {
            Vector si_val(pfes), u0v(pfes);
            Check(rmh_copy_d2d(pfes.ctx, u0v.Write(), u.Block(0), NV));
            smth_indicator->ComputeSmoothnessIndicator(u0v, si_val);
            pfes.SaveGridFunction("si_final.gf", si_val.HostRead(), precision);
         }
      }
      delete dc;
      delete mono_solver; delete smth_indicator; delete fct_solver; delete lo_solver; delete ho_solver;   // remhos.cpp:1484-1489
   }
   rmh_mesh_free(mesh);
   comm.Finalize();
   return rc;
}

} // namespace remhos
