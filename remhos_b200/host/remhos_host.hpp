// Host-side mirror of the Remhos solver interfaces on top of the remhos_b200 C ABI.
//
// Same class names, call shapes and error behaviour as the reference:
//   HOSolver / LocalInverseHOSolver                remhos_ho.hpp:29-42, 63-67
//   LOSolver / DiscreteUpwind / ResidualDistribution / MassBasedAvg
//                                                  remhos_lo.hpp:28-44, 48-66, 68-83, 87-109
//   FCTSolver / FluxBasedFCT / ClipScaleSolver     remhos_fct.hpp:31-90, 92-135, 137-155
//   DofInfo                                        remhos_tools.hpp:114-189
//   LimitedTimeDependentOperator                   remhos_solvers.hpp:25-63
//   AdvectionOperator                              remhos.cpp:115-198
//   ODE solvers (-s 1,2,3,4,6,11,12,13,14,16)      remhos.cpp:486-501, remhos_solvers.hpp
//   int remhos(argc, argv, final_mass_u)           remhos.cpp:210
// MFEM types are replaced by two thin ones: Vector (FP64, device-resident) and
// ParFiniteElementSpace (mesh + order + device context).  This layer is plain C++ (g++): all
// arithmetic happens behind the C ABI in librmh_b200.so; there is no CPU fallback.
#ifndef REMHOS_HOST_HPP
#define REMHOS_HOST_HPP

#include "../../include/remhos_b200.h"

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace remhos
{

// MFEM_VERIFY / MFEM_ABORT terminate the job in the reference; here they throw, and main()
// turns the exception into the message + abort the reference would produce.
[[noreturn]] void Abort(const std::string &msg);
void Verify(bool cond, const std::string &msg);
void Check(int status);   // nonzero C-ABI status -> Abort(rmh_last_error())

class ParFiniteElementSpace;

// What MPI_COMM_WORLD is to the reference (remhos.cpp:212-214): rank / size of a run with one process
// per GPU of one node, and a byte all-gather for the set-up blobs.  Ranks come from the environment
// (RANK, WORLD_SIZE, LOCAL_RANK: torchrun --no-python, or the driver's own -gpus N launcher); the
// all-gather goes through files in a shared directory (RMH_RDZV_DIR, default under /dev/shm).  The
// data path never goes through here: halo puts and reductions live behind the C ABI (rmh_dist_*).
class Communicator
{
   int seq = 0;
   std::string dir;
public:
   int rank = 0, world = 1, local_rank = 0;
   Communicator();
   bool Root() const { return rank == 0; }
   std::vector<std::string> AllGather(const std::string &mine);
   void Barrier() { AllGather("b"); }
   void Finalize();
};

// FP64 vector living in device memory of the space's context
class Vector
{
   const ParFiniteElementSpace *fes = nullptr;
   double *d = nullptr;
   int64_t n = 0;
   int blocks = 1;
public:
   Vector() {}
   // blocks > 1: a BlockVector of that many fields on the space (remhos.cpp:594-598: S = (u, us))
   explicit Vector(const ParFiniteElementSpace &space, int blocks = 1);
   Vector(const Vector &o);
   Vector &operator=(const Vector &o);
   Vector &operator=(double v);
   ~Vector();
   void SetSpace(const ParFiniteElementSpace &space, int blocks = 1);
   // device pointer of block b
   const double *Block(int b) const;
   double *Block(int b);
   int64_t Size() const { return n; }
   const double *Read() const { return d; }
   double *Write() { return d; }
   double *ReadWrite() { return d; }
   void SetFromHost(const std::vector<double> &h);
   std::vector<double> HostRead() const;
   // this += a * x
   void Add(double a, const Vector &x);
   const ParFiniteElementSpace *Space() const { return fes; }
};

// mesh + DG order + the device context (what ParMesh + DG_FECollection + ParFiniteElementSpace
// + the assembled forms are to the reference, remhos.cpp:448-727)
class ParFiniteElementSpace
{
public:
   rmh_mesh *mesh = nullptr;
   rmh_ctx *ctx = nullptr;
   int dim = 0, order = 0, mesh_order = 2, exec_mode = 0, bounds_type = 0, problem = 0;
   std::vector<double> bb_min, bb_max;
   std::vector<double> u0;          // ProjectCoefficient(u0_function) (remhos.cpp:883)
   std::vector<double> xlat;        // lattice points i/p of every element at t = 0 [ne][nd][dim]
   std::vector<int32_t> bdr_dofs, nbr_elem;   // BdrDofs [nfd][nf]; face neighbours [ne][nf]
   std::vector<double> vel_nodes_host;   // remap: mesh velocity at the nodes (the mesh at time t is x0 + t v)
   double dt_cfl = 0.0;
   double t_final = 0.0;
   bool subcells_ready = false;
   bool dt_control_on = false;      // -dtc 1: the LO rate must stay visible to the dt estimate (unfused path)
   // pmesh.Print of the mesh at time t (remap: moved nodes) / u.Save, in MFEM's formats (remhos.cpp:1016-1030)
   void SaveMesh(const std::string &path, double t, int precision = 8, int refine_factor = 1) const;
   void SaveGridFunction(const std::string &path, const std::vector<double> &vals, int precision = 8) const;
   // decomposed runs (comm.world > 1): `mesh` is this rank's owned part, the members below hold the rest
   Communicator *comm = nullptr;
   rmh_mesh *global_mesh = nullptr, *local_mesh = nullptr;   // local = owned + ghost ring
   rmh_halo *halo = nullptr;
   rmh_dplan *dplan = nullptr;
   rmh_dist *dist = nullptr;
   int64_t global_vsize = 0;
   ParFiniteElementSpace(rmh_mesh *m, int problem, int order, int mesh_order, int bounds_type,
                         double &dt, double &t_final, int device, Communicator *comm = nullptr);
   ~ParFiniteElementSpace();
   int64_t GetNE() const;
   int64_t GlobalVSize() const { return dist ? global_vsize : GetVSize(); }
   // MPI_Allreduce of one scalar (op 0 sum, 1 min, 2 max); identity on one rank
   double Reduce(double v, int op) const;
private:
   void SetupDistributed(double &dt, double &t_final, int device);
public:
   int64_t GetVSize() const;
   int GetNDofs() const;
};

struct TimingData   // remhos_tools.hpp: sw_rhs, sw_L2inv, sw_LO, sw_FCT
{
   double sw_rhs = 0.0, sw_L2inv = 0.0, sw_LO = 0.0, sw_FCT = 0.0;
};

class HOSolver
{
protected:
   ParFiniteElementSpace &pfes;
public:
   TimingData *timer = nullptr;
   HOSolver(ParFiniteElementSpace &space) : pfes(space) {}
   virtual ~HOSolver() {}
   virtual void CalcHOSolution(const Vector &u, Vector &du) const = 0;
   virtual int Type() const { return 3; }
};

class LocalInverseHOSolver : public HOSolver
{
public:
   LocalInverseHOSolver(ParFiniteElementSpace &space) : HOSolver(space) {}
   void CalcHOSolution(const Vector &u, Vector &du) const override;
};

// -ho 1 (FA only)
class NeumannHOSolver : public HOSolver
{
public:
   NeumannHOSolver(ParFiniteElementSpace &space);
   void CalcHOSolution(const Vector &u, Vector &du) const override;
   int Type() const override { return 1; }
};

class LOSolver
{
protected:
   ParFiniteElementSpace &pfes;
   double dt = -1.0;
public:
   TimingData *timer = nullptr;
   LOSolver(ParFiniteElementSpace &space) : pfes(space) {}
   virtual ~LOSolver() {}
   virtual void UpdateTimeStep(double dt_new) { dt = dt_new; }
   virtual void CalcLOSolution(const Vector &u, Vector &du) const = 0;
   virtual int Type() const = 0;
};

class DiscreteUpwind : public LOSolver
{
   bool prec;      // -lo 2: upwinding of the preconditioned blocks M_L M^-1 K (remhos.cpp:749-771)
public:
   DiscreteUpwind(ParFiniteElementSpace &space, bool preconditioned = false);
   void CalcLOSolution(const Vector &u, Vector &du) const override;
   int Type() const override { return prec ? 2 : 1; }
};

class ResidualDistribution : public LOSolver
{
public:
   ResidualDistribution(ParFiniteElementSpace &space) : LOSolver(space) {}
   void CalcLOSolution(const Vector &u, Vector &du) const override;
   int Type() const override { return 3; }
};

// -lo 4: the subcell variant; the constructor builds the low-order refined mesh data
// (remhos.cpp:797-868) and hands it to the device
class ResidualDistributionSubcell : public LOSolver
{
public:
   ResidualDistributionSubcell(ParFiniteElementSpace &space);
   void CalcLOSolution(const Vector &u, Vector &du) const override;
   int Type() const override { return 4; }
};

// the subcell mesh data of remhos.cpp:797-868, handed to the device once per space
void SetupSubcells(ParFiniteElementSpace &space);

class MassBasedAvg : public LOSolver
{
   HOSolver &ho_solver;
   mutable const Vector *du_HO = nullptr;   // borrowed for one call (remhos_lo.hpp:93-102)
public:
   MassBasedAvg(ParFiniteElementSpace &space, HOSolver &hos) : LOSolver(space), ho_solver(hos) {}
   void SetHOSolution(Vector &du) { du_HO = &du; }
   void CalcLOSolution(const Vector &u, Vector &du) const override;
   int Type() const override { return 5; }
};

class FCTSolver
{
protected:
   ParFiniteElementSpace &pfes;
   double dt;
public:
   TimingData *timer = nullptr;
   bool verify_bounds = false;
   FCTSolver(ParFiniteElementSpace &space, double dt_) : pfes(space), dt(dt_) {}
   virtual ~FCTSolver() {}
   virtual void UpdateTimeStep(double dt_new) { dt = dt_new; }
   bool NeedsLOProductInput() const { return false; }
   virtual void CalcFCTSolution(const Vector &u, const Vector &m, const Vector &du_ho,
                                const Vector &du_lo, const Vector &u_min, const Vector &u_max,
                                Vector &du) const = 0;
   virtual int Type() const = 0;
};

class FluxBasedFCT : public FCTSolver
{
public:
   FluxBasedFCT(ParFiniteElementSpace &space, double dt_);
   void CalcFCTSolution(const Vector &u, const Vector &m, const Vector &du_ho, const Vector &du_lo,
                        const Vector &u_min, const Vector &u_max, Vector &du) const override;
   int Type() const override { return 1; }
};

class ClipScaleSolver : public FCTSolver
{
public:
   ClipScaleSolver(ParFiniteElementSpace &space, double dt_) : FCTSolver(space, dt_) {}
   void CalcFCTSolution(const Vector &u, const Vector &m, const Vector &du_ho, const Vector &du_lo,
                        const Vector &u_min, const Vector &u_max, Vector &du) const override;
   int Type() const override { return 2; }
};

// remhos_fct.hpp:177-192 (-fct 3): penalty-based flux correction; with a smoothness indicator the
// bounds are relaxed first (remhos_fct.cpp:780-795)
class SmoothnessIndicator;
class NonlinearPenaltySolver : public FCTSolver
{
   SmoothnessIndicator *smth_indicator;
public:
   NonlinearPenaltySolver(ParFiniteElementSpace &space, SmoothnessIndicator *si, double dt_)
      : FCTSolver(space, dt_), smth_indicator(si) {}
   void CalcFCTSolution(const Vector &u, const Vector &m, const Vector &du_ho, const Vector &du_lo,
                        const Vector &u_min, const Vector &u_max, Vector &du) const override;
   int Type() const override { return 3; }
};

// remhos_tools.hpp SmoothnessIndicator (remhos_tools.cpp:24-354; created at remhos.cpp:905-911).
// Any order: H1 order-1 operators on the subcell mesh, assembled into the device context of the space.
class SmoothnessIndicator
{
   ParFiniteElementSpace &pfes;
   int type;
public:
   SmoothnessIndicator(int type_id, ParFiniteElementSpace &pfes_DG);
   ~SmoothnessIndicator();
   // one value per DG dof: the indicator at the dof's vertex, 1 on the domain boundary
   void ComputeSmoothnessIndicator(const Vector &u, Vector &si_vals_u) const;
   // UpdateBounds (remhos_tools.cpp:183-190) on every dof, u_HO = u + dt du_HO
   void UpdateBounds(double dt, const Vector &u, const Vector &du_ho, const Vector &si_vals_u, Vector &u_min,
                     Vector &u_max) const;
};

// remhos_mono.hpp:28-39
class MonolithicSolver
{
protected:
   ParFiniteElementSpace &pfes;
public:
   MonolithicSolver(ParFiniteElementSpace &space) : pfes(space) {}
   virtual ~MonolithicSolver() {}
   virtual void CalcSolution(const Vector &u, Vector &du) const = 0;
};

// remhos_mono.hpp:44-65.  The assembled matrices, the lumped mass, the Assembly object and the
// velocity coefficient of the reference's constructor live in the device context of the space;
// no smoothness indicator yet (-si).
class MonoRDSolver : public MonolithicSolver
{
   bool subcell_scheme, time_dep, mass_lim;
public:
   std::vector<double> scale;      // remhos_mono.cpp:40-57
   MonoRDSolver(ParFiniteElementSpace &space, SmoothnessIndicator *si, bool subcell, bool timedep,
                bool masslim);
   ~MonoRDSolver();
   void CalcSolution(const Vector &u, Vector &du) const override;
};

// remhos_fct.hpp:157-174
class ElementFCTProjection : public FCTSolver
{
public:
   ElementFCTProjection(ParFiniteElementSpace &space, double dt_);
   void CalcFCTSolution(const Vector &u, const Vector &m, const Vector &du_ho, const Vector &du_lo,
                        const Vector &u_min, const Vector &u_max, Vector &du) const override;
   int Type() const override { return 4; }
};

class DofInfo
{
   ParFiniteElementSpace &pfes;
   int bounds_type;
   std::vector<double *> owned;
public:
   Vector xi_min, xi_max;          // per-DOF bounds
   double *xe_min = nullptr, *xe_max = nullptr;   // per-element min/max (device, NE each)
   std::vector<int32_t> BdrDofs, Sub2Ind, NbrDof;
   int numBdrs = 0, numFaceDofs = 0, numSubcells = 0, numDofsSubcell = 0;
   DofInfo(ParFiniteElementSpace &space, int btype);
   ~DofInfo();
   void ComputeElementsMinMax(const Vector &u, double *u_min, double *u_max) const;
   void ComputeBounds(const double *el_min, const double *el_max, Vector &dof_min,
                      Vector &dof_max) const;
};

class LimitedTimeDependentOperator
{
protected:
   double dt = 0.0, t = 0.0;
public:
   virtual ~LimitedTimeDependentOperator() {}
   virtual void SetDt(double dt_) { dt = dt_; }
   double GetDt() const { return dt; }
   virtual void SetTime(double t_) { t = t_; }
   double GetTime() const { return t; }
   virtual void MultUnlimited(const Vector &x, Vector &y) const = 0;
   virtual void LimitMult(const Vector &x, Vector &y) const = 0;
   virtual void Mult(const Vector &x, Vector &y) const { MultUnlimited(x, y); LimitMult(x, y); }
};

class AdvectionOperator : public LimitedTimeDependentOperator
{
   ParFiniteElementSpace &pfes;
   Vector &lumpedM;
   DofInfo &dofs;
   HOSolver *ho_solver;
   LOSolver *lo_solver;
   FCTSolver *fct_solver;
   MonolithicSolver *mono_solver;
public:
   mutable TimingData timer;
   bool verify_bounds = false;
   AdvectionOperator(ParFiniteElementSpace &space, Vector &lumpedM_, DofInfo &dofs_, HOSolver *hos,
                     LOSolver *los, FCTSolver *fct, MonolithicSolver *mos = nullptr);
   void SetDt(double dt_) override;
   void SetTime(double t_) override;
   void MultUnlimited(const Vector &x, Vector &y) const override;
   void LimitMult(const Vector &x, Vector &y) const override;
   void Mult(const Vector &x, Vector &y) const override;
   int HOType() const { return ho_solver ? ho_solver->Type() : 0; }
   int LOType() const { return lo_solver ? lo_solver->Type() : 0; }
   int FCTType() const { return fct_solver ? fct_solver->Type() : 0; }
   ParFiniteElementSpace &Space() const { return pfes; }
   void PrintTimingData(int steps, double stage_seconds) const;
};

// ODESolver::Init / Step (remhos.cpp:486-501): type = the -s value
class ODESolver
{
   int type;
   AdvectionOperator *f = nullptr;
public:
   explicit ODESolver(int ode_solver_type) : type(ode_solver_type) {}
   static bool Known(int t);
   void Init(AdvectionOperator &op) { f = &op; }
   void Step(Vector &x, double &t, double &dt);
   int Stages() const;
};

// VisItDataCollection("Remhos", &pmesh) as the driver uses it (remhos.cpp:1034-1043,1323-1328): every
// Save() writes <name>_<cycle>.mfem_root (JSON index) and <name>_<cycle>/{mesh,<field>}.<rank> in MFEM's
// mesh / GridFunction formats, one domain per rank.
class VisItDataCollection
{
   std::string name;
   ParFiniteElementSpace &pfes;
   std::vector<std::pair<std::string, const Vector *>> fields;
   int cycle = 0, precision = 8;
   double time = 0.0;
public:
   VisItDataCollection(const std::string &collection_name, ParFiniteElementSpace &space)
      : name(collection_name), pfes(space) {}
   void SetPrecision(int p) { precision = p; }
   void RegisterField(const std::string &field_name, const Vector *gf) { fields.emplace_back(field_name, gf); }
   void SetCycle(int c) { cycle = c; }
   void SetTime(double t) { time = t; }
   void Save() const;
};

// The driver (remhos.cpp:210): returns 0 ok, 1 bad flags, 3 unknown ODE solver.
int remhos(int argc, char *argv[], double &final_mass_u);

} // namespace remhos

#endif
