// Multi-GPU layer of the RK-stage path: one process (or context) per GPU, the mesh decomposed into
// per-rank element sets with a vertex-adjacent ghost ring (rmh_halo_*, dplan.cpp).
//
// Replaces ParGridFunction::ExchangeFaceNbrData (remhos.cpp:1813; inside K.Mult), the
// GroupCommunicator min/max reduction of DofInfo::ComputeOverlapBounds (remhos_tools.cpp:463-466)
// and the MPI_Allreduce of mass / min / max (remhos.cpp:1073-1076,1403-1415).
//
// Exchange = ONE kernel per stage, k_halo_put: it gathers the face traces of the stage input that
// the peers' ghost faces need (nfd values per face, already in the receiver's face order) and the
// (min,max) pairs of the ring elements, and stores them straight into the peers' windows over
// NVLink (CUDA-IPC mapped, or plain peer access inside one process); the last block publishes the
// stage epoch in every peer's flag word (release, system scope).  No staging buffer, no message,
// no receive-side unpack.  The receiver's stage kernel (k_stage3c) runs its interior elements first
// and polls the flags once per warp before it touches the first shell element; the other stage
// kernels are preceded by k_halo_wait.  Windows are double-buffered by epoch parity: a peer can be
// at most one stage ahead, because its next put needs this rank's put of the stage in between.
// Scalars go through ncclAllReduce (NCCL is loaded with dlopen when first needed).
#ifndef RMH_DIST_CUH
#define RMH_DIST_CUH

#include <dlfcn.h>
#include <unistd.h>

struct rmh_dplan;
extern "C" int rmh_dplan_sizes(const rmh_dplan *p, int64_t *ne, int64_t *ne_ghost, int64_t *n_slots, int32_t *n_peers);
extern "C" int rmh_dplan_slot_ghosts(const rmh_dplan *p, int32_t *slot_ghost);
extern "C" int64_t rmh_dplan_blob_bytes(const rmh_dplan *p);
extern "C" int rmh_dplan_export(const rmh_dplan *p, void *blob);
extern "C" int rmh_dplan_connect(rmh_dplan *p, int n_blobs, const void *const *blobs, const int64_t *sizes);
extern "C" int rmh_dplan_peer(const rmh_dplan *p, int k, int32_t *rank, int64_t *n_tr, int64_t *n_mm, int32_t *flag_slot);
extern "C" int rmh_dplan_peer_tables(const rmh_dplan *p, int k, int32_t *tr_src, int32_t *tr_dst, int32_t *mm_src,
                                     int32_t *mm_dst);

// ---- the few NCCL entry points used, resolved at run time (torch ships its own libnccl.so.2; a
// link-time dependency could bind to a different copy than the one already in the process)
typedef struct ncclComm *rmh_ncclComm_t;
struct rmh_nccl_id { char internal[128]; };
struct NcclApi
{
   void *lib = nullptr;
   int (*GetUniqueId)(rmh_nccl_id *) = nullptr;
   int (*CommInitRank)(rmh_ncclComm_t *, int, rmh_nccl_id, int) = nullptr;
   int (*AllReduce)(const void *, void *, size_t, int, int, rmh_ncclComm_t, cudaStream_t) = nullptr;
   int (*CommDestroy)(rmh_ncclComm_t) = nullptr;
   const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi *nccl_api()
{
   static NcclApi api;
   static bool tried = false;
   if (!tried)
   {
      tried = true;
      void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
      if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); }
      if (!h) { h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL); }
      if (h)
      {
         api.lib = h;
         api.GetUniqueId = (int (*)(rmh_nccl_id *))dlsym(h, "ncclGetUniqueId");
         api.CommInitRank = (int (*)(rmh_ncclComm_t *, int, rmh_nccl_id, int))dlsym(h, "ncclCommInitRank");
         api.AllReduce = (int (*)(const void *, void *, size_t, int, int, rmh_ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
         api.CommDestroy = (int (*)(rmh_ncclComm_t))dlsym(h, "ncclCommDestroy");
         api.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
         if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) { api.lib = nullptr; }
      }
   }
   return api.lib ? &api : nullptr;
}

// ---- device side
struct PutArgs
{
   int npeers;
   const int32_t *tr_src, *tr_dst, *mm_src, *mm_dst;
   const double *y;
   const double2 *mm_in;                  // owned (min,max) pairs, or null: xe_min / xe_max
   const double *xe_min, *xe_max;
   unsigned int *counter;
   unsigned long long epoch;
   int par;
   PutPeer peer[RMH_MAX_PEERS];
};

__global__ void __launch_bounds__(256) k_halo_put(const __grid_constant__ PutArgs a)
{
   const PutPeer &P = a.peer[blockIdx.y];
   double *gtr = P.gtr[a.par];
   double2 *mm = P.mm[a.par];
   const int64_t stride = (int64_t)gridDim.x * blockDim.x;
   const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   // traces: four independent gather -> remote-store chains per thread and round
   {
      const int32_t *src = a.tr_src + P.tr_off, *dst = a.tr_dst + P.tr_off;
      int64_t i = i0;
      for (; i + 3 * stride < P.tr_n; i += 4 * stride)
      {
         const int32_t s0 = src[i], s1 = src[i + stride], s2 = src[i + 2 * stride], s3 = src[i + 3 * stride];
         const int32_t d0 = dst[i], d1 = dst[i + stride], d2 = dst[i + 2 * stride], d3 = dst[i + 3 * stride];
         const double v0 = a.y[s0], v1 = a.y[s1], v2 = a.y[s2], v3 = a.y[s3];
         gtr[d0] = v0; gtr[d1] = v1; gtr[d2] = v2; gtr[d3] = v3;
      }
      for (; i < P.tr_n; i += stride) { gtr[dst[i]] = a.y[src[i]]; }
   }
   for (int64_t i = i0; i < P.mm_n; i += stride)
   {
      const int64_t k = P.mm_off + i;
      const int32_t e = a.mm_src[k];
      mm[a.mm_dst[k]] = a.mm_in ? a.mm_in[e] : make_double2(a.xe_min[e], a.xe_max[e]);
   }
   // publish: the block barrier orders every thread's stores before thread 0's system-scope fence
   // (cumulativity, as in a grid-wide barrier); the last block to arrive stores the epoch flags
   __syncthreads();
   if (threadIdx.x == 0)
   {
      __threadfence_system();
      const unsigned int total = gridDim.x * gridDim.y;
      const unsigned int t = atomicAdd(a.counter, 1u);
      if (t == total - 1)
      {
         *a.counter = 0;
         __threadfence_system();
         for (int p = 0; p < a.npeers; p++)
         {
            asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(a.peer[p].flag), "l"(a.epoch) : "memory");
         }
      }
   }
}

// wait for the peers' epoch, then unpack the ghost (min,max) pairs for the kernels that read two arrays
__global__ void __launch_bounds__(256) k_halo_wait(const unsigned long long *flags, int n, unsigned long long epoch,
                                                   int64_t n_ghost, const double2 *ghost_mm, double *xe_min_g,
                                                   double *xe_max_g)
{
   wait_peer_flags(flags, n, epoch);
   for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ghost; i += (int64_t)gridDim.x * blockDim.x)
   {
      const double2 v = __ldcg(ghost_mm + i);
      xe_min_g[i] = v.x; xe_max_g[i] = v.y;
   }
}

// ---- host side
struct DistBlobHead
{
   uint64_t magic;
   int32_t rank, world, pid, device;
   uint64_t win_ptr, win_bytes, off_mm[2], off_tr[2];
   int64_t host_id;
   cudaIpcMemHandle_t ipc;
   rmh_nccl_id nccl_id;
   int32_t has_nccl_id, pad;
   int64_t plan_bytes;
};
static const uint64_t DIST_MAGIC = 0x524d484449535431ull;   // "RMHDIST1"

struct rmh_dist
{
   rmh_ctx *c = nullptr;
   rmh_dplan *plan = nullptr;
   int rank = 0, world = 1, npeers = 0;
   bool connected = false;
   std::vector<int32_t> peer_rank;
   std::vector<void *> peer_win;          // mapped windows of the peers
   std::vector<char> peer_ipc;            // 1: opened with cudaIpcOpenMemHandle
   PutArgs put;
   int64_t put_items_max = 0;
   int32_t *d_tr_src = nullptr, *d_tr_dst = nullptr, *d_mm_src = nullptr, *d_mm_dst = nullptr;
   unsigned int *d_counter = nullptr;
   double *d_red = nullptr;
   // fused send of k_stage3c (tables in device memory, see StageSend)
   StageSend *d_send = nullptr;
   unsigned int *d_send_counter = nullptr;
   unsigned int send_groups = 0;
   bool send_ready = false;
   rmh_nccl_id nccl_id;
   bool have_nccl_id = false;
   rmh_ncclComm_t comm = nullptr;
};

// RMH_DEBUG_HALO (timing experiments only, results are WRONG): 1 = no put and no wait, 2 = put, no wait
static int dist_debug_halo()
{
   static int v = -1;
   if (v < 0)
   {
      const char *e = getenv("RMH_DEBUG_HALO");
      v = e ? atoi(e) : 0;
      if (v) { fprintf(stderr, "remhos_b200: RMH_DEBUG_HALO=%d -- the halo exchange is switched OFF, results are wrong (timing experiments only)\n", v); }
   }
   return v;
}

static void dist_stage_args(rmh_ctx *c, rmh::StagePArgs &pa, bool in_kernel_wait)
{
   rmh_dist *d = c->dist;
   if (!d || !d->connected || !in_kernel_wait || d->npeers == 0) { return; }
   if (dist_debug_halo()) { return; }
   pa.flags = c->flags; pa.n_wait = d->npeers; pa.epoch = c->epoch;
   pa.shell_begin = (c->n_split >= 0) ? c->n_split : 0;
   if (c->send_next && d->send_ready)
   {
      pa.send = d->d_send; pa.send_par = (int)((c->epoch + 1) & 1); pa.send_epoch = c->epoch + 1;
      pa.send_groups = d->send_groups;
   }
}

static int64_t host_identity()
{
   char name[256] = {0};
   gethostname(name, sizeof(name) - 1);
   int64_t h = 1469598103934665603ll;
   for (const char *p = name; *p; p++) { h = (h ^ *p) * 1099511628211ll; }
   return h;
}

extern "C" int rmh_dist_create(rmh_ctx *c, rmh_dplan *plan, int rank, int world, int64_t n_interior,
                               rmh_dist **out)
{
   if (!c || !plan || !out) { set_error("rmh_dist_create: null argument"); return 1; }
   int64_t ne = 0, ng = 0, ns = 0;
   int32_t np = 0;
   rmh_dplan_sizes(plan, &ne, &ng, &ns, &np);
   if (ne != c->ne || ng != c->ne_ghost || ns != c->n_gslots)
   { set_error("rmh_dist_create: plan and context describe different decompositions"); return 1; }
   {
      std::vector<int32_t> sg((size_t)ns);
      rmh_dplan_slot_ghosts(plan, sg.data());
      if (sg != c->gs_ghost) { set_error("rmh_dist_create: ghost-face slots of plan and context differ"); return 1; }
   }
   if (np > RMH_MAX_PEERS) { set_error("rmh_dist_create: too many peers"); return 1; }
   if (n_interior < 0 || n_interior > c->ne) { set_error("rmh_dist_create: n_interior out of range"); return 1; }
   CUDA_OK(cudaSetDevice(c->device));
   rmh_dist *d = new rmh_dist;
   d->c = c; d->plan = plan; d->rank = rank; d->world = world; d->npeers = np;
   c->n_split = (n_interior / 8) * 8;     // whole warp groups (1, 2 or 8 elements per warp)
   if (dev_alloc(c, &d->d_counter, 1)) { delete d; return 1; }
   CUDA_OK(cudaMemset(d->d_counter, 0, sizeof(unsigned int)));
   if (dev_alloc(c, &d->d_red, 16)) { delete d; return 1; }
   if (rank == 0 && world > 1)
   {
      NcclApi *na = nccl_api();
      if (na && na->GetUniqueId(&d->nccl_id) == 0) { d->have_nccl_id = true; }
   }
   *out = d;
   return 0;
}

extern "C" int64_t rmh_dist_blob_bytes(const rmh_dist *d)
{
   return (int64_t)sizeof(DistBlobHead) + rmh_dplan_blob_bytes(d->plan);
}

extern "C" int rmh_dist_export(rmh_dist *d, void *blob)
{
   rmh_ctx *c = d->c;
   CUDA_OK(cudaSetDevice(c->device));
   DistBlobHead h;
   memset(&h, 0, sizeof(h));
   h.magic = DIST_MAGIC; h.rank = d->rank; h.world = d->world; h.pid = (int32_t)getpid(); h.device = c->device;
   h.win_ptr = (uint64_t)(uintptr_t)c->win; h.win_bytes = c->win_bytes;
   for (int k = 0; k < 2; k++) { h.off_mm[k] = c->win_off_mm[k]; h.off_tr[k] = c->win_off_tr[k]; }
   h.host_id = host_identity();
   CUDA_OK(cudaIpcGetMemHandle(&h.ipc, c->win));
   h.has_nccl_id = d->have_nccl_id ? 1 : 0;
   if (d->have_nccl_id) { h.nccl_id = d->nccl_id; }
   h.plan_bytes = rmh_dplan_blob_bytes(d->plan);
   memcpy(blob, &h, sizeof(h));
   return rmh_dplan_export(d->plan, (char *)blob + sizeof(h));
}

extern "C" int rmh_dist_connect(rmh_dist *d, int n_blobs, const void *const *blobs, const int64_t *sizes)
{
   rmh_ctx *c = d->c;
   if (n_blobs != d->world) { set_error("rmh_dist_connect: need one blob per rank"); return 1; }
   CUDA_OK(cudaSetDevice(c->device));
   std::vector<DistBlobHead> heads((size_t)n_blobs);
   std::vector<const void *> pb((size_t)n_blobs);
   std::vector<int64_t> ps((size_t)n_blobs);
   for (int r = 0; r < n_blobs; r++)
   {
      if (sizes[r] < (int64_t)sizeof(DistBlobHead)) { set_error("rmh_dist_connect: short blob"); return 1; }
      memcpy(&heads[r], blobs[r], sizeof(DistBlobHead));
      if (heads[r].magic != DIST_MAGIC || heads[r].rank != r) { set_error("rmh_dist_connect: bad blob"); return 1; }
      pb[r] = (const char *)blobs[r] + sizeof(DistBlobHead);
      ps[r] = sizes[r] - (int64_t)sizeof(DistBlobHead);
   }
   if (rmh_dplan_connect(d->plan, n_blobs, pb.data(), ps.data())) { return 1; }
   if (heads[0].has_nccl_id) { d->nccl_id = heads[0].nccl_id; d->have_nccl_id = true; }
   // ---- map the peers' windows, concatenate the tables
   const int np = d->npeers;
   d->peer_rank.resize(np); d->peer_win.assign(np, nullptr); d->peer_ipc.assign(np, 0);
   std::vector<int32_t> tr_src, tr_dst, mm_src, mm_dst;
   memset(&d->put, 0, sizeof(d->put));
   const int64_t my_host = host_identity();
   for (int k = 0; k < np; k++)
   {
      int32_t pr = -1, fslot = -1;
      int64_t ntr = 0, nmm = 0;
      if (rmh_dplan_peer(d->plan, k, &pr, &ntr, &nmm, &fslot)) { return 1; }
      d->peer_rank[k] = pr;
      const DistBlobHead &H = heads[pr];
      if (H.host_id != my_host) { set_error("rmh_dist_connect: peers must be GPUs of one node (NVLink peer memory)"); return 1; }
      void *w = nullptr;
      if (H.pid == (int32_t)getpid())
      {
         w = (void *)(uintptr_t)H.win_ptr;
         if (H.device != c->device)
         {
            int can = 0;
            CUDA_OK(cudaDeviceCanAccessPeer(&can, c->device, H.device));
            if (!can) { set_error("rmh_dist_connect: no peer access between the GPUs"); return 1; }
            cudaError_t e = cudaDeviceEnablePeerAccess(H.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            { set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return 1; }
            cudaGetLastError();
         }
      }
      else
      {
         cudaError_t e = cudaIpcOpenMemHandle(&w, H.ipc, cudaIpcMemLazyEnablePeerAccess);
         if (e != cudaSuccess)
         { set_error(std::string("cudaIpcOpenMemHandle (peer window): ") + cudaGetErrorString(e)); return 1; }
         d->peer_ipc[k] = 1;
      }
      d->peer_win[k] = w;
      PutPeer &P = d->put.peer[k];
      for (int q = 0; q < 2; q++)
      {
         P.gtr[q] = (double *)((char *)w + H.off_tr[q]);
         P.mm[q] = (double2 *)((char *)w + H.off_mm[q]);
      }
      if (fslot < 0 || fslot >= RMH_MAX_PEERS) { set_error("rmh_dist_connect: flag slot out of range"); return 1; }
      P.flag = (unsigned long long *)w + fslot;
      P.tr_off = (int64_t)tr_src.size(); P.tr_n = ntr; P.mm_off = (int64_t)mm_src.size(); P.mm_n = nmm;
      tr_src.resize(tr_src.size() + ntr); tr_dst.resize(tr_dst.size() + ntr);
      mm_src.resize(mm_src.size() + nmm); mm_dst.resize(mm_dst.size() + nmm);
      if (rmh_dplan_peer_tables(d->plan, k, tr_src.data() + P.tr_off, tr_dst.data() + P.tr_off,
                                mm_src.data() + P.mm_off, mm_dst.data() + P.mm_off)) { return 1; }
      d->put_items_max = std::max(d->put_items_max, ntr + nmm);
   }
   if (dev_upload(c, &d->d_tr_src, tr_src.data(), tr_src.size())) { return 1; }
   if (dev_upload(c, &d->d_tr_dst, tr_dst.data(), tr_dst.size())) { return 1; }
   if (dev_upload(c, &d->d_mm_src, mm_src.data(), mm_src.size())) { return 1; }
   if (dev_upload(c, &d->d_mm_dst, mm_dst.data(), mm_dst.size())) { return 1; }
   d->put.npeers = np;
   d->put.tr_src = d->d_tr_src; d->put.tr_dst = d->d_tr_dst; d->put.mm_src = d->d_mm_src; d->put.mm_dst = d->d_mm_dst;
   d->put.counter = d->d_counter;
   // ---- fused send tables: who reads which face of which shell element
   {
      // opt-in (RMH_FUSED_SEND=1): measured on 2 B200s the kernel loses to the rotation and the remote stores
      // what the put launch costs (profiles/r02/README.md), so the stand-alone put kernel is the default
      const char *nf = getenv("RMH_FUSED_SEND");
      const int ND = c->ND, NFD = c->NFD, NF = c->NF, D1 = c->D1, dim = c->dim, pdeg = c->p;
      const int64_t sb = c->n_split, n_shell = c->ne - sb;
      if (c->fold && dim == 3 && np > 0 && n_shell > 0 && (nf && nf[0] == '1'))
      {
         std::vector<int2> face((size_t)n_shell * NF, make_int2(-1, 0));
         std::vector<uint8_t> perm((size_t)n_shell * NF, 0);
         std::vector<int16_t> rperm;
         std::map<std::vector<int16_t>, int> perms;
         std::vector<std::vector<int2>> mm_of((size_t)n_shell);
         bool ok = true;
         for (int k = 0; k < np && ok; k++)
         {
            const PutPeer &P = d->put.peer[k];
            for (int64_t r0 = 0; r0 + NFD <= P.tr_n && ok; r0 += NFD)
            {
               const int32_t *src = tr_src.data() + P.tr_off + r0, *dst = tr_dst.data() + P.tr_off + r0;
               const int64_t le = src[0] / ND;
               const int32_t slot = dst[0] / NFD;
               if (le < sb) { ok = false; break; }
               // my face: the one every requested dof lies on
               int fs = -1;
               for (int f = 0; f < NF && fs < 0; f++)
               {
                  int axis, side;
                  face_axis(dim, f, axis, side);
                  bool all = true;
                  for (int j = 0; j < NFD && all; j++)
                  {
                     int m = src[j] % ND, l[3] = {0, 0, 0};
                     for (int a = 0; a < dim; a++) { l[a] = m % D1; m /= D1; }
                     all = (src[j] / ND == le) && (l[axis] == side * pdeg);
                  }
                  if (all) { fs = f; }
               }
               if (fs < 0) { ok = false; break; }
               int axis, side;
               face_axis(dim, fs, axis, side);
               std::vector<int16_t> key((size_t)NFD, 0);
               for (int j = 0; j < NFD; j++)          // j = the reader's face index
               {
                  int m = src[j] % ND, l[3] = {0, 0, 0}, nat = 0, mul = 1;
                  for (int a = 0; a < dim; a++) { l[a] = m % D1; m /= D1; }
                  for (int a = 0; a < dim; a++) { if (a != axis) { nat += l[a] * mul; mul *= D1; } }
                  if (dst[j] != slot * NFD + j) { ok = false; }
                  key[nat] = (int16_t)j;               // my natural index -> reader's index
               }
               auto it = perms.find(key);
               int id;
               if (it == perms.end())
               {
                  id = (int)perms.size();
                  if (id > 255) { ok = false; break; }
                  perms[key] = id;
                  rperm.insert(rperm.end(), key.begin(), key.end());
               }
               else { id = it->second; }
               face[(size_t)(le - sb) * NF + fs] = make_int2(k, slot);
               perm[(size_t)(le - sb) * NF + fs] = (uint8_t)id;
            }
            for (int64_t i = 0; i < P.mm_n && ok; i++)
            {
               const int64_t e = mm_src[P.mm_off + i];
               if (e < sb) { ok = false; break; }
               mm_of[(size_t)(e - sb)].push_back(make_int2(k, mm_dst[P.mm_off + i]));
            }
         }
         if (ok)
         {
            std::vector<int32_t> mm_off((size_t)n_shell + 1, 0);
            std::vector<int2> mm_flat;
            for (int64_t e = 0; e < n_shell; e++)
            {
               mm_off[e] = (int32_t)mm_flat.size();
               mm_flat.insert(mm_flat.end(), mm_of[e].begin(), mm_of[e].end());
            }
            mm_off[n_shell] = (int32_t)mm_flat.size();
            if (rperm.empty()) { rperm.assign(NFD, 0); }
            if (mm_flat.empty()) { mm_flat.push_back(make_int2(0, 0)); }
            PutPeer *d_peers = nullptr;
            int2 *d_face = nullptr, *d_mm = nullptr;
            uint8_t *d_perm = nullptr;
            int16_t *d_rperm = nullptr;
            int32_t *d_mm_off = nullptr;
            if (dev_upload(c, &d_peers, d->put.peer, (size_t)np) || dev_upload(c, &d_face, face.data(), face.size()) ||
                dev_upload(c, &d_perm, perm.data(), perm.size()) || dev_upload(c, &d_rperm, rperm.data(), rperm.size()) ||
                dev_upload(c, &d_mm_off, mm_off.data(), mm_off.size()) || dev_upload(c, &d_mm, mm_flat.data(), mm_flat.size()) ||
                dev_alloc(c, &d->d_send_counter, 1)) { return 1; }
            CUDA_OK(cudaMemset(d->d_send_counter, 0, sizeof(unsigned int)));
            StageSend hs;
            hs.peer = d_peers; hs.npeers = np; hs.face = d_face; hs.perm = d_perm; hs.rperm = d_rperm;
            hs.mm_off = d_mm_off; hs.mm = d_mm; hs.counter = d->d_send_counter;
            if (dev_upload(c, &d->d_send, &hs, 1)) { return 1; }
            // groups of the kernel: E elements per warp iteration (SmemC::LG)
            const int NL = D1 * D1, LG = NL <= 4 ? 4 : (NL <= 8 ? 8 : (NL <= 16 ? 16 : 32)), E = 32 / LG;
            const int64_t NG = (c->ne + E - 1) / E;
            d->send_groups = (unsigned int)(NG - sb / E);
            d->send_ready = true;
         }
      }
   }
   d->connected = true;
   c->dist = d;
   return 0;
}

extern "C" int rmh_dist_destroy(rmh_dist *d)
{
   if (!d) { return 0; }
   rmh_ctx *c = d->c;
   cudaSetDevice(c->device);
   cudaDeviceSynchronize();
   for (size_t k = 0; k < d->peer_win.size(); k++) { if (d->peer_ipc[k] && d->peer_win[k]) { cudaIpcCloseMemHandle(d->peer_win[k]); } }
   if (d->comm) { NcclApi *na = nccl_api(); if (na) { na->CommDestroy(d->comm); } }
   if (c->dist == d) { c->dist = nullptr; }
   delete d;
   return 0;
}

// true when stage_impl will run k_stage3c<FOLD>, which waits for the halo itself (every other kernel --
// and the entity pass in front of it -- reads ghost (min,max) from xe_min / xe_max: k_halo_wait first)
static bool dist_in_kernel_wait(const rmh_ctx *c, const double *x0, const double *y, const double *out)
{
   return c->fold && c->pipelined && c->dim == 3 && c->all_affine && c->frag && c->op_const && c->npat <= 16 &&
          ((((uintptr_t)y | (uintptr_t)x0 | (uintptr_t)out) & 15) == 0);
}

static int dist_put(rmh_dist *d, const double *y, unsigned long long ep, bool pairs, cudaStream_t s);
static int dist_wait(rmh_dist *d, unsigned long long ep, bool unpack, cudaStream_t s);

// One RK stage on the decomposed mesh: out = a x0 + b (y + dt F(y)); the element min/max of y must be
// in the context (rmh_stage_minmax, or left there by the previous stage)
static int dist_rk_stage(rmh_dist *d, int lo_type, double dt, double a, double b, const double *x0,
                         const double *y, double *out, void *stream, bool send_out)
{
   rmh_ctx *c = d->c;
   if (!d->connected) { set_error("rmh_dist_rk_stage: not connected"); return 1; }
   cudaStream_t s = (cudaStream_t)stream;
   const bool ikw = dist_in_kernel_wait(c, x0, y, out);
   const bool have_halo = (c->sent_ptr == y && c->sent_epoch == c->epoch + 1);
   if (!have_halo && c->sent_epoch == c->epoch + 1) { c->epoch += 2; }    // a stale early send owns that epoch
   const unsigned long long ep = c->epoch + 1;
   if (d->npeers > 0)
   {
      if (!have_halo && dist_debug_halo() != 1) { if (dist_put(d, y, ep, c->fold, s)) { return 1; } }
      if (!ikw) { if (dist_wait(d, ep, !c->fold, s)) { return 1; } }
   }
   // the output of this stage is the next stage's input: let the kernel send its halo (shell groups first)
   c->send_next = send_out && ikw && d->send_ready && d->npeers > 0;
   const int rc = stage_impl(c, lo_type, dt, 1, a, b, x0, y, out, true, c->bounds_type == 0, s);
   if (c->send_next && rc == 0) { c->sent_ptr = out; c->sent_epoch = c->epoch + 1; }
   c->send_next = false;
   return rc;
}

static int dist_put(rmh_dist *d, const double *y, unsigned long long ep, bool pairs, cudaStream_t s)
{
   rmh_ctx *c = d->c;
   const int par = (int)(ep & 1);
   PutArgs &A = d->put;
   A.y = y; A.epoch = ep; A.par = par;
   A.mm_in = pairs ? c->xe_mm2[par] : nullptr;
   A.xe_min = c->xe_min; A.xe_max = c->xe_max;
   const int bs = 256;
   // about two blocks per SM over all peers: one wave, a handful of items per thread
   const int64_t nb = std::max<int64_t>(1, std::min<int64_t>((d->put_items_max + 4 * bs - 1) / (4 * bs),
                                                           std::max<int64_t>(1, 2 * (int64_t)c->num_sms / d->npeers)));
   k_halo_put<<<dim3((unsigned)nb, (unsigned)d->npeers), bs, 0, s>>>(A);
   LAUNCH_OK();
   return 0;
}

static int dist_wait(rmh_dist *d, unsigned long long ep, bool unpack, cudaStream_t s)
{
   rmh_ctx *c = d->c;
   const int bs = 256, par = (int)(ep & 1);
   const int64_t nw = std::max<int64_t>(1, std::min<int64_t>((c->ne_ghost + bs - 1) / bs, 2 * (int64_t)c->num_sms));
   k_halo_wait<<<(unsigned)nw, bs, 0, s>>>(c->flags, d->npeers, ep, unpack ? c->ne_ghost : 0, c->xe_mm2[par] + c->ne,
                                           c->xe_min + c->ne, c->xe_max + c->ne);
   LAUNCH_OK();
   return 0;
}

// ParGridFunction::ExchangeFaceNbrData + the shared min/max of DofInfo for the unfused solver path
// (remhos.cpp:1813, remhos_tools.cpp:399,463-466): every solver entry point that reads neighbour traces
// (HO, DU / RD face terms, Neumann) or ghost element bounds finds them in the window afterwards
static int dist_halo(rmh_ctx *c, const double *u, cudaStream_t s)
{
   rmh_dist *d = c->dist;
   if (!d || !d->connected) { return 0; }
   c->xe_ptr = nullptr;
   launch_elem_min_max(c->ne, c->ND, u, c->xe_min, c->xe_max, s);
   LAUNCH_OK();
   if (c->sent_epoch == c->epoch + 1) { c->epoch += 2; }     // an unused early send owns that epoch (dist_rk_stage)
   c->sent_ptr = nullptr;
   const unsigned long long ep = ++c->epoch;
   if (d->npeers > 0)
   {
      if (dist_put(d, u, ep, false, s)) { return 1; }
      if (dist_wait(d, ep, true, s)) { return 1; }
   }
   c->ughost = c->gtr[ep & 1];
   c->halo_ptr = u;
   return 0;
}

static int dist_traces(rmh_ctx *c, const double *vec, double *dst, cudaStream_t s)
{
   rmh_dist *d = c->dist;
   if (!d || !d->connected) { set_error("dist_traces: no connected rmh_dist"); return 1; }
   if (c->sent_epoch == c->epoch + 1) { c->epoch += 2; }
   c->sent_ptr = nullptr;
   const unsigned long long ep = ++c->epoch;
   if (d->npeers > 0)
   {
      if (dist_put(d, vec, ep, false, s)) { return 1; }
      if (dist_wait(d, ep, false, s)) { return 1; }
   }
   CUDA_OK(cudaMemcpyAsync(dst, c->gtr[ep & 1], (size_t)c->n_gslots * c->NFD * sizeof(double), cudaMemcpyDeviceToDevice, s));
   c->ughost = c->gtr[ep & 1];
   c->halo_ptr = nullptr;
   return 0;
}

extern "C" int rmh_dist_rk_stage(rmh_dist *d, int lo_type, double dt, double a, double b, const double *x0,
                                 const double *y, double *out, void *stream)
{
   d->c->sent_ptr = nullptr;                 // y may have been written by the caller since
   return dist_rk_stage(d, lo_type, dt, a, b, x0, y, out, stream, false);
}

// ODESolver::Step for -s 1/2/3 on the decomposed mesh (cf. rmh_rk_step)
extern "C" int rmh_dist_rk_step(rmh_dist *d, int ode, int lo_type, double *t, double dt, double *u, void *stream)
{
   rmh_ctx *c = d->c;
   cudaStream_t s = (cudaStream_t)stream;
   const bool chain = (c->bounds_type == 0);
   const bool have_xe = chain && c->trust_state && c->xe_ptr == u;
   const bool keep_xe = chain && c->trust_state;
   if (!have_xe) { c->sent_ptr = nullptr; if (stage_minmax(c, u, s)) { return 1; } }
   c->xe_ptr = nullptr;
   // the output of a stage goes out from inside its kernel when the next stage reads it (w1, w2; u if it is trusted)
   const bool snd = chain && c->exec_mode != 1;
   // remap: every stage re-assembles at its own time (rmh_set_time; a no-op in transport mode)
   const double t0 = *t;
   if (ode == 1)
   {
      if (rmh_set_time(c, t0, stream)) { return 1; }
      if (dist_rk_stage(d, lo_type, dt, 0.0, 1.0, u, u, c->w1, stream, false)) { return 1; }
      CUDA_OK(cudaMemcpyAsync(u, c->w1, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToDevice, s));
      if (keep_xe) { c->xe_ptr = u; }
   }
   else if (ode == 2)
   {
      if (rmh_set_time(c, t0, stream)) { return 1; }
      if (dist_rk_stage(d, lo_type, dt, 0.0, 1.0, u, u, c->w1, stream, snd)) { return 1; }
      if (!chain) { if (stage_minmax(c, c->w1, s)) { return 1; } }
      if (rmh_set_time(c, t0 + dt, stream)) { return 1; }
      if (dist_rk_stage(d, lo_type, dt, 0.5, 0.5, u, c->w1, u, stream, snd && keep_xe)) { return 1; }
      if (keep_xe) { c->xe_ptr = u; }
   }
   else if (ode == 3)
   {
      if (rmh_set_time(c, t0, stream)) { return 1; }
      if (dist_rk_stage(d, lo_type, dt, 0.0, 1.0, u, u, c->w1, stream, snd)) { return 1; }
      if (!chain) { if (stage_minmax(c, c->w1, s)) { return 1; } }
      if (rmh_set_time(c, t0 + dt, stream)) { return 1; }
      if (dist_rk_stage(d, lo_type, dt, 0.75, 0.25, u, c->w1, c->w2, stream, snd)) { return 1; }
      if (!chain) { if (stage_minmax(c, c->w2, s)) { return 1; } }
      if (rmh_set_time(c, t0 + dt / 2, stream)) { return 1; }
      if (dist_rk_stage(d, lo_type, dt, 1.0 / 3.0, 2.0 / 3.0, u, c->w2, u, stream, snd && keep_xe)) { return 1; }
      if (keep_xe) { c->xe_ptr = u; }
   }
   else { set_error("rmh_dist_rk_step: ode solver type must be 1, 2 or 3"); return 1; }
   *t += dt;
   return 0;
}

// same with HOST state: H2D of u, one step, D2H (the end-to-end entry point, cf. rmh_rk_step_host)
extern "C" int rmh_dist_rk_step_host(rmh_dist *d, int ode, int lo_type, double *t, double dt, double *u_host)
{
   rmh_ctx *c = d->c;
   const size_t bytes = (size_t)c->N * sizeof(double);
   if (rmh_host_sync(c)) { return 1; }      // queued steps share the stage intermediates with this call
   CUDA_OK(cudaMemcpyAsync(c->w3, u_host, bytes, cudaMemcpyHostToDevice, 0));
   c->xe_ptr = nullptr;
   if (rmh_dist_rk_step(d, ode, lo_type, t, dt, c->w3, nullptr)) { return 1; }
   CUDA_OK(cudaMemcpyAsync(u_host, c->w3, bytes, cudaMemcpyDeviceToHost, 0));
   CUDA_OK(cudaStreamSynchronize(0));
   return 0;
}

// pipelined variant (host_pipe_step in ctx.cu): every rank queues its step; the halo puts and flag waits of
// consecutive calls stay ordered on the compute stream
extern "C" int rmh_dist_rk_step_host_async(rmh_dist *d, int ode, int lo_type, double t, double dt,
                                           const double *u_in_host, double *u_out_host)
{
   return host_pipe_step(d->c, u_in_host, u_out_host, [&](double *u, cudaStream_t s)
   {
      double tt = t;
      return rmh_dist_rk_step(d, ode, lo_type, &tt, dt, u, (void *)s);
   });
}

// in-kernel halo wait statistics of this device since the last reset: out[0] warps that had to wait for a
// peer's flag, out[1] their summed and out[2] longest wait in ns; out[3] warps that reached a shell group
// and out[4] their summed time from there to their end (ns); out[5] all warps of the ghost-aware launches
// and out[6] their summed run time (a diagnostic: which rank waits for whom, what the shell costs)
extern "C" int rmh_halo_wait_stats(rmh_ctx *c, unsigned long long *out, int reset)
{
   CUDA_OK(cudaSetDevice(c->device));
   CUDA_OK(cudaDeviceSynchronize());
   unsigned long long g[9];
   CUDA_OK(cudaMemcpyFromSymbol(g, g_halo_wait, sizeof(g)));
   out[0] = g[0]; out[1] = g[1]; out[2] = g[2];
   // phase clocks: warps that reached a shell group, their summed time behind that point; all warps, summed run time
   out[3] = g[5]; out[4] = g[4] - g[3]; out[5] = g[8]; out[6] = g[7] - g[6];
   if (reset)
   {
      const unsigned long long z[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      CUDA_OK(cudaMemcpyToSymbol(g_halo_wait, z, sizeof(z)));
   }
   return 0;
}

// MPI_Allreduce replacement (remhos.cpp:1073-1076,1403-1415; dt: :538-553): op 0 sum, 1 min, 2 max over the
// ranks, in place on n <= 16 host doubles; collective
extern "C" int rmh_dist_allreduce(rmh_dist *d, int op, double *vals, int n, void *stream)
{
   if (d->world == 1) { return 0; }
   if (n < 1 || n > 16) { set_error("rmh_dist_allreduce: 1 <= n <= 16"); return 1; }
   rmh_ctx *c = d->c;
   CUDA_OK(cudaSetDevice(c->device));
   NcclApi *na = nccl_api();
   if (!na) { set_error("rmh_dist_allreduce: libnccl.so.2 not found"); return 1; }
   if (!d->comm)
   {
      if (!d->have_nccl_id) { set_error("rmh_dist_allreduce: no NCCL id (rank 0 could not create one)"); return 1; }
      const int rc = na->CommInitRank(&d->comm, d->world, d->nccl_id, d->rank);
      if (rc != 0)
      { set_error(std::string("ncclCommInitRank: ") + (na->GetErrorString ? na->GetErrorString(rc) : "error")); return 1; }
   }
   cudaStream_t s = (cudaStream_t)stream;
   CUDA_OK(cudaMemcpyAsync(d->d_red, vals, n * sizeof(double), cudaMemcpyHostToDevice, s));
   const int nccl_double = 8, nccl_op = (op == 0) ? 0 : (op == 1 ? 3 : 2);   // ncclFloat64; ncclSum / ncclMin / ncclMax
   const int rc = na->AllReduce(d->d_red, d->d_red, (size_t)n, nccl_double, nccl_op, d->comm, s);
   if (rc != 0) { set_error(std::string("ncclAllReduce: ") + (na->GetErrorString ? na->GetErrorString(rc) : "error")); return 1; }
   CUDA_OK(cudaMemcpyAsync(vals, d->d_red, n * sizeof(double), cudaMemcpyDeviceToHost, s));
   CUDA_OK(cudaStreamSynchronize(s));
   return 0;
}

#endif
