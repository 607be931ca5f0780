// Host-side exchange plan of the domain-decomposed stage (no CUDA in this file).
//
// Replaces the face-neighbour tables behind ParGridFunction::ExchangeFaceNbrData
// (remhos.cpp:1813; inside K.Mult) and the GroupCommunicator of DofInfo::ComputeOverlapBounds
// (remhos_tools.cpp:463-466).  The reference ships whole DOF blocks of every face-neighbour element
// (remhos_tools.cpp:583,609); here a rank receives, per face with a ghost neighbour, only the nfd
// trace values of that face -- already in the receiving element's natural face order -- plus one
// (min,max) pair per ghost-ring element.
//
// Receiver-driven: every rank numbers its ghost faces ("slots", scan order over (element, face)),
// and publishes, per peer, the list (slot, global id of the ghost element, neighbour-local DOF of
// every face DOF) in a blob; after the blobs have been all-gathered (by the caller: any transport),
// each rank turns the requests addressed to it into gather/scatter tables
//     tr_src (own DOF index) -> tr_dst (index in the peer's ghost trace array)
//     mm_src (own element)   -> mm_dst (index in the peer's (min,max) pair array)
// which the put kernel (dist.cuh) executes with remote stores over NVLink.
#include "../../include/remhos_b200.h"
#include "common.hpp"

#include <algorithm>
#include <cstring>
#include <vector>

using namespace rmh;

namespace
{
struct DPeer
{
   int32_t rank = -1;
   int32_t recv_off = 0;                 // first index of this peer's elements in my ghost list
   // requests this rank makes of the peer
   std::vector<int32_t> req_slot;
   std::vector<int64_t> req_gid;
   std::vector<int16_t> req_loc;         // [n][nfd]
   // what this rank sends to the peer (after connect)
   std::vector<int32_t> tr_src, tr_dst, mm_src, mm_dst;
   int32_t flag_slot = -1;               // position of this rank in the peer's peer list
};

struct BlobHead
{
   uint64_t magic;
   int32_t rank, world, nd, nfd, npeers, pad;
   int64_t ne, ne_ghost, n_slots;
};
struct BlobPeer
{
   int32_t rank, recv_off;
   int64_t n_req;
};
const uint64_t MAGIC = 0x524d4844504c4e31ull;   // "RMHDPLN1"
} // namespace

struct rmh_dplan
{
   int rank = 0, world = 1, nd = 0, nfd = 0;
   int64_t ne = 0, ne_ghost = 0, n_slots = 0;
   std::vector<int32_t> slot_ghost;      // per slot: ghost element (0 .. ne_ghost-1)
   std::vector<DPeer> peers;
   // copies of the halo plan needed at connect
   std::vector<int64_t> owned_sorted, send;
   std::vector<int32_t> owned_pos, send_off;
   bool connected = false;
   int32_t local_of(int64_t g) const
   {
      const auto it = std::lower_bound(owned_sorted.begin(), owned_sorted.end(), g);
      if (it == owned_sorted.end() || *it != g) { return -1; }
      return owned_pos[it - owned_sorted.begin()];
   }
};

extern "C" int rmh_dplan_create(const rmh_halo *h, int rank, int world, int dim, int order,
                                const int32_t *nbr_dof, rmh_dplan **out)
{
   if (!h || !nbr_dof || !out) { set_error("rmh_dplan_create: null argument"); return 1; }
   rmh_dplan *p = new rmh_dplan;
   const int n = order + 1, nf = 2 * dim;
   int nd = 1, nfd = 1;
   for (int a = 0; a < dim; a++) { nd *= n; }
   for (int a = 0; a < dim - 1; a++) { nfd *= n; }
   p->rank = rank; p->world = world; p->nd = nd; p->nfd = nfd;
   p->ne = (int64_t)h->owned.size(); p->ne_ghost = (int64_t)h->ghost.size();
   p->owned_sorted = h->owned_sorted; p->owned_pos = h->owned_pos;
   p->send = h->send; p->send_off = h->send_off;
   p->peers.resize(h->peers.size());
   for (size_t k = 0; k < h->peers.size(); k++)
   {
      p->peers[k].rank = h->peers[k];
      p->peers[k].recv_off = h->recv_off[k];
   }
   std::vector<int> n2r;
   nat2ref_table(order, dim, n2r);
   // ghost element -> peer index (ghosts are ordered by owner)
   auto peer_of_ghost = [&](int64_t g)
   {
      const auto it = std::upper_bound(h->recv_off.begin(), h->recv_off.end(), (int32_t)g);
      return (int)(it - h->recv_off.begin()) - 1;
   };
   for (int64_t e = 0; e < p->ne; e++)
      for (int f = 0; f < nf; f++)
      {
         const int32_t *row = &nbr_dof[((size_t)e * nf + f) * nfd];
         if (row[0] < 0) { continue; }
         const int64_t nb = row[0] / nd;
         if (nb < p->ne) { continue; }
         const int64_t g = nb - p->ne;
         if (g >= p->ne_ghost) { set_error("rmh_dplan_create: neighbour beyond the ghost range"); delete p; return 1; }
         const int k = peer_of_ghost(g);
         if (k < 0 || k >= (int)p->peers.size()) { set_error("rmh_dplan_create: ghost without owner"); delete p; return 1; }
         DPeer &P = p->peers[k];
         P.req_slot.push_back((int32_t)p->n_slots);
         P.req_gid.push_back(h->ghost[g]);
         for (int j = 0; j < nfd; j++)
         {
            const int32_t d = row[n2r[f * nfd + j]];
            if (d < 0 || d / nd != nb) { set_error("rmh_dplan_create: face DOFs must map into one neighbour"); delete p; return 1; }
            P.req_loc.push_back((int16_t)(d - nb * nd));
         }
         p->slot_ghost.push_back((int32_t)g);
         p->n_slots++;
      }
   *out = p;
   return 0;
}

extern "C" int rmh_dplan_free(rmh_dplan *p) { delete p; return 0; }

extern "C" int64_t rmh_dplan_blob_bytes(const rmh_dplan *p)
{
   size_t b = sizeof(BlobHead);
   for (const DPeer &P : p->peers)
   {
      b += sizeof(BlobPeer) + P.req_slot.size() * (sizeof(int32_t) + sizeof(int64_t) + (size_t)p->nfd * sizeof(int16_t));
      b = (b + 7) & ~(size_t)7;
   }
   return (int64_t)b;
}

extern "C" int rmh_dplan_export(const rmh_dplan *p, void *blob)
{
   char *w = (char *)blob;
   BlobHead hd;
   std::memset(&hd, 0, sizeof(hd));
   hd.magic = MAGIC; hd.rank = p->rank; hd.world = p->world; hd.nd = p->nd; hd.nfd = p->nfd;
   hd.npeers = (int32_t)p->peers.size(); hd.ne = p->ne; hd.ne_ghost = p->ne_ghost; hd.n_slots = p->n_slots;
   std::memcpy(w, &hd, sizeof(hd)); w += sizeof(hd);
   for (const DPeer &P : p->peers)
   {
      char *w0 = w;
      BlobPeer bp;
      bp.rank = P.rank; bp.recv_off = P.recv_off; bp.n_req = (int64_t)P.req_slot.size();
      std::memcpy(w, &bp, sizeof(bp)); w += sizeof(bp);
      std::memcpy(w, P.req_gid.data(), P.req_gid.size() * sizeof(int64_t)); w += P.req_gid.size() * sizeof(int64_t);
      std::memcpy(w, P.req_slot.data(), P.req_slot.size() * sizeof(int32_t)); w += P.req_slot.size() * sizeof(int32_t);
      std::memcpy(w, P.req_loc.data(), P.req_loc.size() * sizeof(int16_t)); w += P.req_loc.size() * sizeof(int16_t);
      const size_t used = (size_t)(w - w0), padded = (used + 7) & ~(size_t)7;
      std::memset(w, 0, padded - used); w = w0 + padded;
      // (the head is 8-byte sized, so every peer record starts 8-byte aligned)
   }
   return 0;
}

extern "C" int rmh_dplan_connect(rmh_dplan *p, int n_blobs, const void *const *blobs, const int64_t *sizes)
{
   if (n_blobs != p->world) { set_error("rmh_dplan_connect: need one blob per rank"); return 1; }
   for (DPeer &P : p->peers)
   {
      if (P.rank < 0 || P.rank >= n_blobs) { set_error("rmh_dplan_connect: bad peer rank"); return 1; }
      const char *r = (const char *)blobs[P.rank], *end = r + sizes[P.rank];
      BlobHead hd;
      if (sizes[P.rank] < (int64_t)sizeof(hd)) { set_error("rmh_dplan_connect: short blob"); return 1; }
      std::memcpy(&hd, r, sizeof(hd)); r += sizeof(hd);
      if (hd.magic != MAGIC || hd.rank != P.rank || hd.nd != p->nd || hd.nfd != p->nfd)
      { set_error("rmh_dplan_connect: blob does not belong to this decomposition"); return 1; }
      bool found = false;
      for (int k = 0; k < hd.npeers; k++)
      {
         const char *r0 = r;
         BlobPeer bp;
         if (r + sizeof(bp) > end) { set_error("rmh_dplan_connect: truncated blob"); return 1; }
         std::memcpy(&bp, r, sizeof(bp)); r += sizeof(bp);
         const size_t nreq = (size_t)bp.n_req;
         const char *gid_p = r; r += nreq * sizeof(int64_t);
         const char *slot_p = r; r += nreq * sizeof(int32_t);
         const char *loc_p = r; r += nreq * (size_t)p->nfd * sizeof(int16_t);
         if (r > end) { set_error("rmh_dplan_connect: truncated blob"); return 1; }
         r = r0 + (((size_t)(r - r0) + 7) & ~(size_t)7);
         if (bp.rank != p->rank) { continue; }
         found = true;
         P.flag_slot = k;
         // traces the peer wants from my elements
         P.tr_src.resize(nreq * p->nfd); P.tr_dst.resize(nreq * p->nfd);
         for (size_t i = 0; i < nreq; i++)
         {
            int64_t gid; int32_t slot;
            std::memcpy(&gid, gid_p + i * sizeof(int64_t), sizeof(gid));
            std::memcpy(&slot, slot_p + i * sizeof(int32_t), sizeof(slot));
            const int32_t le = p->local_of(gid);
            if (le < 0) { set_error("rmh_dplan_connect: peer requests an element this rank does not own"); return 1; }
            for (int j = 0; j < p->nfd; j++)
            {
               int16_t loc;
               std::memcpy(&loc, loc_p + (i * p->nfd + j) * sizeof(int16_t), sizeof(loc));
               P.tr_src[i * p->nfd + j] = le * p->nd + loc;
               P.tr_dst[i * p->nfd + j] = slot * p->nfd + j;
            }
         }
         // (min,max) pairs of the ring elements: my send list for this peer and the peer's ghost
         // list of my elements are both ascending in the global id
         size_t kme = 0;
         for (; kme < p->peers.size(); kme++) { if (&p->peers[kme] == &P) { break; } }
         const int32_t s0 = p->send_off[kme], s1 = p->send_off[kme + 1];
         P.mm_src.resize(s1 - s0); P.mm_dst.resize(s1 - s0);
         for (int32_t i = s0; i < s1; i++)
         {
            P.mm_src[i - s0] = p->local_of(p->send[i]);
            P.mm_dst[i - s0] = (int32_t)(hd.ne + bp.recv_off + (i - s0));
         }
      }
      if (!found) { set_error("rmh_dplan_connect: asymmetric peer lists"); return 1; }
   }
   p->connected = true;
   return 0;
}

extern "C" int rmh_dplan_sizes(const rmh_dplan *p, int64_t *ne, int64_t *ne_ghost, int64_t *n_slots, int32_t *n_peers)
{
   if (ne) { *ne = p->ne; }
   if (ne_ghost) { *ne_ghost = p->ne_ghost; }
   if (n_slots) { *n_slots = p->n_slots; }
   if (n_peers) { *n_peers = (int32_t)p->peers.size(); }
   return 0;
}

extern "C" int rmh_dplan_slot_ghosts(const rmh_dplan *p, int32_t *slot_ghost)
{
   std::copy(p->slot_ghost.begin(), p->slot_ghost.end(), slot_ghost);
   return 0;
}

extern "C" int rmh_dplan_peer(const rmh_dplan *p, int k, int32_t *rank, int64_t *n_tr, int64_t *n_mm,
                              int32_t *flag_slot)
{
   if (k < 0 || k >= (int)p->peers.size()) { set_error("rmh_dplan_peer: bad index"); return 1; }
   const DPeer &P = p->peers[k];
   if (rank) { *rank = P.rank; }
   if (n_tr) { *n_tr = (int64_t)P.tr_src.size(); }
   if (n_mm) { *n_mm = (int64_t)P.mm_src.size(); }
   if (flag_slot) { *flag_slot = P.flag_slot; }
   return 0;
}

extern "C" int rmh_dplan_peer_tables(const rmh_dplan *p, int k, int32_t *tr_src, int32_t *tr_dst,
                                     int32_t *mm_src, int32_t *mm_dst)
{
   if (k < 0 || k >= (int)p->peers.size() || !p->connected) { set_error("rmh_dplan_peer_tables: not connected"); return 1; }
   const DPeer &P = p->peers[k];
   if (tr_src) { std::copy(P.tr_src.begin(), P.tr_src.end(), tr_src); }
   if (tr_dst) { std::copy(P.tr_dst.begin(), P.tr_dst.end(), tr_dst); }
   if (mm_src) { std::copy(P.mm_src.begin(), P.mm_src.end(), mm_src); }
   if (mm_dst) { std::copy(P.mm_dst.begin(), P.mm_dst.end(), mm_dst); }
   return 0;
}
