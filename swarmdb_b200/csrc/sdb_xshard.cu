// sdb_xshard.cu - cross-shard delivery of group sends (sm_100a): the receiving half.
//
// Replaces the reference's partitioned Kafka topic (explicit-partition produce M:469-482, one
// consumer per agent over all partitions M:334-345).  Agents are hash-partitioned over the GPUs
// of the box; group tables are replicated, each shard keeping the members it owns and their
// positions in the full list.  A rank exports its batch of group sends ONCE per send
// (descriptor + payload), the caller all-gathers the wire batches over NVLink (NCCL), and
// every shard expands, for every send of every source, the copies for the members it owns:
//
//   k_wire_table      prefix table over the sources (sends, sequence offsets, section offsets)
//   k_wire_measure    per wire send: local member count x record size (scan input)
//   k_bucket_sizes    per (group, source): size of the exporter-built bucket
//   scans             arena offsets of every send's local region; global bucket offsets
//   k_wire_localize   writes the local fan-out descriptor (payload address in the exporter's buffer,
//                     global sequence offset, local member list)
//   k_bucket_fill     concatenates the sources' buckets in rank order (= global send order)
//   then the ordinary fan-out kernel + pull index build run over the localized descriptors.
//
// NVLink carries (64 + payload) bytes per (send, peer) instead of (32 + payload) per recipient:
// 8 x less for 64-way groups on 8 shards.  With the peer-memory transport (sdb_import_wire_ptrs over
// CUDA-IPC mapped export buffers) there is no separate collective at all: these kernels read the
// descriptors, and the fan-out kernel's TMA bulk loads read the payloads, straight out of the
// exporting GPU's memory, so the NVLink transfer overlaps the fan-out send by send and a shard only
// pulls the sends it has recipients for.
#include "sdb_common.cuh"

#define SDB_SCAN_TILE 4096u


__device__ __forceinline__ const sdb_wire_header* wire_hdr(const sdb_import_args& a, uint32_t src) {
  return reinterpret_cast<const sdb_wire_header*>(a.meta[src]);
}

// one thread: read the n_src wire headers (remote when the batch lives in a peer GPU) and publish the prefix table
__global__ void k_wire_table(sdb_import_args a, sdb_src_tab* tab) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t first = 0; uint64_t rb = 0;
  for (uint32_t s = 0; s < a.n_src; ++s) {
    const sdb_wire_header* h = wire_hdr(a, s);
    const bool ok = h->magic == SDB_WIRE_MAGIC;
    const bool ex = ok && h->explicit_seq;
    tab->first[s] = first; tab->rec_base[s] = ex ? h->seq_base : rb; tab->explicit_seq[s] = ex ? 1u : 0u;
    tab->desc_off[s] = ok ? h->desc_off : 0; tab->list_off[s] = ok ? h->list_off : 0; tab->payload_off[s] = ok ? h->payload_off : 0;
    tab->gs_off_off[s] = (ok && h->max_groups == a.max_groups) ? h->gs_off_off : 0; tab->gs_idx_off[s] = ok ? h->gs_idx_off : 0;
    first += ok ? min(h->n_sends, a.max_sends) : 0u;
    rb += ok ? h->total_recs : 0ull;
  }
  for (uint32_t s = a.n_src; s <= SDB_MAX_SRC; ++s) tab->first[s] = first;
}

// global send index -> (source, index inside the source), sources concatenated in rank order
__device__ __forceinline__ bool locate(const sdb_import_args& a, uint32_t gi, uint32_t& src, uint32_t& i,
                                       uint64_t& rec_base) {
  const sdb_src_tab* t = a.tab;
  if (gi >= t->first[a.n_src]) return false;
  uint32_t s = 0;
  while (s + 1 < a.n_src && gi >= t->first[s + 1]) ++s;
  src = s; i = gi - t->first[s]; rec_base = t->rec_base[s];
  return true;
}

__device__ __forceinline__ const sdb_send_desc* wire_desc(const sdb_import_args& a, uint32_t src, uint32_t i) {
  return reinterpret_cast<const sdb_send_desc*>(a.meta[src] + a.tab->desc_off[src]) + i;
}

__device__ __forceinline__ const uint32_t* wire_list(const sdb_import_args& a, uint32_t src) {
  return reinterpret_cast<const uint32_t*>(a.wire[src] + a.tab->list_off[src]);
}

__global__ void __launch_bounds__(256)
k_wire_measure(sdb_import_args a, uint32_t n_total) {
  const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= n_total) return;
  uint32_t src, i; uint64_t rb;
  if (!locate(a, gi, src, i, rb)) { a.w[gi] = 0; a.lw[gi] = 0; return; }
  const uint4* dq = reinterpret_cast<const uint4*>(wire_desc(a, src, i));
  sdb_send_desc dd;                                             // one 64-byte read of the (remote) descriptor
  reinterpret_cast<uint4*>(&dd)[0] = __ldg(dq); reinterpret_cast<uint4*>(&dd)[1] = __ldg(dq + 1);
  reinterpret_cast<uint4*>(&dd)[2] = __ldg(dq + 2); reinterpret_cast<uint4*>(&dd)[3] = __ldg(dq + 3);
  const sdb_send_desc* d = &dd;
  uint32_t lc = 0, own = 0;
  if (d->flags & SDB_DESC_P2P) {                                   // delivered by the receiver's owner only
    own = (d->mstart < a.max_agents && a.shard_of[d->mstart] == a.shard_id) ? 1u : 0u;
    lc = own;
  } else if (d->flags & SDB_DESC_LIST_TEMP) {                      // broadcast: the recipients this shard owns
    const uint32_t* l = wire_list(a, src) + d->mstart;
    for (uint32_t k = 0; k < d->mcount; ++k) { const uint32_t x = l[k]; own += (x < a.max_agents && a.shard_of[x] == a.shard_id); }
    lc = own;
  } else {
    const uint32_t g = d->group;
    lc = g < a.max_groups ? a.lcount[g] : 0u;
  }
  a.w[gi] = lc * d->rgran;
  a.lw[gi] = own;
}

__global__ void __launch_bounds__(256)
k_wire_localize(sdb_import_args a, uint32_t n_total) {
  const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= n_total) return;
  uint32_t src, i; uint64_t rb;
  sdb_send_desc out;
  if (!locate(a, gi, src, i, rb)) {                  // unused capacity: an empty send
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* o = reinterpret_cast<uint4*>(a.descs + gi);
    o[0] = z; o[1] = z; o[2] = z; o[3] = z;
    return;
  }
  {
    const uint4* dq = reinterpret_cast<const uint4*>(wire_desc(a, src, i));
    reinterpret_cast<uint4*>(&out)[0] = __ldg(dq); reinterpret_cast<uint4*>(&out)[1] = __ldg(dq + 1);
    reinterpret_cast<uint4*>(&out)[2] = __ldg(dq + 2); reinterpret_cast<uint4*>(&out)[3] = __ldg(dq + 3);
  }
  const sdb_send_desc wd = out;
  const sdb_send_desc* d = &wd;
  // absolute address (the fan-out kernel is launched with a null payload base): the payload stays where the
  // source rank exported it - possibly in a peer GPU's memory - and is pulled by the fan-out's TMA loads
  out.payload_off = reinterpret_cast<uint64_t>(a.wire[src]) + a.tab->payload_off[src] + d->payload_off;
  out.gran0 = a.w_local[gi] + a.w_tops[gi / SDB_SCAN_TILE];
  out.rec0 = static_cast<uint32_t>(rb + d->rec0);
  const bool abs_seq = a.tab->explicit_seq[src] != 0;
  out.seq_abs = abs_seq ? rb + d->rec0 : 0ull;
  if (d->flags & (SDB_DESC_P2P | SDB_DESC_LIST_TEMP)) {
    const uint32_t own = a.lw[gi];
    uint32_t lo = a.lw_local[gi] + a.lw_tops[gi / SDB_SCAN_TILE];
    const bool fits = static_cast<uint64_t>(lo) + own <= a.list_cap;      // checked on the host before launch
    out.mstart = lo; out.mcount = fits ? own : 0u; out.group = SDB_NO_GROUP;
    if (d->flags & SDB_DESC_P2P) {
      out.flags = SDB_DESC_LIST_TEMP | (abs_seq ? SDB_DESC_ABS_SEQ : 0u);
      if (own && fits) a.tmp_list[lo] = d->mstart;
    } else {
      out.flags = SDB_DESC_LIST_TEMP | SDB_DESC_SHARED_SEQ | (abs_seq ? SDB_DESC_ABS_SEQ : 0u);
      if (fits) {
        const uint32_t* l = wire_list(a, src) + d->mstart;
        for (uint32_t k = 0; k < d->mcount; ++k) { const uint32_t x = l[k]; if (x < a.max_agents && a.shard_of[x] == a.shard_id) a.tmp_list[lo++] = x; }
      }
    }
    a.descs[gi] = out;
    return;
  }
  const uint32_t g = d->group;
  const uint32_t lc = g < a.max_groups ? a.lcount[g] : 0u;
  out.mstart = lc ? a.lstart[g] : 0u;
  out.mcount = lc;
  out.flags = SDB_DESC_SKIP_SENDER | SDB_DESC_PULL | SDB_DESC_POS | (abs_seq ? SDB_DESC_ABS_SEQ : 0u);
  a.descs[gi] = out;
}

// The exporter already bucketed its group sends by group (ascending send index).  The global bucket of a
// group is the concatenation of the sources' buckets in rank order - which is global send order - so no
// atomics and no sort: sizes per (group, source), one scan, one copy.
__device__ __forceinline__ const uint32_t* src_gs_off(const sdb_import_args& a, uint32_t s) {
  return reinterpret_cast<const uint32_t*>(a.meta[s] + a.tab->gs_off_off[s]);
}
__global__ void __launch_bounds__(256)
k_bucket_sizes(sdb_import_args a) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = a.max_groups * a.n_src;
  if (idx > n) return;
  uint32_t c = 0;
  if (idx < n) {
    const uint32_t g = idx / a.n_src, s = idx % a.n_src;
    if (a.lcount[g] && a.tab->gs_off_off[s]) { const uint32_t* o = src_gs_off(a, s); c = o[g + 1] - o[g]; }
  }
  a.gs_cnt[idx] = c;
}
__global__ void __launch_bounds__(256)
k_bucket_fill(sdb_import_args a, const uint32_t* __restrict__ goff, uint32_t* __restrict__ gs_off_out) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = a.max_groups * a.n_src;
  if (idx > n) return;
  if (idx == n) { gs_off_out[a.max_groups] = goff[n]; return; }
  const uint32_t g = idx / a.n_src, s = idx % a.n_src;
  const uint32_t dst = goff[idx];
  if (s == 0) gs_off_out[g] = dst;
  if (!a.lcount[g] || !a.tab->gs_off_off[s]) return;
  const uint32_t* o = src_gs_off(a, s);
  const uint32_t* ix = reinterpret_cast<const uint32_t*>(a.meta[s] + a.tab->gs_idx_off[s]);
  const uint32_t b = o[g], e = o[g + 1], first = a.tab->first[s];
  for (uint32_t k = b; k < e; ++k) a.gs_idx[dst + (k - b)] = first + ix[k];
}

extern "C" cudaError_t sdb_scan_u32(const uint32_t* in, uint32_t* local, uint32_t* tops, uint32_t n,
                                    unsigned long long* total_out, uint32_t* out, cudaStream_t stream);

// phase 1: measure + scans.  totals_dev[0] receives the arena granules this import will write.
extern "C" cudaError_t sdb_launch_import_measure(const sdb_import_args* a, uint32_t n_cap, uint32_t* w_local,
                                                 uint32_t* w_tops, uint32_t* gs_local, uint32_t* gs_tops,
                                                 uint32_t* gs_off_out, uint32_t* lw_local, uint32_t* lw_tops,
                                                 unsigned long long* totals_dev,
                                                 cudaStream_t stream, sdb_profiler* prof, int* n_launches) {
  const int pi = sdb_prof_begin(prof, SDB_PK_XSHARD, stream);
  k_wire_table<<<1, 32, 0, stream>>>(*a, const_cast<sdb_src_tab*>(a->tab));
  k_wire_measure<<<(n_cap + 255) / 256, 256, 0, stream>>>(*a, n_cap);
  cudaError_t e = sdb_scan_u32(a->w, w_local, w_tops, n_cap, totals_dev, nullptr, stream);
  const uint32_t nb = a->max_groups * a->n_src + 1;
  k_bucket_sizes<<<(nb + 255) / 256, 256, 0, stream>>>(*a);
  // gs_off_out here is the materialised scan over (group, source) pairs ("goff", nb entries)
  if (e == cudaSuccess) e = sdb_scan_u32(a->gs_cnt, gs_local, gs_tops, nb, nullptr, gs_off_out, stream);
  if (e == cudaSuccess) e = sdb_scan_u32(a->lw, lw_local, lw_tops, n_cap, totals_dev + 1, nullptr, stream);   // owned recipients
  sdb_prof_end(prof, pi, stream);
  if (n_launches) *n_launches += 9;
  return e != cudaSuccess ? e : cudaGetLastError();
}

// phase 2: localized descriptors + ordered buckets
extern "C" cudaError_t sdb_launch_import_localize(const sdb_import_args* a, uint32_t n_cap, const uint32_t* goff,
                                                  uint32_t* gs_off_out, cudaStream_t stream, sdb_profiler* prof,
                                                  int* n_launches) {
  const int pi = sdb_prof_begin(prof, SDB_PK_XSHARD, stream);
  k_wire_localize<<<(n_cap + 255) / 256, 256, 0, stream>>>(*a, n_cap);
  const uint32_t nb = a->max_groups * a->n_src + 1;
  k_bucket_fill<<<(nb + 255) / 256, 256, 0, stream>>>(*a, goff, gs_off_out);
  sdb_prof_end(prof, pi, stream);
  if (n_launches) *n_launches += 2;
  return cudaGetLastError();
}
