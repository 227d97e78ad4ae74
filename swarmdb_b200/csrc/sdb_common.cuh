// sdb_common.cuh - shared device/host definitions for the sm_100a kernels.
//
// HBM layout of one shard (all sizes powers of two where masks are used):
//
//   arena        uint8 [arena_bytes]      packed message records, 32-byte granules, written
//                                         strictly sequentially in global send order (a log).
//                                         record = sdb_msg_header (32 B) | payload pad32.
//                                         position `apos` (granules, monotonic u64);
//                                         byte address = arena + ((apos & gmask) << 5);
//                                         a batch never straddles the wrap point.
//   ring_hdr     16 B   [max_agents]      {head, tail, ctail, ntomb}: monotonic per-agent cursors, ONE 16-byte
//                                         load per agent for every kernel.  {head, tail} is also one aligned 64-bit
//                                         word: an atomicAdd(1<<32) on it claims a slot and returns head in the
//                                         same L2 transaction (enqueue paths that are not rank-ordered).
//                                         ctail = tail as of the last commit; ntomb = consumed entries still
//                                         inside [head, tail) (holes left by priority receives).
//   ring         8 B    [max_agents][R]   one entry per pending record: {handle = (uint32)apos,
//                                         meta = prio<<14 | record_granules, 0xFFFF = consumed}.  Interleaved so
//                                         the few entries an agent gains/loses per batch share one 32-byte sector.
//   members      uint32 [member_pool]     group member lists (CSR kept on the host)
//
// Ordering contract: after `commit`, every ring is sorted by handle, i.e. by arena
// position, i.e. by global send order (the reference's single-partition Kafka log order,
// SURVEY App. A rule 7).  Enqueue kernels claim slots with atomics in arbitrary order;
// the commit kernel sorts the few entries each agent received in the batch.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/swarmdb_b200.h"

#define SDB_META_TOMB 0xFFFFu
#define SDB_META_GLEN_MASK 0x3FFFu

// per-send descriptor staged to the device (64 B, one per send); q1 is all the index-build
// kernel needs, so it reads 16 bytes per (agent, send) pair
struct __align__(16) sdb_send_desc {
  // q0
  uint64_t payload_off;  // byte offset into the batch payload buffer, 16-B aligned
  double   timestamp;
  // q1
  uint32_t gran0;        // arena granule offset of member 0 relative to the batch arena base
  uint32_t sender;
  uint32_t rgran;        // granules per record = 1 + pad32(len)/32
  uint16_t len;
  uint8_t  prio;
  uint8_t  type;
  // q2
  uint32_t rec0;         // seq offset of member 0 relative to the batch seq base
  uint32_t mstart;       // first member in the member pool (group/list) or receiver idx (p2p)
  uint32_t mcount;       // members (1 for p2p)
  uint32_t group;        // group index or SDB_NO_GROUP
  // q3
  uint32_t flags;        // SDB_DESC_*
  uint32_t pad0;
  uint64_t seq_abs;      // SDB_DESC_ABS_SEQ: absolute sequence number of member position 0 (instead of seq_base + rec0)
};
static_assert(sizeof(sdb_send_desc) == 64, "desc must be 64 bytes");
static_assert(sizeof(sdb_msg_header) == 32, "header must be 32 bytes");

#define SDB_DESC_SKIP_SENDER 1u   // group send: member == sender is skipped (M:1268)
#define SDB_DESC_SHARED_SEQ 2u    // broadcast: every copy carries the same seq (one Message)
#define SDB_DESC_LIST_TEMP 4u     // mstart indexes the per-batch temporary list buffer
#define SDB_DESC_PULL 8u          // ring entries are built by k_pull_index, not by the fan-out kernel
#define SDB_DESC_RANKED 32u       // p2p: pad0 = rank of this send among the batch's sends to the same receiver (ring slot = ctail + rank)
#define SDB_DESC_RANK_LAST 64u    // p2p: highest rank for its receiver (count = rank + 1): this send publishes the ring tail
#define SDB_DESC_POS 16u          // seq offset of member k is member_pos[mstart + k] (sharded: original group position)
#define SDB_DESC_P2P 32u          // wire batches only: point-to-point send, mstart = receiver index
#define SDB_DESC_ABS_SEQ 64u      // seq_abs holds the sequence number (sources that number their own sends)

// cross-shard wire batch (device memory, moved between ranks by the caller):
//   [sdb_wire_header 128 B][n_sends x sdb_send_desc 64 B][group buckets][broadcast lists][payload bytes]
// descriptors use: payload_off (relative to the wire payload), timestamp, sender, group, rgran,
// len, prio, type, rec0 (seq offset inside the source's batch, counted in FULL group sizes)
struct __align__(16) sdb_wire_header {
  uint32_t magic;          // 'SDBW'
  uint32_t n_sends;
  uint64_t total_recs;     // sequence numbers this batch consumes (sum of full group sizes)
  uint64_t payload_bytes;
  uint64_t desc_off;       // byte offsets from the start of the wire batch
  uint64_t payload_off;
  uint32_t max_padlen;
  uint32_t n_other;        // sends that are not group sends (p2p / broadcast lists)
  uint64_t list_off;       // recipient lists of broadcast sends (uint32 agent indices)
  uint32_t n_list;
  uint32_t n_group_sends;
  uint64_t gs_off_off;     // group-send buckets of THIS batch: offsets [max_groups + 1] ...
  uint64_t gs_idx_off;     // ... and send indices, ascending inside a bucket (built by the exporter)
  uint32_t max_groups;
  uint32_t explicit_seq;   // 1: this batch's sends are numbered from seq_base (assigned by the exporter at send time)
  uint64_t seq_base;
  uint32_t pad[8];
};
static_assert(sizeof(sdb_wire_header) == 128, "wire header must be 128 bytes");
#define SDB_WIRE_MAGIC 0x57424453u

// per-batch view for the agent-parallel index build ("pull"): group sends of the batch bucketed
// by group (gs_*), and the inverse of the group table (which groups/positions an agent is in)
struct sdb_pull_view {
  const uint32_t* memb_off;   // [max_agents + 1]
  const uint32_t* memb_grp;   // [memberships]
  const uint32_t* memb_pos;   // [memberships] position of the agent inside that group's member list
  const uint32_t* gs_off;     // [n_src][gs_off_stride] bucket offsets into gs_idx, one table per source
  const uint32_t* gs_idx;     // [n_src][gs_idx_stride] send indices (relative to the source's first send), ascending inside a bucket
  // The group-parallel index build walks the sources' buckets in source order (= global send order), so a
  // cross-shard import never concatenates them.  One source (n_src = 1, first[0] = 0) is the single-GPU case.
  uint32_t n_src, gs_off_stride, gs_idx_stride;
  const uint32_t* first;      // [n_src] global descriptor index of each source's first send (device), or nullptr = {0}
};

struct sdb_dev_counters {   // device-resident, updated with atomics
  unsigned long long enqueued;
  unsigned long long delivered;
  unsigned long long ring_overflow;
  unsigned long long skipped_sender;
  unsigned long long backend_picks;
  unsigned long long ovf_logged;    // records dropped since the overflow log was last read (the first SDB_OVF_LOG are logged)
  unsigned long long pad[2];
};
#define SDB_OVF_LOG 4096u

// everything a kernel needs about the shard, passed by value
struct __align__(16) sdb_ring_hdr { uint32_t head, tail, ctail, ntomb; };
static_assert(sizeof(sdb_ring_hdr) == 16, "ring header must be 16 bytes");

struct sdb_dev_view {
  uint8_t*  arena;
  sdb_ring_hdr* ring_hdr;       // [max_agents]
  uint2*    ring;               // [max_agents][R]  .x = handle, .y = meta (SDB_META_TOMB: consumed)
  const uint32_t* members;
  const uint32_t* member_pos;   // sharded mode: original group position of members[k]; else nullptr
  sdb_dev_counters* ctr;
  uint2*    ovf_log;     // [SDB_OVF_LOG] {agent, arena handle} of records whose ring was full (sdb_overflow_log)
  uint64_t gmask;        // arena granules - 1
  uint32_t ring_slots;   // R
  uint32_t ring_shift;   // log2 R
  uint32_t max_agents;
};

// arguments of one receive call (scratch + outputs), shared by sdb_api.cu and sdb_recv.cu
struct sdb_recv_args {
  const uint32_t* agent_idx;   // [n] or nullptr (identity)
  uint32_t n;
  uint32_t max_messages;
  uint32_t flags;
  uint32_t* cnt;               // [n] records selected per agent                 (multi-kernel path scratch)
  uint32_t* rec_local;         // [n] scan of cnt (block-local part)             (multi-kernel path scratch)
  uint32_t* rec_tops;          // [tiles] scan of cnt (per-block part)           (multi-kernel path scratch)
  uint32_t* rec_off;           // [n] index of the agent's first output record (both paths)
  // per-record plan, indexed by output record number r = rec_off[slot] + rank:
  //   .x arena handle, .y payload offset in the packed output (granules; the multi-kernel path stores the
  //   tile-local part and adds plan_tops[r / tile]), .z payload granules, .w request slot
  uint4* plan;                 // [rec_cap]
  uint32_t* plan_tops;         // [rec tiles] or nullptr (single-pass path: .y is already absolute)
  unsigned long long* totals;  // [0] records delivered, [1] payload granules delivered, [2..] scratch
  uint32_t* big_list;          // [n] request slots that need the warp-per-agent selector
  uint32_t* big_count;         // [1]
  unsigned long long* lb;      // [tiles + 2] single-pass path: decoupled look-back status words (zero between calls),
                               //             then the tile ticket counter and the packed totals (see k_recv_plan)
  // outputs
  uint32_t* count_out;         // [n]
  sdb_msg_header* hdr_out;     // [rec_cap]
  uint8_t* payload_out;        // [rec_cap * pad32(max_payload)]
  uint64_t rec_cap;
};

// ---- asynchronous import (sdb_import_wire_ptrs_async): everything the host used to learn through a sync lives on
// the device.  `sdb_cursor` is the device-resident twin of the host's arena / sequence counters (whoever advanced
// them last marks the other side stale); `sdb_batch_base` carries one import's placement from the kernel that
// computes it (k_import_fused) to the kernels that follow (fan-out, index build, commit).
struct sdb_cursor {
  unsigned long long arena_tail;    // granules, monotonic
  unsigned long long arena_floor;   // granules: everything below is reclaimed
  unsigned long long next_seq;
  unsigned long long error;         // sticky bits: 1 = an import did not fit the arena (dropped whole), 2 = recipient-list pool exhausted
  unsigned long long floor_dist;    // scratch of the floor scan: max distance of a pending record below the tail
  unsigned long long pad[3];
};
struct sdb_batch_base {
  unsigned long long arena_base;    // granule position of the import's first record
  unsigned long long seq_base;      // sequence number the import's relative numbers count from
  uint32_t n_total;                 // localized descriptors (sends) of the import
  uint32_t n_other;                 // point-to-point / broadcast sends among them (their ring entries need the commit sort)
  uint32_t skip;                    // 1: refused (see sdb_cursor.error) - the following kernels do nothing
  uint32_t max_padlen;
  unsigned long long total_grans, total_recs;
};
// grand totals of one localized import, kept between a prefetch (k_import_fused on the prefetch stream, while the previous
// step still runs) and its placement (k_import_place on the shard's stream, after the previous step's commit)
struct sdb_import_totals {
  unsigned long long need, lists, total_recs, explicit_end;
  uint32_t n_total, n_other, maxpad, pad;
};
// mailbox of the low-latency dequeue server (mapped pinned host memory; request and completion words on separate lines)
struct sdb_ls_mailbox {
  uint32_t req_seq, agent, max_messages, flags, quit, pad0[27];
  uint32_t done_seq, pad1[31];
};
static_assert(sizeof(sdb_ls_mailbox) == 256, "latency mailbox must be 256 bytes");

// cross-rank flags of an export buffer (last 128 bytes of the buffer, never touched by an export's copies)
struct sdb_wire_ctrl {
  uint32_t ready;                   // last step whose export into this buffer is complete (written by the owner)
  uint32_t done;                    // last step for which the OWNER, as an importer, finished reading every rank's buffer of this parity
  uint32_t pad[30];
};
static_assert(sizeof(sdb_wire_ctrl) == 128, "wire ctrl block must be 128 bytes");

// arguments of one cross-shard import (sdb_xshard.cu)
#define SDB_MAX_SRC 16
// per-source prefix table built once per import by k_wire_table (so the per-send kernels never
// touch the - possibly remote - wire headers)
struct sdb_src_tab {
  uint32_t first[SDB_MAX_SRC + 1];   // global index of each source's first send
  uint64_t rec_base[SDB_MAX_SRC];    // sequence offset of each source inside the global batch
  uint64_t desc_off[SDB_MAX_SRC], list_off[SDB_MAX_SRC], payload_off[SDB_MAX_SRC];
  uint64_t gs_off_off[SDB_MAX_SRC], gs_idx_off[SDB_MAX_SRC];
  uint32_t explicit_seq[SDB_MAX_SRC];   // rec_base[s] is then an absolute sequence number
};
struct sdb_import_args {
  const sdb_src_tab* tab;
  const uint8_t* meta[SDB_MAX_SRC];   // header + descriptors of each source: a LOCAL copy when the wire batch is remote
  const uint8_t* wire[SDB_MAX_SRC];   // one wire batch per source rank; may point into PEER GPU memory (NVLink)
  uint32_t n_src;
  uint32_t max_sends;         // capacity per wire batch
  const uint32_t* lstart;     // [max_groups] local member list of each group
  const uint32_t* lcount;
  uint32_t max_groups;
  // outputs / scratch
  uint32_t* w;                // [n_src * max_sends] granules written locally per send (scan input)
  uint32_t* gs_cnt;           // [(max_groups + 1) * n_src + 1] bucket sizes per (group, source), then their scan
  sdb_send_desc* descs;       // [n_src * max_sends] localized descriptors
  const uint32_t* w_local;    // scan of w
  const uint32_t* w_tops;
  const uint32_t* gs_off;     // [max_groups + 1] scan of the histogram
  uint32_t* gs_idx;           // [n_src * max_sends]
  // point-to-point and broadcast sends: recipients this shard owns, compacted into a temporary list
  const uint8_t* shard_of;    // [max_agents]
  uint32_t shard_id;
  uint32_t max_agents;
  uint32_t* lw;               // [n_src * max_sends] owned recipients of each non-group send (scan input)
  const uint32_t* lw_local;   // scan of lw
  const uint32_t* lw_tops;
  uint32_t* tmp_list;         // [list_cap]
  uint32_t list_cap;
};

// arguments of the asynchronous import's fused kernel (sdb_xshard.cu: k_import_fused)
struct sdb_import2_args {
  const uint8_t* wire[SDB_MAX_SRC];   // one wire batch per source rank; may point into PEER GPU memory (NVLink)
  uint32_t n_src, max_sends, max_groups, shard_id, max_agents;
  const uint32_t* lcount; const uint32_t* lstart;
  const uint8_t* shard_of;
  sdb_send_desc* descs;               // out [n_src * max_sends]
  uint32_t* gs_off_src;               // out [n_src][max_groups + 1]
  uint32_t* gs_idx_src;               // out [n_src][max_sends]
  uint32_t* first;                    // out [SDB_MAX_SRC + 1]
  uint32_t* tmp_list; uint32_t list_cap;
  unsigned long long* lb;             // [tiles + 2]: look-back status words, tile ticket (zeroed before the launch)
  sdb_cursor* cur; sdb_batch_base* bb;
  unsigned long long arena_grans;
  const sdb_wire_header* hdrs;        // [n_src] local copies of the wire headers (made by k_wire_wait)
  uint32_t* commit_count;             // reset here for the commit sort that follows the fan-out
  sdb_import_totals* totals;          // prefetch mode (non-null): the grand totals go here and k_import_place places the
                                      // import later, on the shard's stream (the cursor is not touched by this kernel)
  uint32_t shared;                    // group sends take the shared-payload layout: lc headers + one payload per send
};

// ---- optional per-kernel timing with CUDA events on the launching stream (bench / roofline) ----
struct sdb_profiler {
  int enabled;
  int n, cap;
  int* kind;
  cudaEvent_t* ev_a;
  cudaEvent_t* ev_b;
};
static inline int sdb_prof_begin(sdb_profiler* p, int kind, cudaStream_t s) {
  if (!p || !p->enabled || p->n >= p->cap) return -1;
  const int i = p->n++;
  p->kind[i] = kind;
  cudaEventRecord(p->ev_a[i], s);
  return i;
}
static inline void sdb_prof_end(sdb_profiler* p, int i, cudaStream_t s) {
  if (i >= 0) cudaEventRecord(p->ev_b[i], s);
}

#ifdef __CUDACC__

__device__ __forceinline__ uint8_t* sdb_arena_ptr(const sdb_dev_view& v, uint64_t apos) {
  return v.arena + ((apos & v.gmask) << 5);
}

// Claim the next slot of agent a's ring and publish (handle, meta) into it.
// Returns false when the ring is full (nothing is written; the commit kernel clamps tail).
// A record whose receiver's ring is full: it exists in the arena but no ring entry refers to it.  The (agent, handle)
// pair is logged so that the host can tell WHICH messages were lost (sdb_overflow_log), not only how many.
__device__ __forceinline__ void sdb_note_overflow(const sdb_dev_view& v, uint32_t a, uint32_t handle) {
  const unsigned long long k = atomicAdd(&v.ctr->ovf_logged, 1ull);
  if (k < SDB_OVF_LOG) v.ovf_log[k] = make_uint2(a, handle);
}
__device__ __forceinline__ bool sdb_ring_append(const sdb_dev_view& v, uint32_t a, uint32_t handle, uint16_t meta) {
  unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(v.ring_hdr + a), 1ull << 32);
  uint32_t e = static_cast<uint32_t>(old >> 32);
  uint32_t head = static_cast<uint32_t>(old);
  if (e - head >= v.ring_slots) { sdb_note_overflow(v, a, handle); return false; }
  size_t slot = (static_cast<size_t>(a) << v.ring_shift) + (e & (v.ring_slots - 1));
  v.ring[slot] = make_uint2(handle, meta);
  return true;
}
__device__ __forceinline__ uint2* sdb_ring_of(const sdb_dev_view& v, uint32_t a) {
  return v.ring + (static_cast<size_t>(a) << v.ring_shift);
}
__device__ __forceinline__ uint32_t sdb_meta(const uint2& e) { return e.y & 0xFFFFu; }
// Shared payloads.  A record is a 32-byte header at `handle` and its payload at handle + 1 + dm1 granules.  dm1 = 0 is
// the classic image (payload right behind its header); group sends above the pull threshold write the members' headers
// back to back and ONE payload behind them (header j of m: dm1 = m - 1 - j), which every member's ring entry points at.
// dm1 travels in the upper half of a ring entry's .y (the lower half is the meta word) and of a plan entry's .z (the
// lower half is the payload size in granules).  The payload sits ABOVE all its headers, so the arena floor - the
// smallest pending handle - never passes a payload that is still referenced.
__device__ __forceinline__ uint32_t sdb_entry_dm1(uint32_t y) { return y >> 16; }
__device__ __forceinline__ uint32_t sdb_plan_z(uint32_t y) { return ((y & SDB_META_GLEN_MASK) - 1u) | (y & 0xFFFF0000u); }   // ring .y -> plan .z
__device__ __forceinline__ const uint8_t* sdb_payload_of(const sdb_dev_view& v, uint32_t handle, uint32_t dm1) {
  return v.arena + (((static_cast<uint64_t>(handle) + 1u + dm1) & v.gmask) << 5);
}

__device__ __forceinline__ uint4 sdb_header_lo(uint64_t seq, double ts) {
  uint4 r;
  r.x = static_cast<uint32_t>(seq);
  r.y = static_cast<uint32_t>(seq >> 32);
  unsigned long long t = __double_as_longlong(ts);
  r.z = static_cast<uint32_t>(t);
  r.w = static_cast<uint32_t>(t >> 32);
  return r;
}
__device__ __forceinline__ uint4 sdb_header_hi(uint32_t sender, uint32_t receiver, uint32_t group,
                                               uint16_t len, uint8_t prio, uint8_t type) {
  uint4 r;
  r.x = sender;
  r.y = receiver;
  r.z = group;
  r.w = static_cast<uint32_t>(len) | (static_cast<uint32_t>(prio) << 16) | (static_cast<uint32_t>(type) << 24);
  return r;
}

// streaming 16-byte accesses: message bytes are touched once, keep them out of L1
__device__ __forceinline__ uint4 sdb_ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void sdb_st_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// L2 eviction-priority policies: message bytes stream through once (evict_first) so that the
// small, hot ring metadata (evict_last) stays L2-resident between enqueue and dequeue
__device__ __forceinline__ uint64_t sdb_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t sdb_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 sdb_ld_stream_pol(const void* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void sdb_st_stream_pol(void* p, const uint4& v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void sdb_st_u32_pol(uint32_t* p, uint32_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void sdb_st_u16_pol(uint16_t* p, uint16_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.u16 [%0], %1, %2;" :: "l"(p), "h"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void sdb_st_u64_pol(uint64_t* p, uint64_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" :: "l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t sdb_ld_u32_pol(const uint32_t* p, uint64_t pol) {
  uint32_t r;
  asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint16_t sdb_ld_u16_pol(const uint16_t* p, uint64_t pol) {
  uint16_t r;
  asm volatile("ld.global.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint64_t sdb_ld_u64_pol(const uint64_t* p, uint64_t pol) {
  uint64_t r;
  asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol));
  return r;
}

// ---- mbarrier + TMA (bulk async copy) wrappers ----------------------------------------------
__device__ __forceinline__ uint32_t sdb_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void sdb_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(sdb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sdb_fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sdb_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void sdb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sdb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sdb_mbar_wait_bounded(uint64_t* bar, uint32_t phase) {
  // bounded spin: a phase-accounting bug must surface as a launch failure (trap), never as a hung GPU.  The bound is
  // wall time (30 s on %globaltimer, polled every 4096 spins), not a spin count: under compute-sanitizer or with a
  // payload arriving from a busy peer GPU a legitimate wait can take many more iterations than on an idle device.
  const uint32_t addr = sdb_smem_u32(bar);
  unsigned long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(addr), "r"(phase) : "memory");
    if (done) return;
    if ((spins & 4095u) == 4095u) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t0 == 0) t0 = t1;
      else if (t1 - t0 > 30ull * 1000000000ull) __trap();
    }
  }
}
// every mbarrier wait in this library is the bounded one: a phase-accounting bug must trap, never hang the GPU
__device__ __forceinline__ void sdb_mbar_wait(uint64_t* bar, uint32_t phase) { sdb_mbar_wait_bounded(bar, phase); }
// global -> shared bulk copy, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void sdb_tma_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(sdb_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(sdb_smem_u32(bar)) : "memory");
}
// shared -> global bulk copy (bulk_group completion)
__device__ __forceinline__ void sdb_tma_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(sdb_smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sdb_tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sdb_tma_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void sdb_tma_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}

// ---- hierarchical decoupled look-back over pairs (a, b) of 31-bit-saturating sums ------------------------------------
// Single-pass kernels (k_recv_plan, k_import_fused) need, per tile, the exclusive prefix of a pair over all earlier
// tiles.  A flat look-back chain advances 32 tiles per L2 round trip, which for thousands of small tiles IS the
// kernel's duration; wider windows drown the L2 in polling.  Two levels instead:
//   * every tile publishes its aggregate at once (no dependency) and bumps its super-tile's counter (32 tiles each);
//   * the tile that completes a super-tile (its "finalizer") sums the 32 aggregates, publishes the super-tile
//     aggregate, looks back over SUPER-TILE words only (tiles / 32 of them: a handful of hops), and publishes the
//     super-tile's exclusive prefix;
//   * a tile's prefix = its super-tile's exclusive prefix + the aggregates of the earlier tiles of its super-tile.
// Words: [63:62] status, [61:31] a, [30:0] b (saturating: a saturated prefix lies beyond every capacity check).
// Layout of the scratch array `lb` (zeroed before the launch), T tiles, S = ceil(T / 32) super-tiles:
//   [0, T) tile aggregates | [T, T+S) super-tile status | [T+S, T+2S) super-tile exclusive prefix | [T+2S, T+3S) counters
//   | [T+3S] tile ticket | [T+3S+1] kernel-specific
// Tiles must be numbered by an atomic ticket (a tile's predecessors are then running or done).  Call with warp 0.
__device__ __forceinline__ unsigned long long sdb_lb_pack(uint32_t st, unsigned long long a, unsigned long long b) {
  const unsigned long long SAT = 0x7FFFFFFFull;
  return (static_cast<unsigned long long>(st) << 62) | ((a < SAT ? a : SAT) << 31) | (b < SAT ? b : SAT);
}
__device__ __forceinline__ unsigned long long sdb_lb_ld(const unsigned long long* p) {
  unsigned long long x;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(x) : "l"(p) : "memory");
  return x;
}
__device__ __forceinline__ void sdb_lb_st(unsigned long long* p, unsigned long long x) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(x) : "memory");
}
__host__ __device__ static inline size_t sdb_lb_words(uint32_t tiles) { return static_cast<size_t>(tiles) + 3 * ((tiles + 31) / 32) + 2; }

// ta/tb: this tile's aggregate.  Returns the exclusive prefix in ea/eb (every lane).  *is_last tells the caller's lane 0
// whether this tile saw the grand total (ga/gb) - exactly one tile does.
__device__ __forceinline__ void sdb_lb_prefix(unsigned long long* lb, uint32_t tile, uint32_t tiles, unsigned long long ta,
                                              unsigned long long tb, uint32_t lane, unsigned long long& ea, unsigned long long& eb,
                                              bool& is_last, unsigned long long& ga, unsigned long long& gb) {
  const unsigned long long SAT = 0x7FFFFFFFull;
  const uint32_t S = (tiles + 31) / 32;
  unsigned long long* agg = lb;
  unsigned long long* sst = lb + tiles;
  unsigned long long* sex = lb + tiles + S;
  unsigned int* scnt = reinterpret_cast<unsigned int*>(lb + tiles + 2 * static_cast<size_t>(S));   // one 64-bit slot per counter
  const uint32_t st = tile >> 5, ti = tile & 31u;
  const uint32_t first = st << 5;
  const uint32_t msize = min(32u, tiles - first);
  is_last = false; ga = 0; gb = 0;
  uint32_t finalizer = 0;
  if (lane == 0) {
    sdb_lb_st(agg + tile, sdb_lb_pack(1u, ta, tb));
    __threadfence();
    finalizer = atomicAdd(scnt + 2 * st, 1u) == msize - 1 ? 1u : 0u;
  }
  finalizer = __shfl_sync(0xFFFFFFFFu, finalizer, 0);
  if (finalizer) {
    __threadfence();
    // every member has published: sum the super-tile
    const unsigned long long w = lane < msize ? sdb_lb_ld(agg + first + lane) : 0ull;
    unsigned long long sa = (w >> 31) & SAT, sb = w & SAT;
    for (int o = 16; o; o >>= 1) { sa += __shfl_xor_sync(0xFFFFFFFFu, sa, o); sb += __shfl_xor_sync(0xFFFFFFFFu, sb, o); }
    unsigned long long xa = 0, xb = 0;
    if (st == 0) {
      if (lane == 0) sdb_lb_st(sst, sdb_lb_pack(2u, sa, sb));
    } else {
      if (lane == 0) sdb_lb_st(sst + st, sdb_lb_pack(1u, sa, sb));
      int32_t idx = static_cast<int32_t>(st) - 1 - static_cast<int32_t>(lane);
      for (;;) {
        unsigned long long w2; uint32_t dm, upto;
        for (;;) {
          w2 = idx >= 0 ? sdb_lb_ld(sst + idx) : sdb_lb_pack(2u, 0, 0);
          const uint32_t s2 = static_cast<uint32_t>(w2 >> 62);
          dm = __ballot_sync(0xFFFFFFFFu, s2 == 2u);
          const uint32_t zm = __ballot_sync(0xFFFFFFFFu, s2 == 0u);
          upto = dm ? static_cast<uint32_t>(__ffs(dm)) : 32u;
          const uint32_t need = upto >= 32u ? 0xFFFFFFFFu : ((1u << upto) - 1u);
          if ((zm & need) == 0) break;
          __nanosleep(40);
        }
        unsigned long long va = lane < upto ? ((w2 >> 31) & SAT) : 0ull, vb = lane < upto ? (w2 & SAT) : 0ull;
        for (int o = 16; o; o >>= 1) { va += __shfl_xor_sync(0xFFFFFFFFu, va, o); vb += __shfl_xor_sync(0xFFFFFFFFu, vb, o); }
        xa += va; xb += vb;
        if (dm) break;
        idx -= 32;
      }
      __threadfence();
      if (lane == 0) sdb_lb_st(sst + st, sdb_lb_pack(2u, xa + sa, xb + sb));
    }
    if (lane == 0) sdb_lb_st(sex + st, sdb_lb_pack(1u, xa, xb));
    if (st == S - 1) { is_last = true; ga = xa + sa; gb = xb + sb; }
  }
  // every tile: super-tile prefix + earlier tiles of the super-tile
  unsigned long long pw = 0;
  for (;;) {
    const unsigned long long w = lane < ti ? sdb_lb_ld(agg + first + lane) : sdb_lb_pack(1u, 0, 0);
    if (lane == 0) pw = sdb_lb_ld(sex + st);
    const bool missing = (w >> 62) == 0 || (lane == 0 && (pw >> 62) == 0);
    if (!__any_sync(0xFFFFFFFFu, missing)) {
      unsigned long long va = (w >> 31) & SAT, vb = w & SAT;
      if (lane == 0) { va += (pw >> 31) & SAT; vb += pw & SAT; }
      for (int o = 16; o; o >>= 1) { va += __shfl_xor_sync(0xFFFFFFFFu, va, o); vb += __shfl_xor_sync(0xFFFFFFFFu, vb, o); }
      ea = va; eb = vb;
      break;
    }
    __nanosleep(40);
  }
  __threadfence();
}

#endif  // __CUDACC__
