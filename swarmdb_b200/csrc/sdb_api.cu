// sdb_api.cu - host side of the C ABI declared in include/swarmdb_b200.h.
//
// Owns: device memory of one shard, the host mirror of the group table (CSR), the global
// sequence counter, arena space accounting (tail / floor / reclaim), pinned staging for
// per-send descriptors, and the launch sequence of every entry point.  No CPU data path:
// every routed byte goes through the kernels in sdb_send.cu / sdb_recv.cu / sdb_balance.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "sdb_common.cuh"

extern "C" {
cudaError_t sdb_launch_p2p(const sdb_dev_view*, const sdb_send_desc*, uint32_t, const uint8_t*, uint64_t, uint64_t, int, cudaStream_t, sdb_profiler*, uint32_t);
cudaError_t sdb_launch_ovf_resolve(const sdb_dev_view*, uint32_t, unsigned long long*, cudaStream_t);
cudaError_t sdb_launch_list_index(const sdb_dev_view*, const sdb_send_desc*, const uint32_t*, const uint32_t*, uint32_t, uint32_t, uint64_t,
                                  cudaStream_t, sdb_profiler*);
cudaError_t sdb_launch_commit_ranked(const sdb_dev_view*, const sdb_send_desc*, uint32_t, uint32_t, cudaStream_t, sdb_profiler*);
cudaError_t sdb_send_prepare_device();
cudaError_t sdb_launch_fanout(const sdb_dev_view*, const sdb_send_desc*, uint32_t, const uint8_t*, const uint32_t*,
                              uint64_t, uint64_t, uint32_t, int, int, cudaStream_t, sdb_profiler*, const sdb_batch_base*, uint32_t);
cudaError_t sdb_launch_fanout_shared(const sdb_dev_view*, const sdb_send_desc*, uint32_t, const uint8_t*, uint64_t, uint64_t, uint32_t, int,
                                     cudaStream_t, sdb_profiler*, const sdb_batch_base*);
cudaError_t sdb_launch_commit(const sdb_dev_view*, uint32_t, uint32_t, uint32_t*, uint32_t*, int, cudaStream_t, sdb_profiler*,
                              const sdb_batch_base*);
cudaError_t sdb_launch_pull(const sdb_dev_view*, const sdb_pull_view*, const sdb_send_desc*, uint32_t, uint32_t, const uint8_t*,
                            uint32_t, uint32_t, const uint32_t*, const uint32_t*, uint64_t, int, cudaStream_t, sdb_profiler*, int*,
                            const sdb_batch_base*, uint32_t);
cudaError_t sdb_launch_receive(const sdb_dev_view*, const sdb_recv_args*, cudaStream_t, int*, sdb_profiler*, int, uint32_t);
cudaError_t sdb_recv_prepare_device();
cudaError_t sdb_launch_receive_small(const sdb_dev_view*, const uint32_t*, uint32_t, uint32_t, uint32_t, uint32_t, uint4*,
                                     uint8_t*, cudaStream_t, sdb_profiler*);
cudaError_t sdb_launch_digest(const sdb_recv_args*, uint32_t, unsigned long long*, int, cudaStream_t);
cudaError_t sdb_launch_arena_floor(const sdb_dev_view*, uint32_t, uint32_t, unsigned long long*, cudaStream_t);
cudaError_t sdb_launch_latency_server(const sdb_dev_view*, sdb_ls_mailbox*, uint8_t*, uint32_t, uint4*, cudaStream_t);
cudaError_t sdb_launch_pick(int mode, uint32_t n_backends, const uint32_t* weight_dev, unsigned long long* load_dev,
                            uint32_t n_req, const uint32_t* cost_dev, uint64_t seed, uint32_t* out_dev,
                            unsigned long long* scratch_dev, const uint32_t* log_tab_dev,
                            cudaStream_t stream, int* n_launches);
void sdb_build_log2_table(uint32_t* tab257);
cudaError_t sdb_launch_agent_loads(const sdb_dev_view*, const uint32_t*, uint32_t, sdb_agent_load*, int, cudaStream_t);
cudaError_t sdb_launch_queue_stats(const sdb_dev_view*, uint32_t, unsigned long long*, cudaStream_t);
cudaError_t sdb_launch_backend_loads_from_queues(const sdb_dev_view*, uint32_t, const uint32_t*, uint32_t, unsigned long long*, cudaStream_t);
cudaError_t sdb_launch_import_measure(const sdb_import_args*, uint32_t, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*,
                                      uint32_t*, uint32_t*, unsigned long long*, cudaStream_t, sdb_profiler*, int*);
cudaError_t sdb_launch_import_localize(const sdb_import_args*, uint32_t, const uint32_t*, uint32_t*, cudaStream_t, sdb_profiler*, int*);
cudaError_t sdb_launch_wire_wait(const uint32_t* const*, uint32_t, uint32_t, cudaStream_t, const void* const*, sdb_wire_header*);
cudaError_t sdb_launch_wire_set(uint32_t*, uint32_t, cudaStream_t);
cudaError_t sdb_launch_import_fused(const sdb_import2_args*, cudaStream_t, sdb_profiler*, int*);
cudaError_t sdb_launch_import_place(const sdb_import_totals*, sdb_cursor*, sdb_batch_base*, unsigned long long, uint32_t, uint32_t*, cudaStream_t);
cudaError_t sdb_launch_arena_floor_cur(const sdb_dev_view*, uint32_t, sdb_cursor*, cudaStream_t);
}

#define SDB_SCAN_TILE 4096u

struct sdb_staged {
  uint32_t kind = 0;          // 0 p2p, 1 group, 2 list
  uint32_t n = 0;
  uint64_t total_recs = 0;    // sequence numbers consumed
  uint64_t total_grans = 0;   // arena granules consumed
  uint32_t max_padlen = 0;
  sdb_send_desc* descs_dev = nullptr;
  uint8_t* payload_dev = nullptr;
  uint32_t* list_dev = nullptr;
  uint32_t* gs_off_dev = nullptr;   // [max_groups + 1] group-send buckets (pull index build)
  uint32_t* gs_idx_dev = nullptr;   // [n]
  bool has_pull = false;      // group sends indexed by k_pull_index
  bool has_atomic = false;    // some sends claim ring slots with atomics -> k_commit must sort
  bool ranked = false;        // pure point-to-point batch ranked per receiver at staging time: no atomics, no sort
  bool shared = false;        // group sends laid out as headers + ONE payload per send (sdb_common.cuh "shared payloads")
  bool list_pull = false;     // pure broadcast batch over pairwise disjoint lists: ring entries built by k_list_index
  uint32_t lp_nd = 0, lp_max_count = 0, lp_words = 0;
  bool owns = false;          // device buffers owned by this object (else the handle's staging)
};

// group batches at least this large build ring entries with the agent-parallel pull kernel
static const uint64_t SDB_PULL_THRESHOLD = 16384;

struct sdb_ctx {
  sdb_config cfg{};
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  sdb_dev_view view{};
  // device buffers
  uint8_t* arena = nullptr;
  sdb_ring_hdr* ring_hdr = nullptr;
  uint2* ring = nullptr;
  uint32_t* members = nullptr;
  sdb_dev_counters* ctr = nullptr;
  // staging for non-staged sends
  sdb_send_desc* descs_host = nullptr;   // pinned [desc_cap]
  uint64_t desc_cap = 0;                 // max_batch_sends + room for the chunks of long broadcast lists
  sdb_staged scratch;                    // device descs/payload/list owned by the handle
  uint32_t* list_host = nullptr;         // pinned
  uint32_t* gs_host = nullptr;           // pinned [max_groups + 1 + max_batch_sends]
  // inverse group table (agent -> memberships), rebuilt lazily from `ghost`
  uint32_t* memb_off_dev = nullptr; uint32_t* memb_grp_dev = nullptr; uint32_t* memb_pos_dev = nullptr;
  bool memb_dirty = true;
  uint8_t* gexcl_dev = nullptr;          // [max_groups] 1: every member of the group belongs to that group only
  uint32_t n_excl_groups = 0, n_shared_agents = 0;
  std::vector<uint32_t> rank_cnt;              // per-receiver counters of the point-to-point ranking (all zero between batches)
  bool shared_cfg = false;                     // SDB_SHARED_PAYLOAD: group sends above the pull threshold share one payload per send
  uint32_t max_lcount = 0;                     // largest local member count of a group (shared payloads need <= 65535)
  std::vector<uint32_t> lp_first;              // broadcast batches: first chunk descriptor of every send
  std::vector<uint64_t> lp_begin, lp_end;      // broadcast batches: where each send's / list's recipients sit in the pool
  uint32_t* lp_host = nullptr; uint32_t* lp_dev = nullptr; uint64_t lp_cap = 0;   // list-parallel index tables (pinned / device)
  uint2* ovf_log = nullptr; unsigned long long* ovf_seq = nullptr;                  // overflow log + resolved sequence numbers
  // sharding (one handle = one shard): owner of each agent, full group lists, local positions
  std::vector<uint8_t> shard_of; bool sharded = false;
  std::vector<std::vector<uint32_t>> gfull;    // full member lists as given by the caller
  std::vector<std::vector<uint32_t>> gpos;     // position in the full list of each kept member
  std::vector<uint32_t> gcount_full;
  uint32_t* member_pos_dev = nullptr;          // [member_pool] parallel to `members`
  uint32_t* lstart_dev = nullptr; uint32_t* lcount_dev = nullptr; bool ltab_dirty = true;
  // import scratch (capacity xs_cap = num_shards * max_batch_sends sends)
  uint32_t xs_cap = 0;
  uint32_t* xs_w = nullptr; uint32_t* xs_w_local = nullptr; uint32_t* xs_w_tops = nullptr;
  uint32_t* xs_gs_cnt = nullptr; uint32_t* xs_gs_local = nullptr; uint32_t* xs_gs_tops = nullptr; uint32_t* xs_gs_cur = nullptr;
  uint32_t* xs_gs_off = nullptr; uint32_t* xs_gs_idx = nullptr; sdb_send_desc* xs_descs = nullptr;
  uint32_t* xs_lw = nullptr; uint32_t* xs_lw_local = nullptr; uint32_t* xs_lw_tops = nullptr;
  sdb_src_tab* xs_tab = nullptr;
  uint8_t* xs_meta = nullptr; uint64_t xs_meta_stride = 0;   // local copies of remote headers + descriptors
  uint8_t* shard_of_dev = nullptr;
  uint32_t* owned_dev = nullptr; uint32_t n_owned = 0;     // agents this shard owns, ascending (SDB_RECV_OWNED)
  // asynchronous import: device-resident twin of {arena_tail, arena_floor, next_seq} + one import's placement
  sdb_cursor* cursor_dev = nullptr; sdb_batch_base* bb_dev = nullptr;
  sdb_cursor* cursor_host = nullptr;          // pinned
  bool host_stale = false;                     // the device cursor moved (async import): host counters must be refreshed first
  bool dev_stale = true;                       // the host counters moved: push them before the next async import
  uint32_t* xs_gs_off_src = nullptr; uint32_t* xs_gs_idx_src = nullptr; uint32_t* xs_first = nullptr;
  unsigned long long* xs_lb = nullptr; sdb_wire_header* xs_hdrs = nullptr;
  // localized-import buffer sets: set 0 aliases the buffers above; set 1 exists once sdb_import_prefetch was used (an
  // import of step k+1 is localized on `pf_stream` while step k's kernels still read the set of the other parity)
  struct XsSet {
    sdb_send_desc* descs = nullptr; uint32_t* gs_off_src = nullptr; uint32_t* gs_idx_src = nullptr; uint32_t* first = nullptr;
    unsigned long long* lb = nullptr; sdb_wire_header* hdrs = nullptr; uint32_t* tmp_list = nullptr;
    sdb_import_totals* totals = nullptr; cudaEvent_t ready = nullptr; uint32_t step = 0; bool pending = false;
    bool shared = false;            // layout the set's descriptors were localized for
  } xset[2];
  cudaStream_t pf_stream = nullptr; cudaEvent_t ev_pf_fork = nullptr;
  uint8_t* wire_host = nullptr;                // pinned: header + descriptors of an export
  sdb_wire_header* hdrs_host = nullptr;        // pinned [num_shards]
  cudaEvent_t staging_free = nullptr;    // previous H2D of pinned staging has completed
  cudaStream_t side = nullptr;           // the index build of a pure group batch runs beside the fan-out
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // receive scratch + outputs
  uint32_t* rx_agent = nullptr; uint32_t* rx_cnt = nullptr; uint32_t* rx_rec_local = nullptr; uint32_t* rx_rec_tops = nullptr;
  uint4* rx_plan = nullptr; uint32_t* rx_rec_off = nullptr; unsigned long long* rx_lb = nullptr;
  uint32_t* rx_plan_tops = nullptr; unsigned long long* rx_totals = nullptr;
  uint32_t* rx_big_list = nullptr; uint32_t* rx_big_count = nullptr;
  uint32_t* rx_count = nullptr; sdb_msg_header* rx_hdr = nullptr; uint8_t* rx_payload = nullptr;
  unsigned long long* totals_host = nullptr;   // pinned [4]
  uint8_t* rx_small = nullptr;                 // device output block of the latency path
  uint8_t* small_host = nullptr;               // pinned mirror
  uint64_t small_bytes = 0;
  uint64_t pay_cap_gran = 0;
  // low-latency dequeue server (sdb_latency_server): mailbox + answer block in mapped pinned host memory
  sdb_ls_mailbox* ls_mbox = nullptr; uint8_t* ls_out = nullptr; uint32_t ls_out_cap = 0; uint4* ls_plan = nullptr;
  cudaStream_t ls_stream = nullptr; bool ls_wanted = false, ls_running = false; uint32_t ls_seq = 0;
  bool stream_busy = false;                    // work was enqueued on `stream` since the last synchronisation
  sdb_recv_args last_rx{}; bool last_rx_valid = false;   // the last bulk receive (sdb_digest_fold reads its device results)
  unsigned long long* digest = nullptr;                  // [max_agents] stream digests, allocated on first use
  // backends
  uint32_t* be_weight = nullptr; unsigned long long* be_load = nullptr; uint32_t n_backends = 0;
  uint32_t* be_req_cost = nullptr; uint32_t* be_out = nullptr; unsigned long long* be_scratch = nullptr;
  uint32_t* be_logtab = nullptr; uint32_t be_req_cap = 0;
  uint32_t* agent_backend_dev = nullptr;      // [max_agents] sticky agent -> backend map on the device (0xFFFFFFFF = none)
  // host state
  uint64_t next_seq = 1;       // ids start at 1 like the deterministic-uuid reference counter
  uint64_t arena_tail = 0;     // granules, monotonic
  uint64_t arena_floor = 0;    // granules
  uint64_t arena_grans = 0;
  uint32_t n_agents = 0;       // watermark
  std::vector<uint64_t> gstart; std::vector<uint32_t> gcount; std::vector<uint8_t> gdefined;
  uint64_t member_used = 0;
  std::vector<std::vector<uint32_t>> ghost;   // authoritative member lists (for pool compaction)
  uint64_t launches = 0;
  uint64_t picks = 0;
  uint64_t seen_overflow = 0;
  sdb_profiler prof{};
  std::string err;
};

namespace {

int fail(sdb_ctx* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
#define CUDA_TRY(h, expr)                                                                       \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(h, SDB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));            \
  } while (0)

bool is_pow2(uint64_t x) { return x && !(x & (x - 1)); }
// header + descriptors + group buckets of a wire batch (the part a remote importer copies locally)
uint64_t wire_meta_bytes(uint32_t max_sends, uint32_t max_groups) {
  const uint64_t b = sizeof(sdb_wire_header) + static_cast<uint64_t>(max_sends) * (sizeof(sdb_send_desc) + 4) + (static_cast<uint64_t>(max_groups) + 1) * 4 + 64;
  return (b + 127) & ~127ull;                         // wire batches are laid out back to back: keep them 128-byte aligned
}
uint32_t ilog2(uint64_t x) { uint32_t r = 0; while ((1ull << r) < x) ++r; return r; }
inline uint32_t pad32(uint32_t len) { return (len + 31u) & ~31u; }

template <typename T>
cudaError_t dmalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)); }

// ---- low-latency dequeue server ----------------------------------------------------------------------------------
int ls_stop(sdb_ctx* h) {
  if (!h->ls_running) return SDB_OK;
  __atomic_store_n(&h->ls_mbox->flags, 0x80000000u, __ATOMIC_RELEASE);     // quit bit, fetched with the request words
  CUDA_TRY(h, cudaStreamSynchronize(h->ls_stream));
  h->ls_running = false;
  return SDB_OK;
}
int ls_start(sdb_ctx* h) {
  if (h->ls_running) return SDB_OK;
  if (!h->ls_mbox) {
    h->ls_out_cap = 64u + 1024u * 64u;                                       // k_recv_small's block: 64-byte head + records
    const uint64_t need = 64ull + 1024ull * (32ull + pad32(h->cfg.max_payload_bytes));
    if (need < (4ull << 20)) h->ls_out_cap = static_cast<uint32_t>(need); else h->ls_out_cap = 4u << 20;
    CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->ls_mbox), sizeof(sdb_ls_mailbox), cudaHostAllocMapped));
    CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->ls_out), h->ls_out_cap, cudaHostAllocMapped));
    CUDA_TRY(h, dmalloc(&h->ls_plan, 1024 + 4));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->ls_stream, cudaStreamNonBlocking));
    std::memset(h->ls_mbox, 0, sizeof(sdb_ls_mailbox));
  }
  h->ls_mbox->quit = 0; h->ls_mbox->flags = 0; h->ls_mbox->req_seq = h->ls_seq; h->ls_mbox->done_seq = h->ls_seq;
  __sync_synchronize();
  sdb_ls_mailbox* mb_dev = nullptr; uint8_t* out_dev = nullptr;
  CUDA_TRY(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&mb_dev), h->ls_mbox, 0));
  CUDA_TRY(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&out_dev), h->ls_out, 0));
  CUDA_TRY(h, sdb_launch_latency_server(&h->view, mb_dev, out_dev, h->ls_out_cap, h->ls_plan, h->ls_stream));
  h->launches += 1;
  h->ls_running = true;
  return SDB_OK;
}

// The asynchronous import advances arena / sequence counters on the device.  Before host code uses its own copies
// it pulls the device's (one small D2H + sync), and reports an import the device had to refuse.
int cursor_to_host(sdb_ctx* h) {
  if (!h->host_stale) return SDB_OK;
  CUDA_TRY(h, cudaMemcpyAsync(h->cursor_host, h->cursor_dev, sizeof(sdb_cursor), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->host_stale = false;
  h->arena_tail = h->cursor_host->arena_tail; h->next_seq = h->cursor_host->next_seq;
  if (h->cursor_host->arena_floor > h->arena_floor) h->arena_floor = h->cursor_host->arena_floor;
  if (h->cursor_host->error) {
    const unsigned long long err = h->cursor_host->error;
    CUDA_TRY(h, cudaMemsetAsync(&h->cursor_dev->error, 0, sizeof(unsigned long long), h->stream));
    if (err & 1ull) return fail(h, SDB_EARENA_FULL, "an asynchronous import did not fit the message arena and was dropped whole "
                                                    "(unconsumed messages pin the log: receive or enlarge arena_bytes)");
    if (err & 2ull) return fail(h, SDB_ECAPACITY, "an asynchronous import's owned recipients exceeded list_pool_entries; it was dropped whole");
    return fail(h, SDB_ECAPACITY, "an asynchronous import carried payloads above 512 bytes (use the synchronous import)");
  }
  return SDB_OK;
}
int cursor_to_dev(sdb_ctx* h) {
  if (!h->dev_stale) return SDB_OK;
  h->cursor_host->arena_tail = h->arena_tail; h->cursor_host->arena_floor = h->arena_floor; h->cursor_host->next_seq = h->next_seq;
  CUDA_TRY(h, cudaMemcpyAsync(h->cursor_dev, h->cursor_host, 3 * sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->staging_free, h->stream));
  h->dev_stale = false;
  return SDB_OK;
}

// Make room for `need` granules in the arena; may run the floor-reclaim kernel (synchronises).
int arena_reserve(sdb_ctx* h, uint64_t need, uint64_t* base_out) {
  const uint64_t G = h->arena_grans;
  if (need > G) return fail(h, SDB_EARENA_FULL, "batch larger than the whole arena");
  uint64_t tail = h->arena_tail;
  if ((tail & (G - 1)) + need > G) tail = (tail + G - 1) & ~(G - 1);     // a batch never straddles the wrap
  if (tail + need - h->arena_floor > G) {
    // reclaim: recompute the floor from the oldest pending record of every ring
    cudaError_t e = sdb_launch_arena_floor(&h->view, h->n_agents, static_cast<uint32_t>(h->arena_tail),
                                           h->rx_totals + 2, h->stream);
    h->launches += 1;
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->totals_host + 2, h->rx_totals + 2, sizeof(unsigned long long),
                                             cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("arena floor: ") + cudaGetErrorString(e));
    h->arena_floor = h->arena_tail - h->totals_host[2];
    if (tail + need - h->arena_floor > G)
      return fail(h, SDB_EARENA_FULL, "message arena full: unconsumed messages pin the log (receive or enlarge arena_bytes)");
  }
  *base_out = tail;
  return SDB_OK;
}

struct SendArrays {
  const uint32_t* sender; const uint32_t* second; const uint8_t* prio; const uint8_t* type; const uint16_t* len;
  const uint64_t* payload_off; const double* timestamp;
  const uint8_t* kind = nullptr;      // mixed batches only
  uint32_t n_lists = 0;               // mixed batches only
  uint32_t* p2p_list = nullptr;       // mixed batches: where single receivers are appended (pinned list staging)
  uint64_t p2p_list_base = 0, p2p_list_cap = 0, p2p_list_used = 0;
  // list batches: recipients of list li are list_pool[lbegin[li] .. lend[li]) (identical lists of a batch are stored once)
  const uint64_t* lbegin = nullptr; const uint64_t* lend = nullptr;
  const uint32_t* list_pool = nullptr;
  uint32_t* lp_out = nullptr; uint64_t lp_cap = 0;     // staging of the list-parallel index tables (nullptr: not offered)
};

// Build descriptors into `out` (host).  kind 0: second = receiver; 1: second = group idx;
// 2: list (list_off/list base offsets supplied).  Returns totals through the staged object.
// A broadcast list longer than SDB_LIST_CHUNK recipients is emitted as several descriptors (same payload, same
// sequence number, consecutive arena regions) so that one warp never owns a million-record fan-out; `out` therefore
// holds s->n >= n descriptors, bounded by `out_cap`.
static const uint32_t SDB_LIST_CHUNK = 1024;
int rebuild_inverse(sdb_ctx* h);

int build_descs(sdb_ctx* h, uint32_t batch_kind, uint32_t n, SendArrays& a, const uint64_t* list_off,
                uint64_t payload_bytes, sdb_send_desc* out, uint64_t out_cap, sdb_staged* s, uint32_t* gs_out = nullptr,
                uint32_t* n_gs_out = nullptr) {
  uint64_t rec = 0, gran = 0, group_recs = 0, list_copies = 0;
  uint32_t n_list_sends = 0;
  std::vector<uint32_t>& first_desc = h->lp_first;    // list sends: index of the send's first chunk descriptor
  uint64_t o = 0;                      // output cursor (descriptors emitted)
  uint32_t max_padlen = 0;
  uint32_t n_group_sends = 0, n_other = 0;
  const uint32_t A = h->cfg.max_agents;
  if (h->sharded) return fail(h, SDB_EINVAL, "sharded handle: sends go through sdb_export_*_batch + sdb_import_wire_*");
  // shared payloads (pure group batches above the pull threshold, exclusive groups, payloads <= 512 bytes, groups of at
  // most 65535 members): a send's region is mcount headers + ONE payload; decided before any arena offset is handed out
  bool shared = false;
  if (h->shared_cfg && batch_kind == 1 && gs_out && h->cfg.fanout_variant >= 2) {
    if (h->memb_dirty) { int rc0 = rebuild_inverse(h); if (rc0 != SDB_OK) return rc0; }
    uint64_t recs = 0; bool ok = h->n_shared_agents == 0 && h->n_excl_groups != 0;
    for (uint32_t i = 0; i < n && ok; ++i) {
      const uint32_t g = a.second[i];
      if (g >= h->cfg.max_groups || !h->gdefined[g] || h->gcount[g] > 65535u || a.len[i] > 512u) ok = false;
      else recs += h->gcount[g];
    }
    shared = ok && recs >= SDB_PULL_THRESHOLD;
  }
  for (uint32_t i = 0; i < n; ++i) {
    sdb_send_desc d;
    std::memset(&d, 0, sizeof(d));
    const uint32_t len = a.len[i];
    if (len > h->cfg.max_payload_bytes) return fail(h, SDB_EINVAL, "payload longer than max_payload_bytes");
    const uint64_t off = a.payload_off ? a.payload_off[i] : 0;
    if (off & 15u) return fail(h, SDB_EINVAL, "payload_off must be 16-byte aligned");
    if (off + len > payload_bytes) return fail(h, SDB_EINVAL, "payload_off + len beyond payload_bytes");
    if (a.sender[i] >= A) return fail(h, SDB_EINVAL, "sender index out of range");
    if ((a.prio ? a.prio[i] : 1) > 3) return fail(h, SDB_EINVAL, "priority out of range");
    const uint32_t pl = pad32(len);
    max_padlen = std::max(max_padlen, pl);
    d.payload_off = off;
    d.timestamp = a.timestamp ? a.timestamp[i] : 0.0;
    d.sender = a.sender[i];
    d.len = static_cast<uint16_t>(len);
    d.prio = a.prio ? a.prio[i] : 1;
    d.type = a.type ? a.type[i] : 0;
    d.rgran = 1u + pl / SDB_GRANULE;
    d.group = SDB_NO_GROUP;
    if (rec > 0xFFFFFFFFull || gran > 0xFFFFFFFFull) return fail(h, SDB_ECAPACITY, "batch exceeds 2^32 records/granules");
    d.rec0 = static_cast<uint32_t>(rec);
    d.gran0 = static_cast<uint32_t>(gran);
    uint32_t kind = batch_kind;
    if (batch_kind == 3) {
      kind = a.kind[i];
      if (kind > 2) return fail(h, SDB_EINVAL, "kind must be 0, 1 or 2");
      if (kind == 0) {      // p2p inside a mixed batch: a one-entry temporary list through the fan-out kernel
        if (a.second[i] >= A) return fail(h, SDB_EINVAL, "receiver index out of range");
        if (a.p2p_list_used >= a.p2p_list_cap) return fail(h, SDB_ECAPACITY, "recipient lists exceed list_pool_entries");
        a.p2p_list[a.p2p_list_used] = a.second[i];
        h->n_agents = std::max(h->n_agents, a.second[i] + 1);
        d.mstart = static_cast<uint32_t>(a.p2p_list_base + a.p2p_list_used); d.mcount = 1; d.flags = SDB_DESC_LIST_TEMP;
        a.p2p_list_used++;
        rec += 1; gran += d.rgran;
        ++n_other;
        if (o >= out_cap) return fail(h, SDB_ECAPACITY, "descriptor staging exhausted");
        out[o++] = d;
        continue;
      }
      if (kind == 2 && a.second[i] >= a.n_lists) return fail(h, SDB_EINVAL, "list number out of range");
    }
    if (kind == 0) {
      if (a.second[i] >= A) return fail(h, SDB_EINVAL, "receiver index out of range");
      h->n_agents = std::max(h->n_agents, a.second[i] + 1);     // auto-registration of the receiver (M:423-427)
      d.mstart = a.second[i]; d.mcount = 1; d.flags = 0;
      rec += 1; gran += d.rgran;
    } else if (kind == 1) {
      const uint32_t g = a.second[i];
      if (g >= h->cfg.max_groups || !h->gdefined[g]) return fail(h, SDB_ENOTFOUND, "unknown group index");
      if (h->sharded) return fail(h, SDB_EINVAL, "sharded handle: use sdb_export_group_batch + sdb_import_wire_batches");
      d.mstart = static_cast<uint32_t>(h->gstart[g]); d.mcount = h->gcount[g]; d.group = g;
      d.flags = SDB_DESC_SKIP_SENDER;
      rec += d.mcount;
      gran += shared ? (d.mcount ? static_cast<uint64_t>(d.mcount) + d.rgran - 1u : 0ull) : static_cast<uint64_t>(d.mcount) * d.rgran;
      group_recs += d.mcount; ++n_group_sends;
    } else {
      const uint32_t li = batch_kind == 3 ? a.second[i] : i;
      const uint64_t b = a.lbegin ? a.lbegin[li] : list_off[li], e = a.lend ? a.lend[li] : list_off[li + 1];
      list_copies += e - b; ++n_list_sends;
      if (first_desc.size() < n) first_desc.resize(n);
      first_desc[i] = static_cast<uint32_t>(o);
      if (e < b || e - b > 0xFFFFFFFFull) return fail(h, SDB_EINVAL, "bad list_off");
      d.flags = SDB_DESC_SHARED_SEQ | SDB_DESC_LIST_TEMP;
      const uint64_t cnt = e - b;
      rec += 1; ++n_other;
      for (uint64_t c0 = 0; c0 == 0 || c0 < cnt; c0 += SDB_LIST_CHUNK) {      // at least one descriptor, even for an empty list
        if (gran > 0xFFFFFFFFull) return fail(h, SDB_ECAPACITY, "batch exceeds 2^32 granules");
        d.mstart = static_cast<uint32_t>(b + c0);
        d.mcount = static_cast<uint32_t>(std::min<uint64_t>(SDB_LIST_CHUNK, cnt - c0));
        d.gran0 = static_cast<uint32_t>(gran);
        gran += static_cast<uint64_t>(d.mcount) * d.rgran;
        if (o >= out_cap) return fail(h, SDB_ECAPACITY, "descriptor staging exhausted");
        out[o++] = d;
      }
      continue;
    }
    if (kind != 1) ++n_other;
    if (o >= out_cap) return fail(h, SDB_ECAPACITY, "descriptor staging exhausted");
    out[o++] = d;
  }
  const uint32_t n_out = static_cast<uint32_t>(o);
  s->kind = batch_kind; s->n = n_out; s->total_recs = rec; s->total_grans = gran; s->max_padlen = max_padlen;
  s->has_pull = false; s->has_atomic = true; s->ranked = false; s->list_pull = false; s->shared = shared;
  if (n_list_sends == n && n && list_copies >= SDB_PULL_THRESHOLD && a.lp_out && a.lbegin && a.list_pool) {
    // pure broadcast batch: if the lists it names are pairwise disjoint (and free of repeats), every member of a list
    // receives exactly the sends naming it, in send order -> owner-computes index build (k_list_index), no atomics.
    // Lists are identified by where they start in the pool (identical lists were stored once by the caller).
    std::vector<std::pair<uint64_t, uint32_t>> key(n);                    // (list begin, send)
    for (uint32_t i = 0; i < n; ++i) { const uint32_t li = batch_kind == 3 ? a.second[i] : i; key[i] = {a.lbegin[li], i}; }
    std::stable_sort(key.begin(), key.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    uint32_t nd = 0;
    for (uint32_t k = 0; k < n; ++k) if (k == 0 || key[k].first != key[k - 1].first) ++nd;
    const uint64_t words = 3ull * nd + 1 + n;
    bool ok = nd <= 65535u && words <= a.lp_cap;
    if (ok) {
      uint32_t* start = a.lp_out; uint32_t* count = start + nd; uint32_t* off = count + nd; uint32_t* idx = off + nd + 1;
      if (h->rank_cnt.size() != A) h->rank_cnt.assign(A, 0u);
      std::vector<uint32_t>& seen = h->rank_cnt;                            // all zero between batches
      uint32_t dcur = 0, max_count = 0;
      std::vector<std::pair<uint64_t, uint64_t>> marked;
      for (uint32_t k = 0; k < n && ok; ++k) {
        const uint32_t i = key[k].second, li = batch_kind == 3 ? a.second[i] : i;
        if (k == 0 || key[k].first != key[k - 1].first) {
          const uint64_t b = a.lbegin[li], e = a.lend[li];
          if (k) ++dcur;
          start[dcur] = static_cast<uint32_t>(b); count[dcur] = static_cast<uint32_t>(e - b); off[dcur] = k;
          max_count = std::max(max_count, count[dcur]);
          marked.push_back({b, e});
          for (uint64_t x = b; x < e; ++x) { uint32_t& c = seen[a.list_pool[x]]; if (c) { ok = false; marked.back().second = x; break; } c = 1; }
        } else if (a.lend[li] - a.lbegin[li] != count[dcur]) ok = false;     // same start, different length: not the same list
        idx[k] = first_desc[i];
      }
      off[nd] = n;
      for (const auto& m : marked) for (uint64_t x = m.first; x < m.second; ++x) seen[a.list_pool[x]] = 0;
      if (ok) {
        for (uint32_t i = 0; i < n_out; ++i) out[i].flags |= SDB_DESC_PULL;
        s->list_pull = true; s->has_atomic = false; s->lp_nd = nd; s->lp_max_count = max_count; s->lp_words = static_cast<uint32_t>(words);
      }
    }
  }
  if (batch_kind == 0 && n_out) {
    // pure point-to-point batch: number the sends of every receiver 0, 1, 2 .. in send order, so that the enqueue
    // kernel places ring entries at ctail + rank without atomics and in final order (no commit sort)
    if (h->rank_cnt.size() != A) h->rank_cnt.assign(A, 0u);
    std::vector<uint32_t>& cnt = h->rank_cnt;
    for (uint32_t i = 0; i < n_out; ++i) { out[i].pad0 = cnt[out[i].mstart]++; out[i].flags |= SDB_DESC_RANKED; }
    for (uint32_t i = n_out; i-- > 0;) {
      uint32_t& c = cnt[out[i].mstart];
      if (c) { out[i].flags |= SDB_DESC_RANK_LAST; c = 0; }      // first seen from the back = highest rank; resets the table
    }
    s->ranked = true;
  }
  if (gs_out && group_recs >= SDB_PULL_THRESHOLD) {
    // bucket the group sends by group (counting sort, ascending send index inside a bucket)
    const uint32_t G = h->cfg.max_groups;
    uint32_t* off = gs_out;            // [G + 1]
    uint32_t* idx = gs_out + G + 1;    // [n_group_sends]
    std::memset(off, 0, (static_cast<size_t>(G) + 1) * sizeof(uint32_t));
    for (uint32_t i = 0; i < n_out; ++i) if (out[i].flags & SDB_DESC_SKIP_SENDER) off[out[i].group + 1]++;
    for (uint32_t g = 0; g < G; ++g) off[g + 1] += off[g];
    std::vector<uint32_t> cur(off, off + G);
    for (uint32_t i = 0; i < n_out; ++i) if (out[i].flags & SDB_DESC_SKIP_SENDER) { idx[cur[out[i].group]++] = i; out[i].flags |= SDB_DESC_PULL; }
    s->has_pull = true; s->has_atomic = n_other != 0;
    *n_gs_out = n_group_sends;
  }
  return SDB_OK;
}

// agent -> (group, position) lists, CSR over agents, from the authoritative host group table
int rebuild_inverse(sdb_ctx* h) {
  const uint32_t A = h->cfg.max_agents, G = h->cfg.max_groups;
  std::vector<uint32_t> off(static_cast<size_t>(A) + 1, 0);
  uint64_t total = 0;
  for (uint32_t g = 0; g < G; ++g) if (h->gdefined[g]) { for (uint32_t m : h->ghost[g]) off[m + 1]++; total += h->ghost[g].size(); }
  if (total > h->cfg.member_pool_entries) return fail(h, SDB_ECAPACITY, "memberships exceed member_pool_entries");
  for (uint32_t a = 0; a < A; ++a) off[a + 1] += off[a];
  std::vector<uint32_t> grp(total ? total : 1), pos(total ? total : 1), cur(off.begin(), off.end() - 1);
  for (uint32_t g = 0; g < G; ++g) if (h->gdefined[g]) {
    const std::vector<uint32_t>& ms = h->ghost[g];
    for (uint32_t j = 0; j < ms.size(); ++j) { const uint32_t c = cur[ms[j]]++; grp[c] = g; pos[c] = j; }
  }
  // a group is "exclusive" when each of its members belongs to that group only (and appears in it once): all members
  // then receive exactly the group's sends and the index build runs group-parallel (k_pull_index_group)
  std::vector<uint8_t> excl(G, 0);
  uint32_t n_excl = 0, n_shared = 0;
  for (uint32_t g = 0; g < G; ++g) if (h->gdefined[g]) {
    bool ex = true;
    for (uint32_t m : h->ghost[g]) if (off[m + 1] - off[m] != 1) { ex = false; break; }
    excl[g] = ex ? 1 : 0; n_excl += ex ? 1 : 0;
  }
  for (uint32_t a = 0; a < A; ++a) {
    const uint32_t nk = off[a + 1] - off[a];
    if (nk > 1 || (nk == 1 && !excl[grp[off[a]]])) ++n_shared;
  }
  h->n_excl_groups = n_excl; h->n_shared_agents = n_shared;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaMemcpy(h->gexcl_dev, excl.data(), G, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->memb_off_dev, off.data(), off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  if (total) {
    CUDA_TRY(h, cudaMemcpy(h->memb_grp_dev, grp.data(), total * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->memb_pos_dev, pos.data(), total * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  h->memb_dirty = false;
  return SDB_OK;
}

// device copy of the local group table (start, count per group): index build and cross-shard import read it
int ensure_ltab(sdb_ctx* h) {
  if (!h->ltab_dirty) return SDB_OK;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  std::vector<uint32_t> st(h->cfg.max_groups), ct(h->cfg.max_groups);
  h->max_lcount = 0;
  for (uint32_t g = 0; g < h->cfg.max_groups; ++g) {
    st[g] = static_cast<uint32_t>(h->gstart[g]); ct[g] = h->gdefined[g] ? h->gcount[g] : 0;
    h->max_lcount = std::max(h->max_lcount, ct[g]);
  }
  CUDA_TRY(h, cudaMemcpy(h->lstart_dev, st.data(), st.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->lcount_dev, ct.data(), ct.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  h->ltab_dirty = false;
  return SDB_OK;
}

int submit(sdb_ctx* h, const sdb_staged* s, uint64_t* seq_base_out) {
  { int rcc = cursor_to_host(h); if (rcc != SDB_OK) return rcc; }
  if (seq_base_out) *seq_base_out = h->next_seq;
  if (s->n == 0) return SDB_OK;
  if (s->has_pull && h->memb_dirty) { int rc0 = rebuild_inverse(h); if (rc0 != SDB_OK) return rc0; }
  if (s->shared && (h->n_shared_agents != 0 || h->n_excl_groups == 0))
    return fail(h, SDB_EINVAL, "batch staged before a membership change (its groups are no longer exclusive): stage it again");
  if (s->has_pull) { int rc1 = ensure_ltab(h); if (rc1 != SDB_OK) return rc1; }
  uint64_t base = 0;
  int rc = arena_reserve(h, s->total_grans, &base);
  if (rc != SDB_OK) return rc;
  cudaError_t e;
  // a pure group batch has two independent halves: the payload fan-out (arena) and the index build (rings);
  // the latter is latency-bound and runs on a side stream beside the former
  // (measured: the two kernels contend for issue slots and HBM - 0.995 vs 1.006 ms per step - and the
  // fan-out's own duration stretches by 50 %, so the overlap stays off unless SDB_OVERLAP_INDEX=1)
  static const bool allow_overlap = getenv("SDB_OVERLAP_INDEX") != nullptr;
  const bool overlap = allow_overlap && s->has_pull && !s->has_atomic && s->kind != 0;
  if (overlap) {
    CUDA_TRY(h, cudaEventRecord(h->ev_fork, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    sdb_pull_view pv{h->memb_off_dev, h->memb_grp_dev, h->memb_pos_dev, s->gs_off_dev, s->gs_idx_dev, 1u, 0u, 0u, nullptr};
    int nlp = 0;
    e = sdb_launch_pull(&h->view, &pv, s->descs_dev, h->n_agents, h->cfg.max_groups, h->gexcl_dev, h->n_excl_groups,
                        h->n_shared_agents, h->lstart_dev, h->lcount_dev, base, 1, h->side, &h->prof, &nlp, nullptr, s->shared ? 1u : 0u);
    if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("index launch: ") + cudaGetErrorString(e));
    CUDA_TRY(h, cudaEventRecord(h->ev_join, h->side));
    h->launches += nlp;
  }
  if (s->kind == 0) {
    e = sdb_launch_p2p(&h->view, s->descs_dev, s->n, s->payload_dev, h->next_seq, base, h->sm_count, h->stream, &h->prof, s->max_padlen);
    if (e == cudaSuccess && s->ranked) {      // entries are in send order already: publish them, nothing to sort
      e = sdb_launch_commit_ranked(&h->view, s->descs_dev, s->n, h->n_agents, h->stream, &h->prof);
      h->launches += 1;
    }
  } else if (s->shared) {
    e = sdb_launch_fanout_shared(&h->view, s->descs_dev, s->n, s->payload_dev, h->next_seq, base, 0u, h->sm_count, h->stream, &h->prof, nullptr);
  } else {
    e = sdb_launch_fanout(&h->view, s->descs_dev, s->n, s->payload_dev, s->list_dev, h->next_seq, base,
                          s->max_padlen, static_cast<int>(h->cfg.fanout_variant), h->sm_count, h->stream, &h->prof, nullptr, 0u);
  }
  h->launches += 1;
  if (e == cudaSuccess && overlap) {
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
  } else if (e == cudaSuccess && s->has_pull) {
    sdb_pull_view pv{h->memb_off_dev, h->memb_grp_dev, h->memb_pos_dev, s->gs_off_dev, s->gs_idx_dev, 1u, 0u, 0u, nullptr};
    int nlp = 0;
    e = sdb_launch_pull(&h->view, &pv, s->descs_dev, h->n_agents, h->cfg.max_groups, h->gexcl_dev, h->n_excl_groups,
                        h->n_shared_agents, h->lstart_dev, h->lcount_dev, base, s->has_atomic ? 0 : 1, h->stream, &h->prof, &nlp, nullptr,
                        s->shared ? 1u : 0u);
    h->launches += nlp;
  }
  if (e == cudaSuccess && s->list_pull) {
    e = sdb_launch_list_index(&h->view, s->descs_dev, s->list_dev, h->lp_dev, s->lp_nd, s->lp_max_count, base, h->stream, &h->prof);
    h->launches += 1;
  }
  if (e == cudaSuccess && s->has_atomic && !s->ranked) {
    e = sdb_launch_commit(&h->view, h->n_agents, static_cast<uint32_t>(base), h->rx_big_list, h->rx_big_count + 1, h->sm_count, h->stream, &h->prof, nullptr);
    h->launches += 2;
  }
  if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("enqueue launch: ") + cudaGetErrorString(e));
  h->next_seq += s->total_recs;
  h->arena_tail = base + s->total_grans;
  h->dev_stale = true;
  h->stream_busy = true;
  return SDB_OK;
}

int send_common(sdb_ctx* h, uint32_t kind, uint32_t n, SendArrays& a, const uint64_t* list_off,
                const uint32_t* list_idx, const uint8_t* payload, uint64_t payload_bytes, uint64_t* seq_base_out) {
  if (!h) return SDB_EINVAL;
  if (n == 0) { int rcc = cursor_to_host(h); if (seq_base_out) *seq_base_out = h->next_seq; return rcc; }
  if (!a.sender || !a.len || (kind != 2 && !a.second) || (kind == 3 && !a.kind)) return fail(h, SDB_EINVAL, "null array");
  if (n > h->cfg.max_batch_sends) return fail(h, SDB_ECAPACITY, "n exceeds max_batch_sends");
  if (payload_bytes > h->cfg.max_batch_payload) return fail(h, SDB_ECAPACITY, "payload_bytes exceeds max_batch_payload");
  if (payload_bytes && !payload) return fail(h, SDB_EINVAL, "null payload");
  CUDA_TRY(h, cudaEventSynchronize(h->staging_free));     // pinned descriptor staging is reusable
  sdb_staged* s = &h->scratch;
  uint64_t list_total = 0;
  if (kind == 2 || kind == 3) {
    const uint32_t nl = kind == 2 ? n : a.n_lists;
    list_total = (nl && list_off) ? list_off[nl] : 0;
    const uint64_t cap = h->cfg.list_pool_entries;
    // identical consecutive lists (the same "everybody" list passed with every broadcast) are stored and uploaded once
    h->lp_begin.assign(nl, 0); h->lp_end.assign(nl, 0);
    uint64_t used = 0, prev_b = 0, prev_n = 0; bool have_prev = false;
    for (uint32_t li = 0; li < nl; ++li) {
      const uint64_t b = list_off[li], e = list_off[li + 1];
      if (e < b || e > list_total) return fail(h, SDB_EINVAL, "bad list_off");
      const uint64_t cnt = e - b;
      if (have_prev && cnt == prev_n && cnt && std::memcmp(list_idx + b, h->list_host + prev_b, cnt * sizeof(uint32_t)) == 0) {
        h->lp_begin[li] = prev_b; h->lp_end[li] = prev_b + cnt;
        continue;
      }
      for (uint64_t k = b; k < e; ++k) {
        if (list_idx[k] >= h->cfg.max_agents) return fail(h, SDB_EINVAL, "recipient index out of range");
        h->n_agents = std::max(h->n_agents, list_idx[k] + 1);
      }
      if (used + cnt > cap) return fail(h, SDB_ECAPACITY, "recipient lists exceed list_pool_entries");
      if (cnt) std::memcpy(h->list_host + used, list_idx + b, cnt * sizeof(uint32_t));
      h->lp_begin[li] = used; h->lp_end[li] = used + cnt;
      prev_b = used; prev_n = cnt; have_prev = true;
      used += cnt;
    }
    list_total = used;
    a.lbegin = h->lp_begin.data(); a.lend = h->lp_end.data(); a.list_pool = h->list_host;
    if (!h->lp_host) {
      h->lp_cap = 4ull * h->cfg.max_batch_sends + 8;
      CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->lp_host), h->lp_cap * sizeof(uint32_t), cudaHostAllocDefault));
      CUDA_TRY(h, dmalloc(&h->lp_dev, h->lp_cap));
    }
    a.lp_out = h->lp_host; a.lp_cap = h->lp_cap;
    a.p2p_list = h->list_host + list_total; a.p2p_list_base = list_total; a.p2p_list_cap = cap - list_total;
  }
  uint32_t n_gs = 0;
  int rc = build_descs(h, kind, n, a, list_off, payload_bytes, h->descs_host, h->desc_cap, s, h->gs_host, &n_gs);
  if (rc != SDB_OK) return rc;
  if (s->has_pull) {
    const size_t G1 = static_cast<size_t>(h->cfg.max_groups) + 1;
    CUDA_TRY(h, cudaMemcpyAsync(s->gs_off_dev, h->gs_host, G1 * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(s->gs_idx_dev, h->gs_host + G1, static_cast<size_t>(n_gs) * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  }
  if (s->list_pull)
    CUDA_TRY(h, cudaMemcpyAsync(h->lp_dev, h->lp_host, static_cast<size_t>(s->lp_words) * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  list_total += a.p2p_list_used;
  if (list_total)
    CUDA_TRY(h, cudaMemcpyAsync(s->list_dev, h->list_host, list_total * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(s->descs_dev, h->descs_host, static_cast<size_t>(s->n) * sizeof(sdb_send_desc),
                              cudaMemcpyHostToDevice, h->stream));
  if (payload_bytes)   // straight from the caller's buffer (pinned memory makes this a true async DMA)
    CUDA_TRY(h, cudaMemcpyAsync(s->payload_dev, payload, payload_bytes, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->staging_free, h->stream));
  return submit(h, s, seq_base_out);
}

}  // namespace

// =============================================================================================
extern "C" {

int sdb_abi_version(void) { return static_cast<int>(SDB_ABI_VERSION); }

const char* sdb_last_error(sdb_handle h) { return h ? h->err.c_str() : "null handle"; }

int sdb_create(const sdb_config* cfg, sdb_handle* out) {
  if (!cfg || !out) return SDB_EINVAL;
  *out = nullptr;
  if (cfg->struct_bytes != sizeof(sdb_config)) return SDB_EINVAL;
  sdb_ctx* h = new (std::nothrow) sdb_ctx();
  if (!h) return SDB_ENOMEM;
  *out = h;                       // returned even on failure so sdb_last_error works; caller destroys
  h->cfg = *cfg;
  sdb_config& c = h->cfg;
  if (c.num_shards == 0) c.num_shards = 1;
  if (!is_pow2(c.ring_slots) || c.ring_slots < 2 || c.ring_slots > (1u << 24)) return fail(h, SDB_EINVAL, "ring_slots must be a power of two in [2, 2^24]");
  if (!is_pow2(c.arena_bytes) || c.arena_bytes < (1u << 16) || c.arena_bytes > (1ull << 37)) return fail(h, SDB_EINVAL, "arena_bytes must be a power of two in [64 KiB, 128 GiB]");
  if (c.max_agents == 0 || c.max_agents > (1u << 30)) return fail(h, SDB_EINVAL, "max_agents out of range");
  if (c.max_payload_bytes == 0 || c.max_payload_bytes > 65504) return fail(h, SDB_EINVAL, "max_payload_bytes must be in [1, 65504]");
  if (c.max_groups == 0) c.max_groups = 1;
  if (c.member_pool_entries == 0) c.member_pool_entries = 1;
  if (c.max_backends == 0) c.max_backends = 256;
  if (c.max_batch_sends == 0) c.max_batch_sends = 65536;
  if (c.max_batch_payload == 0) c.max_batch_payload = static_cast<uint64_t>(c.max_batch_sends) * pad32(std::min(c.max_payload_bytes, 256u));
  if (c.max_recv_records == 0) c.max_recv_records = 1u << 20;
  if (c.max_recv_records >= 0x7FFFFFF0ull) return fail(h, SDB_EINVAL, "max_recv_records too large (< 2^31)");
  if (c.max_recv_payload == 0) c.max_recv_payload = c.max_recv_records * 256ull;
  if (c.list_pool_entries == 0) c.list_pool_entries = 2ull * c.max_agents + 1024;
  if (c.fanout_variant > 3) return fail(h, SDB_EINVAL, "fanout_variant must be 0, 1, 2 or 3");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, SDB_ECUDA, "no CUDA device: swarmdb_b200 has no CPU fallback");
  if (c.device < 0 || c.device >= ndev) return fail(h, SDB_EINVAL, "device ordinal out of range");
  CUDA_TRY(h, cudaSetDevice(c.device));
  {
    // L2 fetch granularity 64 bytes (device-wide limit of this process; SDB_L2_FETCH=128 restores the default, 0 leaves
    // the limit alone).  The receive gather reads 288-byte records at 32-byte alignment in receiver order: with the
    // default 128-byte granularity every record costs three full lines (1.37x its bytes, ncu); at 64 bytes the
    // gather's DRAM reads drop from 1.66 GB to 1.40 GB per c2 step and it runs 4.5 % faster (0.464 -> 0.443 ms), the
    // fan-out loses 1 % - measured on B200, profiles/README.md.
    const char* fg = getenv("SDB_L2_FETCH");
    const int gran = fg ? atoi(fg) : 64;
    if (gran > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, static_cast<size_t>(gran));
  }
  {
    // shared payloads for group sends above the pull threshold (sdb_common.cuh): SDB_SHARED_PAYLOAD=0 keeps one
    // payload copy per recipient in the arena log
    const char* sp = getenv("SDB_SHARED_PAYLOAD");
    h->shared_cfg = sp ? atoi(sp) != 0 : true;
  }
  CUDA_TRY(h, sdb_send_prepare_device());
  CUDA_TRY(h, sdb_recv_prepare_device());
  cudaDeviceProp prop;
  CUDA_TRY(h, cudaGetDeviceProperties(&prop, c.device));
  h->sm_count = prop.multiProcessorCount;
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->staging_free, cudaEventDisableTiming));
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));

  const size_t A = c.max_agents, R = c.ring_slots;
  h->arena_grans = c.arena_bytes / SDB_GRANULE;
  CUDA_TRY(h, dmalloc(&h->arena, c.arena_bytes));
  CUDA_TRY(h, dmalloc(&h->ring_hdr, A));
  CUDA_TRY(h, dmalloc(&h->ring, A * R));
  CUDA_TRY(h, dmalloc(&h->members, c.member_pool_entries));
  CUDA_TRY(h, dmalloc(&h->ctr, 1));
  CUDA_TRY(h, cudaMemsetAsync(h->ring_hdr, 0, A * sizeof(sdb_ring_hdr), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->ring, 0xFF, A * R * sizeof(uint2), h->stream));          // every slot "consumed"
  CUDA_TRY(h, cudaMemsetAsync(h->ctr, 0, sizeof(sdb_dev_counters), h->stream));
  CUDA_TRY(h, dmalloc(&h->ovf_log, SDB_OVF_LOG)); CUDA_TRY(h, dmalloc(&h->ovf_seq, SDB_OVF_LOG));

  // send staging
  h->desc_cap = static_cast<uint64_t>(c.max_batch_sends) + c.list_pool_entries / SDB_LIST_CHUNK + 1;
  CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->descs_host), static_cast<size_t>(h->desc_cap) * sizeof(sdb_send_desc), cudaHostAllocDefault));
  CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->list_host), c.list_pool_entries * sizeof(uint32_t), cudaHostAllocDefault));
  CUDA_TRY(h, dmalloc(&h->scratch.descs_dev, h->desc_cap));
  CUDA_TRY(h, dmalloc(&h->scratch.payload_dev, c.max_batch_payload + 64));
  CUDA_TRY(h, dmalloc(&h->scratch.list_dev, c.list_pool_entries));
  CUDA_TRY(h, dmalloc(&h->scratch.gs_off_dev, static_cast<size_t>(c.max_groups) + 1));
  CUDA_TRY(h, dmalloc(&h->scratch.gs_idx_dev, c.max_batch_sends));
  CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->gs_host), (static_cast<size_t>(c.max_groups) + 1 + c.max_batch_sends) * sizeof(uint32_t), cudaHostAllocDefault));
  CUDA_TRY(h, dmalloc(&h->memb_off_dev, static_cast<size_t>(c.max_agents) + 1));
  CUDA_TRY(h, dmalloc(&h->memb_grp_dev, c.member_pool_entries));
  CUDA_TRY(h, dmalloc(&h->memb_pos_dev, c.member_pool_entries));
  CUDA_TRY(h, dmalloc(&h->gexcl_dev, c.max_groups));
  CUDA_TRY(h, cudaMemsetAsync(h->gexcl_dev, 0, c.max_groups, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->scratch.payload_dev, 0, c.max_batch_payload + 64, h->stream));

  // receive scratch
  const size_t tiles = (A + SDB_SCAN_TILE - 1) / SDB_SCAN_TILE + 1;
  CUDA_TRY(h, dmalloc(&h->rx_agent, A)); CUDA_TRY(h, dmalloc(&h->rx_cnt, A));
  CUDA_TRY(h, dmalloc(&h->rx_rec_local, A)); CUDA_TRY(h, dmalloc(&h->rx_rec_tops, tiles));
  const size_t rtiles = (c.max_recv_records + SDB_SCAN_TILE - 1) / SDB_SCAN_TILE + 1;
  CUDA_TRY(h, dmalloc(&h->rx_plan, c.max_recv_records + 4)); CUDA_TRY(h, dmalloc(&h->rx_rec_off, A));
  CUDA_TRY(h, dmalloc(&h->rx_plan_tops, rtiles));
  CUDA_TRY(h, dmalloc(&h->rx_lb, sdb_lb_words(static_cast<uint32_t>(A / 64 + 2)) + 4));
  CUDA_TRY(h, dmalloc(&h->rx_totals, 8));
  CUDA_TRY(h, dmalloc(&h->rx_big_list, A)); CUDA_TRY(h, dmalloc(&h->rx_big_count, 4));
  CUDA_TRY(h, dmalloc(&h->rx_count, A));
  CUDA_TRY(h, dmalloc(&h->rx_hdr, c.max_recv_records));
  h->pay_cap_gran = (c.max_recv_payload + SDB_GRANULE - 1) / SDB_GRANULE;
  if (h->pay_cap_gran >= 0x7FFFFFF0ull) return fail(h, SDB_EINVAL, "max_recv_payload too large (< 64 GiB per call)");
  CUDA_TRY(h, dmalloc(&h->rx_payload, h->pay_cap_gran * SDB_GRANULE));
  CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->totals_host), 8 * sizeof(unsigned long long), cudaHostAllocDefault));
  h->small_bytes = 64 + 1024ull * (32 + pad32(c.max_payload_bytes));
  CUDA_TRY(h, dmalloc(&h->rx_small, h->small_bytes));
  CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->small_host), h->small_bytes, cudaHostAllocDefault));

  // backends
  CUDA_TRY(h, dmalloc(&h->be_weight, c.max_backends));
  CUDA_TRY(h, dmalloc(&h->be_load, c.max_backends));
  CUDA_TRY(h, dmalloc(&h->be_scratch, static_cast<size_t>(c.max_backends) * 4 + 16));
  CUDA_TRY(h, dmalloc(&h->be_logtab, 257));
  {
    uint32_t tab[257];
    sdb_build_log2_table(tab);
    CUDA_TRY(h, cudaMemcpyAsync(h->be_logtab, tab, sizeof(tab), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  }

  h->gstart.assign(c.max_groups, 0); h->gcount.assign(c.max_groups, 0); h->gdefined.assign(c.max_groups, 0);
  h->ghost.resize(c.max_groups); h->gfull.resize(c.max_groups); h->gpos.resize(c.max_groups);
  h->gcount_full.assign(c.max_groups, 0);
  h->shard_of.assign(c.max_agents, static_cast<uint8_t>(c.shard_id));
  if (c.shard_id >= c.num_shards || c.num_shards > 255) return fail(h, SDB_EINVAL, "shard_id / num_shards out of range");
  CUDA_TRY(h, dmalloc(&h->member_pos_dev, c.member_pool_entries));
  CUDA_TRY(h, dmalloc(&h->lstart_dev, c.max_groups)); CUDA_TRY(h, dmalloc(&h->lcount_dev, c.max_groups));
  CUDA_TRY(h, cudaMemset(h->lcount_dev, 0, static_cast<size_t>(c.max_groups) * sizeof(uint32_t)));
  CUDA_TRY(h, cudaMemset(h->lstart_dev, 0, static_cast<size_t>(c.max_groups) * sizeof(uint32_t)));
  {
    const uint64_t cap64 = static_cast<uint64_t>(c.num_shards) * c.max_batch_sends;
    if (cap64 > 0x7FFFFFFFull) return fail(h, SDB_EINVAL, "num_shards * max_batch_sends too large");
    h->xs_cap = static_cast<uint32_t>(cap64);
    const size_t n = h->xs_cap, G1 = static_cast<size_t>(c.max_groups) + 1;
    const size_t GB = G1 * c.num_shards + 2;                 // (group, source) pairs
    const size_t wt = n / SDB_SCAN_TILE + 2, gt = GB / SDB_SCAN_TILE + 2;
    CUDA_TRY(h, dmalloc(&h->xs_w, n)); CUDA_TRY(h, dmalloc(&h->xs_w_local, n)); CUDA_TRY(h, dmalloc(&h->xs_w_tops, wt));
    CUDA_TRY(h, dmalloc(&h->xs_gs_cnt, GB)); CUDA_TRY(h, dmalloc(&h->xs_gs_local, GB)); CUDA_TRY(h, dmalloc(&h->xs_gs_tops, gt));
    CUDA_TRY(h, dmalloc(&h->xs_gs_cur, GB)); CUDA_TRY(h, dmalloc(&h->xs_gs_off, G1)); CUDA_TRY(h, dmalloc(&h->xs_gs_idx, n));
    CUDA_TRY(h, dmalloc(&h->xs_descs, n));
    CUDA_TRY(h, dmalloc(&h->xs_lw, n)); CUDA_TRY(h, dmalloc(&h->xs_lw_local, n)); CUDA_TRY(h, dmalloc(&h->xs_lw_tops, wt));
    CUDA_TRY(h, dmalloc(&h->xs_tab, 1));
    CUDA_TRY(h, dmalloc(&h->xs_gs_off_src, G1 * c.num_shards)); CUDA_TRY(h, dmalloc(&h->xs_gs_idx_src, n));
    CUDA_TRY(h, dmalloc(&h->xs_first, SDB_MAX_SRC + 1)); CUDA_TRY(h, dmalloc(&h->xs_lb, sdb_lb_words(static_cast<uint32_t>(n / 256 + 2)) + 4));
    CUDA_TRY(h, dmalloc(&h->cursor_dev, 1)); CUDA_TRY(h, dmalloc(&h->bb_dev, 1)); CUDA_TRY(h, dmalloc(&h->xs_hdrs, SDB_MAX_SRC));
    CUDA_TRY(h, cudaMemset(h->cursor_dev, 0, sizeof(sdb_cursor))); CUDA_TRY(h, cudaMemset(h->bb_dev, 0, sizeof(sdb_batch_base)));
    {
      auto& S0 = h->xset[0];
      S0.descs = h->xs_descs; S0.gs_off_src = h->xs_gs_off_src; S0.gs_idx_src = h->xs_gs_idx_src; S0.first = h->xs_first;
      S0.lb = h->xs_lb; S0.hdrs = h->xs_hdrs; S0.tmp_list = h->scratch.list_dev;
      CUDA_TRY(h, dmalloc(&S0.totals, 1));
      CUDA_TRY(h, cudaEventCreateWithFlags(&S0.ready, cudaEventDisableTiming));
    }
    CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->cursor_host), sizeof(sdb_cursor), cudaHostAllocDefault));
    std::memset(h->cursor_host, 0, sizeof(sdb_cursor));
    h->xs_meta_stride = wire_meta_bytes(c.max_batch_sends, c.max_groups);
    CUDA_TRY(h, dmalloc(&h->xs_meta, h->xs_meta_stride * c.num_shards));
    CUDA_TRY(h, dmalloc(&h->shard_of_dev, c.max_agents));
    CUDA_TRY(h, cudaMemset(h->shard_of_dev, static_cast<int>(c.shard_id), c.max_agents));
    CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->wire_host), wire_meta_bytes(c.max_batch_sends, c.max_groups), cudaHostAllocDefault));
    CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->hdrs_host), static_cast<size_t>(c.num_shards) * sizeof(sdb_wire_header), cudaHostAllocDefault));
  }

  sdb_dev_view& v = h->view;
  v.arena = h->arena; v.ring_hdr = h->ring_hdr; v.ring = h->ring; v.members = h->members; v.member_pos = h->member_pos_dev; v.ctr = h->ctr;
  v.ovf_log = h->ovf_log;
  v.gmask = h->arena_grans - 1; v.ring_slots = c.ring_slots; v.ring_shift = ilog2(c.ring_slots); v.max_agents = c.max_agents;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return SDB_OK;
}

int sdb_destroy(sdb_handle h) {
  if (!h) return SDB_EINVAL;
  cudaSetDevice(h->cfg.device);
  ls_stop(h);
  if (h->ls_stream) cudaStreamDestroy(h->ls_stream);
  if (h->ls_mbox) cudaFreeHost(h->ls_mbox);
  if (h->ls_out) cudaFreeHost(h->ls_out);
  if (h->ls_plan) cudaFree(h->ls_plan);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->pf_stream) { cudaStreamSynchronize(h->pf_stream); cudaStreamDestroy(h->pf_stream); }
  if (h->ev_pf_fork) cudaEventDestroy(h->ev_pf_fork);
  {
    auto& S1 = h->xset[1];
    void* own[] = {S1.descs, S1.gs_off_src, S1.gs_idx_src, S1.first, S1.lb, S1.hdrs, S1.tmp_list, S1.totals, h->xset[0].totals};
    for (void* p : own) if (p) cudaFree(p);
    for (auto& S : h->xset) if (S.ready) cudaEventDestroy(S.ready);
  }
  void* dev[] = {h->arena, h->ring_hdr, h->ring, h->members, h->ctr, h->gexcl_dev, h->rx_rec_off, h->rx_lb,
                 h->xs_gs_off_src, h->xs_gs_idx_src, h->xs_first, h->xs_lb, h->cursor_dev, h->bb_dev, h->xs_hdrs, h->owned_dev, h->agent_backend_dev,
                 h->scratch.descs_dev, h->scratch.payload_dev, h->scratch.list_dev, h->scratch.gs_off_dev,
                 h->scratch.gs_idx_dev, h->memb_off_dev, h->memb_grp_dev, h->memb_pos_dev, h->member_pos_dev,
                 h->lstart_dev, h->lcount_dev, h->xs_w, h->xs_w_local, h->xs_w_tops, h->xs_gs_cnt, h->xs_gs_local,
                 h->xs_gs_tops, h->xs_gs_cur, h->xs_gs_off, h->xs_gs_idx, h->xs_descs, h->xs_lw, h->xs_lw_local,
                 h->xs_lw_tops, h->xs_tab, h->xs_meta, h->shard_of_dev, h->rx_agent, h->rx_cnt,
                 h->rx_rec_local, h->rx_rec_tops, h->rx_plan, h->rx_plan_tops,
                 h->rx_totals, h->rx_big_list, h->rx_big_count, h->rx_count, h->rx_hdr, h->rx_payload,
                 h->be_weight, h->be_load, h->be_scratch, h->be_logtab, h->be_req_cost, h->be_out, h->digest};
  for (void* p : dev) if (p) cudaFree(p);
  if (h->descs_host) cudaFreeHost(h->descs_host);
  if (h->list_host) cudaFreeHost(h->list_host);
  if (h->lp_host) cudaFreeHost(h->lp_host);
  if (h->lp_dev) cudaFree(h->lp_dev);
  if (h->ovf_log) cudaFree(h->ovf_log);
  if (h->ovf_seq) cudaFree(h->ovf_seq);
  if (h->gs_host) cudaFreeHost(h->gs_host);
  if (h->wire_host) cudaFreeHost(h->wire_host);
  if (h->hdrs_host) cudaFreeHost(h->hdrs_host);
  if (h->cursor_host) cudaFreeHost(h->cursor_host);
  if (h->totals_host) cudaFreeHost(h->totals_host);
  if (h->small_host) cudaFreeHost(h->small_host);
  if (h->rx_small) cudaFree(h->rx_small);
  if (h->staging_free) cudaEventDestroy(h->staging_free);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->prof.cap) {
    for (int i = 0; i < h->prof.cap; ++i) { cudaEventDestroy(h->prof.ev_a[i]); cudaEventDestroy(h->prof.ev_b[i]); }
    delete[] h->prof.kind; delete[] h->prof.ev_a; delete[] h->prof.ev_b;
  }
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return SDB_OK;
}

int sdb_set_stream(sdb_handle h, void* cuda_stream) {
  if (!h) return SDB_EINVAL;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  return SDB_OK;
}

int sdb_sync(sdb_handle h) {
  if (!h) return SDB_EINVAL;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->stream_busy = false;
  return SDB_OK;
}

int sdb_profile(sdb_handle h, int enable) {
  if (!h) return SDB_EINVAL;
  sdb_profiler& p = h->prof;
  if (enable && !p.cap) {
    p.cap = 8192;
    p.kind = new int[p.cap]; p.ev_a = new cudaEvent_t[p.cap]; p.ev_b = new cudaEvent_t[p.cap];
    for (int i = 0; i < p.cap; ++i) { CUDA_TRY(h, cudaEventCreate(&p.ev_a[i])); CUDA_TRY(h, cudaEventCreate(&p.ev_b[i])); }
  }
  p.enabled = enable ? 1 : 0;
  if (enable) p.n = 0;
  return SDB_OK;
}

int sdb_profile_read(sdb_handle h, double* ms_out, uint64_t* count_out) {
  if (!h || !ms_out || !count_out) return SDB_EINVAL;
  for (int k = 0; k < SDB_PK_N; ++k) { ms_out[k] = 0.0; count_out[k] = 0; }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  sdb_profiler& p = h->prof;
  for (int i = 0; i < p.n; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.ev_a[i], p.ev_b[i]) == cudaSuccess && p.kind[i] < SDB_PK_N) {
      ms_out[p.kind[i]] += ms; count_out[p.kind[i]] += 1;
    }
  }
  p.n = 0;
  return SDB_OK;
}

int sdb_get_stats(sdb_handle h, sdb_stats* out) {
  if (!h || !out) return SDB_EINVAL;
  { int rcc = cursor_to_host(h); if (rcc != SDB_OK) return rcc; }
  sdb_dev_counters c;
  CUDA_TRY(h, cudaMemcpyAsync(&c, h->ctr, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  out->next_seq = h->next_seq; out->enqueued = c.enqueued; out->delivered = c.delivered;
  out->ring_overflow = c.ring_overflow; out->skipped_sender = c.skipped_sender;
  out->arena_tail_bytes = h->arena_tail * SDB_GRANULE; out->arena_floor_bytes = h->arena_floor * SDB_GRANULE;
  out->n_agents = h->n_agents; out->kernel_launches = h->launches; out->backend_picks = h->picks;
  return SDB_OK;
}

int sdb_debug_set_arena_pos(sdb_handle h, uint64_t granules) {
  if (!h) return SDB_EINVAL;
  sdb_stats st;
  int rc = sdb_get_stats(h, &st);
  if (rc != SDB_OK) return rc;
  if (st.enqueued != st.delivered + 0 && st.enqueued - st.delivered != 0) return fail(h, SDB_EINVAL, "messages pending");
  h->arena_tail = granules; h->arena_floor = granules;
  h->dev_stale = true;
  return SDB_OK;
}

int sdb_advance_seq(sdb_handle h, uint64_t next_seq) {
  if (!h) return SDB_EINVAL;
  { int rcc = cursor_to_host(h); if (rcc != SDB_OK) return rcc; }
  if (next_seq > h->next_seq) { h->next_seq = next_seq; h->dev_stale = true; }
  return SDB_OK;
}

int sdb_register_agents(sdb_handle h, uint32_t n, const uint32_t* agent_idx) {
  if (!h || (n && !agent_idx)) return SDB_EINVAL;
  for (uint32_t i = 0; i < n; ++i) {
    if (agent_idx[i] >= h->cfg.max_agents) return fail(h, SDB_EINVAL, "agent index >= max_agents");
    h->n_agents = std::max(h->n_agents, agent_idx[i] + 1);
  }
  return SDB_OK;
}

int sdb_deregister_agents(sdb_handle h, uint32_t n, const uint32_t* agent_idx) {
  if (!h || (n && !agent_idx)) return SDB_EINVAL;
  for (uint32_t i = 0; i < n; ++i)
    if (agent_idx[i] >= h->cfg.max_agents) return fail(h, SDB_EINVAL, "agent index >= max_agents");
  return SDB_OK;   // ring and read position are kept (consumer-group offsets survive close, App. A rule 11)
}

// (re)compute the members this shard owns for group g and their positions in the full list
static void localize_group(sdb_ctx* h, uint32_t g) {
  const std::vector<uint32_t>& full = h->gfull[g];
  std::vector<uint32_t>& loc = h->ghost[g];
  std::vector<uint32_t>& pos = h->gpos[g];
  loc.clear(); pos.clear();
  for (uint32_t j = 0; j < full.size(); ++j)
    if (!h->sharded || h->shard_of[full[j]] == h->cfg.shard_id) { loc.push_back(full[j]); pos.push_back(j); }
  h->gcount_full[g] = static_cast<uint32_t>(full.size());
}

// write every defined group's local list into the device pools from scratch
static int upload_all_groups(sdb_ctx* h) {
  uint64_t used = 0;
  for (uint32_t k = 0; k < h->cfg.max_groups; ++k) if (h->gdefined[k]) used += h->ghost[k].size();
  if (used > h->cfg.member_pool_entries) return fail(h, SDB_ECAPACITY, "member pool exhausted");
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  std::vector<uint32_t> flat, fpos; flat.reserve(used); fpos.reserve(used);
  for (uint32_t k = 0; k < h->cfg.max_groups; ++k) if (h->gdefined[k]) {
    h->gstart[k] = flat.size(); h->gcount[k] = static_cast<uint32_t>(h->ghost[k].size());
    flat.insert(flat.end(), h->ghost[k].begin(), h->ghost[k].end());
    fpos.insert(fpos.end(), h->gpos[k].begin(), h->gpos[k].end());
  }
  if (!flat.empty()) {
    CUDA_TRY(h, cudaMemcpy(h->members, flat.data(), flat.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->member_pos_dev, fpos.data(), fpos.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  h->member_used = flat.size();
  h->memb_dirty = true; h->ltab_dirty = true;
  return SDB_OK;
}

int sdb_create_group(sdb_handle h, uint32_t g, uint32_t n_members, const uint32_t* member_idx) {
  if (!h || (n_members && !member_idx)) return SDB_EINVAL;
  if (g >= h->cfg.max_groups) return fail(h, SDB_EINVAL, "group index >= max_groups");
  for (uint32_t i = 0; i < n_members; ++i)
    if (member_idx[i] >= h->cfg.max_agents) return fail(h, SDB_EINVAL, "member index >= max_agents");
  h->memb_dirty = true; h->ltab_dirty = true;
  h->gfull[g].assign(member_idx, member_idx + n_members);
  localize_group(h, g);
  h->gdefined[g] = 1;
  const std::vector<uint32_t>& loc = h->ghost[g];
  for (uint32_t m : loc) h->n_agents = std::max(h->n_agents, m + 1);   // members will receive: keep them inside the sweeps
  if (h->member_used + loc.size() > h->cfg.member_pool_entries) return upload_all_groups(h);   // compact the pool
  h->gstart[g] = h->member_used; h->gcount[g] = static_cast<uint32_t>(loc.size());
  if (!loc.empty()) {
    // pageable source: cudaMemcpyAsync stages it before returning, ordered after earlier kernels on the stream
    CUDA_TRY(h, cudaMemcpyAsync(h->members + h->member_used, loc.data(), loc.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->member_pos_dev + h->member_used, h->gpos[g].data(), loc.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  }
  h->member_used += loc.size();
  if (h->member_used > 0xFFFFFFFFull) return fail(h, SDB_ECAPACITY, "member pool beyond 2^32 entries");
  return SDB_OK;
}

uint64_t sdb_wire_bytes(sdb_handle h, uint32_t max_sends, uint64_t max_payload_bytes) {
  // header + descriptors + group buckets + payload (+ slack); broadcast recipient lists count towards max_payload_bytes
  const uint32_t G = h ? h->cfg.max_groups : 0;
  // ... + the cross-rank control block (sdb_wire_ctrl) in the last 128 bytes, outside everything an export writes
  return wire_meta_bytes(max_sends, G) + ((max_payload_bytes + 256 + 127) & ~127ull) + sizeof(sdb_wire_ctrl);
}

int sdb_set_agent_shards(sdb_handle h, uint32_t n, const uint8_t* shard_of) {
  if (!h || (n && !shard_of)) return SDB_EINVAL;
  if (n > h->cfg.max_agents) return fail(h, SDB_EINVAL, "n > max_agents");
  for (uint32_t i = 0; i < n; ++i) {
    if (shard_of[i] >= h->cfg.num_shards) return fail(h, SDB_EINVAL, "shard id >= num_shards");
    h->shard_of[i] = shard_of[i];
  }
  h->sharded = h->cfg.num_shards > 1;
  if (h->sharded) h->n_agents = h->cfg.max_agents;   // receivers are named by other ranks: sweeps cover the whole index space
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaMemcpy(h->shard_of_dev, h->shard_of.data(), h->cfg.max_agents, cudaMemcpyHostToDevice));
  {
    std::vector<uint32_t> owned;
    for (uint32_t i = 0; i < h->cfg.max_agents; ++i) if (h->shard_of[i] == h->cfg.shard_id) owned.push_back(i);
    if (!h->owned_dev) CUDA_TRY(h, dmalloc(&h->owned_dev, h->cfg.max_agents));
    if (!owned.empty()) CUDA_TRY(h, cudaMemcpy(h->owned_dev, owned.data(), owned.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    h->n_owned = static_cast<uint32_t>(owned.size());
  }
  // groups created earlier are re-filtered against the new ownership map
  bool any = false;
  for (uint32_t g = 0; g < h->cfg.max_groups; ++g) if (h->gdefined[g]) { localize_group(h, g); any = true; }
  return any ? upload_all_groups(h) : SDB_OK;
}

// kind == nullptr: every send is a group send (target = group index)
static int export_common(sdb_ctx* h, uint32_t n, const uint32_t* sender, const uint8_t* kind, const uint32_t* target,
                         uint32_t n_lists, const uint64_t* list_off, const uint32_t* list_idx, const uint8_t* prio,
                         const uint8_t* type, const uint16_t* len, const uint64_t* payload_off, const uint8_t* payload,
                         uint64_t payload_bytes, const double* timestamp, void* wire_dev, uint64_t wire_cap,
                         uint64_t explicit_seq_base = 0) {
  if (!h || !wire_dev) return SDB_EINVAL;
  if (n && (!sender || !target || !len)) return fail(h, SDB_EINVAL, "null array");
  if (n > h->cfg.max_batch_sends) return fail(h, SDB_ECAPACITY, "n exceeds max_batch_sends");
  if (payload_bytes && !payload) return fail(h, SDB_EINVAL, "null payload");
  const uint64_t n_list = (n_lists && list_off) ? list_off[n_lists] : 0;
  if (n_list > h->cfg.list_pool_entries) return fail(h, SDB_ECAPACITY, "recipient lists exceed list_pool_entries");
  const uint32_t G = h->cfg.max_groups;
  const uint64_t desc_off = sizeof(sdb_wire_header);
  const uint64_t gso_off = desc_off + static_cast<uint64_t>(n) * sizeof(sdb_send_desc);
  const uint64_t gsi_off = gso_off + (static_cast<uint64_t>(G) + 1) * 4;
  const uint64_t l_off = gsi_off + static_cast<uint64_t>(n) * 4;            // room for every send being a group send
  const uint64_t pay_off = (l_off + n_list * sizeof(uint32_t) + 63) & ~63ull;
  if (pay_off + payload_bytes + 64 + sizeof(sdb_wire_ctrl) > wire_cap) return fail(h, SDB_ECAPACITY, "wire buffer too small (sdb_wire_bytes)");
  CUDA_TRY(h, cudaEventSynchronize(h->staging_free));
  sdb_wire_header* wh = reinterpret_cast<sdb_wire_header*>(h->wire_host);
  sdb_send_desc* wd = reinterpret_cast<sdb_send_desc*>(h->wire_host + desc_off);
  uint32_t* gso = reinterpret_cast<uint32_t*>(h->wire_host + gso_off);
  uint32_t* gsi = reinterpret_cast<uint32_t*>(h->wire_host + gsi_off);
  std::memset(gso, 0, (static_cast<size_t>(G) + 1) * 4);
  uint32_t n_group = 0;
  std::memset(wh, 0, sizeof(*wh));
  uint64_t rec = 0; uint32_t max_padlen = 0, n_other = 0;
  for (uint64_t k = 0; k < n_list; ++k) if (list_idx[k] >= h->cfg.max_agents) return fail(h, SDB_EINVAL, "recipient index out of range");
  for (uint32_t i = 0; i < n; ++i) {
    sdb_send_desc d; std::memset(&d, 0, sizeof(d));
    const uint32_t k = kind ? kind[i] : 1u;
    if (sender[i] >= h->cfg.max_agents) return fail(h, SDB_EINVAL, "sender index out of range");
    if (len[i] > h->cfg.max_payload_bytes) return fail(h, SDB_EINVAL, "payload longer than max_payload_bytes");
    const uint64_t off = payload_off ? payload_off[i] : 0;
    if ((off & 15u) || off + len[i] > payload_bytes) return fail(h, SDB_EINVAL, "bad payload_off");
    if ((prio ? prio[i] : 1) > 3) return fail(h, SDB_EINVAL, "priority out of range");
    if (rec > 0xFFFFFFFFull) return fail(h, SDB_ECAPACITY, "batch exceeds 2^32 records");
    const uint32_t pl = pad32(len[i]);
    max_padlen = std::max(max_padlen, pl);
    d.payload_off = off; d.timestamp = timestamp ? timestamp[i] : 0.0;
    d.sender = sender[i]; d.group = SDB_NO_GROUP; d.len = len[i]; d.prio = prio ? prio[i] : 1; d.type = type ? type[i] : 0;
    d.rgran = 1u + pl / SDB_GRANULE; d.rec0 = static_cast<uint32_t>(rec);
    if (k == 1) {
      const uint32_t g = target[i];
      if (g >= h->cfg.max_groups || !h->gdefined[g]) return fail(h, SDB_ENOTFOUND, "unknown group index");
      d.group = g; rec += h->gcount_full[g];
      gso[g + 1]++; ++n_group;
    } else if (k == 0) {
      if (target[i] >= h->cfg.max_agents) return fail(h, SDB_EINVAL, "receiver index out of range");
      d.mstart = target[i]; d.mcount = 1; d.flags = SDB_DESC_P2P; rec += 1; ++n_other;
    } else if (k == 2) {
      if (target[i] >= n_lists) return fail(h, SDB_EINVAL, "list number out of range");
      const uint64_t b = list_off[target[i]], e = list_off[target[i] + 1];
      if (e < b || e > n_list) return fail(h, SDB_EINVAL, "bad list_off");
      d.mstart = static_cast<uint32_t>(b); d.mcount = static_cast<uint32_t>(e - b); d.flags = SDB_DESC_LIST_TEMP; rec += 1; ++n_other;
    } else return fail(h, SDB_EINVAL, "kind must be 0, 1 or 2");
    wd[i] = d;
  }
  // bucket the group sends by group: counting sort, ascending send index inside a bucket
  for (uint32_t g = 0; g < G; ++g) gso[g + 1] += gso[g];
  {
    std::vector<uint32_t> cur(gso, gso + G);
    for (uint32_t i = 0; i < n; ++i) if (wd[i].group != SDB_NO_GROUP) gsi[cur[wd[i].group]++] = i;
  }
  wh->magic = SDB_WIRE_MAGIC; wh->n_sends = n; wh->total_recs = rec; wh->payload_bytes = payload_bytes;
  wh->desc_off = desc_off; wh->payload_off = pay_off; wh->max_padlen = max_padlen; wh->n_other = n_other;
  wh->list_off = l_off; wh->n_list = static_cast<uint32_t>(n_list);
  wh->n_group_sends = n_group; wh->gs_off_off = gso_off; wh->gs_idx_off = gsi_off; wh->max_groups = G;
  wh->explicit_seq = explicit_seq_base ? 1u : 0u; wh->seq_base = explicit_seq_base;
  uint8_t* w = static_cast<uint8_t*>(wire_dev);
  CUDA_TRY(h, cudaMemcpyAsync(w, h->wire_host, l_off, cudaMemcpyHostToDevice, h->stream));
  if (n_list) {
    std::memcpy(h->list_host, list_idx, n_list * sizeof(uint32_t));
    CUDA_TRY(h, cudaMemcpyAsync(w + l_off, h->list_host, n_list * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  }
  if (payload_bytes) CUDA_TRY(h, cudaMemcpyAsync(w + pay_off, payload, payload_bytes, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->staging_free, h->stream));
  return SDB_OK;
}

int sdb_export_group_batch(sdb_handle h, uint32_t n, const uint32_t* sender, const uint32_t* group_idx, const uint8_t* prio,
                           const uint8_t* type, const uint16_t* len, const uint64_t* payload_off, const uint8_t* payload,
                           uint64_t payload_bytes, const double* timestamp, void* wire_dev, uint64_t wire_cap) {
  return export_common(h, n, sender, nullptr, group_idx, 0, nullptr, nullptr, prio, type, len, payload_off, payload, payload_bytes,
                       timestamp, wire_dev, wire_cap);
}

int sdb_export_mixed_batch(sdb_handle h, uint32_t n, const uint32_t* sender, const uint8_t* kind, const uint32_t* target,
                           uint32_t n_lists, const uint64_t* list_off, const uint32_t* list_idx, const uint8_t* prio,
                           const uint8_t* type, const uint16_t* len, const uint64_t* payload_off, const uint8_t* payload,
                           uint64_t payload_bytes, const double* timestamp, void* wire_dev, uint64_t wire_cap) {
  if (h && n && !kind) return fail(h, SDB_EINVAL, "null kind array");
  return export_common(h, n, sender, kind, target, n_lists, list_off, list_idx, prio, type, len, payload_off, payload, payload_bytes,
                       timestamp, wire_dev, wire_cap);
}

int sdb_export_mixed_batch_seq(sdb_handle h, uint64_t seq_base, uint32_t n, const uint32_t* sender, const uint8_t* kind,
                               const uint32_t* target, uint32_t n_lists, const uint64_t* list_off, const uint32_t* list_idx,
                               const uint8_t* prio, const uint8_t* type, const uint16_t* len, const uint64_t* payload_off,
                               const uint8_t* payload, uint64_t payload_bytes, const double* timestamp, void* wire_dev,
                               uint64_t wire_cap) {
  if (h && n && !kind) return fail(h, SDB_EINVAL, "null kind array");
  if (h && seq_base == 0) return fail(h, SDB_EINVAL, "explicit seq_base must be non-zero");
  return export_common(h, n, sender, kind, target, n_lists, list_off, list_idx, prio, type, len, payload_off, payload, payload_bytes,
                       timestamp, wire_dev, wire_cap, seq_base);
}

// ---- peer-memory transport: export buffers that other ranks map with CUDA IPC ---------------------
int sdb_wire_alloc(sdb_handle h, uint64_t bytes, void** dev_out, void* ipc_handle_out) {
  if (!h || !dev_out || bytes == 0) return SDB_EINVAL;
  void* p = nullptr;
  CUDA_TRY(h, cudaMalloc(&p, bytes));
  CUDA_TRY(h, cudaMemset(p, 0, bytes));
  if (ipc_handle_out) {
    cudaIpcMemHandle_t hd;
    CUDA_TRY(h, cudaIpcGetMemHandle(&hd, p));
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(ipc_handle_out, &hd, sizeof(hd));
  }
  *dev_out = p;
  return SDB_OK;
}

int sdb_wire_open(sdb_handle h, const void* ipc_handle, void** dev_out) {
  if (!h || !ipc_handle || !dev_out) return SDB_EINVAL;
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, ipc_handle, sizeof(hd));
  CUDA_TRY(h, cudaIpcOpenMemHandle(dev_out, hd, cudaIpcMemLazyEnablePeerAccess));
  return SDB_OK;
}

int sdb_wire_close(sdb_handle h, void* dev, int opened) {
  if (!h || !dev) return SDB_EINVAL;
  ls_stop(h);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (opened) CUDA_TRY(h, cudaIpcCloseMemHandle(dev)); else CUDA_TRY(h, cudaFree(dev));
  return SDB_OK;
}

int sdb_import_wire_batches(sdb_handle h, uint32_t n_src, const void* wire_dev_all, uint64_t wire_stride, uint64_t* seq_base_out) {
  if (!h || !wire_dev_all || n_src == 0 || n_src > SDB_MAX_SRC) return SDB_EINVAL;
  const void* ptrs[SDB_MAX_SRC];
  for (uint32_t s = 0; s < n_src; ++s) ptrs[s] = static_cast<const uint8_t*>(wire_dev_all) + static_cast<uint64_t>(s) * wire_stride;
  return sdb_import_wire_ptrs(h, n_src, ptrs, seq_base_out);
}

int sdb_import_wire_ptrs(sdb_handle h, uint32_t n_src, const void* const* wire_ptrs, uint64_t* seq_base_out) {
  if (!h || !wire_ptrs || n_src == 0) return SDB_EINVAL;
  { int rcc = cursor_to_host(h); if (rcc != SDB_OK) return rcc; }
  if (seq_base_out) *seq_base_out = h->next_seq;
  if (n_src > h->cfg.num_shards || n_src > SDB_MAX_SRC) return fail(h, SDB_EINVAL, "n_src > num_shards (or > 16)");
  { int rc1 = ensure_ltab(h); if (rc1 != SDB_OK) return rc1; }
  if (h->memb_dirty) { int rc0 = rebuild_inverse(h); if (rc0 != SDB_OK) return rc0; }
  const uint32_t n_cap = n_src * h->cfg.max_batch_sends;
  sdb_import_args a{};
  for (uint32_t k = 0; k < n_src; ++k) {
    if (!wire_ptrs[k]) return fail(h, SDB_EINVAL, "null wire pointer");
    a.wire[k] = static_cast<const uint8_t*>(wire_ptrs[k]);
    a.meta[k] = a.wire[k];
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, wire_ptrs[k]) == cudaSuccess && pa.type == cudaMemoryTypeDevice && pa.device != h->cfg.device) {
      // header + descriptors of a peer's batch: one DMA copy over NVLink instead of per-thread remote reads;
      // the payload (the bulk) stays remote and is pulled by the fan-out kernel's TMA loads
      uint8_t* dst = h->xs_meta + static_cast<uint64_t>(k) * h->xs_meta_stride;
      CUDA_TRY(h, cudaMemcpyAsync(dst, wire_ptrs[k], h->xs_meta_stride, cudaMemcpyDeviceToDevice, h->stream));
      a.meta[k] = dst;
    }
  }
  a.n_src = n_src; a.max_sends = h->cfg.max_batch_sends; a.tab = h->xs_tab;
  a.lstart = h->lstart_dev; a.lcount = h->lcount_dev; a.max_groups = h->cfg.max_groups;
  a.w = h->xs_w; a.gs_cnt = h->xs_gs_cnt; a.descs = h->xs_descs; a.w_local = h->xs_w_local; a.w_tops = h->xs_w_tops;
  a.gs_off = h->xs_gs_off; a.gs_idx = h->xs_gs_idx;
  a.shard_of = h->shard_of_dev; a.shard_id = h->cfg.shard_id; a.max_agents = h->cfg.max_agents;
  a.lw = h->xs_lw; a.lw_local = h->xs_lw_local; a.lw_tops = h->xs_lw_tops;
  a.tmp_list = h->scratch.list_dev; a.list_cap = static_cast<uint32_t>(std::min<uint64_t>(h->cfg.list_pool_entries, 0xFFFFFFFFull));
  int nl = 0;
  cudaError_t e = sdb_launch_import_measure(&a, n_cap, h->xs_w_local, h->xs_w_tops, h->xs_gs_local, h->xs_gs_tops,
                                            h->xs_gs_cur /* goff */, h->xs_lw_local, h->xs_lw_tops, h->rx_totals + 4, h->stream, &h->prof, &nl);
  if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("import measure: ") + cudaGetErrorString(e));
  // the host needs the arena footprint and the sequence numbers consumed: one small sync
  CUDA_TRY(h, cudaMemcpyAsync(h->totals_host + 4, h->rx_totals + 4, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  for (uint32_t k = 0; k < n_src; ++k)
    CUDA_TRY(h, cudaMemcpyAsync(h->hdrs_host + k, wire_ptrs[k], sizeof(sdb_wire_header), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  uint64_t total_recs = 0, n_other = 0, explicit_end = 0; uint32_t max_padlen = 0;
  if (h->totals_host[5] > a.list_cap) return fail(h, SDB_ECAPACITY, "owned recipients of this import exceed list_pool_entries");
  for (uint32_t s = 0; s < n_src; ++s) {
    const sdb_wire_header& wh = h->hdrs_host[s];
    n_other += wh.n_other;
    if (wh.magic == SDB_WIRE_MAGIC && wh.explicit_seq) explicit_end = std::max<uint64_t>(explicit_end, wh.seq_base + wh.total_recs);
    if (wh.magic != SDB_WIRE_MAGIC) return fail(h, SDB_EINVAL, "wire batch without magic (not exported by sdb_export_group_batch?)");
    if (wh.n_sends > h->cfg.max_batch_sends) return fail(h, SDB_ECAPACITY, "wire batch larger than max_batch_sends");
    total_recs += wh.total_recs; max_padlen = std::max(max_padlen, wh.max_padlen);
  }
  const uint64_t total_grans = h->totals_host[4];
  uint64_t base = 0;
  int rc = arena_reserve(h, total_grans, &base);
  if (rc != SDB_OK) return rc;
  e = sdb_launch_import_localize(&a, n_cap, h->xs_gs_cur /* goff */, h->xs_gs_off, h->stream, &h->prof, &nl);
  static const bool allow_overlap_x = getenv("SDB_OVERLAP_INDEX") != nullptr;
  const bool overlap = allow_overlap_x && n_other == 0;   // pure group traffic: index build beside the fan-out (off by default)
  if (e == cudaSuccess && overlap) {
    CUDA_TRY(h, cudaEventRecord(h->ev_fork, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    sdb_pull_view pv{h->memb_off_dev, h->memb_grp_dev, h->memb_pos_dev, h->xs_gs_off, h->xs_gs_idx, 1u, 0u, 0u, nullptr};
    e = sdb_launch_pull(&h->view, &pv, h->xs_descs, h->n_agents, h->cfg.max_groups, h->gexcl_dev, h->n_excl_groups,
                        h->n_shared_agents, h->lstart_dev, h->lcount_dev, base, 1, h->side, &h->prof, &nl, nullptr, 0u);
    if (e == cudaSuccess) CUDA_TRY(h, cudaEventRecord(h->ev_join, h->side));
  }
  if (e == cudaSuccess)
    e = sdb_launch_fanout(&h->view, h->xs_descs, n_cap, nullptr, h->scratch.list_dev, h->next_seq, base, max_padlen,
                          h->cfg.fanout_variant >= 2 ? 3 : static_cast<int>(h->cfg.fanout_variant), h->sm_count, h->stream, &h->prof, nullptr, 0u);
  if (e == cudaSuccess && overlap) {
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
  } else if (e == cudaSuccess) {
    sdb_pull_view pv{h->memb_off_dev, h->memb_grp_dev, h->memb_pos_dev, h->xs_gs_off, h->xs_gs_idx, 1u, 0u, 0u, nullptr};
    e = sdb_launch_pull(&h->view, &pv, h->xs_descs, h->n_agents, h->cfg.max_groups, h->gexcl_dev, h->n_excl_groups,
                        h->n_shared_agents, h->lstart_dev, h->lcount_dev, base, n_other ? 0 : 1, h->stream, &h->prof, &nl, nullptr, 0u);
  }
  if (e == cudaSuccess && n_other) {     // p2p / broadcast copies claimed their slots with atomics: sort them into place
    e = sdb_launch_commit(&h->view, h->n_agents, static_cast<uint32_t>(base), h->rx_big_list, h->rx_big_count + 1, h->sm_count, h->stream, &h->prof, nullptr);
    h->launches += 2;
  }
  h->launches += nl + 1;
  if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("import launch: ") + cudaGetErrorString(e));
  h->next_seq = std::max<uint64_t>(h->next_seq + (explicit_end ? 0 : total_recs), explicit_end);
  h->arena_tail = base + total_grans;
  h->dev_stale = true;
  h->stream_busy = true;
  return SDB_OK;
}

// ---- flags + asynchronous import -------------------------------------------------------------------------------
static uint32_t* ctrl_word(const void* wire, uint64_t wire_bytes, int which) {
  uint8_t* c = static_cast<uint8_t*>(const_cast<void*>(wire)) + wire_bytes - sizeof(sdb_wire_ctrl);
  return reinterpret_cast<uint32_t*>(c) + which;           // 0 = ready, 1 = done
}

int sdb_wire_publish(sdb_handle h, void* wire_dev, uint64_t wire_bytes, uint32_t step) {
  if (!h || !wire_dev || wire_bytes < sizeof(sdb_wire_ctrl)) return SDB_EINVAL;
  CUDA_TRY(h, sdb_launch_wire_set(ctrl_word(wire_dev, wire_bytes, 0), step, h->stream));
  h->launches += 1;
  return SDB_OK;
}

int sdb_wire_wait_done(sdb_handle h, uint32_t n_src, const void* const* wire_ptrs, uint64_t wire_bytes, uint32_t step) {
  if (!h || !wire_ptrs || n_src == 0 || n_src > SDB_MAX_SRC || wire_bytes < sizeof(sdb_wire_ctrl)) return SDB_EINVAL;
  const uint32_t* flags[SDB_MAX_SRC];
  for (uint32_t k = 0; k < n_src; ++k) flags[k] = ctrl_word(wire_ptrs[k], wire_bytes, 1);
  CUDA_TRY(h, sdb_launch_wire_wait(flags, n_src, step, h->stream, nullptr, nullptr));
  h->launches += 1;
  return SDB_OK;
}

// traffic shapes the device-only import handles (everything else takes the synchronous body between the same flags)
static bool async_fast(const sdb_ctx* h) {
  return pad32(h->cfg.max_payload_bytes) <= 512 && h->n_shared_agents == 0 && h->n_excl_groups != 0 && h->cfg.fanout_variant >= 2;
}

// wait for the sources' flags + fused localize into buffer set S, on stream `st`; `defer`: leave the placement to
// k_import_place (the cursor belongs to the shard's stream)
static int xs_localize(sdb_ctx* h, sdb_ctx::XsSet& S, uint32_t n_src, const void* const* wire_ptrs, uint64_t wire_bytes,
                       uint32_t step, cudaStream_t st, bool defer) {
  const uint32_t* ready[SDB_MAX_SRC];
  for (uint32_t k = 0; k < n_src; ++k) ready[k] = ctrl_word(wire_ptrs[k], wire_bytes, 0);
  const int pw = sdb_prof_begin(&h->prof, SDB_PK_XWAIT, st);
  CUDA_TRY(h, sdb_launch_wire_wait(ready, n_src, step, st, wire_ptrs, S.hdrs));
  sdb_prof_end(&h->prof, pw, st);
  sdb_import2_args a{};
  for (uint32_t k = 0; k < n_src; ++k) a.wire[k] = static_cast<const uint8_t*>(wire_ptrs[k]);
  a.n_src = n_src; a.max_sends = h->cfg.max_batch_sends; a.max_groups = h->cfg.max_groups; a.shard_id = h->cfg.shard_id;
  a.max_agents = h->cfg.max_agents; a.lcount = h->lcount_dev; a.lstart = h->lstart_dev; a.shard_of = h->shard_of_dev;
  a.descs = S.descs; a.gs_off_src = S.gs_off_src; a.gs_idx_src = S.gs_idx_src; a.first = S.first;
  a.tmp_list = S.tmp_list; a.list_cap = static_cast<uint32_t>(std::min<uint64_t>(h->cfg.list_pool_entries, 0x7FFFFFF0ull));
  a.lb = S.lb; a.cur = h->cursor_dev; a.bb = h->bb_dev; a.arena_grans = h->arena_grans;
  a.hdrs = S.hdrs; a.commit_count = h->rx_big_count + 1; a.totals = defer ? S.totals : nullptr;
  // shared payloads: a send's local region is lc headers + one payload (the fast shape already guarantees payloads
  // <= 512 bytes and exclusive groups; local groups must fit the 16-bit header-to-payload distance)
  S.shared = h->shared_cfg && h->max_lcount <= 65535u;
  a.shared = S.shared ? 1u : 0u;
  int nl = 1;
  cudaError_t e = sdb_launch_import_fused(&a, st, &h->prof, &nl);
  h->launches += nl;
  if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("import localize launch: ") + cudaGetErrorString(e));
  return SDB_OK;
}

static int import_args_check(sdb_handle h, uint32_t n_src, const void* const* wire_ptrs, uint64_t wire_bytes) {
  if (!h || !wire_ptrs || n_src == 0 || wire_bytes < sizeof(sdb_wire_ctrl)) return SDB_EINVAL;
  if (n_src > h->cfg.num_shards || n_src > SDB_MAX_SRC) return fail(h, SDB_EINVAL, "n_src > num_shards (or > 16)");
  for (uint32_t k = 0; k < n_src; ++k) if (!wire_ptrs[k]) return fail(h, SDB_EINVAL, "null wire pointer");
  { int rc1 = ensure_ltab(h); if (rc1 != SDB_OK) return rc1; }
  if (h->memb_dirty) { int rc0 = rebuild_inverse(h); if (rc0 != SDB_OK) return rc0; }
  return SDB_OK;
}

int sdb_import_prefetch(sdb_handle h, uint32_t n_src, const void* const* wire_ptrs, uint64_t wire_bytes, uint32_t step) {
  { int rc = import_args_check(h, n_src, wire_ptrs, wire_bytes); if (rc != SDB_OK) return rc; }
  if (!async_fast(h)) return SDB_OK;                  // the synchronous body has nothing to run ahead
  if (!h->pf_stream) {
    auto& S1 = h->xset[1];
    const size_t n = h->xs_cap, G1 = static_cast<size_t>(h->cfg.max_groups) + 1;
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->pf_stream, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_pf_fork, cudaEventDisableTiming));
    CUDA_TRY(h, dmalloc(&S1.descs, n)); CUDA_TRY(h, dmalloc(&S1.gs_off_src, G1 * h->cfg.num_shards)); CUDA_TRY(h, dmalloc(&S1.gs_idx_src, n));
    CUDA_TRY(h, dmalloc(&S1.first, SDB_MAX_SRC + 1)); CUDA_TRY(h, dmalloc(&S1.lb, sdb_lb_words(static_cast<uint32_t>(n / 256 + 2)) + 4));
    CUDA_TRY(h, dmalloc(&S1.hdrs, SDB_MAX_SRC)); CUDA_TRY(h, dmalloc(&S1.tmp_list, h->cfg.list_pool_entries));
    CUDA_TRY(h, dmalloc(&S1.totals, 1));
    CUDA_TRY(h, cudaEventCreateWithFlags(&S1.ready, cudaEventDisableTiming));
  }
  auto& S = h->xset[step & 1u];
  if (S.pending) return S.step == step ? SDB_OK : fail(h, SDB_EINVAL, "a prefetched import of this parity has not been imported yet");
  // the set's previous user (the import two steps ago) and the table uploads above are ordered before the prefetch;
  // what the caller enqueues on the shard's stream AFTER this call (the receive of the current step) runs beside it
  CUDA_TRY(h, cudaEventRecord(h->ev_pf_fork, h->stream));
  CUDA_TRY(h, cudaStreamWaitEvent(h->pf_stream, h->ev_pf_fork, 0));
  { int rc = xs_localize(h, S, n_src, wire_ptrs, wire_bytes, step, h->pf_stream, true); if (rc != SDB_OK) return rc; }
  CUDA_TRY(h, cudaEventRecord(S.ready, h->pf_stream));
  h->launches += 1;
  S.pending = true; S.step = step;
  return SDB_OK;
}

int sdb_import_wire_ptrs_async(sdb_handle h, uint32_t n_src, const void* const* wire_ptrs, uint64_t wire_bytes, uint32_t step) {
  { int rc = import_args_check(h, n_src, wire_ptrs, wire_bytes); if (rc != SDB_OK) return rc; }
  uint32_t* my_done = ctrl_word(wire_ptrs[h->cfg.shard_id < n_src ? h->cfg.shard_id : 0], wire_bytes, 1);
  if (!async_fast(h)) {
    // general traffic shapes (payloads above 512 bytes, agents in several groups): the synchronous import body
    // 1. wait (on the stream, not on the host) until every source's export of this step is complete
    const uint32_t* ready[SDB_MAX_SRC];
    for (uint32_t k = 0; k < n_src; ++k) ready[k] = ctrl_word(wire_ptrs[k], wire_bytes, 0);
    const int pw = sdb_prof_begin(&h->prof, SDB_PK_XWAIT, h->stream);
    CUDA_TRY(h, sdb_launch_wire_wait(ready, n_src, step, h->stream, wire_ptrs, h->xs_hdrs));
    sdb_prof_end(&h->prof, pw, h->stream);
    h->launches += 1;
    int rc = sdb_import_wire_ptrs(h, n_src, wire_ptrs, nullptr);
    if (rc != SDB_OK) return rc;
    CUDA_TRY(h, sdb_launch_wire_set(my_done, step, h->stream));
    h->launches += 1;
    return SDB_OK;
  }
  auto& S = h->xset[h->pf_stream ? (step & 1u) : 0u];
  const bool prefetched = S.pending && S.step == step;
  // everything happens on the device
  { int rc2 = cursor_to_dev(h); if (rc2 != SDB_OK) return rc2; }
  CUDA_TRY(h, sdb_launch_arena_floor_cur(&h->view, h->n_agents, h->cursor_dev, h->stream));
  if (S.pending) { CUDA_TRY(h, cudaStreamWaitEvent(h->stream, S.ready, 0)); S.pending = false; }
  if (prefetched) {
    // 1. the flags were awaited and the batch localized ahead of time (sdb_import_prefetch): place it now
    CUDA_TRY(h, sdb_launch_import_place(S.totals, h->cursor_dev, h->bb_dev, h->arena_grans,
                                        static_cast<uint32_t>(std::min<uint64_t>(h->cfg.list_pool_entries, 0x7FFFFFF0ull)),
                                        h->rx_big_count + 1, h->stream));
    h->launches += 1;
  } else {
    // 1. wait (on the stream, not on the host) until every source's export of this step is complete; 2. localize + place
    int rc = xs_localize(h, S, n_src, wire_ptrs, wire_bytes, step, h->stream, false);
    if (rc != SDB_OK) return rc;
    h->launches += 1;
  }
  int nl = 0;
  const uint32_t n_cap = n_src * h->cfg.max_batch_sends;
  cudaError_t e = cudaSuccess;
  if (S.shared) {            // group sends: headers + one payload per send; the span kernel only serves p2p / broadcast descriptors
    e = sdb_launch_fanout_shared(&h->view, S.descs, n_cap, nullptr, 0, 0, h->cfg.num_shards >= 4 ? 1u : 0u, h->sm_count, h->stream,
                                 &h->prof, h->bb_dev);
    ++nl;
  }
  if (e == cudaSuccess)
    e = sdb_launch_fanout(&h->view, S.descs, n_cap, nullptr, S.tmp_list, 0, 0, pad32(h->cfg.max_payload_bytes), 3,
                          h->sm_count, h->stream, &h->prof, h->bb_dev, S.shared ? 1u : 0u);
  if (e == cudaSuccess) {
    sdb_pull_view pv{h->memb_off_dev, h->memb_grp_dev, h->memb_pos_dev, S.gs_off_src, S.gs_idx_src, n_src,
                     h->cfg.max_groups + 1, h->cfg.max_batch_sends, S.first};
    e = sdb_launch_pull(&h->view, &pv, S.descs, h->n_agents, h->cfg.max_groups, h->gexcl_dev, h->n_excl_groups, 0,
                        h->lstart_dev, h->lcount_dev, 0, 0, h->stream, &h->prof, &nl, h->bb_dev, S.shared ? 1u : 0u);
  }
  // point-to-point / broadcast copies claimed their slots with atomics: sort them into place (the kernels return at
  // once when the import carried none); the group-parallel build leaves ctail to the commit
  if (e == cudaSuccess)
    e = sdb_launch_commit(&h->view, h->n_agents, 0, h->rx_big_list, h->rx_big_count + 1, h->sm_count, h->stream, &h->prof, h->bb_dev);
  h->launches += nl + 3;
  if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("async import launch: ") + cudaGetErrorString(e));
  h->host_stale = true; h->stream_busy = true;
  // 3. tell the exporters that this rank is done reading their buffers of this parity
  CUDA_TRY(h, sdb_launch_wire_set(my_done, step, h->stream));
  h->launches += 1;
  return SDB_OK;
}

int sdb_overflow_log(sdb_handle h, uint32_t cap, uint32_t* agent_out, uint64_t* seq_out, uint32_t* n_out, uint64_t* dropped_out) {
  if (!h || !n_out || (cap && (!agent_out || !seq_out))) return SDB_EINVAL;
  *n_out = 0;
  if (dropped_out) *dropped_out = 0;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  unsigned long long logged = 0;
  CUDA_TRY(h, cudaMemcpy(&logged, &h->ctr->ovf_logged, sizeof(logged), cudaMemcpyDeviceToHost));
  if (dropped_out) *dropped_out = logged;
  if (logged == 0) return SDB_OK;
  const uint32_t n = static_cast<uint32_t>(std::min<unsigned long long>(std::min<unsigned long long>(logged, SDB_OVF_LOG), cap));
  if (n) {
    cudaError_t e = sdb_launch_ovf_resolve(&h->view, n, h->ovf_seq, h->stream);
    if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("overflow log: ") + cudaGetErrorString(e));
    std::vector<uint2> log(n);
    CUDA_TRY(h, cudaMemcpyAsync(log.data(), h->ovf_log, n * sizeof(uint2), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(seq_out, h->ovf_seq, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (uint32_t i = 0; i < n; ++i) agent_out[i] = log[i].x;
  }
  CUDA_TRY(h, cudaMemsetAsync(&h->ctr->ovf_logged, 0, sizeof(unsigned long long), h->stream));
  *n_out = n;
  return SDB_OK;
}

int sdb_send_batch(sdb_handle h, uint32_t n, const uint32_t* sender, const uint32_t* receiver, const uint8_t* prio,
                   const uint8_t* type, const uint16_t* len, const uint64_t* payload_off, const uint8_t* payload,
                   uint64_t payload_bytes, const double* timestamp, uint64_t* seq_base_out) {
  SendArrays a{sender, receiver, prio, type, len, payload_off, timestamp};
  return send_common(h, 0, n, a, nullptr, nullptr, payload, payload_bytes, seq_base_out);
}

int sdb_send_group_batch(sdb_handle h, uint32_t n, const uint32_t* sender, const uint32_t* group_idx, const uint8_t* prio,
                         const uint8_t* type, const uint16_t* len, const uint64_t* payload_off, const uint8_t* payload,
                         uint64_t payload_bytes, const double* timestamp, uint64_t* seq_base_out) {
  SendArrays a{sender, group_idx, prio, type, len, payload_off, timestamp};
  return send_common(h, 1, n, a, nullptr, nullptr, payload, payload_bytes, seq_base_out);
}

int sdb_send_list_batch(sdb_handle h, uint32_t n, const uint32_t* sender, const uint64_t* list_off, const uint32_t* list_idx,
                        const uint8_t* prio, const uint8_t* type, const uint16_t* len, const uint64_t* payload_off,
                        const uint8_t* payload, uint64_t payload_bytes, const double* timestamp, uint64_t* seq_base_out) {
  if (n && (!list_off || (list_off[n] && !list_idx))) return fail(h, SDB_EINVAL, "null list arrays");
  SendArrays a{sender, nullptr, prio, type, len, payload_off, timestamp};
  return send_common(h, 2, n, a, list_off, list_idx, payload, payload_bytes, seq_base_out);
}

int sdb_send_mixed_batch(sdb_handle h, uint32_t n, const uint32_t* sender, const uint8_t* kind, const uint32_t* target,
                         uint32_t n_lists, const uint64_t* list_off, const uint32_t* list_idx, const uint8_t* prio,
                         const uint8_t* type, const uint16_t* len, const uint64_t* payload_off, const uint8_t* payload,
                         uint64_t payload_bytes, const double* timestamp, uint64_t* seq_base_out) {
  if (!h) return SDB_EINVAL;
  if (n_lists && (!list_off || (list_off[n_lists] && !list_idx))) return fail(h, SDB_EINVAL, "null list arrays");
  SendArrays a{sender, target, prio, type, len, payload_off, timestamp};
  a.kind = kind; a.n_lists = n_lists;
  return send_common(h, 3, n, a, list_off, list_idx, payload, payload_bytes, seq_base_out);
}

int sdb_stage_batch(sdb_handle h, uint32_t kind, uint32_t n, const uint32_t* sender, const uint32_t* second,
                    const uint8_t* prio, const uint8_t* type, const uint16_t* len, const uint64_t* payload_off,
                    const uint8_t* payload, uint64_t payload_bytes, const double* timestamp, sdb_staged_t* out) {
  if (!h || !out || kind > 1) return SDB_EINVAL;
  *out = nullptr;
  if (n == 0 || !sender || !second || !len) return fail(h, SDB_EINVAL, "empty or null batch");
  sdb_staged* s = new (std::nothrow) sdb_staged();
  if (!s) return SDB_ENOMEM;
  s->owns = true;
  std::vector<sdb_send_desc> descs(n);
  std::vector<uint32_t> gs(static_cast<size_t>(h->cfg.max_groups) + 1 + n);
  uint32_t n_gs = 0;
  SendArrays a{sender, second, prio, type, len, payload_off, timestamp};
  int rc = build_descs(h, kind, n, a, nullptr, payload_bytes, descs.data(), n, s, gs.data(), &n_gs);
  if (rc != SDB_OK) { delete s; return rc; }
  cudaError_t e = dmalloc(&s->descs_dev, n);
  if (e == cudaSuccess && s->has_pull) {
    const size_t G1 = static_cast<size_t>(h->cfg.max_groups) + 1;
    e = dmalloc(&s->gs_off_dev, G1);
    if (e == cudaSuccess) e = dmalloc(&s->gs_idx_dev, n_gs ? n_gs : 1);
    if (e == cudaSuccess) e = cudaMemcpy(s->gs_off_dev, gs.data(), G1 * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_gs) e = cudaMemcpy(s->gs_idx_dev, gs.data() + G1, static_cast<size_t>(n_gs) * sizeof(uint32_t), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = dmalloc(&s->payload_dev, payload_bytes + 64);
  if (e == cudaSuccess) e = cudaMemset(s->payload_dev, 0, payload_bytes + 64);
  if (e == cudaSuccess) e = cudaMemcpy(s->descs_dev, descs.data(), static_cast<size_t>(n) * sizeof(sdb_send_desc), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && payload_bytes) e = cudaMemcpy(s->payload_dev, payload, payload_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (s->descs_dev) cudaFree(s->descs_dev);
    if (s->payload_dev) cudaFree(s->payload_dev);
    if (s->gs_off_dev) cudaFree(s->gs_off_dev);
    if (s->gs_idx_dev) cudaFree(s->gs_idx_dev);
    delete s;
    return fail(h, SDB_ECUDA, std::string("stage: ") + cudaGetErrorString(e));
  }
  *out = s;
  return SDB_OK;
}

int sdb_submit_staged(sdb_handle h, sdb_staged_t s, uint64_t* seq_base_out) {
  if (!h || !s) return SDB_EINVAL;
  return submit(h, s, seq_base_out);
}

int sdb_free_staged(sdb_handle h, sdb_staged_t s) {
  if (!h || !s) return SDB_EINVAL;
  ls_stop(h);
  cudaStreamSynchronize(h->stream);
  if (s->owns) {
    cudaFree(s->descs_dev); cudaFree(s->payload_dev);
    if (s->gs_off_dev) cudaFree(s->gs_off_dev);
    if (s->gs_idx_dev) cudaFree(s->gs_idx_dev);
  }
  delete s;
  return SDB_OK;
}

int sdb_latency_server(sdb_handle h, int enable) {
  if (!h) return SDB_EINVAL;
  h->ls_wanted = enable != 0;
  return enable ? ls_start(h) : ls_stop(h);
}

// unpack the latency paths' output block (u64 total | u64 granules | u32 counts[8] | pad to 64 B | records) into the caller's arrays
static void unpack_small(const uint8_t* block, uint32_t n_agents, uint32_t* count_out, sdb_msg_header* hdr_out, uint8_t* payload_out,
                         uint64_t* total_out, uint64_t* payload_bytes_out) {
  const unsigned long long* t = reinterpret_cast<const unsigned long long*>(block);
  const uint64_t total = t[0];
  if (count_out) std::memcpy(count_out, block + 16, static_cast<size_t>(n_agents) * sizeof(uint32_t));
  const uint8_t* p = block + 64;
  uint64_t poff = 0;
  for (uint64_t r = 0; r < total; ++r) {
    const sdb_msg_header* hd = reinterpret_cast<const sdb_msg_header*>(p);
    hdr_out[r] = *hd;
    const uint32_t pl = pad32(hd->len);
    std::memcpy(payload_out + poff, p + 32, pl);
    poff += pl; p += 32 + pl;
  }
  if (total_out) *total_out = total;
  if (payload_bytes_out) *payload_bytes_out = poff;
}

int sdb_receive_batch(sdb_handle h, uint32_t n_agents, const uint32_t* agent_idx, uint32_t max_messages, uint32_t flags,
                      uint32_t* count_out, sdb_msg_header* hdr_out, uint64_t hdr_cap, uint8_t* payload_out,
                      uint64_t payload_cap, uint64_t* total_out, uint64_t* payload_bytes_out) {
  if (!h) return SDB_EINVAL;
  if (total_out) *total_out = 0;
  if (payload_bytes_out) *payload_bytes_out = 0;
  const bool owned = (flags & SDB_RECV_OWNED) && !agent_idx && h->sharded && h->owned_dev;
  if (!agent_idx) n_agents = owned ? h->n_owned : h->n_agents;
  if (n_agents == 0 || max_messages == 0) return SDB_OK;
  if (n_agents > h->cfg.max_agents) return fail(h, SDB_EINVAL, "n_agents > max_agents");
  if (agent_idx) {
    for (uint32_t i = 0; i < n_agents; ++i)
      if (agent_idx[i] >= h->cfg.max_agents) return fail(h, SDB_EINVAL, "agent index out of range");
  }
  // ---- latency path: few agents, host outputs -> one launch, one D2H, one sync
  if (agent_idx && n_agents <= 8 && hdr_out && payload_out &&
      static_cast<uint64_t>(n_agents) * max_messages <= 1024) {
    const uint64_t max_rec_bytes = pad32(h->cfg.max_payload_bytes);
    uint64_t rec_cap = std::min<uint64_t>(1024, std::min<uint64_t>(hdr_cap, payload_cap / max_rec_bytes));
    if (rec_cap == 0) return fail(h, SDB_EOUTPUT, "output buffers cannot hold a single maximum-size record");
    h->last_rx_valid = false;
    bool served = false;
    if (h->ls_wanted && n_agents == 1) {
      { int rcl = ls_start(h); if (rcl != SDB_OK) return rcl; }
      const uint64_t want = std::min<uint64_t>(max_messages, rec_cap);
      served = 64ull + want * (32 + max_rec_bytes) <= h->ls_out_cap;      // the whole answer must fit the pinned block
    }
    if (served) {
      // ---- the persistent server: post the request in mapped pinned memory, read the answer from pinned memory
      if (h->stream_busy) { CUDA_TRY(h, cudaStreamSynchronize(h->stream)); h->stream_busy = false; }   // earlier sends are complete and visible
      sdb_ls_mailbox* mb = h->ls_mbox;
      mb->agent = agent_idx[0]; mb->max_messages = static_cast<uint32_t>(std::min<uint64_t>(max_messages, rec_cap));
      mb->flags = flags & 0x7FFFFFFFu;
      const uint32_t seq = ++h->ls_seq;
      __atomic_store_n(&mb->req_seq, seq, __ATOMIC_RELEASE);
      uint64_t spins = 0;
      while (__atomic_load_n(&mb->done_seq, __ATOMIC_ACQUIRE) != seq) {
        if (++spins > (1ull << 33)) return fail(h, SDB_ECUDA, "latency server did not answer");
        __builtin_ia32_pause();
      }
      unpack_small(h->ls_out, 1, count_out, hdr_out, payload_out, total_out, payload_bytes_out);
      return SDB_OK;
    }
    cudaError_t e = sdb_launch_receive_small(&h->view, agent_idx, n_agents, max_messages, flags, static_cast<uint32_t>(rec_cap),
                                             h->rx_plan, h->rx_small, h->stream, &h->prof);
    h->launches += 1;
    if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("receive launch: ") + cudaGetErrorString(e));
    const uint64_t window = std::min<uint64_t>(h->small_bytes, 16384);
    CUDA_TRY(h, cudaMemcpyAsync(h->small_host, h->rx_small, window, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const unsigned long long* t = reinterpret_cast<const unsigned long long*>(h->small_host);
    const uint64_t total = t[0], bytes = 64 + t[1] * SDB_GRANULE;
    if (bytes > window) {
      CUDA_TRY(h, cudaMemcpyAsync(h->small_host + window, h->rx_small + window, bytes - window, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    unpack_small(h->small_host, n_agents, count_out, hdr_out, payload_out, total_out, payload_bytes_out);
    (void)total;
    return SDB_OK;
  }
  if (agent_idx) {
    CUDA_TRY(h, cudaMemcpyAsync(h->rx_agent, agent_idx, static_cast<size_t>(n_agents) * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  }
  const bool async = (flags & SDB_RECV_ASYNC) != 0;
  if (async && (count_out || hdr_out || payload_out)) return fail(h, SDB_EINVAL, "SDB_RECV_ASYNC takes no host output buffers");
  sdb_recv_args r{};
  r.agent_idx = agent_idx ? h->rx_agent : (owned ? h->owned_dev : nullptr); r.n = n_agents; r.max_messages = max_messages;
  r.flags = flags & (SDB_RECV_PRIORITY | SDB_RECV_PEEK);
  r.cnt = h->rx_cnt; r.rec_local = h->rx_rec_local; r.rec_tops = h->rx_rec_tops;
  r.rec_off = h->rx_rec_off; r.plan = h->rx_plan; r.lb = h->rx_lb;
  r.plan_tops = h->rx_plan_tops; r.totals = h->rx_totals; r.big_list = h->rx_big_list; r.big_count = h->rx_big_count;
  r.count_out = h->rx_count; r.hdr_out = h->rx_hdr; r.payload_out = h->rx_payload;
  // record capacity: bounded so that even maximum-size payloads fit the payload buffers
  const uint64_t max_rec_bytes = pad32(h->cfg.max_payload_bytes);
  uint64_t rec_cap = std::min<uint64_t>(h->cfg.max_recv_records, h->pay_cap_gran * SDB_GRANULE / max_rec_bytes);
  if (hdr_out) rec_cap = std::min<uint64_t>(rec_cap, hdr_cap);
  if (payload_out) rec_cap = std::min<uint64_t>(rec_cap, payload_cap / max_rec_bytes);
  if (rec_cap == 0) return fail(h, SDB_EOUTPUT, "output buffers cannot hold a single maximum-size record");
  r.rec_cap = rec_cap;
  int nl = 0;
  cudaError_t e = sdb_launch_receive(&h->view, &r, h->stream, &nl, &h->prof, h->sm_count, 32u + pad32(h->cfg.max_payload_bytes));
  h->launches += nl;
  if (!(r.flags & SDB_RECV_PRIORITY)) r.plan_tops = nullptr;       // single-pass path: plan offsets are absolute
  h->last_rx = r; h->last_rx_valid = (e == cudaSuccess);
  if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("receive launch: ") + cudaGetErrorString(e));
  h->stream_busy = async;
  if (async) return SDB_OK;          // the caller consumes on the device (stream order) or asks sdb_last_receive_totals later
  CUDA_TRY(h, cudaMemcpyAsync(h->totals_host, h->rx_totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  if (count_out)
    CUDA_TRY(h, cudaMemcpyAsync(count_out, h->rx_count, static_cast<size_t>(n_agents) * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  const uint64_t total = h->totals_host[0], pbytes = h->totals_host[1] * SDB_GRANULE;
  if (total_out) *total_out = total;
  if (payload_bytes_out) *payload_bytes_out = pbytes;
  if (total) {
    if (hdr_out) CUDA_TRY(h, cudaMemcpyAsync(hdr_out, h->rx_hdr, total * sizeof(sdb_msg_header), cudaMemcpyDeviceToHost, h->stream));
    if (payload_out && pbytes) CUDA_TRY(h, cudaMemcpyAsync(payload_out, h->rx_payload, pbytes, cudaMemcpyDeviceToHost, h->stream));
    if (hdr_out || payload_out) CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  }
  return SDB_OK;
}

int sdb_last_receive_dev(sdb_handle h, const uint32_t** count_dev, const sdb_msg_header** hdr_dev, const uint8_t** payload_dev) {
  if (!h) return SDB_EINVAL;
  if (count_dev) *count_dev = h->rx_count;
  if (hdr_dev) *hdr_dev = h->rx_hdr;
  if (payload_dev) *payload_dev = h->rx_payload;
  return SDB_OK;
}

int sdb_last_receive_totals(sdb_handle h, uint64_t* total_out, uint64_t* payload_bytes_out) {
  if (!h) return SDB_EINVAL;
  CUDA_TRY(h, cudaMemcpyAsync(h->totals_host, h->rx_totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (total_out) *total_out = h->totals_host[0];
  if (payload_bytes_out) *payload_bytes_out = h->totals_host[1] * SDB_GRANULE;
  return SDB_OK;
}

// ---- N3: inbox / load queries from the rings ---------------------------------------------------
int sdb_agent_loads(sdb_handle h, uint32_t n, const uint32_t* agent_idx, sdb_agent_load* out) {
  if (!h || (n && !out)) return SDB_EINVAL;
  if (n == 0) return SDB_OK;
  if (n > h->cfg.max_agents) return fail(h, SDB_EINVAL, "n > max_agents");
  if (agent_idx) {
    for (uint32_t i = 0; i < n; ++i) if (agent_idx[i] >= h->cfg.max_agents) return fail(h, SDB_EINVAL, "agent index out of range");
    CUDA_TRY(h, cudaMemcpyAsync(h->rx_agent, agent_idx, static_cast<size_t>(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  }
  // the receive plan buffer doubles as scratch (32 bytes per agent; it holds max_recv_records + 4 16-byte entries)
  sdb_agent_load* tmp = nullptr;
  CUDA_TRY(h, cudaMallocAsync(reinterpret_cast<void**>(&tmp), static_cast<size_t>(n) * sizeof(sdb_agent_load), h->stream));
  cudaError_t e = sdb_launch_agent_loads(&h->view, agent_idx ? h->rx_agent : nullptr, n, tmp, h->sm_count, h->stream);
  h->launches += 1;
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, tmp, static_cast<size_t>(n) * sizeof(sdb_agent_load), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaFreeAsync(tmp, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("agent loads: ") + cudaGetErrorString(e));
  return SDB_OK;
}

int sdb_queue_stats(sdb_handle h, sdb_queue_summary* out) {
  if (!h || !out) return SDB_EINVAL;
  unsigned long long* acc = h->be_scratch;                 // >= 16 words, free between picks
  cudaError_t e = sdb_launch_queue_stats(&h->view, h->n_agents, acc, h->stream);
  h->launches += 1;
  unsigned long long host[10];
  if (e == cudaSuccess) e = cudaMemcpyAsync(host, acc, sizeof(host), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("queue stats: ") + cudaGetErrorString(e));
  out->agents_with_pending = host[0]; out->pending = host[1];
  for (int k = 0; k < 4; ++k) out->pending_by_prio[k] = host[2 + k];
  out->pending_granules = host[6]; out->received = host[7];
  out->max_pending = host[8] >> 32; out->max_pending_agent = host[8] ? 0xFFFFFFFFull - (host[8] & 0xFFFFFFFFull) : 0;
  return SDB_OK;
}

int sdb_assign_agent_backends(sdb_handle h, uint32_t n, const uint32_t* agent_idx, const uint32_t* backend_idx) {
  if (!h || (n && (!agent_idx || !backend_idx))) return SDB_EINVAL;
  if (!h->agent_backend_dev) {
    CUDA_TRY(h, dmalloc(&h->agent_backend_dev, h->cfg.max_agents));
    CUDA_TRY(h, cudaMemsetAsync(h->agent_backend_dev, 0xFF, static_cast<size_t>(h->cfg.max_agents) * sizeof(uint32_t), h->stream));
  }
  for (uint32_t i = 0; i < n; ++i) {
    if (agent_idx[i] >= h->cfg.max_agents) return fail(h, SDB_EINVAL, "agent index out of range");
    if (backend_idx[i] != 0xFFFFFFFFu && backend_idx[i] >= h->cfg.max_backends) return fail(h, SDB_ENOTFOUND, "backend index out of range");
    // assignments are rare and small (one per agent when it first asks): element-wise stream-ordered copies
    CUDA_TRY(h, cudaMemcpyAsync(h->agent_backend_dev + agent_idx[i], backend_idx + i, sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  }
  if (n) CUDA_TRY(h, cudaStreamSynchronize(h->stream));     // the source array is the caller's
  return SDB_OK;
}

int sdb_backend_loads_from_queues(sdb_handle h) {
  if (!h) return SDB_EINVAL;
  if (h->n_backends == 0) return fail(h, SDB_ENOTFOUND, "no backends configured (sdb_set_backends)");
  if (!h->agent_backend_dev) return fail(h, SDB_ENOTFOUND, "no agent is assigned to a backend (sdb_assign_agent_backends)");
  cudaError_t e = sdb_launch_backend_loads_from_queues(&h->view, h->n_agents, h->agent_backend_dev, h->n_backends, h->be_load, h->stream);
  h->launches += 1;
  if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("backend loads: ") + cudaGetErrorString(e));
  return SDB_OK;
}

// ---- stream digests -------------------------------------------------------------------------
static int digest_ensure(sdb_ctx* h) {
  if (h->digest) return SDB_OK;
  CUDA_TRY(h, dmalloc(&h->digest, h->cfg.max_agents));
  CUDA_TRY(h, cudaMemsetAsync(h->digest, 0, static_cast<size_t>(h->cfg.max_agents) * sizeof(unsigned long long), h->stream));
  return SDB_OK;
}

int sdb_digest_reset(sdb_handle h) {
  if (!h) return SDB_EINVAL;
  int rc = digest_ensure(h);
  if (rc != SDB_OK) return rc;
  CUDA_TRY(h, cudaMemsetAsync(h->digest, 0, static_cast<size_t>(h->cfg.max_agents) * sizeof(unsigned long long), h->stream));
  return SDB_OK;
}

int sdb_digest_fold(sdb_handle h) {
  if (!h) return SDB_EINVAL;
  if (!h->last_rx_valid) return fail(h, SDB_EINVAL, "sdb_digest_fold: no bulk receive to fold (the <= 8-agent latency path keeps no device results)");
  int rc = digest_ensure(h);
  if (rc != SDB_OK) return rc;
  cudaError_t e = sdb_launch_digest(&h->last_rx, h->cfg.max_agents, h->digest, h->sm_count, h->stream);
  h->launches += 1;
  if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("digest launch: ") + cudaGetErrorString(e));
  return SDB_OK;
}

int sdb_digest_read(sdb_handle h, uint32_t n, const uint32_t* agent_idx, uint64_t* digest_out) {
  if (!h || (n && !digest_out)) return SDB_EINVAL;
  if (n == 0) return SDB_OK;
  int rc = digest_ensure(h);
  if (rc != SDB_OK) return rc;
  if (!agent_idx) {
    if (n > h->cfg.max_agents) return fail(h, SDB_EINVAL, "n > max_agents");
    CUDA_TRY(h, cudaMemcpyAsync(digest_out, h->digest, static_cast<size_t>(n) * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SDB_OK;
  }
  std::vector<unsigned long long> all(h->cfg.max_agents);
  CUDA_TRY(h, cudaMemcpyAsync(all.data(), h->digest, all.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (uint32_t i = 0; i < n; ++i) {
    if (agent_idx[i] >= h->cfg.max_agents) return fail(h, SDB_EINVAL, "agent index out of range");
    digest_out[i] = all[agent_idx[i]];
  }
  return SDB_OK;
}

// ---- backends -----------------------------------------------------------------------------
int sdb_set_backends(sdb_handle h, uint32_t n, const uint32_t* weight, const uint64_t* load0) {
  if (!h || !weight || n == 0) return SDB_EINVAL;
  if (n > h->cfg.max_backends) return fail(h, SDB_ECAPACITY, "n > max_backends");
  std::vector<unsigned long long> l(n, 0);
  for (uint32_t i = 0; i < n; ++i) {
    if (weight[i] == 0 || weight[i] > 65535) return fail(h, SDB_EINVAL, "backend weight must be in [1, 65535]");
    if (load0) { if (load0[i] >= (1ull << 40)) return fail(h, SDB_EINVAL, "backend load must be < 2^40"); l[i] = load0[i]; }
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaMemcpy(h->be_weight, weight, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->be_load, l.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  h->n_backends = n;
  return SDB_OK;
}

int sdb_get_backend_loads(sdb_handle h, uint32_t n, uint64_t* load_out) {
  if (!h || !load_out || n > h->n_backends) return SDB_EINVAL;
  CUDA_TRY(h, cudaMemcpyAsync(load_out, h->be_load, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return SDB_OK;
}

static int ensure_req_cap(sdb_ctx* h, uint32_t n) {
  if (n <= h->be_req_cap) return SDB_OK;
  { int rcl = ls_stop(h); if (rcl != SDB_OK) return rcl; }      // cudaFree below waits for running kernels
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (h->be_req_cost) cudaFree(h->be_req_cost);
  if (h->be_out) cudaFree(h->be_out);
  h->be_req_cost = nullptr; h->be_out = nullptr; h->be_req_cap = 0;
  CUDA_TRY(h, dmalloc(&h->be_req_cost, n));
  CUDA_TRY(h, dmalloc(&h->be_out, n));
  h->be_req_cap = n;
  return SDB_OK;
}

int sdb_release_backends(sdb_handle h, uint32_t n, const uint32_t* backend, const uint32_t* cost) {
  if (!h || (n && !backend)) return SDB_EINVAL;
  if (n == 0) return SDB_OK;
  // completion reports are rare and small: fold them on the host, apply as one device update
  std::vector<unsigned long long> l(h->n_backends);
  CUDA_TRY(h, cudaMemcpyAsync(l.data(), h->be_load, h->n_backends * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (uint32_t i = 0; i < n; ++i) {
    if (backend[i] >= h->n_backends) return fail(h, SDB_ENOTFOUND, "backend index out of range");
    const unsigned long long c = cost ? cost[i] : 1;
    l[backend[i]] = l[backend[i]] >= c ? l[backend[i]] - c : 0;
  }
  CUDA_TRY(h, cudaMemcpy(h->be_load, l.data(), h->n_backends * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  return SDB_OK;
}

int sdb_select_backend_batch(sdb_handle h, uint32_t n_req, const uint32_t* cost, uint32_t mode, uint64_t seed, uint32_t* backend_out) {
  if (!h || !backend_out) return SDB_EINVAL;
  if (mode > 1) return fail(h, SDB_EINVAL, "mode must be 0 (least-load) or 1 (weighted random)");
  if (h->n_backends == 0) return fail(h, SDB_ENOTFOUND, "no backends configured (sdb_set_backends)");
  if (n_req == 0) return SDB_OK;
  int rc = ensure_req_cap(h, n_req);
  if (rc != SDB_OK) return rc;
  if (cost) CUDA_TRY(h, cudaMemcpyAsync(h->be_req_cost, cost, static_cast<size_t>(n_req) * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  int nl = 0;
  cudaError_t e = sdb_launch_pick(static_cast<int>(mode), h->n_backends, h->be_weight, h->be_load, n_req,
                                  cost ? h->be_req_cost : nullptr, seed, h->be_out, h->be_scratch, h->be_logtab,
                                  h->stream, &nl);
  h->launches += nl;
  h->picks += n_req;
  if (e != cudaSuccess) return fail(h, SDB_ECUDA, std::string("backend pick: ") + cudaGetErrorString(e));
  CUDA_TRY(h, cudaMemcpyAsync(backend_out, h->be_out, static_cast<size_t>(n_req) * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return SDB_OK;
}

}  // extern "C"
