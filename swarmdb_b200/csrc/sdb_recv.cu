// sdb_recv.cu - dequeue kernels (sm_100a): receive_messages (reference M:521-601; the drain
// loop M:553-601 and its filter M:579-585 are replaced by per-agent rings that only ever
// hold records addressed to that agent).
//
// Pipeline of one sdb_receive_batch call (all stream-ordered, no host round trip inside):
//   k_recv_count    cnt[q]   = min(max_messages, live(agent q))
//   scan            rec_off  = exclusive scan of cnt             (k_scan_local + k_scan_tops)
//   k_recv_select   per agent: choose WHICH pending entries are delivered
//                     stream order      -> the first cnt live entries (contiguous when the
//                                          window holds no consumed entries: nothing to do);
//                     priority order    -> segmented radix-select on the 2-bit priority:
//                                          pass 1 4-bin histogram of the pending window with
//                                          warp ballots/popc, pick the cut level and residual,
//                                          pass 2 stable compaction (ballot prefix ranks) of
//                                          the selected ring positions in (prio desc, arrival);
//                   also payload granules per agent and the would-be new head / tombstone count
//   scan            pay_off  = exclusive scan of payload granules per agent
//   k_recv_gather   per agent: copy header + payload of each selected record from the arena to
//                   the packed output, then retire the entries (advance head / tombstone)
//
// Roofline: HBM-bound.  Algorithmic bytes per agent-call: P*1 priority bytes scanned (we scan
// 2-byte ring_meta entries) + 2*k*(L+H) gather+emit.
#include "sdb_common.cuh"

#define SDB_SCAN_TILE 4096u   // elements per scan block (1024 threads x 4)
#define SDB_MODE_LIST 0x80000000u


// ------------------------------------------------------------------------------------------
// generic exclusive scan over uint32 (two tiny kernels; consumers add tops[i / TILE])
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_scan_local(const uint32_t* __restrict__ in, uint32_t mask, uint32_t* __restrict__ local,
             uint32_t* __restrict__ tops, uint32_t n) {
  __shared__ uint32_t s_warp[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t base = blockIdx.x * SDB_SCAN_TILE + tid * 4u;
  uint32_t x[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = (base + k < n) ? (in[base + k] & mask) : 0u;
  uint32_t tsum = x[0] + x[1] + x[2] + x[3];
  uint32_t incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane];
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
      if (lane >= o) wi += y;
    }
    s_warp[lane] = wi - w;              // exclusive warp offsets
    if (lane == 31) tops[blockIdx.x] = wi;
  }
  __syncthreads();
  uint32_t run = s_warp[warp] + incl - tsum;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (base + k < n) local[base + k] = run;
    run += x[k];
  }
}

__global__ void __launch_bounds__(1024)
k_scan_tops(uint32_t* __restrict__ tops, uint32_t n_tiles, unsigned long long* __restrict__ total_out) {
  // single block; n_tiles is small (n / 4096)
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_tiles; base += 1024) {
    const uint32_t i = base + tid;
    const uint32_t x = i < n_tiles ? tops[i] : 0u;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
        if (lane >= o) wi += y;
      }
      s_warp[lane] = wi - w;
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    const uint32_t excl = carry + s_warp[warp] + incl - x;
    if (i < n_tiles) tops[i] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + x;
    __syncthreads();
  }
  if (tid == 0 && total_out) *total_out = s_carry;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_recv_count(sdb_dev_view v, sdb_recv_args r) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.n) return;
  const uint32_t a = r.agent_idx ? r.agent_idx[q] : q;
  uint32_t c = 0;
  if (a < v.max_agents) {
    const uint64_t st = v.ring_state[a];
    const uint32_t live = static_cast<uint32_t>(st >> 32) - static_cast<uint32_t>(st) - v.ntomb[a];
    c = min(live, r.max_messages);
  }
  r.cnt[q] = c;
}

// ------------------------------------------------------------------------------------------
// select: one warp owns 32 consecutive agents of the request list.  Agents whose selection is
// a short contiguous run are finished by their own lane; the rest are processed one at a
// time by the whole warp (window scans with ballots).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_prefix(uint32_t ballot, uint32_t lane) {
  return __popc(ballot & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256)
k_recv_select(sdb_dev_view v, sdb_recv_args r) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = q < r.n;
  const uint32_t R = v.ring_slots, mask = R - 1;

  uint32_t a = 0, head = 0, tail = 0, nt = 0, cnt = 0, roff = 0;
  if (valid) {
    a = r.agent_idx ? r.agent_idx[q] : q;
    cnt = r.cnt[q];
    roff = r.rec_local[q] + r.rec_tops[q / SDB_SCAN_TILE];
    if (static_cast<uint64_t>(roff) + cnt > r.rec_cap) cnt = 0;     // whole-agent truncation, lossless
    if (cnt) {
      const uint64_t st = v.ring_state[a];
      head = static_cast<uint32_t>(st); tail = static_cast<uint32_t>(st >> 32);
      nt = v.ntomb[a];
    }
  }
  const bool prio_mode = (r.flags & SDB_RECV_PRIORITY) != 0;
  const bool contiguous = !prio_mode && nt == 0;
  constexpr uint32_t SMALL = 8;
  bool done = !valid || cnt == 0;
  if (valid && cnt == 0) { r.cnt[q] = 0; r.pay[q] = 0; }
  if (!done && contiguous && cnt <= SMALL) {
    const uint16_t* ms = v.ring_meta + (static_cast<size_t>(a) << v.ring_shift);
    uint32_t g = 0;
    for (uint32_t j = 0; j < cnt; ++j) g += (ms[(head + j) & mask] & SDB_META_GLEN_MASK) - 1u;
    r.pay[q] = g; r.cnt[q] = cnt; r.old_head[q] = head; r.new_head[q] = head + cnt; r.new_ntomb[q] = 0;
    done = true;
  }
  uint32_t todo = __ballot_sync(0xFFFFFFFFu, !done);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t A = __shfl_sync(0xFFFFFFFFu, a, src);
    const uint32_t H = __shfl_sync(0xFFFFFFFFu, head, src);
    const uint32_t T = __shfl_sync(0xFFFFFFFFu, tail, src);
    const uint32_t NT = __shfl_sync(0xFFFFFFFFu, nt, src);
    const uint32_t C = __shfl_sync(0xFFFFFFFFu, cnt, src);
    const uint32_t RO = __shfl_sync(0xFFFFFFFFu, roff, src);
    const uint32_t Q = (q - lane) + src;
    const uint16_t* ms = v.ring_meta + (static_cast<size_t>(A) << v.ring_shift);

    if (!prio_mode && NT == 0) {
      // long contiguous run: just add up payload granules
      uint32_t g = 0;
      for (uint32_t j = lane; j < C; j += 32) g += (ms[(H + j) & mask] & SDB_META_GLEN_MASK) - 1u;
      for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xFFFFFFFFu, g, o);
      if (lane == 0) { r.pay[Q] = g; r.cnt[Q] = C; r.old_head[Q] = H; r.new_head[Q] = H + C; r.new_ntomb[Q] = 0; }
      continue;
    }
    // ---- pass 1: histogram of live entries per priority level over the window [H, T)
    uint32_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;
    if (prio_mode) {
      for (uint32_t p = H + lane; static_cast<int32_t>(T - p) > 0; p += 32) {
        const uint16_t m = ms[p & mask];
        if (m == SDB_META_TOMB) continue;
        const uint32_t L = m >> 14;
        h0 += (L == 0); h1 += (L == 1); h2 += (L == 2); h3 += (L == 3);
      }
      // 4 counters, each < 2^16 per lane only if window < 2M entries: reduce as two packed u64
      unsigned long long lo = (static_cast<unsigned long long>(h1) << 32) | h0;
      unsigned long long hi = (static_cast<unsigned long long>(h3) << 32) | h2;
      for (int o = 16; o; o >>= 1) {
        lo += __shfl_xor_sync(0xFFFFFFFFu, lo, o);
        hi += __shfl_xor_sync(0xFFFFFFFFu, hi, o);
      }
      h0 = static_cast<uint32_t>(lo); h1 = static_cast<uint32_t>(lo >> 32);
      h2 = static_cast<uint32_t>(hi); h3 = static_cast<uint32_t>(hi >> 32);
    } else {
      h0 = (T - H) - NT;     // single level: every live entry
    }
    // ---- cut: take every entry of levels above `cut`, and the first `resid` of level `cut`
    uint32_t hist[4] = {h0, h1, h2, h3};
    uint32_t quota[4] = {0, 0, 0, 0}, basek[4] = {0, 0, 0, 0};
    {
      uint32_t need = C, acc = 0;
      for (int L = 3; L >= 0; --L) {
        const uint32_t take = min(hist[L], need);
        quota[L] = take; basek[L] = acc; acc += take; need -= take;
      }
    }
    // ---- pass 2: stable compaction of the selected positions, window order within a level
    uint32_t taken[4] = {0, 0, 0, 0};
    uint32_t first_unsel = T;      // first live entry left behind
    uint32_t g = 0, got = 0;
    uint32_t p0 = H;
    for (; static_cast<int32_t>(T - p0) > 0 && got < C; p0 += 32) {
      const uint32_t p = p0 + lane;
      const bool in = static_cast<int32_t>(T - p) > 0;
      const uint16_t m = in ? ms[p & mask] : SDB_META_TOMB;
      const bool live = m != SDB_META_TOMB;
      const uint32_t L = (live && prio_mode) ? (m >> 14) : 0u;
      bool sel = false; uint32_t rank = 0;
#pragma unroll
      for (uint32_t lev = 0; lev < 4; ++lev) {
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, live && L == lev);
        if (live && L == lev) {
          const uint32_t k = taken[lev] + lane_prefix(b, lane);
          if (k < quota[lev]) { sel = true; rank = basek[lev] + k; }
        }
        taken[lev] += __popc(b);
      }
      if (sel) {
        r.sel_pos[RO + rank] = p;
        g += (m & SDB_META_GLEN_MASK) - 1u;
      }
      const uint32_t bs = __ballot_sync(0xFFFFFFFFu, sel);
      got += __popc(bs);
      const uint32_t bu = __ballot_sync(0xFFFFFFFFu, live && !sel);
      if (bu && first_unsel == T) first_unsel = p0 + (__ffs(bu) - 1);
    }
    const uint32_t scan_end = static_cast<int32_t>(T - p0) > 0 ? p0 : T;
    uint32_t nh = first_unsel;
    if (static_cast<int32_t>(nh - scan_end) > 0) nh = scan_end;
    for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xFFFFFFFFu, g, o);
    if (lane == 0) {
      r.pay[Q] = g; r.cnt[Q] = got | SDB_MODE_LIST; r.old_head[Q] = H;
      r.new_head[Q] = nh; r.new_ntomb[Q] = NT + got - (nh - H);
    }
  }
}

// ------------------------------------------------------------------------------------------
// gather: one warp per agent.  Copies each selected record (32-byte header -> hdr_out,
// padded payload -> payload_out) with 16-byte streaming accesses, then retires the entries.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_recv_gather(sdb_dev_view v, sdb_recv_args r) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= r.n) return;
  const uint32_t cm = r.cnt[q];
  const uint32_t cnt = cm & ~SDB_MODE_LIST;
  const bool list = (cm & SDB_MODE_LIST) != 0;
  if (cnt == 0) { if (lane == 0) r.count_out[q] = 0; return; }
  const uint32_t roff = r.rec_local[q] + r.rec_tops[q / SDB_SCAN_TILE];
  const uint64_t poff = static_cast<uint64_t>(r.pay_local[q]) + r.pay_tops[q / SDB_SCAN_TILE];
  const uint32_t pay = r.pay[q];
  if (poff + pay > r.pay_cap_gran) { if (lane == 0) r.count_out[q] = 0; return; }   // does not fit: stays queued

  const uint32_t a = r.agent_idx ? r.agent_idx[q] : q;
  const uint32_t R = v.ring_slots, mask = R - 1;
  const uint32_t H = r.old_head[q];
  uint32_t* hs = v.ring_handle + (static_cast<size_t>(a) << v.ring_shift);
  uint16_t* ms = v.ring_meta + (static_cast<size_t>(a) << v.ring_shift);
  // arena position of a 32-bit handle: handles of live entries are within 2^32 granules of each
  // other and the arena is at most 2^32 granules, so the low bits select the slot directly.
  uint32_t run = 0;                       // payload granules emitted so far for this agent
  for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
    const uint32_t j = j0 + lane;
    uint32_t pos = 0, handle = 0, pg = 0;
    if (j < cnt) {
      pos = list ? r.sel_pos[roff + j] : H + j;
      handle = hs[pos & mask];
      pg = (ms[pos & mask] & SDB_META_GLEN_MASK) - 1u;
      if (list) ms[pos & mask] = SDB_META_TOMB;
    }
    uint32_t incl = pg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += y;
    }
    const uint32_t excl = run + incl - pg;
    run += __shfl_sync(0xFFFFFFFFu, incl, 31);
    const uint32_t m = min(32u, cnt - j0);
    for (uint32_t t = 0; t < m; ++t) {
      const uint32_t hd = __shfl_sync(0xFFFFFFFFu, handle, t);
      const uint32_t g = __shfl_sync(0xFFFFFFFFu, pg, t);
      const uint32_t ex = __shfl_sync(0xFFFFFFFFu, excl, t);
      const uint8_t* src = v.arena + ((static_cast<uint64_t>(hd) & v.gmask) << 5);
      uint8_t* hdst = reinterpret_cast<uint8_t*>(r.hdr_out + roff + j0 + t);
      uint8_t* pdst = r.payload_out + ((poff + ex) << 5);
      const uint32_t nchunk = 2u + (g << 1);
      for (uint32_t c = lane; c < nchunk; c += 32) {
        const uint4 x = sdb_ld_stream(src + (c << 4));
        if (c < 2) sdb_st_stream(hdst + (c << 4), x);
        else sdb_st_stream(pdst + ((c - 2u) << 4), x);
      }
    }
  }
  if (lane == 0) {
    r.count_out[q] = cnt;
    // retire: the low word of ring_state is head (little endian); no enqueue runs concurrently
    reinterpret_cast<uint32_t*>(v.ring_state + a)[0] = r.new_head[q];
    v.ntomb[a] = r.new_ntomb[q];
    atomicAdd(&v.ctr->delivered, static_cast<unsigned long long>(cnt));
    atomicMax(r.totals + 0, static_cast<unsigned long long>(roff) + cnt);
    atomicMax(r.totals + 1, poff + pay);
  }
}

// ------------------------------------------------------------------------------------------
extern "C" cudaError_t sdb_launch_receive(const sdb_dev_view* v, const sdb_recv_args* r, cudaStream_t stream,
                                          int* n_launches, sdb_profiler* prof) {
  if (r->n == 0) return cudaSuccess;
  const uint32_t n = r->n;
  const uint32_t tiles = (n + SDB_SCAN_TILE - 1) / SDB_SCAN_TILE;
  cudaMemsetAsync(r->totals, 0, 2 * sizeof(unsigned long long), stream);
  int pi = sdb_prof_begin(prof, SDB_PK_RECV_COUNT, stream);
  k_recv_count<<<(n + 255) / 256, 256, 0, stream>>>(*v, *r);
  sdb_prof_end(prof, pi, stream);
  pi = sdb_prof_begin(prof, SDB_PK_RECV_SCAN, stream);
  k_scan_local<<<tiles, 1024, 0, stream>>>(r->cnt, 0xFFFFFFFFu, r->rec_local, r->rec_tops, n);
  k_scan_tops<<<1, 1024, 0, stream>>>(r->rec_tops, tiles, nullptr);
  sdb_prof_end(prof, pi, stream);
  pi = sdb_prof_begin(prof, SDB_PK_RECV_SELECT, stream);
  k_recv_select<<<(n + 255) / 256, 256, 0, stream>>>(*v, *r);
  sdb_prof_end(prof, pi, stream);
  pi = sdb_prof_begin(prof, SDB_PK_RECV_SCAN, stream);
  k_scan_local<<<tiles, 1024, 0, stream>>>(r->pay, 0xFFFFFFFFu, r->pay_local, r->pay_tops, n);
  k_scan_tops<<<1, 1024, 0, stream>>>(r->pay_tops, tiles, nullptr);
  sdb_prof_end(prof, pi, stream);
  const uint64_t threads = static_cast<uint64_t>(n) * 32;
  pi = sdb_prof_begin(prof, SDB_PK_RECV_GATHER, stream);
  k_recv_gather<<<static_cast<uint32_t>((threads + 255) / 256), 256, 0, stream>>>(*v, *r);
  sdb_prof_end(prof, pi, stream);
  if (n_launches) *n_launches += 7;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// arena floor: smallest arena position still referenced by a pending ring entry.
// One thread per agent; distance below the current arena tail is maximised with atomicMax.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_arena_floor(sdb_dev_view v, uint32_t n_agents, uint32_t tail32, unsigned long long* max_dist) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_agents) return;
  const uint64_t st = v.ring_state[a];
  const uint32_t head = static_cast<uint32_t>(st), tail = static_cast<uint32_t>(st >> 32);
  const uint32_t mask = v.ring_slots - 1;
  const uint16_t* ms = v.ring_meta + (static_cast<size_t>(a) << v.ring_shift);
  const uint32_t* hs = v.ring_handle + (static_cast<size_t>(a) << v.ring_shift);
  for (uint32_t p = head; p != tail; ++p) {
    if (ms[p & mask] != SDB_META_TOMB) {
      const uint32_t dist = tail32 - hs[p & mask];      // granules below the arena tail (mod 2^32)
      atomicMax(max_dist, static_cast<unsigned long long>(dist));
      return;                                           // rings are sorted: first live entry is the oldest
    }
  }
}

extern "C" cudaError_t sdb_launch_arena_floor(const sdb_dev_view* v, uint32_t n_agents, uint32_t tail32,
                                              unsigned long long* max_dist_dev, cudaStream_t stream) {
  cudaMemsetAsync(max_dist_dev, 0, sizeof(unsigned long long), stream);
  if (n_agents) k_arena_floor<<<(n_agents + 255) / 256, 256, 0, stream>>>(*v, n_agents, tail32, max_dist_dev);
  return cudaGetLastError();
}
