// sdb_recv.cu - dequeue kernels (sm_100a): receive_messages (reference M:521-601; the drain
// loop M:553-601 and its filter M:579-585 are replaced by per-agent rings that only ever
// hold records addressed to that agent).
//
// Pipeline of one sdb_receive_batch call (all stream-ordered, no host round trip inside):
//   k_recv_count    cnt[q]   = min(max_messages, live(agent q))
//   scan            rec_off  = exclusive scan of cnt             (k_scan_local + k_scan_tops)
//   k_recv_select   per agent: choose WHICH pending entries are delivered, write the per-record
//                   plan (arena handle, payload size) at the record's output index, retire them
//                     stream order      -> the first cnt live entries (contiguous when the
//                                          window holds no consumed entries: nothing to do);
//                     priority order    -> segmented radix-select on the 2-bit priority:
//                                          pass 1 4-bin histogram of the pending window with
//                                          warp ballots/popc, pick the cut level and residual,
//                                          pass 2 stable compaction (ballot prefix ranks) of
//                                          the selected ring positions in (prio desc, arrival);
//   scan            pay_off  = exclusive scan of payload granules over the planned RECORDS
//   k_recv_gather   flat over records: copy header + payload from the arena to the packed output
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

// out[i] = local[i] + tops[i / TILE]: materialise a scan as a plain array
__global__ void __launch_bounds__(256)
k_scan_apply(const uint32_t* __restrict__ local, const uint32_t* __restrict__ tops, uint32_t* __restrict__ out, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = local[i] + tops[i / SDB_SCAN_TILE];
}

// exclusive scan of n uint32 (block-local part in `local`, per-block part in `tops`), optional
// grand total and optional materialised output; usable from the other translation units
extern "C" cudaError_t sdb_scan_u32(const uint32_t* in, uint32_t* local, uint32_t* tops, uint32_t n,
                                    unsigned long long* total_out, uint32_t* out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const uint32_t tiles = (n + SDB_SCAN_TILE - 1) / SDB_SCAN_TILE;
  k_scan_local<<<tiles, 1024, 0, stream>>>(in, 0xFFFFFFFFu, local, tops, n);
  k_scan_tops<<<1, 1024, 0, stream>>>(tops, tiles, total_out);
  if (out) k_scan_apply<<<(n + 255) / 256, 256, 0, stream>>>(local, tops, out, n);
  return cudaGetLastError();
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
// select + retire: one warp owns 32 consecutive agents of the request list.  Agents whose
// selection is a short contiguous run are finished by their own lane; the rest are processed
// one at a time by the whole warp (window scans with ballots).  For every selected record the
// plan (arena handle, payload granules) is written at its output index, the ring entry is
// retired (head advanced / tombstoned), and count_out is final.  Agents that would overflow
// the record capacity are left untouched (whole-agent truncation: nothing is ever lost).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_prefix(uint32_t ballot, uint32_t lane) {
  return __popc(ballot & ((1u << lane) - 1u));
}

// one agent, whole warp: choose the C entries to deliver from the window [H, T), write their plan at
// output index RO.., retire them (advance head / tombstone).  Stream order when !prio_mode.
__device__ __forceinline__ void select_agent_warp(const sdb_dev_view& v, uint32_t* plan_handle, uint32_t* plan_glen,
                                                  bool prio_mode, uint32_t A, uint32_t H, uint32_t T, uint32_t NT,
                                                  uint32_t C, uint32_t RO, uint32_t lane, bool retire) {
  const uint32_t mask = v.ring_slots - 1;
    uint16_t* ms = v.ring_meta + (static_cast<size_t>(A) << v.ring_shift);
    const uint32_t* hs = v.ring_handle + (static_cast<size_t>(A) << v.ring_shift);

    if (!prio_mode && NT == 0) {
      // long contiguous run [H, H+C)
      for (uint32_t j = lane; j < C; j += 32) {
        plan_handle[RO + j] = hs[(H + j) & mask];
        plan_glen[RO + j] = (ms[(H + j) & mask] & SDB_META_GLEN_MASK) - 1u;
      }
      if (lane == 0 && retire) reinterpret_cast<uint32_t*>(v.ring_state + A)[0] = H + C;
      return;
    }
    // ---- pass 1: histogram of live entries per priority level over the window [H, T)
    uint32_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;
    if (prio_mode) {
      if (v.ring_slots >= 8) {
        // 8 metas (16 bytes) per lane per load, 4 independent loads in flight per lane (1024 entries per
        // warp step), aligned groups of the circular window
        for (uint32_t b8 = (H & ~7u) + (lane << 3); static_cast<int32_t>(T - b8) > 0; b8 += 1024) {
          uint4 q[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t p8 = b8 + u * 256;
            q[u] = static_cast<int32_t>(T - p8) > 0 ? *reinterpret_cast<const uint4*>(ms + (p8 & mask)) : make_uint4(~0u, ~0u, ~0u, ~0u);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t p8 = b8 + u * 256;
            const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t p = p8 + k;
              const uint32_t m = (w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
              if (static_cast<int32_t>(p - H) >= 0 && static_cast<int32_t>(T - p) > 0 && m != SDB_META_TOMB) {
                const uint32_t L = m >> 14;
                h0 += (L == 0); h1 += (L == 1); h2 += (L == 2); h3 += (L == 3);
              }
            }
          }
        }
      } else {
        for (uint32_t p = H + lane; static_cast<int32_t>(T - p) > 0; p += 32) {
          const uint16_t m = ms[p & mask];
          if (m == SDB_META_TOMB) continue;
          const uint32_t L = m >> 14;
          h0 += (L == 0); h1 += (L == 1); h2 += (L == 2); h3 += (L == 3);
        }
      }
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
    // ---- cut: take every entry of the levels above the cut, and the first `quota` of the cut level
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
    uint32_t got = 0;
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
        plan_handle[RO + rank] = hs[p & mask];
        plan_glen[RO + rank] = (m & SDB_META_GLEN_MASK) - 1u;
        if (retire) ms[p & mask] = SDB_META_TOMB;
      }
      const uint32_t bs = __ballot_sync(0xFFFFFFFFu, sel);
      got += __popc(bs);
      const uint32_t bu = __ballot_sync(0xFFFFFFFFu, live && !sel);
      if (bu && first_unsel == T) first_unsel = p0 + (__ffs(bu) - 1);
    }
    const uint32_t scan_end = static_cast<int32_t>(T - p0) > 0 ? p0 : T;
    uint32_t nh = first_unsel;
    if (static_cast<int32_t>(nh - scan_end) > 0) nh = scan_end;
    if (lane == 0 && retire) {
      reinterpret_cast<uint32_t*>(v.ring_state + A)[0] = nh;
      v.ntomb[A] = NT + got - (nh - H);
    }
}

__global__ void __launch_bounds__(256)
k_recv_select(sdb_dev_view v, sdb_recv_args r) {
  // one lane per requested agent: short contiguous runs (the common drain case) are planned and
  // retired here; everything else is queued for k_recv_select_big (one warp per agent)
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = q < r.n;
  const uint32_t R = v.ring_slots, mask = R - 1;

  uint32_t a = 0, head = 0, nt = 0, cnt = 0, roff = 0;
  if (valid) {
    a = r.agent_idx ? r.agent_idx[q] : q;
    cnt = r.cnt[q];
    roff = r.rec_local[q] + r.rec_tops[q / SDB_SCAN_TILE];
    if (cnt && static_cast<uint64_t>(roff) + cnt > r.rec_cap) {      // does not fit: stays queued
      atomicMin(r.totals, static_cast<unsigned long long>(roff));
      cnt = 0;
    }
    r.count_out[q] = cnt;
    if (cnt) {
      head = static_cast<uint32_t>(v.ring_state[a]);
      nt = v.ntomb[a];
    }
  }
  const bool prio_mode = (r.flags & SDB_RECV_PRIORITY) != 0;
  constexpr uint32_t SMALL = 64, CH = 8;
  const bool small = valid && cnt && !prio_mode && nt == 0 && cnt <= SMALL;
  if (small) {
    const uint64_t pol = sdb_policy_evict_last();
    const uint16_t* ms = v.ring_meta + (static_cast<size_t>(a) << v.ring_shift);
    const uint32_t* hs = v.ring_handle + (static_cast<size_t>(a) << v.ring_shift);
    for (uint32_t b = 0; b < cnt; b += CH) {
      uint32_t hv[CH]; uint16_t mv[CH];
#pragma unroll
      for (uint32_t j = 0; j < CH; ++j)       // all loads of the chunk first, then the stores
        if (b + j < cnt) { hv[j] = sdb_ld_u32_pol(hs + ((head + b + j) & mask), pol); mv[j] = sdb_ld_u16_pol(ms + ((head + b + j) & mask), pol); }
#pragma unroll
      for (uint32_t j = 0; j < CH; ++j)
        if (b + j < cnt) { r.plan_handle[roff + b + j] = hv[j]; r.plan_glen[roff + b + j] = (mv[j] & SDB_META_GLEN_MASK) - 1u; }
    }
    if (!(r.flags & SDB_RECV_PEEK)) reinterpret_cast<uint32_t*>(v.ring_state + a)[0] = head + cnt;    // low word = head (little endian)
  }
  // everything else (long runs, holes, priority order) goes to the warp-per-agent kernel
  const bool big = valid && cnt && !small;
  const uint32_t todo = __ballot_sync(0xFFFFFFFFu, big);
  uint32_t n_deliv = (valid && !(r.flags & SDB_RECV_PEEK)) ? cnt : 0;
  if (todo) {
    uint32_t basew = 0;
    if (lane == 0) basew = atomicAdd(r.big_count, static_cast<uint32_t>(__popc(todo)));
    basew = __shfl_sync(0xFFFFFFFFu, basew, 0);
    if (big) r.big_list[basew + lane_prefix(todo, lane)] = q;
  }
  for (int o = 16; o; o >>= 1) n_deliv += __shfl_xor_sync(0xFFFFFFFFu, n_deliv, o);
  if (lane == 0 && n_deliv) atomicAdd(&v.ctr->delivered, static_cast<unsigned long long>(n_deliv));
}

// Single-pass bounded selection (priority order, or stream order with holes) for C <= SDB_CAND:
// walk the window once in arrival order keeping, per level, the first C live positions (stable,
// ballot-prefix ranks) in shared memory; stop as soon as the top level alone fills the quota -
// with mixed priorities that happens after a few hundred entries instead of the whole window.
#define SDB_CAND 128u
__device__ __forceinline__ void select_agent_warp_bounded(const sdb_dev_view& v, uint32_t* plan_handle, uint32_t* plan_glen,
                                                          bool prio_mode, uint32_t A, uint32_t H, uint32_t T, uint32_t NT,
                                                          uint32_t C, uint32_t RO, uint32_t lane, bool retire,
                                                          uint32_t (*cand)[SDB_CAND]) {
  const uint32_t mask = v.ring_slots - 1;
  uint16_t* ms = v.ring_meta + (static_cast<size_t>(A) << v.ring_shift);
  const uint32_t* hs = v.ring_handle + (static_cast<size_t>(A) << v.ring_shift);
  const uint32_t top = prio_mode ? 3u : 0u;
  uint32_t cnt[4] = {0, 0, 0, 0};
  uint32_t p0 = H;
  while (static_cast<int32_t>(T - p0) > 0 && cnt[top] < C) {
    // four chunks of 32 entries in flight per step (the loads do not depend on the running counts)
    uint16_t mm[4];
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u) {
      const uint32_t p = p0 + u * 32u + lane;
      mm[u] = static_cast<int32_t>(T - p) > 0 ? ms[p & mask] : SDB_META_TOMB;
    }
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u) {
      if (static_cast<int32_t>(T - p0) <= 0 || cnt[top] >= C) break;       // warp-uniform
      const uint32_t p = p0 + lane;
      const uint16_t m = mm[u];
      const bool live = m != SDB_META_TOMB;
      const uint32_t L = (live && prio_mode) ? (m >> 14) : 0u;
#pragma unroll
      for (uint32_t lev = 0; lev < 4; ++lev) {
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, live && L == lev);
        if (live && L == lev) {
          const uint32_t k = cnt[lev] + lane_prefix(b, lane);
          if (k < C) cand[lev][k] = p;
        }
        cnt[lev] += __popc(b);
      }
      p0 += 32;
    }
  }
  const uint32_t scan_end = static_cast<int32_t>(T - p0) > 0 ? p0 : T;
  __syncwarp();
  // quotas: everything of the higher levels first, then the first entries of the cut level
  uint32_t need = C, acc = 0, got = 0;
  for (int lev = 3; lev >= 0; --lev) {
    const uint32_t have = min(cnt[lev], C);
    const uint32_t take = min(have, need);
    for (uint32_t k = lane; k < take; k += 32) {
      const uint32_t p = cand[lev][k];
      plan_handle[RO + acc + k] = hs[p & mask];
      plan_glen[RO + acc + k] = (ms[p & mask] & SDB_META_GLEN_MASK) - 1u;
      if (retire) ms[p & mask] = SDB_META_TOMB;
    }
    acc += take; need -= take; got += take;
  }
  if (!retire) return;
  __syncwarp();
  // new head: first live entry left behind inside the scanned range (the tombstones just written count as consumed)
  uint32_t nh = scan_end;
  for (uint32_t q0 = H; static_cast<int32_t>(scan_end - q0) > 0; q0 += 32) {
    const uint32_t p = q0 + lane;
    const bool live = static_cast<int32_t>(scan_end - p) > 0 && ms[p & mask] != SDB_META_TOMB;
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, live);
    if (b) { nh = q0 + (__ffs(b) - 1); break; }
  }
  if (lane == 0) {
    reinterpret_cast<uint32_t*>(v.ring_state + A)[0] = nh;
    v.ntomb[A] = NT + got - (nh - H);
  }
}

// warp per listed agent (persistent grid-stride over the worklist built by k_recv_select)
__global__ void __launch_bounds__(256)
k_recv_select_big(sdb_dev_view v, sdb_recv_args r) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
  __shared__ uint32_t s_cand[8][4][SDB_CAND];            // per warp: first C positions of each level
  const uint32_t n_big = *r.big_count;
  const bool prio_mode = (r.flags & SDB_RECV_PRIORITY) != 0;
  const bool retire = !(r.flags & SDB_RECV_PEEK);
  for (uint32_t w = gw; w < n_big; w += nw) {
    const uint32_t q = r.big_list[w];
    const uint32_t a = r.agent_idx ? r.agent_idx[q] : q;
    const uint32_t cnt = r.count_out[q];
    const uint32_t roff = r.rec_local[q] + r.rec_tops[q / SDB_SCAN_TILE];
    const uint64_t st = v.ring_state[a];
    const uint32_t H = static_cast<uint32_t>(st), T = static_cast<uint32_t>(st >> 32), NT = v.ntomb[a];
    if (cnt <= SDB_CAND && (prio_mode || NT != 0))
      select_agent_warp_bounded(v, r.plan_handle, r.plan_glen, prio_mode, a, H, T, NT, cnt, roff, lane, retire,
                                s_cand[threadIdx.x >> 5]);
    else
      select_agent_warp(v, r.plan_handle, r.plan_glen, prio_mode, a, H, T, NT, cnt, roff, lane, retire);
    __syncwarp();
  }
}

// scan over the per-record payload sizes; the element count lives on the device (totals[0])
__global__ void __launch_bounds__(1024)
k_scan_plan(const uint32_t* __restrict__ in, uint32_t* __restrict__ local, uint32_t* __restrict__ tops,
            const unsigned long long* __restrict__ n_dev) {
  __shared__ uint32_t s_warp[32];
  const uint32_t n = static_cast<uint32_t>(*n_dev);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t base = blockIdx.x * SDB_SCAN_TILE + tid * 4u;
  if (blockIdx.x * SDB_SCAN_TILE >= n) { if (tid == 0) tops[blockIdx.x] = 0; return; }
  uint32_t x[4];
  if (base + 3 < n) {
    const uint4 q = *reinterpret_cast<const uint4*>(in + base);
    x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = (base + k < n) ? in[base + k] : 0u;
  }
  const uint32_t tsum = x[0] + x[1] + x[2] + x[3];
  uint32_t incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = s_warp[lane];
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
      if (lane >= o) wi += y;
    }
    s_warp[lane] = wi - w;
    if (lane == 31) tops[blockIdx.x] = wi;
  }
  __syncthreads();
  uint32_t run = s_warp[warp] + incl - tsum;
  uint4 o4; o4.x = run; o4.y = run + x[0]; o4.z = o4.y + x[1]; o4.w = o4.z + x[2];
  if (base + 3 < n) *reinterpret_cast<uint4*>(local + base) = o4;
  else {
    if (base < n) local[base] = o4.x;
    if (base + 1 < n) local[base + 1] = o4.y;
    if (base + 2 < n) local[base + 2] = o4.z;
  }
}

// ------------------------------------------------------------------------------------------
// gather: flat over output records, 8 lanes per record (4 records per warp step), all loads of a
// step issued before its stores.  Pure indexed copy: arena record -> hdr_out[r] + payload_out.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_recv_gather(sdb_dev_view v, sdb_recv_args r) {
  const uint32_t total = static_cast<uint32_t>(r.totals[0]);
  const uint32_t lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
  const uint64_t pol = sdb_policy_evict_first();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r0 = gw * 4u; r0 < total; r0 += nw * 4u) {
    const uint32_t rec = r0 + sub;
    if (rec >= total) continue;
    const uint32_t handle = r.plan_handle[rec];
    const uint32_t g = r.plan_glen[rec];
    const uint64_t po = static_cast<uint64_t>(r.plan_local[rec]) + r.plan_tops[rec / SDB_SCAN_TILE];
    const uint8_t* src = v.arena + ((static_cast<uint64_t>(handle) & v.gmask) << 5);
    uint8_t* hdst = reinterpret_cast<uint8_t*>(r.hdr_out + rec);
    uint8_t* pdst = r.payload_out + (po << 5) - 32;          // chunk c >= 2 lands at pdst + 16 c
    const uint32_t nchunk = 2u + (g << 1);
    for (uint32_t c = l8; c < nchunk; c += 24) {
      const uint32_t c1 = c + 8, c2 = c + 16;
      uint4 x0 = sdb_ld_stream_pol(src + (c << 4), pol), x1, x2;
      if (c1 < nchunk) x1 = sdb_ld_stream_pol(src + (c1 << 4), pol);
      if (c2 < nchunk) x2 = sdb_ld_stream_pol(src + (c2 << 4), pol);
      sdb_st_stream_pol((c < 2 ? hdst : pdst) + (c << 4), x0, pol);
      if (c1 < nchunk) sdb_st_stream_pol(pdst + (c1 << 4), x1, pol);
      if (c2 < nchunk) sdb_st_stream_pol(pdst + (c2 << 4), x2, pol);
    }
  }
}

// ------------------------------------------------------------------------------------------
// stream digests (definition in include/swarmdb_b200.h): one warp per request slot folds the records the last
// receive delivered to that agent, in delivery order, into digest[agent].  Lanes hash the 64-bit words of a
// record in parallel (the per-word terms are position-keyed and summed), lane 0 chains the records.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long sdb_fmix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
  return x;
}
__global__ void __launch_bounds__(256)
k_recv_digest(sdb_recv_args r, uint32_t max_agents, unsigned long long* __restrict__ digest) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
  const unsigned long long K = 0x9E3779B97F4A7C15ull;
  for (uint32_t q = gw; q < r.n; q += nw) {
    const uint32_t cnt = r.count_out[q];
    if (cnt == 0) continue;
    const uint32_t a = r.agent_idx ? r.agent_idx[q] : q;
    if (a >= max_agents) continue;
    const uint32_t roff = r.rec_local[q] + r.rec_tops[q / SDB_SCAN_TILE];
    unsigned long long d = digest[a];
    for (uint32_t j = 0; j < cnt; ++j) {
      const uint32_t rec = roff + j;
      const unsigned long long* hw = reinterpret_cast<const unsigned long long*>(r.hdr_out + rec);
      const uint32_t len = r.hdr_out[rec].len;
      const uint32_t nwords = 4u + (((len + 31u) & ~31u) >> 3);
      const unsigned long long po = (static_cast<unsigned long long>(r.plan_local[rec]) + r.plan_tops[rec / SDB_SCAN_TILE]) << 5;
      const unsigned long long* pw = reinterpret_cast<const unsigned long long*>(r.payload_out + po);
      unsigned long long acc = 0;
      for (uint32_t k = lane; k < nwords; k += 32) {
        const unsigned long long w = k < 4u ? hw[k] : pw[k - 4u];
        acc += sdb_fmix64(w ^ (static_cast<unsigned long long>(k + 1u) * K));
      }
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
      d = (((d << 5) | (d >> 59)) ^ acc) * K;
    }
    if (lane == 0) digest[a] = d;
  }
}

extern "C" cudaError_t sdb_launch_digest(const sdb_recv_args* r, uint32_t max_agents, unsigned long long* digest,
                                         int sm_count, cudaStream_t stream) {
  if (r->n == 0) return cudaSuccess;
  uint32_t grid = static_cast<uint32_t>(sm_count) * 8u;
  const uint32_t need = (r->n + 7) / 8;
  if (grid > need) grid = need;
  k_recv_digest<<<grid, 256, 0, stream>>>(*r, max_agents, digest);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// latency path: up to 8 agents, up to 1024 records, ONE launch (count, scan, select, retire, gather).
// Output block: u64 total records | u64 total granules | u32 counts[8] | pad to 64 B | the records
// back to back in arena format (32-B header + padded payload), so the host needs one D2H.
// ------------------------------------------------------------------------------------------
#define SDB_SMALL_AGENTS 8u
#define SDB_SMALL_RECS 1024u
struct sdb_small_agents { uint32_t idx[SDB_SMALL_AGENTS]; };

__global__ void __launch_bounds__(256)
k_recv_small(sdb_dev_view v, sdb_small_agents ag, uint32_t n, uint32_t max_messages, uint32_t flags, uint32_t rec_cap,
             uint32_t* __restrict__ plan_handle, uint32_t* __restrict__ plan_glen, uint8_t* __restrict__ out) {
  __shared__ uint32_t s_cnt[SDB_SMALL_AGENTS], s_roff[SDB_SMALL_AGENTS + 1], s_goff[SDB_SMALL_RECS + 1];
  __shared__ uint32_t s_total;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool mine = warp < n;
  const uint32_t a = mine ? ag.idx[warp] : 0u;
  uint32_t head = 0, tail = 0, nt = 0;
  if (lane == 0) {
    uint32_t c = 0;
    if (mine && a < v.max_agents) {
      const uint64_t st = v.ring_state[a];
      head = static_cast<uint32_t>(st); tail = static_cast<uint32_t>(st >> 32); nt = v.ntomb[a];
      c = min(tail - head - nt, max_messages);
    }
    s_cnt[warp] = c;
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t acc = 0, total = 0; bool open = true;
    for (uint32_t w = 0; w < SDB_SMALL_AGENTS; ++w) {
      s_roff[w] = acc;
      const uint32_t c = s_cnt[w];
      if (open && acc + c <= rec_cap) total = acc + c; else { open = false; s_cnt[w] = 0; }   // whole-agent truncation
      acc += c;
    }
    s_total = total;
  }
  __syncthreads();
  const uint32_t cnt = s_cnt[warp], roff = s_roff[warp];
  if (lane == 0 && mine) reinterpret_cast<uint32_t*>(out + 16)[warp] = cnt;
  if (cnt) {
    const uint32_t H = __shfl_sync(0xFFFFFFFFu, head, 0), T = __shfl_sync(0xFFFFFFFFu, tail, 0);
    const uint32_t NT = __shfl_sync(0xFFFFFFFFu, nt, 0);
    select_agent_warp(v, plan_handle, plan_glen, (flags & SDB_RECV_PRIORITY) != 0, a, H, T, NT, cnt, roff, lane,
                      !(flags & SDB_RECV_PEEK));
  }
  __syncthreads();
  const uint32_t total = s_total;
  if (warp == 0) {                       // record offsets (granules, header included)
    uint32_t run = 0;
    for (uint32_t r0 = 0; r0 < total; r0 += 32) {
      const uint32_t r = r0 + lane;
      const uint32_t g = r < total ? plan_glen[r] + 1u : 0u;
      uint32_t incl = g;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
      if (r < total) s_goff[r] = run + incl - g;
      run += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) {
      s_goff[total] = run;
      reinterpret_cast<unsigned long long*>(out)[0] = total;
      reinterpret_cast<unsigned long long*>(out)[1] = run;
      if (total && !(flags & SDB_RECV_PEEK)) atomicAdd(&v.ctr->delivered, static_cast<unsigned long long>(total));
    }
  }
  __syncthreads();
  const uint32_t l8 = tid & 7;
  for (uint32_t r = tid >> 3; r < total; r += 32) {
    const uint32_t g = plan_glen[r];
    const uint8_t* src = v.arena + ((static_cast<uint64_t>(plan_handle[r]) & v.gmask) << 5);
    uint8_t* dst = out + 64 + (static_cast<size_t>(s_goff[r]) << 5);
    const uint32_t nchunk = 2u + (g << 1);
    for (uint32_t c = l8; c < nchunk; c += 8) sdb_st_stream(dst + (c << 4), sdb_ld_stream(src + (c << 4)));
  }
}

extern "C" cudaError_t sdb_launch_receive_small(const sdb_dev_view* v, const uint32_t* agents_host, uint32_t n,
                                                uint32_t max_messages, uint32_t flags, uint32_t rec_cap,
                                                uint32_t* plan_handle, uint32_t* plan_glen, uint8_t* out,
                                                cudaStream_t stream, sdb_profiler* prof) {
  sdb_small_agents ag{};
  for (uint32_t i = 0; i < n && i < SDB_SMALL_AGENTS; ++i) ag.idx[i] = agents_host[i];
  const int pi = sdb_prof_begin(prof, SDB_PK_RECV_GATHER, stream);
  k_recv_small<<<1, 256, 0, stream>>>(*v, ag, n, max_messages, flags, rec_cap < SDB_SMALL_RECS ? rec_cap : SDB_SMALL_RECS,
                                      plan_handle, plan_glen, out);
  sdb_prof_end(prof, pi, stream);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
extern "C" cudaError_t sdb_launch_receive(const sdb_dev_view* v, const sdb_recv_args* r, cudaStream_t stream,
                                          int* n_launches, sdb_profiler* prof, int sm_count) {
  if (r->n == 0) return cudaSuccess;
  const uint32_t n = r->n;
  const uint32_t tiles = (n + SDB_SCAN_TILE - 1) / SDB_SCAN_TILE;
  int pi = sdb_prof_begin(prof, SDB_PK_RECV_COUNT, stream);
  k_recv_count<<<(n + 255) / 256, 256, 0, stream>>>(*v, *r);
  sdb_prof_end(prof, pi, stream);
  pi = sdb_prof_begin(prof, SDB_PK_RECV_SCAN, stream);
  k_scan_local<<<tiles, 1024, 0, stream>>>(r->cnt, 0xFFFFFFFFu, r->rec_local, r->rec_tops, n);
  k_scan_tops<<<1, 1024, 0, stream>>>(r->rec_tops, tiles, r->totals);          // totals[0] = records requested
  sdb_prof_end(prof, pi, stream);
  pi = sdb_prof_begin(prof, SDB_PK_RECV_SELECT, stream);
  cudaMemsetAsync(r->big_count, 0, sizeof(uint32_t), stream);
  k_recv_select<<<(n + 255) / 256, 256, 0, stream>>>(*v, *r);                   // may lower totals[0] (capacity)
  {
    uint32_t bg = static_cast<uint32_t>(sm_count) * 8u;                         // 8 warps per CTA, 8 CTAs per SM
    const uint32_t need = (n + 7) / 8;
    if (bg > need) bg = need;
    k_recv_select_big<<<bg, 256, 0, stream>>>(*v, *r);
  }
  sdb_prof_end(prof, pi, stream);
  // upper bound on records: min(rec_cap, n * max_messages); grids sized from it, kernels read the true count
  uint64_t bound = static_cast<uint64_t>(n) * r->max_messages;
  if (bound > r->rec_cap) bound = r->rec_cap;
  const uint32_t rtiles = static_cast<uint32_t>((bound + SDB_SCAN_TILE - 1) / SDB_SCAN_TILE);
  pi = sdb_prof_begin(prof, SDB_PK_RECV_SCAN, stream);
  k_scan_plan<<<rtiles, 1024, 0, stream>>>(r->plan_glen, r->plan_local, r->plan_tops, r->totals);
  k_scan_tops<<<1, 1024, 0, stream>>>(r->plan_tops, rtiles, r->totals + 1);    // totals[1] = payload granules
  sdb_prof_end(prof, pi, stream);
  uint64_t gwarps = (bound + 3) / 4;
  uint64_t gblocks = (gwarps + 7) / 8;
  const uint64_t cap = static_cast<uint64_t>(sm_count) * 8 * 4;                // grid-stride beyond ~4 waves
  if (gblocks > cap) gblocks = cap;
  if (gblocks == 0) gblocks = 1;
  pi = sdb_prof_begin(prof, SDB_PK_RECV_GATHER, stream);
  k_recv_gather<<<static_cast<uint32_t>(gblocks), 256, 0, stream>>>(*v, *r);
  sdb_prof_end(prof, pi, stream);
  if (n_launches) *n_launches += 8;
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
