// sdb_recv.cu - dequeue kernels (sm_100a): receive_messages (reference M:521-601; the drain
// loop M:553-601 and its filter M:579-585 are replaced by per-agent rings that only ever
// hold records addressed to that agent).
//
// Stream order (the reference's behaviour) takes the SINGLE-PASS path: k_recv_plan + k_recv_gather.
//   k_recv_plan     one thread per requested agent: one 16-byte ring-header load, the agent's entries (8 bytes
//                   each, one sector for the usual handful), then ONE decoupled look-back over 256-agent tiles
//                   carrying {records, payload granules} together gives every record its output index and
//                   payload offset; the per-record plan is written, heads advance.  No separate count / scan /
//                   select / scan kernels and no second read of the ring.
//   k_recv_gather   flat over records: copy header + payload from the arena to the packed output
// Priority order (extension, SURVEY App. A rule 9) keeps the multi-kernel pipeline (all stream-ordered, no host
// round trip inside):
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
#include <cstdlib>
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
    const uint4 h = *reinterpret_cast<const uint4*>(v.ring_hdr + a);      // head, tail, ctail, ntomb
    c = min(h.y - h.x - h.w, r.max_messages);
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
// output index RO.., retire them (advance head / tombstone).  Stream order when !prio_mode.  The plan's payload
// offsets (.y) are filled later (record-level scan / fill_offsets_warp).
__device__ __forceinline__ void select_agent_warp(const sdb_dev_view& v, uint4* plan, uint32_t slot,
                                                  bool prio_mode, uint32_t A, uint32_t H, uint32_t T, uint32_t NT,
                                                  uint32_t C, uint32_t RO, uint32_t lane, bool retire) {
  const uint32_t mask = v.ring_slots - 1;
  uint2* rs = sdb_ring_of(v, A);

  if (!prio_mode && NT == 0) {
    // long contiguous run [H, H+C)
    for (uint32_t j = lane; j < C; j += 32) {
      const uint2 e = rs[(H + j) & mask];
      plan[RO + j] = make_uint4(e.x, 0u, sdb_plan_z(e.y), slot);
    }
    if (lane == 0 && retire) v.ring_hdr[A].head = H + C;
    return;
  }
  // ---- pass 1: histogram of live entries per priority level over the window [H, T)
  uint32_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;
  if (prio_mode) {
    // 2 entries (16 bytes) per lane per load, 4 independent loads in flight per lane (256 entries per warp
    // step), aligned pairs of the circular window
    for (uint32_t b2 = (H & ~1u) + (lane << 1); static_cast<int32_t>(T - b2) > 0; b2 += 256) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t p2 = b2 + u * 64;
        q[u] = static_cast<int32_t>(T - p2) > 0 ? *reinterpret_cast<const uint4*>(rs + (p2 & mask)) : make_uint4(0u, ~0u, 0u, ~0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t p2 = b2 + u * 64;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint32_t p = p2 + k;
          const uint32_t m = (k ? q[u].w : q[u].y) & 0xFFFFu;
          if (static_cast<int32_t>(p - H) >= 0 && static_cast<int32_t>(T - p) > 0 && m != SDB_META_TOMB) {
            const uint32_t L = m >> 14;
            h0 += (L == 0); h1 += (L == 1); h2 += (L == 2); h3 += (L == 3);
          }
        }
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
    const uint2 e = in ? rs[p & mask] : make_uint2(0u, SDB_META_TOMB);
    const uint32_t m = sdb_meta(e);
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
      plan[RO + rank] = make_uint4(e.x, 0u, sdb_plan_z(e.y), slot);
      if (retire) rs[p & mask].y = SDB_META_TOMB;
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
    v.ring_hdr[A].head = nh;
    v.ring_hdr[A].ntomb = NT + got - (nh - H);
  }
}

__global__ void __launch_bounds__(256)
k_recv_select(sdb_dev_view v, sdb_recv_args r) {
  // one lane per requested agent: short contiguous runs are planned and retired here; everything else is
  // queued for k_recv_select_big (one warp per agent)
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
    r.rec_off[q] = roff;
    if (cnt) {
      const uint4 h = *reinterpret_cast<const uint4*>(v.ring_hdr + a);
      head = h.x; nt = h.w;
    }
  }
  const bool prio_mode = (r.flags & SDB_RECV_PRIORITY) != 0;
  constexpr uint32_t SMALL = 64, CH = 8;
  const bool small = valid && cnt && !prio_mode && nt == 0 && cnt <= SMALL;
  if (small) {
    const uint64_t pol = sdb_policy_evict_last();
    const uint2* rs = sdb_ring_of(v, a);
    for (uint32_t b = 0; b < cnt; b += CH) {
      uint64_t ev[CH];
#pragma unroll
      for (uint32_t j = 0; j < CH; ++j)       // all loads of the chunk first, then the stores
        if (b + j < cnt) ev[j] = sdb_ld_u64_pol(reinterpret_cast<const uint64_t*>(rs + ((head + b + j) & mask)), pol);
#pragma unroll
      for (uint32_t j = 0; j < CH; ++j)
        if (b + j < cnt)
          r.plan[roff + b + j] = make_uint4(static_cast<uint32_t>(ev[j]), 0u, sdb_plan_z(static_cast<uint32_t>(ev[j] >> 32)), q);
    }
    if (!(r.flags & SDB_RECV_PEEK)) v.ring_hdr[a].head = head + cnt;
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
__device__ __forceinline__ void select_agent_warp_bounded(const sdb_dev_view& v, uint4* plan, uint32_t slot,
                                                          bool prio_mode, uint32_t A, uint32_t H, uint32_t T, uint32_t NT,
                                                          uint32_t C, uint32_t RO, uint32_t lane, bool retire,
                                                          uint32_t (*cand)[SDB_CAND]) {
  const uint32_t mask = v.ring_slots - 1;
  uint2* rs = sdb_ring_of(v, A);
  const uint32_t top = prio_mode ? 3u : 0u;
  uint32_t cnt[4] = {0, 0, 0, 0};
  uint32_t p0 = H;
  while (static_cast<int32_t>(T - p0) > 0 && cnt[top] < C) {
    // four chunks of 32 entries in flight per step (the loads do not depend on the running counts)
    uint32_t mm[4];
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u) {
      const uint32_t p = p0 + u * 32u + lane;
      mm[u] = static_cast<int32_t>(T - p) > 0 ? (rs[p & mask].y & 0xFFFFu) : SDB_META_TOMB;
    }
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u) {
      if (static_cast<int32_t>(T - p0) <= 0 || cnt[top] >= C) break;       // warp-uniform
      const uint32_t p = p0 + lane;
      const uint32_t m = mm[u];
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
      const uint2 e = rs[p & mask];
      plan[RO + acc + k] = make_uint4(e.x, 0u, sdb_plan_z(e.y), slot);
      if (retire) rs[p & mask].y = SDB_META_TOMB;
    }
    acc += take; need -= take; got += take;
  }
  if (!retire) return;
  __syncwarp();
  // new head: first live entry left behind inside the scanned range (the tombstones just written count as consumed)
  uint32_t nh = scan_end;
  for (uint32_t q0 = H; static_cast<int32_t>(scan_end - q0) > 0; q0 += 32) {
    const uint32_t p = q0 + lane;
    const bool live = static_cast<int32_t>(scan_end - p) > 0 && (rs[p & mask].y & 0xFFFFu) != SDB_META_TOMB;
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, live);
    if (b) { nh = q0 + (__ffs(b) - 1); break; }
  }
  if (lane == 0) {
    v.ring_hdr[A].head = nh;
    v.ring_hdr[A].ntomb = NT + got - (nh - H);
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
    const uint4 hd = *reinterpret_cast<const uint4*>(v.ring_hdr + a);
    const uint32_t H = hd.x, T = hd.y, NT = hd.w;
    if (cnt <= SDB_CAND && (prio_mode || NT != 0))
      select_agent_warp_bounded(v, r.plan, q, prio_mode, a, H, T, NT, cnt, roff, lane, retire, s_cand[threadIdx.x >> 5]);
    else
      select_agent_warp(v, r.plan, q, prio_mode, a, H, T, NT, cnt, roff, lane, retire);
    __syncwarp();
  }
}

// scan over the per-record payload sizes (plan[].z) of the multi-kernel path; the tile-local exclusive prefix is
// stored back into plan[].y and the per-tile totals into `tops`.  The element count lives on the device (totals[0]).
__global__ void __launch_bounds__(1024)
k_scan_plan(uint4* __restrict__ plan, uint32_t* __restrict__ tops, const unsigned long long* __restrict__ n_dev) {
  __shared__ uint32_t s_warp[32];
  const uint32_t n = static_cast<uint32_t>(*n_dev);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t base = blockIdx.x * SDB_SCAN_TILE + tid * 4u;
  if (blockIdx.x * SDB_SCAN_TILE >= n) { if (tid == 0) tops[blockIdx.x] = 0; return; }
  uint32_t x[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = (base + k < n) ? (plan[base + k].z & 0xFFFFu) : 0u;
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
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (base + k < n) plan[base + k].y = run;
    run += x[k];
  }
}

// ------------------------------------------------------------------------------------------
// SINGLE-PASS receive plan for stream order (the reference's order): count + both scans + select + retire in one
// kernel.  One thread per request slot, 256 slots per tile, tiles ordered by an atomic ticket; one hierarchical
// decoupled look-back (sdb_common.cuh: sdb_lb_prefix) carries {records, payload granules} as a 31+31-bit pair, so a
// record's output index and its payload offset come out of the same pass.
// Agents whose window holds consumed entries (holes left by earlier priority receives) are walked by their warp.
// Capacity: an agent whose records would run past rec_cap is left untouched together with every agent after it
// (whole-agent truncation, nothing is ever lost); the totals are lowered with an atomic min on the packed pair.
//   r.lb               look-back scratch (layout in sdb_common.cuh, zeroed by the launcher), then the tile ticket, then
//                      ~(records << 32 | granules) of the call, maintained with atomicMax (= min of the value)
// ------------------------------------------------------------------------------------------
#ifndef SDB_PLAN_WARPS
#define SDB_PLAN_WARPS 8u
#endif
#define SDB_PLAN_TILE (32u * SDB_PLAN_WARPS)    // request slots per tile (one per thread)
#define SDB_PLAN_STAGE 256u                     // plan entries staged per warp before a coalesced write
// sum of the payload granules of the first C live entries of the window [H, T), whole warp
__device__ __forceinline__ uint32_t holey_sum_warp(const uint2* rs, uint32_t mask, uint32_t H, uint32_t T, uint32_t C, uint32_t lane) {
  uint32_t taken = 0, sum = 0;
  for (uint32_t p0 = H; static_cast<int32_t>(T - p0) > 0 && taken < C; p0 += 32) {
    const uint32_t p = p0 + lane;
    const uint32_t m = static_cast<int32_t>(T - p) > 0 ? (rs[p & mask].y & 0xFFFFu) : SDB_META_TOMB;
    const bool live = m != SDB_META_TOMB;
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, live);
    if (live && taken + lane_prefix(b, lane) < C) sum += (m & SDB_META_GLEN_MASK) - 1u;
    taken += __popc(b);
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
  return sum;
}
// payload offsets of C freshly planned records [RO, RO + C): running sum of .z from g0, whole warp
__device__ __forceinline__ void fill_offsets_warp(uint4* plan, uint32_t RO, uint32_t C, uint32_t g0, uint32_t lane) {
  uint32_t run = g0;
  for (uint32_t j0 = 0; j0 < C; j0 += 32) {
    const uint32_t j = j0 + lane;
    const uint32_t g = j < C ? (plan[RO + j].z & 0xFFFFu) : 0u;
    uint32_t incl = g;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
    if (j < C) plan[RO + j].y = run + incl - g;
    run += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
}

__global__ void __launch_bounds__(SDB_PLAN_TILE)
k_recv_plan(sdb_dev_view v, sdb_recv_args r, uint32_t tiles, uint32_t ticketed) {
  __shared__ uint32_t s_tile;
  __shared__ unsigned long long s_wr[SDB_PLAN_WARPS], s_wg[SDB_PLAN_WARPS];
  __shared__ unsigned long long s_br, s_bg;
  __shared__ uint4 s_stage[SDB_PLAN_WARPS][SDB_PLAN_STAGE];                // per warp: plan entries on their way out (32 KB)
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t ticket_at = static_cast<size_t>(tiles) + 3 * ((tiles + 31) / 32), totals_at = ticket_at + 1;
  // tile id: a ticket (dispatch order by construction) or, SDB_PLAN_TICKET=0, the block index (the hardware dispatches
  // the blocks of a 1-D grid in index order, so earlier tiles are resident or finished: saves one L2 round trip)
  if (ticketed) {
    if (tid == 0) s_tile = atomicAdd(reinterpret_cast<unsigned int*>(r.lb + ticket_at), 1u);
    __syncthreads();
  }
  const uint32_t tile = ticketed ? s_tile : blockIdx.x;
  const uint32_t q = tile * SDB_PLAN_TILE + tid;
  const bool valid = q < r.n;
  const bool retire = !(r.flags & SDB_RECV_PEEK);
  const uint32_t R = v.ring_slots, mask = R - 1;
  constexpr uint32_t KEEP = 8;
  const uint64_t pol = sdb_policy_evict_last();

  uint32_t a = 0, head = 0, tail = 0, nt = 0, cnt = 0;
  uint4 h4 = make_uint4(0, 0, 0, 0);
  if (valid) {
    a = r.agent_idx ? r.agent_idx[q] : q;
    if (a < v.max_agents) {
      h4 = *reinterpret_cast<const uint4*>(v.ring_hdr + a);                 // head, tail, ctail, ntomb
      head = h4.x; tail = h4.y; nt = h4.w;
      cnt = min(tail - head - nt, r.max_messages);
    }
  }
  const bool holey = cnt && nt != 0;
  const bool simple = cnt && !holey;
  const uint2* rs = sdb_ring_of(v, valid && a < v.max_agents ? a : 0u);
  uint64_t ev[KEEP];
  uint32_t gsum = 0;
  if (simple) {
    if (cnt <= KEEP) {
#pragma unroll
      for (uint32_t j = 0; j < KEEP; ++j)
        if (j < cnt) ev[j] = sdb_ld_u64_pol(reinterpret_cast<const uint64_t*>(rs + ((head + j) & mask)), pol);
#pragma unroll
      for (uint32_t j = 0; j < KEEP; ++j)
        if (j < cnt) gsum += (static_cast<uint32_t>(ev[j] >> 32) & SDB_META_GLEN_MASK) - 1u;
    } else {
      for (uint32_t j = 0; j < cnt; ++j) gsum += (rs[(head + j) & mask].y & SDB_META_GLEN_MASK) - 1u;
    }
  }
  // agents with holes: their warp walks the window (rare: stream-order receive after priority receives)
  for (uint32_t hb = __ballot_sync(0xFFFFFFFFu, holey); hb; hb &= hb - 1) {
    const uint32_t src = __ffs(hb) - 1;
    const uint32_t A_ = __shfl_sync(0xFFFFFFFFu, a, src), H_ = __shfl_sync(0xFFFFFFFFu, head, src);
    const uint32_t T_ = __shfl_sync(0xFFFFFFFFu, tail, src), C_ = __shfl_sync(0xFFFFFFFFu, cnt, src);
    const uint32_t sum = holey_sum_warp(sdb_ring_of(v, A_), mask, H_, T_, C_, lane);
    if (lane == src) gsum = sum;
  }

  // ---- tile scan of {records, granules} (two 64-bit sums: a pathological tile may exceed 32 bits before truncation)
  unsigned long long ir = cnt, ig = gsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long yr = __shfl_up_sync(0xFFFFFFFFu, ir, o), yg = __shfl_up_sync(0xFFFFFFFFu, ig, o);
    if (lane >= o) { ir += yr; ig += yg; }
  }
  if (lane == 31) { s_wr[warp] = ir; s_wg[warp] = ig; }
  __syncthreads();
  if (warp == 0) {
    const unsigned long long wr = lane < SDB_PLAN_WARPS ? s_wr[lane] : 0ull, wg = lane < SDB_PLAN_WARPS ? s_wg[lane] : 0ull;
    unsigned long long cr = wr, cg = wg;
#pragma unroll
    for (int o = 1; o < static_cast<int>(SDB_PLAN_WARPS); o <<= 1) {
      const unsigned long long yr = __shfl_up_sync(0xFFFFFFFFu, cr, o), yg = __shfl_up_sync(0xFFFFFFFFu, cg, o);
      if (lane >= o) { cr += yr; cg += yg; }
    }
    const unsigned long long tr = __shfl_sync(0xFFFFFFFFu, cr, SDB_PLAN_WARPS - 1), tg = __shfl_sync(0xFFFFFFFFu, cg, SDB_PLAN_WARPS - 1);   // tile aggregate
    if (lane < SDB_PLAN_WARPS) { s_wr[lane] = cr - wr; s_wg[lane] = cg - wg; }                                           // exclusive warp offsets
    // ---- hierarchical decoupled look-back (sdb_common.cuh): exclusive prefix of {records, granules} over earlier tiles
    unsigned long long er, eg, ga, gb; bool last;
    sdb_lb_prefix(r.lb, tile, tiles, tr, tg, lane, er, eg, last, ga, gb);
    if (lane == 0) {
      s_br = er; s_bg = eg;
      if (last)                    // grand total of the call (before truncation; truncation only lowers it)
        atomicMax(r.lb + totals_at, ~((min(ga, 0xFFFFFFFFull) << 32) | min(gb, 0xFFFFFFFFull)));
    }
  }
  __syncthreads();
  const unsigned long long pr = s_br + s_wr[warp] + ir - cnt, pg = s_bg + s_wg[warp] + ig - gsum;   // exclusive prefix of this slot
  const uint32_t roff = static_cast<uint32_t>(min(pr, 0xFFFFFFFFull));
  const uint32_t goff = static_cast<uint32_t>(min(pg, 0xFFFFFFFFull));
  if (cnt && pr + cnt > r.rec_cap) {                                            // does not fit: stays queued
    atomicMax(r.lb + totals_at, ~((static_cast<unsigned long long>(roff) << 32) | goff));
    cnt = 0;
  }
  if (valid) { r.count_out[q] = cnt; r.rec_off[q] = roff; }

  // ---- plan entries of the warp's simple agents: staged in shared memory, written as whole 32-byte sectors
  // (per-thread 16-byte stores to scattered places would make the L2 fetch every sector before merging it)
  {
    const uint32_t mycnt = (cnt && simple) ? cnt : 0u;
    uint32_t winc = mycnt;                                                      // inclusive scan over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, winc, o); if (lane >= o) winc += y; }
    const uint32_t wtotal = __shfl_sync(0xFFFFFFFFu, winc, 31);
    // the warp's simple records are NOT contiguous in the plan when it also holds holey or truncated agents; they are
    // contiguous per agent, so stage (destination, entry) pairs: destination index travels in a side array
    __shared__ uint32_t s_dst[SDB_PLAN_WARPS][SDB_PLAN_STAGE];
    const uint32_t lbase = winc - mycnt;                                        // local index of this lane's first entry
    for (uint32_t c0 = 0; c0 < wtotal; c0 += SDB_PLAN_STAGE) {
      uint32_t g0 = goff;
      if (mycnt) {
        for (uint32_t j = 0; j < mycnt; ++j) {
          uint64_t e;
          if (cnt <= KEEP) {
            e = ev[0];
#pragma unroll
            for (uint32_t u = 1; u < KEEP; ++u) if (j == u) e = ev[u];
          } else {
            e = sdb_ld_u64_pol(reinterpret_cast<const uint64_t*>(rs + ((head + j) & mask)), pol);
          }
          const uint32_t g = (static_cast<uint32_t>(e >> 32) & SDB_META_GLEN_MASK) - 1u;
          const uint32_t li = lbase + j;
          if (li >= c0 && li < c0 + SDB_PLAN_STAGE) {
            s_stage[warp][li - c0] = make_uint4(static_cast<uint32_t>(e), g0, g | (static_cast<uint32_t>(e >> 32) & 0xFFFF0000u), q);
            s_dst[warp][li - c0] = roff + j;
          }
          g0 += g;
        }
      }
      __syncwarp();
      const uint32_t nst = min(SDB_PLAN_STAGE, wtotal - c0);
      for (uint32_t i = lane; i < nst; i += 32) r.plan[s_dst[warp][i]] = s_stage[warp][i];
      __syncwarp();
    }
    // the whole 16-byte header goes back in one store: neighbouring agents' headers then leave as full sectors
    // (a 4-byte store per agent is a partial-sector write the L2 has to merge)
    if (mycnt && retire) { h4.x = head + cnt; *reinterpret_cast<uint4*>(v.ring_hdr + a) = h4; }
  }
  for (uint32_t hb = __ballot_sync(0xFFFFFFFFu, holey && cnt); hb; hb &= hb - 1) {
    const uint32_t src = __ffs(hb) - 1;
    const uint32_t A_ = __shfl_sync(0xFFFFFFFFu, a, src), H_ = __shfl_sync(0xFFFFFFFFu, head, src);
    const uint32_t T_ = __shfl_sync(0xFFFFFFFFu, tail, src), C_ = __shfl_sync(0xFFFFFFFFu, cnt, src);
    const uint32_t N_ = __shfl_sync(0xFFFFFFFFu, nt, src), RO_ = __shfl_sync(0xFFFFFFFFu, roff, src);
    const uint32_t G_ = __shfl_sync(0xFFFFFFFFu, goff, src), Q_ = __shfl_sync(0xFFFFFFFFu, q, src);
    select_agent_warp(v, r.plan, Q_, false, A_, H_, T_, N_, C_, RO_, lane, retire);
    __threadfence_block();
    __syncwarp();
    fill_offsets_warp(r.plan, RO_, C_, G_, lane);
  }
  uint32_t n_deliv = retire ? cnt : 0u;
  for (int o = 16; o; o >>= 1) n_deliv += __shfl_xor_sync(0xFFFFFFFFu, n_deliv, o);
  if (lane == 0 && n_deliv) atomicAdd(&v.ctr->delivered, static_cast<unsigned long long>(n_deliv));
}

// ------------------------------------------------------------------------------------------
// gather: flat over output records, 8 lanes per record (4 records per warp step), all loads of a
// step issued before its stores.  Pure indexed copy: arena record -> hdr_out[r] + payload_out.
// The first thread also publishes the call's totals where the host (and the digest kernel) read them.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void gather_record(const sdb_dev_view& v, const sdb_recv_args& r, uint32_t rec, uint32_t l8, uint64_t pol) {
  const uint4 pe = __ldg(r.plan + rec);                        // handle, payload offset, payload granules, slot
  const uint32_t g = pe.z & 0xFFFFu;
  const uint64_t po = static_cast<uint64_t>(pe.y) + (r.plan_tops ? r.plan_tops[rec / SDB_SCAN_TILE] : 0u);
  const uint8_t* src = v.arena + ((static_cast<uint64_t>(pe.x) & v.gmask) << 5);
  const uint8_t* psrc = sdb_payload_of(v, pe.x, sdb_entry_dm1(pe.z)) - 32;   // payload chunk c >= 2 comes from psrc + 16 c
  uint8_t* hdst = reinterpret_cast<uint8_t*>(r.hdr_out + rec);
  uint8_t* pdst = r.payload_out + (po << 5) - 32;          // chunk c >= 2 lands at pdst + 16 c
  const uint32_t nchunk = 2u + (g << 1);
  for (uint32_t c = l8; c < nchunk; c += 24) {
    const uint32_t c1 = c + 8, c2 = c + 16;
    uint4 x0 = sdb_ld_stream_pol((c < 2 ? src : psrc) + (c << 4), pol), x1, x2;
    if (c1 < nchunk) x1 = sdb_ld_stream_pol(psrc + (c1 << 4), pol);
    if (c2 < nchunk) x2 = sdb_ld_stream_pol(psrc + (c2 << 4), pol);
    sdb_st_stream_pol((c < 2 ? hdst : pdst) + (c << 4), x0, pol);
    if (c1 < nchunk) sdb_st_stream_pol(pdst + (c1 << 4), x1, pol);
    if (c2 < nchunk) sdb_st_stream_pol(pdst + (c2 << 4), x2, pol);
  }
}
__device__ __forceinline__ uint32_t gather_total(const sdb_recv_args& r, const unsigned long long* packed_inv) {
  if (!packed_inv) return static_cast<uint32_t>(r.totals[0]);
  const unsigned long long pk = ~*packed_inv;
  if (blockIdx.x == 0 && threadIdx.x == 0) { r.totals[0] = pk >> 32; r.totals[1] = pk & 0xFFFFFFFFull; }
  return static_cast<uint32_t>(pk >> 32);
}

__global__ void __launch_bounds__(256)
k_recv_gather(sdb_dev_view v, sdb_recv_args r, const unsigned long long* __restrict__ packed_inv) {
  const uint32_t total = gather_total(r, packed_inv);
  const uint32_t lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
  const uint64_t pol = sdb_policy_evict_first();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r0 = gw * 4u; r0 < total; r0 += nw * 4u) {
    const uint32_t rec = r0 + sub;
    if (rec < total) gather_record(v, r, rec, l8, pol);
  }
}

// The same copy with the TMA engine doing the moving: every lane owns one record per step and issues ONE bulk load of the
// whole record (header + padded payload, 32-byte granules) into its shared-memory slot, completion counted on the
// warp's mbarrier; when a step has landed the lane issues two bulk stores (header -> hdr_out[r], payload -> the packed
// stream).  Two stages per warp: while step i is being stored, step i + 1 is already in flight.  A lane never touches
// the bytes, so the kernel needs few registers, and a CTA keeps stage_bytes x 32 lanes x 2 stages x warps in flight -
// several times what 16-byte loads held in registers can.  Used when a record fits the per-lane slot.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_recv_gather_tma(sdb_dev_view v, sdb_recv_args r, const unsigned long long* __restrict__ packed_inv, uint32_t slot_bytes) {
  extern __shared__ __align__(128) uint8_t s_dyn[];          // [WARPS][2][32][slot_bytes]
  __shared__ __align__(8) uint64_t s_bar[WARPS][2];
  const uint32_t total = gather_total(r, packed_inv);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* const wbase = s_dyn + static_cast<size_t>(warp) * 2u * 32u * slot_bytes;
  if (lane == 0) { sdb_mbar_init(&s_bar[warp][0], 1); sdb_mbar_init(&s_bar[warp][1], 1); sdb_fence_barrier_init(); }
  __syncwarp();
  const uint32_t gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
  uint32_t phase = 0;                                        // bit s = parity of stage s
  // per-lane state of the two stages.  The 32-byte header travels through registers (two 16-byte loads now, two stores
  // one step later: the lanes' headers are consecutive in hdr_out, so a warp stores 1 KB contiguous); the payload - right
  // behind the header, or shared by all recipients of a group send (sdb_common.cuh) - takes ONE bulk load and ONE
  // bulk store.  Two bulk operations per record instead of three: the copy engine, not DRAM, was the limit once the
  // payloads came out of the L2.
  uint32_t g0 = 0, g1 = 0, rc0 = 0, rc1 = 0; uint64_t po0 = 0, po1 = 0; bool hv0 = false, hv1 = false;
  uint4 ha0 = make_uint4(0, 0, 0, 0), hb0 = ha0, ha1 = ha0, hb1 = ha0;
  uint4 pe_ahead = make_uint4(0, 0, 0, 0);                   // plan entry of the step after the one being requested
  auto issue = [&](uint32_t st, uint32_t r0) {               // all lanes: request step r0 into stage st
    const uint32_t rec = r0 + lane;
    const bool have = rec < total;
    uint32_t bytes = 0; uint4 pe = make_uint4(0, 0, 0, 0);
    if (have) { pe = pe_ahead; bytes = (pe.z & 0xFFFFu) << 5; }
    { const uint32_t ra = rec + nw * 32u; if (ra < total) pe_ahead = __ldg(r.plan + ra); }   // consumed one step later
    uint32_t sum = bytes;
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if (lane == 0) sdb_mbar_expect_tx(&s_bar[warp][st], sum);
    __syncwarp();
    uint4 ha = make_uint4(0, 0, 0, 0), hb = ha;
    if (have) {
      const uint8_t* src = v.arena + ((static_cast<uint64_t>(pe.x) & v.gmask) << 5);
      ha = sdb_ld_stream(src); hb = sdb_ld_stream(src + 16);
      if (bytes) sdb_tma_load(wbase + (static_cast<size_t>(st) * 32u + lane) * slot_bytes, sdb_payload_of(v, pe.x, sdb_entry_dm1(pe.z)), bytes, &s_bar[warp][st]);
    }
    const uint64_t po = static_cast<uint64_t>(pe.y) + ((have && r.plan_tops) ? r.plan_tops[rec / SDB_SCAN_TILE] : 0u);
    if (st) { g1 = pe.z & 0xFFFFu; rc1 = rec; hv1 = have; po1 = po; ha1 = ha; hb1 = hb; }
    else { g0 = pe.z & 0xFFFFu; rc0 = rec; hv0 = have; po0 = po; ha0 = ha; hb0 = hb; }
  };
  uint32_t r0 = gw * 32u;
  if (r0 + lane < total) pe_ahead = __ldg(r.plan + r0 + lane);
  if (r0 < total) issue(0, r0);
  uint32_t st = 0;
  for (; r0 < total; r0 += nw * 32u, st ^= 1u) {
    const uint32_t rn = r0 + nw * 32u;
    // the other stage was last read by the bulk stores of the step before this one: they must have finished reading it
    sdb_tma_wait_read<0>();
    __syncwarp();
    if (rn < total) issue(st ^ 1u, rn);
    if (st ? hv1 : hv0) {
      uint4* hdst = reinterpret_cast<uint4*>(r.hdr_out + (st ? rc1 : rc0));
      sdb_st_stream(hdst, st ? ha1 : ha0); sdb_st_stream(hdst + 1, st ? hb1 : hb0);
    }
    sdb_mbar_wait_bounded(&s_bar[warp][st], (phase >> st) & 1u);
    phase ^= 1u << st;
    {
      const bool hv = st ? hv1 : hv0;
      const uint32_t g = st ? g1 : g0;
      // payload offsets are cumulative in record order: when all 32 payloads fill their slots exactly, the warp's output
      // is ONE contiguous run and the slots are contiguous too - one bulk store instead of 32
      if (__all_sync(0xFFFFFFFFu, hv && (g << 5) == slot_bytes)) {
        if (lane == 0) sdb_tma_store(r.payload_out + ((st ? po1 : po0) << 5), wbase + static_cast<size_t>(st) * 32u * slot_bytes, 32u * slot_bytes);
      } else if (hv && g) {
        sdb_tma_store(r.payload_out + ((st ? po1 : po0) << 5), wbase + (static_cast<size_t>(st) * 32u + lane) * slot_bytes, g << 5);
      }
    }
    sdb_tma_commit();
  }
  sdb_tma_wait_all<0>();
}

// The same gather with the payloads brought in by the whole warp with 16-byte asynchronous copies (cp.async, LDGSTS)
// instead of one bulk load per lane.  A per-lane bulk copy costs the warp ~12 issue slots (the copy engine takes its
// operands from uniform registers, so the 32 lanes issue one after the other); with the payloads coming out of the L2
// that issue loop - not memory - looked like the limit (ncu: 17 % issue active, long_scoreboard 11).  MEASURED: it is not -
// this form runs the c2 gather in 0.382 ms against 0.357 ms for the bulk-load form (SDB_GATHER_LDGSTS=1 selects it; both
// pass the whole parity suite); the gather moves 1.6 GB of physical DRAM traffic at 4.5 TB/s either way.  Here one warp instruction
// moves 512 bytes: lane l copies chunk (flat % cpr) of record (flat / cpr), flat = 32 f + l, the records' source
// addresses travelling by shuffle.  Completion by cp.async groups (one per step, also when the step is empty), then a
// proxy fence, then ONE bulk store of the warp's 32 slots when they are all full (else one per lane).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_recv_gather_cpa(sdb_dev_view v, sdb_recv_args r, const unsigned long long* __restrict__ packed_inv, uint32_t slot_bytes) {
  extern __shared__ __align__(128) uint8_t s_dyn[];          // [WARPS][2][32][slot_bytes]
  const uint32_t total = gather_total(r, packed_inv);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* const wbase = s_dyn + static_cast<size_t>(warp) * 2u * 32u * slot_bytes;
  const uint32_t wbase_sa = sdb_smem_u32(wbase);
  const uint32_t gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
  const uint32_t cpr = slot_bytes >> 4;                      // 16-byte chunks per slot
  uint32_t g0 = 0, g1 = 0, rc0 = 0, rc1 = 0; uint64_t po0 = 0, po1 = 0; bool hv0 = false, hv1 = false;
  uint4 ha0 = make_uint4(0, 0, 0, 0), hb0 = ha0, ha1 = ha0, hb1 = ha0;
  auto issue = [&](uint32_t st, uint32_t r0) {               // all lanes: request step r0 into stage st
    const uint32_t rec = r0 + lane;
    const bool have = rec < total;
    uint32_t bytes = 0; uint4 pe = make_uint4(0, 0, 0, 0);
    if (have) { pe = __ldg(r.plan + rec); bytes = (pe.z & 0xFFFFu) << 5; }
    uint4 ha = make_uint4(0, 0, 0, 0), hb = ha;
    unsigned long long psrc = 0;
    if (have) {
      const uint8_t* src = v.arena + ((static_cast<uint64_t>(pe.x) & v.gmask) << 5);
      ha = sdb_ld_stream(src); hb = sdb_ld_stream(src + 16);
      psrc = reinterpret_cast<unsigned long long>(sdb_payload_of(v, pe.x, sdb_entry_dm1(pe.z)));
    }
    const uint32_t stage_sa = wbase_sa + st * 32u * slot_bytes;
    for (uint32_t f = 0; f < cpr; ++f) {
      const uint32_t flat = (f << 5) + lane;
      const uint32_t k = flat / cpr, c = flat - k * cpr;     // record of the step, chunk of its payload
      const unsigned long long sk = __shfl_sync(0xFFFFFFFFu, psrc, k);
      const uint32_t bk = __shfl_sync(0xFFFFFFFFu, bytes, k);
      if ((c << 4) < bk)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(stage_sa + k * slot_bytes + (c << 4)), "l"(sk + (c << 4)) : "memory");
    }
    const uint64_t po = static_cast<uint64_t>(pe.y) + ((have && r.plan_tops) ? r.plan_tops[rec / SDB_SCAN_TILE] : 0u);
    if (st) { g1 = pe.z & 0xFFFFu; rc1 = rec; hv1 = have; po1 = po; ha1 = ha; hb1 = hb; }
    else { g0 = pe.z & 0xFFFFu; rc0 = rec; hv0 = have; po0 = po; ha0 = ha; hb0 = hb; }
  };
  uint32_t r0 = gw * 32u;
  if (r0 < total) issue(0, r0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  uint32_t st = 0;
  for (; r0 < total; r0 += nw * 32u, st ^= 1u) {
    const uint32_t rn = r0 + nw * 32u;
    sdb_tma_wait_read<0>();                                  // the other stage's bulk stores have finished reading it
    __syncwarp();
    if (rn < total) issue(st ^ 1u, rn);
    asm volatile("cp.async.commit_group;" ::: "memory");     // one group per step, empty or not: the wait below counts groups
    const bool hv = st ? hv1 : hv0;
    if (hv) {
      uint4* hdst = reinterpret_cast<uint4*>(r.hdr_out + (st ? rc1 : rc0));
      sdb_st_stream(hdst, st ? ha1 : ha0); sdb_st_stream(hdst + 1, st ? hb1 : hb0);
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");     // everything but the newest group: this stage has landed
    __syncwarp();
    sdb_fence_proxy_async();                                 // copies written through the generic proxy -> visible to the bulk store
    const uint32_t g = st ? g1 : g0;
    if (__all_sync(0xFFFFFFFFu, hv && (g << 5) == slot_bytes)) {
      if (lane == 0) sdb_tma_store(r.payload_out + ((st ? po1 : po0) << 5), wbase + static_cast<size_t>(st) * 32u * slot_bytes, 32u * slot_bytes);
    } else if (hv && g) {
      sdb_tma_store(r.payload_out + ((st ? po1 : po0) << 5), wbase + (static_cast<size_t>(st) * 32u + lane) * slot_bytes, g << 5);
    }
    sdb_tma_commit();
  }
  sdb_tma_wait_all<0>();
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
    const uint32_t roff = r.rec_off[q];
    unsigned long long d = digest[a];
    for (uint32_t j = 0; j < cnt; ++j) {
      const uint32_t rec = roff + j;
      const unsigned long long* hw = reinterpret_cast<const unsigned long long*>(r.hdr_out + rec);
      const uint32_t len = r.hdr_out[rec].len;
      const uint32_t nwords = 4u + (((len + 31u) & ~31u) >> 3);
      const unsigned long long po = (static_cast<unsigned long long>(r.plan[rec].y) + (r.plan_tops ? r.plan_tops[rec / SDB_SCAN_TILE] : 0u)) << 5;
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
             uint4* __restrict__ plan, uint8_t* __restrict__ out) {
  __shared__ uint32_t s_cnt[SDB_SMALL_AGENTS], s_roff[SDB_SMALL_AGENTS + 1], s_goff[SDB_SMALL_RECS + 1];
  __shared__ uint32_t s_total;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool mine = warp < n;
  const uint32_t a = mine ? ag.idx[warp] : 0u;
  uint32_t head = 0, tail = 0, nt = 0;
  if (lane == 0) {
    uint32_t c = 0;
    if (mine && a < v.max_agents) {
      const uint4 h = *reinterpret_cast<const uint4*>(v.ring_hdr + a);
      head = h.x; tail = h.y; nt = h.w;
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
    select_agent_warp(v, plan, warp, (flags & SDB_RECV_PRIORITY) != 0, a, H, T, NT, cnt, roff, lane,
                      !(flags & SDB_RECV_PEEK));
  }
  __syncthreads();
  const uint32_t total = s_total;
  if (warp == 0) {                       // record offsets (granules, header included)
    uint32_t run = 0;
    for (uint32_t r0 = 0; r0 < total; r0 += 32) {
      const uint32_t r = r0 + lane;
      const uint32_t g = r < total ? (plan[r].z & 0xFFFFu) + 1u : 0u;
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
    const uint4 pe = plan[r];
    const uint32_t g = pe.z & 0xFFFFu;
    const uint8_t* src = v.arena + ((static_cast<uint64_t>(pe.x) & v.gmask) << 5);
    const uint8_t* psrc = sdb_payload_of(v, pe.x, sdb_entry_dm1(pe.z)) - 32;
    uint8_t* dst = out + 64 + (static_cast<size_t>(s_goff[r]) << 5);
    const uint32_t nchunk = 2u + (g << 1);
    for (uint32_t c = l8; c < nchunk; c += 8) sdb_st_stream(dst + (c << 4), sdb_ld_stream((c < 2 ? src : psrc) + (c << 4)));
  }
}

// ------------------------------------------------------------------------------------------
// Low-latency dequeue SERVER: the same single-launch receive as k_recv_small for one agent, as a persistent CTA that
// polls a mailbox in mapped pinned host memory and writes its answer (the k_recv_small output block) into pinned host
// memory.  Request: {req_seq, agent, max_messages, flags}; the host bumps req_seq last.  Answer complete when
// done_seq == req_seq.  quit != 0 ends the kernel.  Mailbox layout: sdb_ls_mailbox (sdb_common.cuh).
// A persistent kernel sees other kernels' writes only through the L2: every request starts with a gpu-scope fence,
// which also invalidates this SM's L1 (B300_MICROARCH: fence scope >= cluster emits CCTL.IVALL), and the ring header
// is read with volatile loads.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
k_latency_server(sdb_dev_view v, sdb_ls_mailbox* mb, uint8_t* out, uint32_t out_cap_bytes, uint4* plan) {
  __shared__ uint32_t s_req[4];         // seq, agent, max_messages, flags
  __shared__ uint32_t s_quit, s_cnt, s_total;
  __shared__ uint32_t s_goff[SDB_SMALL_RECS + 1];
  __shared__ uint2 s_fast[32];          // fast path: {arena handle, payload granules} of up to 32 records
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t served = 0;
  if (tid == 0) served = *reinterpret_cast<volatile uint32_t*>(&mb->done_seq);
  for (;;) {
    if (tid == 0) {
      // ONE 16-byte read per poll fetches the whole request {req_seq, agent, max_messages, flags}: every access to the
      // mailbox is a PCIe round trip, so the fields must not be fetched one by one.  flags bit 31 = quit.
      uint4 rq;
      for (;;) {
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rq.x), "=r"(rq.y), "=r"(rq.z), "=r"(rq.w) : "l"(mb) : "memory");
        if (rq.x != served || (rq.w & 0x80000000u)) break;
        __nanosleep(20);
      }
      s_req[0] = rq.x; s_req[1] = rq.y; s_req[2] = rq.z; s_req[3] = rq.w & 0x7FFFFFFFu;
      s_quit = rq.w >> 31;
    }
    __syncthreads();
    if (s_quit) return;
    __threadfence();                                    // fresh view of the rings and the arena (L1 invalidated)
    const uint32_t a = s_req[1], max_messages = s_req[2], flags = s_req[3];
    const uint32_t rec_cap = min(SDB_SMALL_RECS, out_cap_bytes / 64u);
    uint32_t head = 0, tail = 0, nt = 0;
    if (tid < 32) {                                     // warp 0: header, then (short stream-order answers) the entries themselves
      uint32_t c = 0;
      if (a < v.max_agents) {
        const uint4 hd = __ldcv(reinterpret_cast<const uint4*>(v.ring_hdr + a));
        head = hd.x; tail = hd.y; nt = hd.w;
        c = min(min(tail - head - nt, max_messages), rec_cap);
      }
      const bool fast = c && c <= 32u && nt == 0 && !(flags & SDB_RECV_PRIORITY);
      if (fast) {
        // lane j owns record j: entry -> arena handle and size; offsets by a shuffle scan; nothing goes through memory
        uint2 e = make_uint2(0u, 1u);
        if (lane < c) e = __ldcv(sdb_ring_of(v, a) + ((head + lane) & (v.ring_slots - 1)));
        const uint32_t g = lane < c ? (sdb_meta(e) & SDB_META_GLEN_MASK) : 0u;       // granules incl. the header
        uint32_t incl = g;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
        if (lane < c) { s_goff[lane] = incl - g; s_fast[lane] = make_uint2(e.x, (g - 1u) | (e.y & 0xFFFF0000u)); }
        if (lane == c - 1) s_goff[c] = incl;
        if (lane == 0) {
          s_total = c; s_cnt = 0xFFFFFFFFu;                                          // marker: plan lives in shared memory
          if (!(flags & SDB_RECV_PEEK)) v.ring_hdr[a].head = head + c;
        }
      } else if (lane == 0) {
        s_cnt = c;
      }
    }
    __syncthreads();
    const bool fastpath = s_cnt == 0xFFFFFFFFu;
    const uint32_t cnt = fastpath ? 0u : s_cnt;
    if (warp == 0 && cnt) {
      const uint32_t H = __shfl_sync(0xFFFFFFFFu, head, 0), T = __shfl_sync(0xFFFFFFFFu, tail, 0);
      const uint32_t NT = __shfl_sync(0xFFFFFFFFu, nt, 0);
      select_agent_warp(v, plan, 0u, (flags & SDB_RECV_PRIORITY) != 0, a, H, T, NT, cnt, 0u, lane, !(flags & SDB_RECV_PEEK));
      __threadfence_block();
      __syncwarp();
      // record offsets (granules, header included)
      uint32_t run = 0;
      for (uint32_t r0 = 0; r0 < cnt; r0 += 32) {
        const uint32_t r = r0 + lane;
        const uint32_t g = r < cnt ? (plan[r].z & 0xFFFFu) + 1u : 0u;
        uint32_t incl = g;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
        if (r < cnt) s_goff[r] = run + incl - g;
        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
      }
      if (lane == 0) { s_goff[cnt] = run; s_total = cnt; }
    }
    if (tid == 0 && !cnt && !fastpath) { s_total = 0; s_goff[0] = 0; }
    __syncthreads();
    const uint32_t total = s_total;
    const uint32_t l8 = tid & 7;
    for (uint32_t r = tid >> 3; r < total; r += 32) {
      uint32_t handle, gz;
      if (fastpath) { handle = s_fast[r].x; gz = s_fast[r].y; } else { const uint4 pe = plan[r]; handle = pe.x; gz = pe.z; }
      const uint32_t g = gz & 0xFFFFu;
      const uint8_t* src = v.arena + ((static_cast<uint64_t>(handle) & v.gmask) << 5);
      const uint8_t* psrc = sdb_payload_of(v, handle, sdb_entry_dm1(gz)) - 32;
      uint8_t* dst = out + 64 + (static_cast<size_t>(s_goff[r]) << 5);
      const uint32_t nchunk = 2u + (g << 1);
      for (uint32_t c = l8; c < nchunk; c += 8) *reinterpret_cast<uint4*>(dst + (c << 4)) = sdb_ld_stream((c < 2 ? src : psrc) + (c << 4));
    }
    if (tid == 0) {
      // totals | granules | count: one 32-byte store group at the head of the answer block
      reinterpret_cast<unsigned long long*>(out)[0] = total;
      reinterpret_cast<unsigned long long*>(out)[1] = s_goff[total];
      reinterpret_cast<uint32_t*>(out + 16)[0] = total;
      if (total && !(flags & SDB_RECV_PEEK)) atomicAdd(&v.ctr->delivered, static_cast<unsigned long long>(total));
    }
    __threadfence_system();                              // the answer is in host memory before the completion flag
    __syncthreads();
    if (tid == 0) {
      *reinterpret_cast<volatile uint32_t*>(&mb->done_seq) = s_req[0];
      __threadfence_system();
    }
    served = s_req[0];
    __syncthreads();
  }
}

extern "C" cudaError_t sdb_launch_latency_server(const sdb_dev_view* v, sdb_ls_mailbox* mb_dev, uint8_t* out_dev, uint32_t out_cap,
                                                 uint4* plan, cudaStream_t stream) {
  k_latency_server<<<1, 256, 0, stream>>>(*v, mb_dev, out_dev, out_cap, plan);
  return cudaGetLastError();
}

extern "C" cudaError_t sdb_launch_receive_small(const sdb_dev_view* v, const uint32_t* agents_host, uint32_t n,
                                                uint32_t max_messages, uint32_t flags, uint32_t rec_cap,
                                                uint4* plan, uint8_t* out,
                                                cudaStream_t stream, sdb_profiler* prof) {
  sdb_small_agents ag{};
  for (uint32_t i = 0; i < n && i < SDB_SMALL_AGENTS; ++i) ag.idx[i] = agents_host[i];
  const int pi = sdb_prof_begin(prof, SDB_PK_RECV_GATHER, stream);
  k_recv_small<<<1, 256, 0, stream>>>(*v, ag, n, max_messages, flags, rec_cap < SDB_SMALL_RECS ? rec_cap : SDB_SMALL_RECS,
                                      plan, out);
  sdb_prof_end(prof, pi, stream);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// max_rec_bytes: largest record (32-byte header + padded payload) this handle can hold; the TMA gather keeps one
// shared-memory slot of that size per lane and stage, so it is used for records up to 544 bytes (512-byte payloads)
static cudaError_t launch_gather(const sdb_dev_view* v, const sdb_recv_args* r, const unsigned long long* packed_inv,
                                 uint64_t bound, uint32_t max_rec_bytes, int sm_count, cudaStream_t stream,
                                 sdb_profiler* prof, int* n_launches) {
  static int use_tma = -1;
  if (use_tma < 0) { const char* e = getenv("SDB_GATHER_TMA"); use_tma = e ? atoi(e) : 1; }
  const int pi = sdb_prof_begin(prof, SDB_PK_RECV_GATHER, stream);
  if (use_tma && max_rec_bytes <= 544 && bound >= 4096) {
    constexpr int WARPS = 4;
    const uint32_t slot = max_rec_bytes > 64u ? max_rec_bytes - 32u : 32u;          // a slot stages the payload only
    const size_t smem = static_cast<size_t>(WARPS) * 2u * 32u * slot;
    uint32_t per_sm = static_cast<uint32_t>((226u * 1024u) / (smem + 1024 + 128));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    static int waves = -1;
    if (waves < 0) { const char* e = getenv("SDB_GATHER_WAVES"); waves = e ? atoi(e) : 8; if (waves < 1) waves = 1; }
    uint64_t grid = static_cast<uint64_t>(sm_count) * per_sm * waves;              // a few waves: the tail evens out
    const uint64_t need = (bound + WARPS * 32 - 1) / (WARPS * 32);
    if (grid > need) grid = need;
    static int use_cpa = -1;
    if (use_cpa < 0) { const char* e = getenv("SDB_GATHER_LDGSTS"); use_cpa = e ? atoi(e) : 0; }   // measured slower: off
    if (use_cpa) k_recv_gather_cpa<WARPS><<<static_cast<uint32_t>(grid), WARPS * 32, smem, stream>>>(*v, *r, packed_inv, slot);
    else k_recv_gather_tma<WARPS><<<static_cast<uint32_t>(grid), WARPS * 32, smem, stream>>>(*v, *r, packed_inv, slot);
  } else {
    uint64_t gwarps = (bound + 3) / 4;
    uint64_t gblocks = (gwarps + 7) / 8;
    const uint64_t cap = static_cast<uint64_t>(sm_count) * 8 * 4;                // grid-stride beyond ~4 waves
    if (gblocks > cap) gblocks = cap;
    if (gblocks == 0) gblocks = 1;
    k_recv_gather<<<static_cast<uint32_t>(gblocks), 256, 0, stream>>>(*v, *r, packed_inv);
  }
  sdb_prof_end(prof, pi, stream);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

extern "C" cudaError_t sdb_recv_prepare_device() {
  cudaError_t e = cudaFuncSetAttribute(k_recv_gather_cpa<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2 * 32 * 544);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_recv_gather_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2 * 32 * 544);
}

extern "C" cudaError_t sdb_launch_receive(const sdb_dev_view* v, const sdb_recv_args* r0, cudaStream_t stream,
                                          int* n_launches, sdb_profiler* prof, int sm_count, uint32_t max_rec_bytes) {
  if (r0->n == 0) return cudaSuccess;
  sdb_recv_args ra = *r0;
  sdb_recv_args* r = &ra;
  const uint32_t n = r->n;
  // upper bound on records: min(rec_cap, n * max_messages); grids sized from it, kernels read the true count
  uint64_t bound = static_cast<uint64_t>(n) * r->max_messages;
  if (bound > r->rec_cap) bound = r->rec_cap;
  int pi;
  if (!(r->flags & SDB_RECV_PRIORITY)) {
    // ---- stream order: single-pass plan + gather
    const uint32_t tiles = (n + SDB_PLAN_TILE - 1) / SDB_PLAN_TILE;
    r->plan_tops = nullptr;
    pi = sdb_prof_begin(prof, SDB_PK_RECV_SELECT, stream);
    cudaMemsetAsync(r->lb, 0, sdb_lb_words(tiles) * sizeof(unsigned long long), stream);
    static const uint32_t ticketed = (getenv("SDB_PLAN_TICKET") && atoi(getenv("SDB_PLAN_TICKET")) == 0) ? 0u : 1u;
    k_recv_plan<<<tiles, SDB_PLAN_TILE, 0, stream>>>(*v, *r, tiles, ticketed);
    sdb_prof_end(prof, pi, stream);
    if (n_launches) *n_launches += 1;
    return launch_gather(v, r, r->lb + sdb_lb_words(tiles) - 1, bound, max_rec_bytes, sm_count, stream, prof, n_launches);
  }
  const uint32_t tiles = (n + SDB_SCAN_TILE - 1) / SDB_SCAN_TILE;
  pi = sdb_prof_begin(prof, SDB_PK_RECV_COUNT, stream);
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
  const uint32_t rtiles = static_cast<uint32_t>((bound + SDB_SCAN_TILE - 1) / SDB_SCAN_TILE);
  pi = sdb_prof_begin(prof, SDB_PK_RECV_SCAN, stream);
  k_scan_plan<<<rtiles, 1024, 0, stream>>>(r->plan, r->plan_tops, r->totals);
  k_scan_tops<<<1, 1024, 0, stream>>>(r->plan_tops, rtiles, r->totals + 1);    // totals[1] = payload granules
  sdb_prof_end(prof, pi, stream);
  if (n_launches) *n_launches += 7;
  return launch_gather(v, r, nullptr, bound, max_rec_bytes, sm_count, stream, prof, n_launches);
}

// ------------------------------------------------------------------------------------------
// arena floor: smallest arena position still referenced by a pending ring entry.
// One thread per agent; distance below the current arena tail is maximised with atomicMax.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_arena_floor(sdb_dev_view v, uint32_t n_agents, uint32_t tail32, unsigned long long* max_dist) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_agents) return;
  const uint4 h = *reinterpret_cast<const uint4*>(v.ring_hdr + a);
  const uint32_t head = h.x, tail = h.y;
  const uint32_t mask = v.ring_slots - 1;
  const uint2* rs = sdb_ring_of(v, a);
  for (uint32_t p = head; p != tail; ++p) {
    const uint2 e = rs[p & mask];
    if (sdb_meta(e) != SDB_META_TOMB) {
      const uint32_t dist = tail32 - e.x;               // granules below the arena tail (mod 2^32)
      atomicMax(max_dist, static_cast<unsigned long long>(dist));
      return;                                           // rings are sorted: first live entry is the oldest
    }
  }
}

// same scan for the asynchronous import: the tail comes from the device cursor, the result goes to cursor.floor_dist
__global__ void __launch_bounds__(256)
k_arena_floor_cur(sdb_dev_view v, uint32_t n_agents, sdb_cursor* cur) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_agents) return;
  const uint4 h = *reinterpret_cast<const uint4*>(v.ring_hdr + a);
  const uint32_t head = h.x, tail = h.y;
  if (head == tail) return;
  const uint32_t tail32 = static_cast<uint32_t>(cur->arena_tail);
  const uint32_t mask = v.ring_slots - 1;
  const uint2* rs = sdb_ring_of(v, a);
  for (uint32_t p = head; p != tail; ++p) {
    const uint2 e = rs[p & mask];
    if (sdb_meta(e) != SDB_META_TOMB) { atomicMax(&cur->floor_dist, static_cast<unsigned long long>(tail32 - e.x)); return; }
  }
}
extern "C" cudaError_t sdb_launch_arena_floor_cur(const sdb_dev_view* v, uint32_t n_agents, sdb_cursor* cur, cudaStream_t stream) {
  cudaMemsetAsync(&cur->floor_dist, 0, sizeof(unsigned long long), stream);
  if (n_agents) k_arena_floor_cur<<<(n_agents + 255) / 256, 256, 0, stream>>>(*v, n_agents, cur);
  return cudaGetLastError();
}

extern "C" cudaError_t sdb_launch_arena_floor(const sdb_dev_view* v, uint32_t n_agents, uint32_t tail32,
                                              unsigned long long* max_dist_dev, cudaStream_t stream) {
  cudaMemsetAsync(max_dist_dev, 0, sizeof(unsigned long long), stream);
  if (n_agents) k_arena_floor<<<(n_agents + 255) / 256, 256, 0, stream>>>(*v, n_agents, tail32, max_dist_dev);
  return cudaGetLastError();
}
