// sdb_balance.cu - K6: batched LLM-backend selection over live load counters (sm_100a).
//
// The reference only stores a flag and a per-agent dict (set_llm_load_balancing /
// assign_llm_backend / get_llm_backend, M:1281-1325) - it has no pick algorithm, so parity
// here is against the definition in include/swarmdb_b200.h, restated by oracle/cpu_ref.c.
//
// mode 0, weighted least-load.  Sequential definition: request t goes to
//   argmin_b (load[b] / weight[b])   (exact rational compare, ties -> lowest index)
// and adds its cost to that backend.  For unit costs the greedy sequence is exactly the
// sorted merge of the per-backend progressions {(L_b + j) / w_b : j >= 0} (each progression
// is increasing, so the smallest remaining element is always a progression head).  Hence
//   pick[t] = backend of the element of global rank t,
//   rank(b, j) = j + sum_{b' != b} #{ j' >= 0 : (L_b' + j')/w_b'  <  (L_b + j)/w_b   [<= if b' < b] }
// which every element computes independently with integer divisions:
//   k_ll_plan     one thread per backend: binary-search n_b = #{j : rank(b, j) < T}
//   k_ll_scatter  one thread per element (b, j): pick[rank(b, j)] = b
// Arbitrary costs have a true sequential dependence; they run on one warp (k_ll_seq).
//
// mode 1, weighted random = one-item weighted reservoir (Efraimidis-Spirakis): backend b
// draws u_b = hash(seed, t, b) and the exponential-race key -log2(u_b)/w_b; the smallest key
// wins, so P(b) ~ w_b.  -log2 is evaluated in Q24 fixed point from a 257-entry table built
// with integer arithmetic only, and keys are compared by cross-multiplication, so results
// are bit-identical on any machine.  Loads are then bumped with shared-memory histograms +
// one global atomic per (CTA, backend).
//
// Not HBM-bound: the table is < 10 KB (L1/shared-resident); report picks/s.
#include "sdb_common.cuh"

#define SDB_MAX_BACKENDS_SMEM 4096

// ---- fixed-point -log2 table (host, integer only) --------------------------------------------
extern "C" void sdb_build_log2_table(uint32_t* tab257) {
  for (uint32_t i = 0; i <= 256; ++i) {
    if (i == 256) { tab257[i] = 1u << 24; break; }
    // x = 1 + i/256 in Q62; extract 26 fractional bits of log2(x) by repeated squaring
    unsigned __int128 x = static_cast<unsigned __int128>(256 + i) << 54;
    const unsigned __int128 two = static_cast<unsigned __int128>(2) << 62;
    uint32_t r = 0;
    for (int k = 0; k < 26; ++k) {
      x = (x * x) >> 62;
      r <<= 1;
      if (x >= two) { r |= 1u; x >>= 1; }
    }
    tab257[i] = (r + 2u) >> 2;     // round Q26 -> Q24
  }
}

__device__ __forceinline__ uint64_t sdb_mix(uint64_t seed, uint32_t t, uint32_t b) {
  uint64_t x = seed ^ (static_cast<uint64_t>(t) + 1ull) * 0x9E3779B97F4A7C15ull;
  x ^= (static_cast<uint64_t>(b) + 1ull) * 0xD1B54A32D192ED03ull;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

// -log2(u / 2^32) in Q24 for u in [1, 2^32)
__device__ __forceinline__ uint32_t sdb_neglog2_q24(uint32_t u, const uint32_t* tab) {
  const uint32_t lz = __clz(u);
  const uint32_t un = u << lz;                       // bit 31 set
  const uint32_t idx = (un >> 23) & 0xFFu;
  const uint32_t frac = un & 0x7FFFFFu;
  const uint32_t t0 = tab[idx], t1 = tab[idx + 1];
  const uint32_t lg = t0 + static_cast<uint32_t>((static_cast<uint64_t>(t1 - t0) * frac) >> 23);   // log2(mantissa), Q24
  // log2(u) = (31 - lz) + lg ; -log2(u/2^32) = 32 - log2(u) = (1 + lz) - lg
  return ((1u + lz) << 24) - lg;
}

// ---- mode 0 -----------------------------------------------------------------------------------
// number of progression elements of backend bp that precede element (b, j) in pick order
__device__ __forceinline__ unsigned long long ll_before(unsigned long long Lb_j, uint32_t wb, uint32_t b,
                                                        unsigned long long Lbp, uint32_t wbp, uint32_t bp) {
  const unsigned long long X = Lb_j * wbp;
  const unsigned long long fl = X / wb;
  if (bp < b) {                       // ties go to the lower index: count (L'+j')*w <= X
    return fl >= Lbp ? fl - Lbp + 1ull : 0ull;
  } else {                            // strict: (L'+j')*w < X  <=>  L'+j' <= ceil(X/w) - 1
    const unsigned long long ce = fl + ((X % wb) ? 1ull : 0ull);
    return ce > Lbp ? ce - Lbp : 0ull;
  }
}

__device__ unsigned long long ll_rank(uint32_t b, unsigned long long j, uint32_t B, const unsigned long long* L,
                                      const uint32_t* w) {
  unsigned long long r = j;
  const unsigned long long Lb_j = L[b] + j;
  const uint32_t wb = w[b];
  for (uint32_t bp = 0; bp < B; ++bp)
    if (bp != b) r += ll_before(Lb_j, wb, b, L[bp], w[bp], bp);
  return r;
}

// scratch layout (unsigned long long): [0,B) start offsets, [B,2B) L0 (loads before the batch), [2B] total
__global__ void __launch_bounds__(1024)
k_ll_plan(uint32_t B, const uint32_t* __restrict__ weight, unsigned long long* __restrict__ load, uint32_t T,
          unsigned long long* __restrict__ scratch) {
  extern __shared__ unsigned long long s_mem[];
  unsigned long long* sL = s_mem;                       // [B]
  unsigned long long* sN = s_mem + B;                   // [B]
  uint32_t* sW = reinterpret_cast<uint32_t*>(s_mem + 2 * B);   // [B]
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) { sL[b] = load[b]; sW[b] = weight[b]; }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) {
    // smallest j with rank(b, j) >= T  ==  number of picks of backend b
    unsigned long long lo = 0, hi = T;
    while (lo < hi) {
      const unsigned long long mid = (lo + hi) >> 1;
      if (ll_rank(b, mid, B, sL, sW) >= T) hi = mid; else lo = mid + 1;
    }
    sN[b] = lo;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long acc = 0;
    for (uint32_t b = 0; b < B; ++b) { scratch[b] = acc; acc += sN[b]; }
    scratch[2 * B] = acc;
  }
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) {
    scratch[B + b] = sL[b];
    load[b] = sL[b] + sN[b];
  }
}

__global__ void __launch_bounds__(256)
k_ll_scatter(uint32_t B, const uint32_t* __restrict__ weight, const unsigned long long* __restrict__ scratch,
             uint32_t T, uint32_t* __restrict__ out) {
  extern __shared__ unsigned long long s_mem[];
  unsigned long long* sL = s_mem;                       // [B]
  unsigned long long* sS = s_mem + B;                   // [B+1]
  uint32_t* sW = reinterpret_cast<uint32_t*>(s_mem + 2 * B + 1);
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) { sL[b] = scratch[B + b]; sS[b] = scratch[b]; sW[b] = weight[b]; }
  if (threadIdx.x == 0) sS[B] = scratch[2 * B];
  __syncthreads();
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= T || e >= sS[B]) return;
  uint32_t lo = 0, hi = B;                              // largest b with start[b] <= e
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (sS[mid] <= e) lo = mid; else hi = mid; }
  // skip empty backends that share the same start
  uint32_t b = lo;
  while (b + 1 < B && sS[b + 1] <= e) ++b;
  const unsigned long long j = e - sS[b];
  const unsigned long long t = ll_rank(b, j, B, sL, sW);
  if (t < T) out[t] = b;
}

// arbitrary costs: one warp, backends strided over lanes, exact sequential greedy
__global__ void __launch_bounds__(32)
k_ll_seq(uint32_t B, const uint32_t* __restrict__ weight, unsigned long long* __restrict__ load, uint32_t T,
         const uint32_t* __restrict__ cost, uint32_t* __restrict__ out) {
  extern __shared__ unsigned long long s_mem[];
  unsigned long long* sL = s_mem;
  uint32_t* sW = reinterpret_cast<uint32_t*>(s_mem + B);
  const uint32_t lane = threadIdx.x;
  for (uint32_t b = lane; b < B; b += 32) { sL[b] = load[b]; sW[b] = weight[b]; }
  __syncwarp();
  for (uint32_t t = 0; t < T; ++t) {
    unsigned long long bl = ~0ull; uint32_t bw = 1, bi = 0xFFFFFFFFu;
    for (uint32_t b = lane; b < B; b += 32) {
      const unsigned long long l = sL[b]; const uint32_t w = sW[b];
      // l/w < bl/bw  <=>  l*bw < bl*w   (bl == ~0 means "none yet")
      if (bi == 0xFFFFFFFFu || l * bw < bl * w) { bl = l; bw = w; bi = b; }
    }
    for (int o = 16; o; o >>= 1) {
      const unsigned long long ol = __shfl_xor_sync(0xFFFFFFFFu, bl, o);
      const uint32_t ow = __shfl_xor_sync(0xFFFFFFFFu, bw, o);
      const uint32_t oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
      if (oi != 0xFFFFFFFFu) {
        bool better;
        if (bi == 0xFFFFFFFFu) better = true;
        else {
          const unsigned long long a = ol * bw, c = bl * ow;
          better = (a < c) || (a == c && oi < bi);
        }
        if (better) { bl = ol; bw = ow; bi = oi; }
      }
    }
    if (lane == 0) out[t] = bi;
    if ((bi & 31u) == lane) sL[bi] += cost[t];
    __syncwarp();
  }
  for (uint32_t b = lane; b < B; b += 32) load[b] = sL[b];
}

// ---- mode 1 -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_wr_pick(uint32_t B, const uint32_t* __restrict__ weight, unsigned long long* __restrict__ load, uint32_t T,
          const uint32_t* __restrict__ cost, uint64_t seed, const uint32_t* __restrict__ logtab,
          uint32_t* __restrict__ out) {
  extern __shared__ unsigned long long s_mem[];
  unsigned long long* sH = s_mem;                                   // [B] cost histogram of this CTA
  uint32_t* sW = reinterpret_cast<uint32_t*>(s_mem + B);            // [B]
  uint32_t* sT = sW + B;                                            // [257]
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) { sH[b] = 0; sW[b] = weight[b]; }
  for (uint32_t i = threadIdx.x; i < 257; i += blockDim.x) sT[i] = logtab[i];
  __syncthreads();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) {
    uint32_t best = 0, bw = 1; unsigned long long bk = 0; bool have = false;
    for (uint32_t b = 0; b < B; ++b) {
      const uint32_t u = static_cast<uint32_t>(sdb_mix(seed, t, b) >> 32) | 1u;
      const unsigned long long k = sdb_neglog2_q24(u, sT);
      const uint32_t w = sW[b];
      // k/w < bk/bw  <=>  k*bw < bk*w ; strict, so ties keep the lower index
      if (!have || k * bw < bk * w) { bk = k; bw = w; best = b; have = true; }
    }
    out[t] = best;
    atomicAdd(&sH[best], static_cast<unsigned long long>(cost ? cost[t] : 1u));
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x)
    if (sH[b]) atomicAdd(&load[b], sH[b]);
}

extern "C" cudaError_t sdb_launch_pick(int mode, uint32_t B, const uint32_t* weight_dev, unsigned long long* load_dev,
                                       uint32_t n_req, const uint32_t* cost_dev, uint64_t seed, uint32_t* out_dev,
                                       unsigned long long* scratch_dev, const uint32_t* log_tab_dev,
                                       cudaStream_t stream, int* n_launches) {
  if (B > SDB_MAX_BACKENDS_SMEM) return cudaErrorInvalidValue;
  if (mode == 0) {
    if (cost_dev == nullptr) {
      const size_t sm1 = (2 * static_cast<size_t>(B)) * 8 + static_cast<size_t>(B) * 4;
      k_ll_plan<<<1, 1024, sm1, stream>>>(B, weight_dev, load_dev, n_req, scratch_dev);
      const size_t sm2 = (2 * static_cast<size_t>(B) + 1) * 8 + static_cast<size_t>(B) * 4;
      k_ll_scatter<<<(n_req + 255) / 256, 256, sm2, stream>>>(B, weight_dev, scratch_dev, n_req, out_dev);
      if (n_launches) *n_launches += 2;
    } else {
      const size_t sm = static_cast<size_t>(B) * 12;
      k_ll_seq<<<1, 32, sm, stream>>>(B, weight_dev, load_dev, n_req, cost_dev, out_dev);
      if (n_launches) *n_launches += 1;
    }
  } else {
    const size_t sm = static_cast<size_t>(B) * 12 + 257 * 4;
    k_wr_pick<<<(n_req + 255) / 256, 256, sm, stream>>>(B, weight_dev, load_dev, n_req, cost_dev, seed, log_tab_dev, out_dev);
    if (n_launches) *n_launches += 1;
  }
  return cudaGetLastError();
}
