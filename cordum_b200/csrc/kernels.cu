// kernels.cu — sm_100a kernels of the policy-gate + dispatch path.
//
//   worker_chunk_kernel + worker_merge_kernel   per heartbeat epoch: load score / overload per worker (strategy_least_loaded.go:157-159,
//                        :177-193), pool kept sorted by load, per-label bitmaps over that order, per-pool argmin
//   policy_kernel        per job batch: first-match over the rule set (safety_policy.go:187-206, 259-294), decision
//                        mapping + tenant MCP + effective-config overlay (kernel.go:187-248), scheduler post-step
//                        (engine.go:528-530)
//   route_kernel         per job batch: pool filter + least-loaded pick (strategy_least_loaded.go:40-136) for the
//                        jobs that may dispatch (engine.go:298-347)
//
// Integer / bit work only: no tensor cores (north star).  Mapping to the hardware:
//   * job columns are column-major; a warp loads 32 consecutive jobs with one coalesced 128 B request per u32
//     column (256 B per u64 column) through the read-only, no-L1-allocate path
//   * rule predicates live in bit-rows ("pass-rows"); rule bits are laid out so that a topic touches few 128-bit
//     words, and a warp walks the (job, word) items of its 32 jobs densely: one 128-bit gather per row per lane,
//     7-12 independent gathers in flight, all lanes on one instruction stream
//   * the tables are a few MB: L2-resident, hot rows L1-resident; DRAM traffic is the job columns and records
//   * first match / argmin = ballot, ffs, shuffles and shared-memory atomicMin; worker pools are kept sorted by
//     load (bitonic sort in shared memory) so label-constrained picks are bitmap ANDs
//   * decision records are written back coalesced, 16 B per lane
// IEEE float32 with explicit _rn intrinsics, no fast-math: scores compare exactly like Go's.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/cordum_b200.h"
#include "kernels.h"

#define FULL 0xFFFFFFFFu
#define KEY_NONE 0xFFFFFFFFFFFFFFFFull

namespace {

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_row(const Row16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint4 and4(uint4 a, uint4 b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }
__device__ __forceinline__ uint4 or4(uint4 a, uint4 b) { return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }
__device__ __forceinline__ uint64_t shfl64(unsigned mask, uint64_t v, int src) {
  uint32_t lo = __shfl_sync(mask, (uint32_t)v, src), hi = __shfl_sync(mask, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl64_xor(unsigned mask, uint64_t v, int lanemask) {
  uint32_t lo = __shfl_xor_sync(mask, (uint32_t)v, lanemask), hi = __shfl_xor_sync(mask, (uint32_t)(v >> 32), lanemask);
  return ((uint64_t)hi << 32) | lo;
}

// float32 -> uint32 whose unsigned order equals the float order (no NaN by construction)
__device__ __forceinline__ uint32_t orderable(float s) {
  if (s == 0.0f) s = 0.0f;   // -0 -> +0: Go's `<` treats them as equal
  uint32_t b = __float_as_uint(s);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// A worker's key: (orderable load score << 32) | rank.  Overloaded workers keep their rank but carry
// an all-ones score field, so they sort last and are never selected; KEY_NONE = no candidate / padding.
__device__ __forceinline__ bool key_over(uint64_t k) { return (uint32_t)(k >> 32) == 0xFFFFFFFFu; }

// (key, count-at-minimum-score) pairs form a commutative monoid under this merge
__device__ __forceinline__ void merge_best(uint64_t& key, uint32_t& cnt, uint64_t k2, uint32_t c2) {
  uint32_t s1 = (uint32_t)(key >> 32), s2 = (uint32_t)(k2 >> 32);
  if (key_over(k2)) return;
  if (key_over(key) || s2 < s1) { key = k2; cnt = c2; }
  else if (s2 == s1) { key = k2 < key ? k2 : key; cnt += c2; }
}

}  // namespace

// ------------------------------------------------------------------ K2: worker-table refresh, one CTA per pool chunk
// Per heartbeat epoch: load score + overload test per worker, then the pool's keys in ascending order (the
// load-sorted view lets label-constrained jobs find the least-loaded worker carrying their labels from per-label
// bitmaps).  A pool is cut into chunks of CORDUM_POOL_CHUNK workers so that the whole GPU takes part in the refresh
// instead of one SM per pool: worker_chunk_kernel sorts one chunk per CTA in shared memory (and finishes pools that
// have a single chunk); worker_merge_kernel places every worker of a multi-chunk pool at
//     its index in its own chunk + the number of smaller keys in each sibling chunk        (keys are unique: rank)
// Pools above CORDUM_POOL_SORT_MAX workers keep only min / count-at-min; their jobs scan (route_kernel S').
__device__ __forceinline__ uint64_t worker_key(const DeviceTables& T, uint32_t pos) {
  const Load16 L = T.loads[T.pos_slot[pos]];
  bool over = false;
  if (L.max_parallel > 0) over = __fdiv_rn(__int2float_rn(L.active), __int2float_rn(L.max_parallel)) >= 0.9f;   // :177-184
  over = over || L.cpu >= 90.0f || L.gpu >= 90.0f;                                                               // :185-191
  const float score = __fadd_rn(__fadd_rn(__int2float_rn(L.active), __fdiv_rn(L.cpu, 100.0f)), __fdiv_rn(L.gpu, 100.0f));   // :157-159
  return ((uint64_t)(over ? 0xFFFFFFFFu : orderable(score)) << 32) | T.pos_rank[pos];
}

template <uint32_t NT>
__global__ void __launch_bounds__(NT) worker_chunk_kernel(DeviceTables T) {
  constexpr uint32_t CH = CORDUM_POOL_CHUNK;
  __shared__ uint64_t sk[CH];
  __shared__ uint32_t s_cnt, s_nok;
  const uint32_t g = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const uint32_t p = T.chunk_pool[g], c = g - T.pool_chunk0[p], m = T.pool_chunk0[p + 1] - T.pool_chunk0[p];
  const uint32_t a = T.pool_off[p], n = T.pool_off[p + 1] - a;
  const uint32_t cs = c * CH, len = n - cs < CH ? n - cs : CH;   // an empty pool has one empty chunk
  const bool sortable = n <= CORDUM_POOL_SORT_MAX;
  uint32_t n_pad = 1;
  while (n_pad < len) n_pad <<= 1;
  if (tid == 0) { s_cnt = 0; s_nok = 0; }
  for (uint32_t i = tid; i < n_pad; i += NT) {
    uint64_t key = KEY_NONE;
    if (i < len) { key = worker_key(T, a + cs + i); T.pos_key[a + cs + i] = key; }
    sk[i] = key;
  }
  if (!sortable) return;   // chunk 0 of the pool reduces min / count in worker_merge_kernel
  __syncthreads();
  // bitonic sort, one compare-exchange per (thread, pair): pair q touches i = q with a zero inserted at bit log2(jj)
  for (uint32_t k = 2; k <= n_pad; k <<= 1)
    for (uint32_t jj = k >> 1; jj > 0; jj >>= 1) {
      for (uint32_t q = tid; q < (n_pad >> 1); q += NT) {
        const uint32_t i = ((q & ~(jj - 1)) << 1) | (q & (jj - 1)), x = i | jj;
        const uint64_t ki = sk[i], kx = sk[x];
        if ((ki > kx) == ((i & k) == 0)) { sk[i] = kx; sk[x] = ki; }
      }
      __syncthreads();
    }
  const uint32_t words = (n + 31) >> 5, nbits = T.place_bits;
  uint32_t* bm = T.lbm + T.lbm_off[p];
  if (m > 1) {
    // hand the sorted chunk to worker_merge_kernel; clear what it accumulates into
    for (uint32_t i = tid; i < len; i += NT) T.ckey[a + cs + i] = sk[i];
    const uint64_t total = (uint64_t)nbits * words;
    for (uint64_t i = total * c / m + tid; i < total * (c + 1) / m; i += NT) bm[i] = 0;
    if (c == 0 && tid == 0) { T.pool_mincnt[p] = 0; T.pool_nok[p] = 0; }
    return;
  }
  // single chunk: sorted view + label bitmaps; a warp takes 32 consecutive sorted workers (one bitmap word per label bit)
  const uint64_t k0 = n ? sk[0] : KEY_NONE;
  const bool none = n == 0 || key_over(k0);
  uint32_t cnt = 0, ok = 0;
  for (uint32_t w = tid >> 5; w < words; w += NT >> 5) {
    const uint32_t i = w * 32 + lane;
    uint64_t k = KEY_NONE, llo = 0, lhi = 0;
    if (i < n) {
      k = sk[i];
      const uint32_t src = T.rank_pos[(uint32_t)(k & 0xFFFFFFFFu)];   // every key (overloaded ones too) carries its rank
      llo = T.pos_label_lo[src]; lhi = T.pos_label_hi[src];
      T.skey[a + i] = k; T.slab_lo[a + i] = llo; T.slab_hi[a + i] = lhi;
      cnt += (!none && (uint32_t)(k >> 32) == (uint32_t)(k0 >> 32)) ? 1u : 0u;
      ok += key_over(k) ? 0u : 1u;
    }
    for (uint32_t b0 = 0; b0 < nbits; b0 += 32) {   // lane t keeps the word of label bit b0+t, then one strided store each
      uint32_t mine = 0;
      const uint32_t lim = nbits - b0 < 32 ? nbits - b0 : 32;
      for (uint32_t t = 0; t < lim; ++t) {
        const uint32_t bit = b0 + t;
        const unsigned bal = __ballot_sync(FULL, ((bit < 64 ? llo >> bit : lhi >> (bit - 64)) & 1ull) != 0);
        if (lane == t) mine = bal;
      }
      if (lane < lim) bm[(size_t)(b0 + lane) * words + w] = mine;
    }
  }
  cnt = __reduce_add_sync(FULL, cnt); ok = __reduce_add_sync(FULL, ok);
  if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
  if (lane == 0 && ok) atomicAdd(&s_nok, ok);
  __syncthreads();
  if (tid == 0) { T.pool_best[p] = none ? KEY_NONE : k0; T.pool_mincnt[p] = s_cnt; T.pool_sorted[p] = 1; T.pool_nok[p] = s_nok; }
}

__global__ void __launch_bounds__(256) worker_merge_kernel(DeviceTables T) {
  constexpr uint32_t CH = CORDUM_POOL_CHUNK, NT = 256;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_cnt, s_nok;
  __shared__ uint64_t s_best;
  uint64_t* sp = reinterpret_cast<uint64_t*>(smem_raw);   // the pool's keys, chunk by chunk, each chunk ascending
  const uint32_t g = T.merge_list[blockIdx.x], tid = threadIdx.x, lane = tid & 31;
  const uint32_t p = T.chunk_pool[g], c = g - T.pool_chunk0[p], m = T.pool_chunk0[p + 1] - T.pool_chunk0[p];
  const uint32_t a = T.pool_off[p], n = T.pool_off[p + 1] - a;
  if (tid == 0) { s_cnt = 0; s_nok = 0; s_best = KEY_NONE; }
  __syncthreads();
  if (n > CORDUM_POOL_SORT_MAX) {   // unsorted pool: min / count-at-min by reduction (listed once, as chunk 0)
    uint64_t best = KEY_NONE;
    for (uint32_t i = tid; i < n; i += NT) { const uint64_t k = T.pos_key[a + i]; if (!key_over(k) && k < best) best = k; }
    for (int o = 16; o; o >>= 1) { const uint64_t v = shfl64_xor(FULL, best, o); best = v < best ? v : best; }
    if (lane == 0) atomicMin(reinterpret_cast<unsigned long long*>(&s_best), (unsigned long long)best);
    __syncthreads();
    best = s_best;
    uint32_t cnt = 0;
    for (uint32_t i = tid; i < n; i += NT) { const uint64_t k = T.pos_key[a + i]; cnt += (!key_over(k) && (uint32_t)(k >> 32) == (uint32_t)(best >> 32)) ? 1u : 0u; }
    cnt = __reduce_add_sync(FULL, cnt);
    if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (tid == 0) { T.pool_best[p] = best; T.pool_mincnt[p] = best == KEY_NONE ? 0 : s_cnt; T.pool_sorted[p] = 0; }
    return;
  }
  for (uint32_t i = tid; i < n; i += NT) sp[i] = T.ckey[a + i];
  __syncthreads();
  uint64_t k0 = KEY_NONE;
  for (uint32_t cc = 0; cc < m; ++cc) { const uint64_t h = sp[cc * CH]; k0 = h < k0 ? h : k0; }
  const bool none = key_over(k0);
  const uint32_t cs = c * CH, len = n - cs < CH ? n - cs : CH;
  const uint32_t words = (n + 31) >> 5;
  uint32_t* bm = T.lbm + T.lbm_off[p];
  uint32_t cnt = 0, ok = 0;
  for (uint32_t i = tid; i < len; i += NT) {
    const uint64_t k = sp[cs + i];
    uint32_t idx = i;
    for (uint32_t cc = 0; cc < m; ++cc) {
      if (cc == c) continue;
      const uint64_t* q = sp + cc * CH;
      uint32_t lo = 0, hi = n - cc * CH < CH ? n - cc * CH : CH;   // lower bound of k in the sibling chunk
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (q[mid] < k) lo = mid + 1; else hi = mid; }
      idx += lo;
    }
    const uint32_t src = T.rank_pos[(uint32_t)(k & 0xFFFFFFFFu)];
    uint64_t llo = T.pos_label_lo[src], lhi = T.pos_label_hi[src];
    T.skey[a + idx] = k; T.slab_lo[a + idx] = llo; T.slab_hi[a + idx] = lhi;
    const uint32_t w = idx >> 5, bitv = 1u << (idx & 31);
    while (llo) { const uint32_t b = __ffsll((long long)llo) - 1; llo &= llo - 1; atomicOr(&bm[(size_t)b * words + w], bitv); }
    while (lhi) { const uint32_t b = 64 + __ffsll((long long)lhi) - 1; lhi &= lhi - 1; atomicOr(&bm[(size_t)b * words + w], bitv); }
    cnt += (!none && (uint32_t)(k >> 32) == (uint32_t)(k0 >> 32)) ? 1u : 0u;
    ok += key_over(k) ? 0u : 1u;
  }
  cnt = __reduce_add_sync(FULL, cnt); ok = __reduce_add_sync(FULL, ok);
  if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
  if (lane == 0 && ok) atomicAdd(&s_nok, ok);
  __syncthreads();
  if (tid == 0) {
    if (s_cnt) atomicAdd(&T.pool_mincnt[p], s_cnt);
    if (s_nok) atomicAdd(&T.pool_nok[p], s_nok);
    if (c == 0) { T.pool_best[p] = none ? KEY_NONE : k0; T.pool_sorted[p] = 1; }
  }
}

// ------------------------------------------------------------------ policy: first match + decision
// A warp owns a tile of 32 consecutive jobs (lane = job for the coalesced column loads, the scalar
// decision logic and the 16 B record store).  Needs no worker state, so it runs concurrently with the
// heartbeat exchange and the worker-table refresh.
//   P  the per-topic word lists of the tile's 32 jobs are walked 32 (job, word) items at a time: each lane ANDs one
//      128-bit word of the 7-12 pass-rows its job selects (rule bits are permuted so that a topic touches few
//      words); surviving bits map back to original rule indices; first match = min per job.
//   D  decision mapping, tenant MCP, effective-config overlay, approval flags, scheduler post-step:
//      thread per job.
template <int MINB, int IU>
__global__ void __launch_bounds__(256, MINB) policy_kernel(KParams P) {
  const DeviceTables& T = P.t;
  const JobColumns& C = P.cols;
  const unsigned lane = threadIdx.x & 31;
  const uint32_t n_tiles = (P.n_jobs + 31u) >> 5;
  const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t rowu4 = T.row_u4;

  for (uint32_t tile = warp_id; tile < n_tiles; tile += warps_total) {
    const uint32_t j = tile * 32 + lane;
    const bool valid = j < P.n_jobs;
    // only what the rule scan needs stays in registers; the rest is fetched where it is used
    uint32_t c_tenant = 0, c_topic = 0, c_cap = 0, c_pack = 0, c_actor = 0, c_flags = 0;
    uint64_t c_risk = 0;
    if (valid) {
      c_flags = ld_stream_u32(C.flags + j); c_topic = ld_stream_u32(C.topic + j);
      c_tenant = ld_stream_u32(C.tenant + j);
      c_cap = ld_stream_u32(C.capability + j); c_pack = ld_stream_u32(C.pack + j); c_actor = ld_stream_u32(C.actor + j);
      c_risk = ld_stream_u64(C.risk_mask + j);
    }

    // =================================================================== P: first matching rule
    // Rule bits are permuted (and multi-pattern rules duplicated) so that a topic's pass-row is non-zero in only a
    // few 128-bit words: its word list.  The word lists of the tile's 32 jobs are laid end to end and the 32 lanes
    // walk that sequence 32 items at a time: lane = one (job, word) item.  It ANDs that word of the 7-12 rows the
    // job selects and turns surviving bits back into ORIGINAL rule indices (pos2rule); the first match is the
    // minimum per job (shared-memory atomicMin; survivors are rare).  Every lane is busy in every step and all
    // lanes run one instruction stream.
    const bool bypass = P.honor_approved && (c_flags & JF_APPROVED);                                    // engine.go:484-522
    const bool early = (c_flags & (JF_TOPIC_MISSING | JF_TOPIC_UNSUPPORTED)) != 0;                        // kernel.go:171-176
    const bool eval = valid && !bypass && !early;
    __shared__ uint32_t s_best[8][32];
    uint32_t* my_best = s_best[threadIdx.x >> 5];
    int first = -1;
    {
      const uint32_t o_combo = CORDUM_COMBO_INDEX(c_flags) * rowu4, o_tenant = c_tenant * rowu4, o_topic = c_topic * rowu4,
                     o_cap = c_cap * rowu4, o_pack = c_pack * rowu4, o_actor = c_actor * rowu4;
      uint32_t c_twoff = 0, c_twcnt = 0;
      if (eval) { c_twoff = __ldg(T.tw_off + c_topic); c_twcnt = __ldg(T.tw_cnt + c_topic); }
      uint32_t incl = c_twcnt;   // inclusive prefix sum of the word counts over the tile
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += t; }
      const uint32_t excl = incl - c_twcnt, total = __shfl_sync(FULL, incl, 31);
      my_best[lane] = 0xFFFFFFFFu;
      __syncwarp();
      const uint32_t* chk_words = reinterpret_cast<const uint32_t*>(T.row_check);
      for (uint32_t q0 = 0; q0 < total; q0 += 32) {
        const uint32_t q = q0 + lane;
        const bool have = q < total;
        // owner of item q = the last lane whose exclusive prefix is <= q (binary search over the warp's registers)
        int jl = 0;
#pragma unroll
        for (int st = 16; st; st >>= 1) { const uint32_t e = __shfl_sync(FULL, excl, jl + st); if (e <= q) jl += st; }
        const uint32_t li = q - __shfl_sync(FULL, excl, jl);
        const uint32_t fl = __shfl_sync(FULL, c_flags, jl);
        const uint32_t twoff = __shfl_sync(FULL, c_twoff, jl);   // every lane takes part in every shuffle
        // wi: the item's first 128-bit word within a row (item 0: valid address for idle lanes)
        const uint32_t wi = have ? (uint32_t)__ldg(T.tw_list + twoff + li) * (uint32_t)IU : 0u;
        const Row16* p_combo = T.row_combo + __shfl_sync(FULL, o_combo, jl) + wi;
        const Row16* p_tenant = T.row_tenant + __shfl_sync(FULL, o_tenant, jl) + wi;
        const Row16* p_topic = T.row_topic + __shfl_sync(FULL, o_topic, jl) + wi;
        const Row16* p_cap = T.row_cap + __shfl_sync(FULL, o_cap, jl) + wi;
        const Row16* p_pack = T.row_pack + __shfl_sync(FULL, o_pack, jl) + wi;
        const Row16* p_actor = T.row_actor + __shfl_sync(FULL, o_actor, jl) + wi;
        uint4 acc[IU];
#pragma unroll
        for (int u = 0; u < IU; ++u)
          acc[u] = and4(and4(and4(ld_row(p_combo + u), ld_row(p_tenant + u)), and4(ld_row(p_topic + u), ld_row(p_cap + u))),
                        and4(ld_row(p_pack + u), ld_row(p_actor + u)));
        {   // risk tags: containsAny = OR over the job's tags (:308-318).  Branch-free: row 0 = "no referenced tag",
            // row 1+b = tag b, and lanes that ran out of tags read the all-zero row.
          uint64_t m = shfl64(FULL, c_risk, jl);
          uint32_t idx = (uint32_t)__ffsll((long long)m);   // 0 when the job has no referenced tag, else 1 + lowest bit
          m &= m - 1;
          uint4 rk[IU];
#pragma unroll
          for (int u = 0; u < IU; ++u) rk[u] = ld_row(T.row_risk + (idx * rowu4 + wi + u));
          while (__any_sync(FULL, m != 0)) {
            idx = m ? (uint32_t)__ffsll((long long)m) : T.risk_zero_row;
            m &= m - 1;
#pragma unroll
            for (int u = 0; u < IU; ++u) rk[u] = or4(rk[u], ld_row(T.row_risk + (idx * rowu4 + wi + u)));
          }
#pragma unroll
          for (int u = 0; u < IU; ++u) acc[u] = and4(acc[u], rk[u]);
        }
        const uint32_t jsrc = tile * 32 + (uint32_t)jl;   // the job this lane works for (MCP ids / masks are read on demand)
        const bool mcp_used = have && (fl & JF_MCP_USED);
        if (__any_sync(FULL, mcp_used)) {   // mcpMatch (:365-382); lanes whose job carries no MCP labels read the all-ones row
#pragma unroll
          for (int qf = 0; qf < 4; ++qf) {
            const uint32_t id = mcp_used ? __ldg(C.mcp[qf] + jsrc) : T.mcp_ones_row[qf];
#pragma unroll
            for (int u = 0; u < IU; ++u) acc[u] = and4(acc[u], ld_row(T.row_mcp[qf] + (id * rowu4 + wi + u)));
          }
        }
        // Surviving bits -> original rule index.  Inside an item the positions ascend with the rule index, so the
        // lowest surviving bit is the item's first match; a further bit is looked at only when that rule carries a
        // requires / labels subset test (containsAll :320-330, labelsMatch :332-345) and the test fails.
        uint32_t best = 0xFFFFFFFFu;
        uint32_t w[4 * IU];
#pragma unroll
        for (int u = 0; u < IU; ++u) {
          w[4 * u] = have ? acc[u].x : 0u; w[4 * u + 1] = have ? acc[u].y : 0u;
          w[4 * u + 2] = have ? acc[u].z : 0u; w[4 * u + 3] = have ? acc[u].w : 0u;
        }
        bool alive;   // this lane still has a surviving bit whose rule has not passed its subset test
        do {
          uint32_t sel = 0, base = 0;   // lowest non-zero 32-bit word of the item
#pragma unroll
          for (int k = 4 * IU - 1; k >= 0; --k) if (w[k]) { sel = w[k]; base = 32u * (uint32_t)k; }
          const bool nz = sel != 0;
          alive = false;
          if (__any_sync(FULL, nz)) {
            const uint32_t pos = nz ? wi * 128u + base + (uint32_t)__ffs((int)sel) - 1u : 0u;
            const uint32_t r = __ldg(T.pos2rule + pos);
            bool ok = nz;
            if (nz && ((__ldg(chk_words + (pos >> 5)) >> (pos & 31)) & 1u)) {
              const uint64_t req = __ldg(C.req_mask + jsrc), lab = __ldg(C.lab_mask + jsrc);
              const uint64_t need = __ldg(T.rule_req_need + r), ln = __ldg(T.rule_lab_need + r);
              ok = ((need & ~req) == 0) && (ln == 0 || ((fl & JF_HAS_LABELS) && (ln & ~lab) == 0));
            }
            if (ok) best = r;
            alive = nz && !ok;
            if (alive) {   // drop the bit just handled and look at the next one (rare)
              const uint32_t drop = sel & (sel - 1);
#pragma unroll
              for (int k = 0; k < 4 * IU; ++k) if (base == 32u * (uint32_t)k) w[k] = drop;
            }
          }
        } while (__any_sync(FULL, alive));
        if (best != 0xFFFFFFFFu) atomicMin(&my_best[jl], best);
      }
      __syncwarp();
      first = (int)my_best[lane];   // 0xFFFFFFFF -> -1: no rule matched
      __syncwarp();
    }

    // =================================================================== D: decision (thread per job)
    uint32_t dec = CORDUM_DEC_UNSPECIFIED, sched = CORDUM_DEC_UNSPECIFIED, rflags = 0, reason = 0;
    int rule = -1;
    if (valid) {
      if (bypass) { dec = sched = CORDUM_DEC_ALLOW; reason = CORDUM_REASON_APPROVAL_GRANTED; rflags = CORDUM_F_APPROVED_BYPASS; }
      else if (c_flags & JF_TOPIC_MISSING) { dec = sched = CORDUM_DEC_DENY; reason = CORDUM_REASON_MISSING_TOPIC; }
      else if (c_flags & JF_TOPIC_UNSUPPORTED) { dec = sched = CORDUM_DEC_DENY; reason = CORDUM_REASON_UNSUPPORTED_TOPIC; }
      else {
        const bool mcp_used = c_flags & JF_MCP_USED;
        const uint32_t c_tpol = ld_stream_u32(C.tenant_pol + j), c_eff = ld_stream_u32(C.effcfg + j);
        uint32_t mid[4] = {0, 0, 0, 0};
        if (mcp_used) { mid[0] = __ldg(C.mcp[0] + j); mid[1] = __ldg(C.mcp[1] + j); mid[2] = __ldg(C.mcp[2] + j); mid[3] = __ldg(C.mcp[3] + j); }
        rule = first;
        uint32_t code = CORDUM_DEC_ALLOW;
        bool hascons = false;
        if (first >= 0) { uint8_t rd = __ldg(T.rule_dec + first); code = rd & 0x7Fu; hascons = rd & 0x80u; }
        const bool rule_approval = code == CORDUM_DEC_REQUIRE_HUMAN;   // safety_policy.go:200
        uint32_t tm = 0;                                               // tenant MCP lists, kernel.go:190-195
        if (mcp_used && c_tpol) {
          const uint8_t* base = T.tenant_mcp + (size_t)(c_tpol - 1) * 4 * T.mcp_stride;
          for (int q = 0; q < 4 && !tm; ++q) { uint8_t v = __ldg(base + q * T.mcp_stride + mid[q]); if (v) tm = 1 + q * 2 + (v - 1); }
          if (tm) code = CORDUM_DEC_DENY;
        }
        dec = CORDUM_DEC_ALLOW;                                        // kernel.go:198-215
        if (code == CORDUM_DEC_DENY) { dec = CORDUM_DEC_DENY; reason = tm ? CORDUM_REASON_TENANT_MCP + (tm - 1) : CORDUM_REASON_RULE; }
        else if (code == CORDUM_DEC_REQUIRE_HUMAN) { dec = CORDUM_DEC_REQUIRE_HUMAN; reason = CORDUM_REASON_RULE; }
        else if (code == CORDUM_DEC_THROTTLE) { dec = CORDUM_DEC_THROTTLE; reason = CORDUM_REASON_RULE; }
        else if (code == CORDUM_DEC_ALLOW_WITH_CONSTRAINTS || hascons) dec = CORDUM_DEC_ALLOW_WITH_CONSTRAINTS;
        if (c_eff) {                                                   // kernel.go:218-231
          const uint8_t tb = __ldg(T.eff_topic + (size_t)c_eff * T.topic_stride + c_topic);
          if (tb & 1) { dec = CORDUM_DEC_DENY; reason = CORDUM_REASON_EFF_DENIED_TOPIC; }
          if (tb & 2) { dec = CORDUM_DEC_DENY; reason = CORDUM_REASON_EFF_NOT_ALLOWED_TOPIC; }
          if (mcp_used) {
            const uint8_t* base = T.eff_mcp + (size_t)c_eff * 4 * T.mcp_stride;
            uint32_t em = 0;
            for (int q = 0; q < 4 && !em; ++q) { uint8_t v = __ldg(base + q * T.mcp_stride + mid[q]); if (v) em = 1 + q * 2 + (v - 1); }
            if (em) { dec = CORDUM_DEC_DENY; reason = CORDUM_REASON_EFF_MCP + (em - 1); }
          }
        }
        const bool approval = rule_approval || dec == CORDUM_DEC_REQUIRE_HUMAN;   // kernel.go:233
        rflags = CORDUM_F_HAS_SNAPSHOT | (approval ? CORDUM_F_APPROVAL_REQUIRED : 0) | (hascons ? CORDUM_F_CONSTRAINTS : 0);
        sched = dec;                                                   // engine.go:528-530
        if (approval && (dec == CORDUM_DEC_ALLOW || dec == CORDUM_DEC_ALLOW_WITH_CONSTRAINTS)) sched = CORDUM_DEC_REQUIRE_HUMAN;
      }
    }

    if (valid) {   // coalesced 16 B store per lane; the route fields are filled by route_kernel
      const uint32_t head = dec | (sched << 8) | (rflags << 16);
      reinterpret_cast<uint4*>(P.out)[j] = make_uint4(head, reason & 0xFFu, (uint32_t)rule, 0xFFFFFFFFu);
    }
    if (P.route_list) {   // compact the jobs that may dispatch (engine.go:298-347): one atomic per tile
      const bool dr = valid && (sched == CORDUM_DEC_ALLOW || sched == CORDUM_DEC_ALLOW_WITH_CONSTRAINTS);
      const unsigned m = __ballot_sync(FULL, dr);
      if (m) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(P.route_count, (uint32_t)__popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (dr) P.route_list[base + __popc(m & ((1u << lane) - 1u))] = j;
      }
    }
  }
}


// ------------------------------------------------------------------ route: pool filter + least-loaded pick
// Works on the compacted list of jobs that may dispatch (written by policy_kernel; in ROUTE_ONLY mode: all jobs).
//   D  topic->pools, preferred_pool, requires filter, preferred_worker_id and, for label-free jobs, the
//      merge of per-pool best keys: thread per job.
//   S  jobs with placement labels: 4 jobs per warp step, 8 lanes each.  For every eligible pool the group ANDs the
//      pool's label bitmaps over its load-sorted view (lane = 32 workers); first set bit in the non-overloaded
//      prefix = argmin, the next match decides the tie flag, popcount = candidate total.
//   S' jobs that touch a pool larger than K2's sort buffer: warp per job, coalesced scan (rare).
template <bool ROUTE_ONLY, int MINB = 4>
__global__ void __launch_bounds__(256, MINB) route_kernel(KParams P) {
  const DeviceTables& T = P.t;
  const JobColumns& C = P.cols;
  const unsigned lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
  const uint32_t n_items = ROUTE_ONLY ? P.n_jobs : *P.route_count;
  const uint32_t n_tiles = (n_items + 31u) >> 5;
  const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;

  for (uint32_t tile = warp_id; tile < n_tiles; tile += warps_total) {
    const uint32_t item = tile * 32 + lane;
    const bool valid = item < n_items;
    const uint32_t j = valid ? (ROUTE_ONLY ? item : P.route_list[item]) : 0u;
    uint4 rec = make_uint4(0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu);
    uint32_t c_flags = 0, c_topic = 0, c_ppool = 0, c_pwork = 0;
    uint64_t c_req = 0, c_plo = 0, c_phi = 0;
    if (valid) {
      if (!ROUTE_ONLY) rec = reinterpret_cast<const uint4*>(P.out)[j];
      c_flags = __ldg(C.flags + j); c_topic = __ldg(C.topic + j);
      c_ppool = __ldg(C.pref_pool + j); c_pwork = __ldg(C.pref_worker + j);
      c_req = __ldg(C.req_mask + j); c_plo = __ldg(C.place_lo + j); c_phi = __ldg(C.place_hi + j);
    }
    uint32_t rflags = 0, route = CORDUM_ROUTE_NOT_ATTEMPTED;
    int slot = -1;

    // =================================================================== D: pool filter, thread per job
    bool need_scan = false, slow = false;
    uint32_t r_off = 0, r_cnt = 0;
    int r_single = -1;
    uint64_t best = KEY_NONE;
    uint32_t bcnt = 0, total = 0;
    if (valid) {
      if (c_flags & JF_TOPIC_RAW_EMPTY) route = CORDUM_ROUTE_MISSING_TOPIC;   // :41-43
      else {
        r_off = __ldg(T.topic_pool_off + c_topic);
        r_cnt = __ldg(T.topic_pool_cnt + c_topic);
        if (c_ppool) {                                                         // preferred_pool (:50-55)
          bool found = false;
          if (c_ppool != CORDUM_PREF_UNKNOWN)
            for (uint32_t k = 0; k < r_cnt; ++k) found |= __ldg(T.pool_list + r_off + k) == c_ppool - 1;
          if (!found) route = CORDUM_ROUTE_NO_POOL_PREFERRED;
          else { r_single = (int)(c_ppool - 1); r_cnt = 1; }
        }
        if (route == CORDUM_ROUTE_NOT_ATTEMPTED && r_cnt == 0) route = CORDUM_ROUTE_NO_POOL_TOPIC;   // :56-58
        if (route == CORDUM_ROUTE_NOT_ATTEMPTED) {
          const bool req_any = c_flags & JF_REQ_NONEMPTY, req_unknown = c_flags & JF_REQ_UNKNOWN;
          const uint64_t need_req = c_req & ~T.req_blank_mask;
          const bool unsat = c_flags & JF_PLACE_UNSAT;
          const bool labelled = (c_plo | c_phi) != 0 || unsat;
          int pw_pos = -1;
          uint32_t pw_pool = 0xFFFFFFFFu;
          if (c_pwork && c_pwork != CORDUM_PREF_UNKNOWN) {
            uint32_t p1 = __ldg(T.slot_pos + (c_pwork - 1));
            if (p1) { pw_pos = (int)p1 - 1; pw_pool = __ldg(T.pos_pool + pw_pos); }
          }
          uint32_t n_elig = 0;
          bool pw_in_set = false;
          for (uint32_t k = 0; k < r_cnt; ++k) {
            const uint32_t pid = r_single >= 0 ? (uint32_t)r_single : __ldg(T.pool_list + r_off + k);
            // poolSatisfies (:241-265)
            if (req_any && !(__ldg(T.pool_req_nonempty + pid) && !req_unknown && (need_req & ~__ldg(T.pool_req_mask + pid)) == 0)) continue;
            n_elig++;
            pw_in_set |= pid == pw_pool;
            if (!labelled) {
              total += __ldg(T.pool_off + pid + 1) - __ldg(T.pool_off + pid);
              merge_best(best, bcnt, T.pool_best[pid], T.pool_mincnt[pid]);
            } else slow |= T.pool_sorted[pid] == 0;
          }
          if (n_elig == 0) route = CORDUM_ROUTE_NO_POOL_REQUIRES;              // :64-66
          else {
            bool took_pref = false;
            if (pw_pos >= 0 && pw_in_set && !unsat) {                           // :73-87
              uint64_t llo = __ldg(T.pos_label_lo + pw_pos), lhi = __ldg(T.pos_label_hi + pw_pos);
              if ((llo & c_plo) == c_plo && (lhi & c_phi) == c_phi && !key_over(T.pos_key[pw_pos])) {
                took_pref = true; route = CORDUM_ROUTE_OK_PREFERRED; slot = (int)(c_pwork - 1);
              }
            }
            if (!took_pref) need_scan = labelled && !unsat;   // label-free and unsatisfiable jobs are final already
            if (!took_pref && !need_scan) {
              if (best != KEY_NONE) {
                route = CORDUM_ROUTE_OK;
                slot = (int)__ldg(T.rank_slot + (uint32_t)(best & 0xFFFFFFFFu));
                if (bcnt > 1) rflags |= CORDUM_F_TIE;
              } else route = total > 0 ? CORDUM_ROUTE_POOL_OVERLOADED : CORDUM_ROUTE_NO_WORKERS;   // :114-119
            }
          }
        }
      }
    }

    // =================================================================== S: label picks, 4 jobs x 8 lanes per step
    for (unsigned todo = __ballot_sync(FULL, need_scan && !slow); todo;) {
      int own[4];   // the (up to) four jobs of this step: group q serves lane own[q]
#pragma unroll
      for (int q = 0; q < 4; ++q) { own[q] = todo ? __ffs(todo) - 1 : -1; todo &= todo - 1; }
      const int src = g == 0 ? own[0] : g == 1 ? own[1] : g == 2 ? own[2] : own[3];
      const bool act = src >= 0;
      const int s_ = act ? src : 0;
      const uint32_t off = __shfl_sync(FULL, r_off, s_), fl = __shfl_sync(FULL, c_flags, s_);
      const uint32_t cnt_all = __shfl_sync(FULL, r_cnt, s_);   // every lane takes part in every shuffle
      const uint32_t cnt = act ? cnt_all : 0u;
      const int single = __shfl_sync(FULL, r_single, s_);
      const uint64_t need_req = shfl64(FULL, c_req, s_) & ~T.req_blank_mask;
      const uint64_t need_lo = shfl64(FULL, c_plo, s_), need_hi = shfl64(FULL, c_phi, s_);
      const bool req_any = fl & JF_REQ_NONEMPTY, req_unknown = fl & JF_REQ_UNKNOWN;
      // required label bits -> bitmap row numbers, once per job (not per pool and word): four 8-bit row numbers in lbp
      uint32_t lbp = 0, nlb = 0;
      bool more;
      {
        uint64_t rl = need_lo, rh = need_hi;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint32_t bit = 0;
          bool got = false;
          if (rl) { bit = (uint32_t)__ffsll((long long)rl) - 1u; rl &= rl - 1; got = true; }
          else if (rh) { bit = 64u + (uint32_t)__ffsll((long long)rh) - 1u; rh &= rh - 1; got = true; }
          if (got) { nlb = (uint32_t)t + 1u; lbp |= bit << (8 * t); }
        }
        more = (rl | rh) != 0;
      }
      uint64_t b = KEY_NONE;
      uint32_t bc = 0, tot = 0;
      for (uint32_t k = 0; __any_sync(FULL, k < cnt); ++k) {
        const bool kin = k < cnt;
        const uint32_t pid = kin ? (single >= 0 ? (uint32_t)single : __ldg(T.pool_list + off + k)) : 0u;
        const bool elig = kin && (!req_any || (__ldg(T.pool_req_nonempty + pid) && !req_unknown && (need_req & ~__ldg(T.pool_req_mask + pid)) == 0));
        const uint32_t a = __ldg(T.pool_off + pid), e = __ldg(T.pool_off + pid + 1);
        const uint32_t words = elig ? (e - a + 31) >> 5 : 0u, nok = T.pool_nok[pid];
        const uint32_t* bm = T.lbm + __ldg(T.lbm_off + pid);
        uint64_t kstar = KEY_NONE;
        uint32_t kc = 0;
        bool tie_done = false;
        for (uint32_t w0 = 0; __any_sync(FULL, w0 < words && !tie_done); w0 += CORDUM_GROUP) {
          const uint32_t w = w0 + sub;
          uint32_t v = 0;
          if (w < words && !tie_done) {   // matchesLabels (:161-175) for 32 load-sorted workers at once
            // the first four required label bits are independent predicated loads; more than four is rare
            const uint32_t x0 = nlb > 0 ? bm[(lbp & 0xFFu) * words + w] : 0xFFFFFFFFu;
            const uint32_t x1 = nlb > 1 ? bm[((lbp >> 8) & 0xFFu) * words + w] : 0xFFFFFFFFu;
            const uint32_t x2 = nlb > 2 ? bm[((lbp >> 16) & 0xFFu) * words + w] : 0xFFFFFFFFu;
            const uint32_t x3 = nlb > 3 ? bm[(lbp >> 24) * words + w] : 0xFFFFFFFFu;
            v = (x0 & x1) & (x2 & x3);
            if (more) {   // AND is idempotent: walk all required bits again
              for (uint64_t m = need_lo; m; m &= m - 1) v &= bm[(uint32_t)(__ffsll((long long)m) - 1) * words + w];
              for (uint64_t m = need_hi; m; m &= m - 1) v &= bm[(uint32_t)(64 + __ffsll((long long)m) - 1) * words + w];
            }
          }
          tot += __popc(v);                                     // label-matching candidates, overloaded ones included
          const uint32_t lo = w * 32;
          uint32_t okv = lo >= nok ? 0u : (nok - lo >= 32 ? v : v & ((1u << (nok - lo)) - 1u));   // non-overloaded prefix
          while (true) {   // at most two rounds per group: the argmin, then the next match for the tie flag
            const unsigned hit = __ballot_sync(FULL, okv != 0 && !tie_done);
            if (!hit) break;
            const unsigned gb = (hit >> (g * 8)) & 0xFFu;
            const int L = (int)(g * 8) + (gb ? __ffs(gb) - 1 : 0);
            const uint32_t pos = __shfl_sync(FULL, lo + (uint32_t)(okv ? __ffs(okv) - 1 : 0), L);
            if (gb) {
              const uint64_t kk = T.skey[a + pos];
              if (kstar == KEY_NONE) { kstar = kk; kc = 1; }
              else { if ((uint32_t)(kk >> 32) == (uint32_t)(kstar >> 32)) kc = 2; tie_done = true; }
              if ((int)lane == L) okv &= okv - 1;
            }
          }
        }
        merge_best(b, bc, kstar, kc);
      }
      tot += __shfl_xor_sync(FULL, tot, 1);
      tot += __shfl_xor_sync(FULL, tot, 2);
      tot += __shfl_xor_sync(FULL, tot, 4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {   // hand each job's result to its owner lane
        const uint64_t vb = shfl64(FULL, b, q * 8);
        const uint32_t vc = __shfl_sync(FULL, bc, q * 8), vt = __shfl_sync(FULL, tot, q * 8);
        if ((int)lane == own[q]) { best = vb; bcnt = vc; total = vt; }
      }
    }

    // =================================================================== S': pools beyond the sort buffer, warp per job
    for (unsigned todo = __ballot_sync(FULL, need_scan && slow); todo; todo &= todo - 1) {
      const int i = __ffs(todo) - 1;
      const uint32_t off = __shfl_sync(FULL, r_off, i), cnt = __shfl_sync(FULL, r_cnt, i), fl = __shfl_sync(FULL, c_flags, i);
      const int single = __shfl_sync(FULL, r_single, i);
      const uint64_t need_req = shfl64(FULL, c_req, i) & ~T.req_blank_mask;
      const uint64_t need_lo = shfl64(FULL, c_plo, i), need_hi = shfl64(FULL, c_phi, i);
      const bool req_any = fl & JF_REQ_NONEMPTY, req_unknown = fl & JF_REQ_UNKNOWN;
      uint64_t b = KEY_NONE;
      uint32_t bc = 0, tot = 0;
      for (uint32_t k = 0; k < cnt; ++k) {
        const uint32_t pid = single >= 0 ? (uint32_t)single : __ldg(T.pool_list + off + k);
        if (req_any && !(__ldg(T.pool_req_nonempty + pid) && !req_unknown && (need_req & ~__ldg(T.pool_req_mask + pid)) == 0)) continue;
        const uint32_t a = __ldg(T.pool_off + pid), e = __ldg(T.pool_off + pid + 1);
        for (uint32_t pos = a + lane; pos < e; pos += 32) {   // matchesLabels (:161-175), coalesced
          const uint64_t llo = __ldg(T.pos_label_lo + pos), lhi = __ldg(T.pos_label_hi + pos);
          if ((llo & need_lo) != need_lo || (lhi & need_hi) != need_hi) continue;
          tot++;
          merge_best(b, bc, T.pos_key[pos], 1);
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const uint64_t k2 = shfl64_xor(FULL, b, o);
        const uint32_t c2 = __shfl_xor_sync(FULL, bc, o);
        merge_best(b, bc, k2, c2);
        tot += __shfl_xor_sync(FULL, tot, o);
      }
      if ((int)lane == i) { best = b; bcnt = bc; total = tot; }
    }

    if (need_scan) {
      if (best != KEY_NONE) {
        route = CORDUM_ROUTE_OK;
        slot = (int)__ldg(T.rank_slot + (uint32_t)(best & 0xFFFFFFFFu));
        if (bcnt > 1) rflags |= CORDUM_F_TIE;
      } else route = total > 0 ? CORDUM_ROUTE_POOL_OVERLOADED : CORDUM_ROUTE_NO_WORKERS;
    }
    if (valid) {
      rec.x = (rec.x & 0x00FFFFFFu) | (rflags << 16) | (route << 24);   // rflags only adds CORDUM_F_TIE
      rec.w = (uint32_t)slot;
      reinterpret_cast<uint4*>(P.out)[j] = rec;
    }
  }
}

// ------------------------------------------------------------------ launchers (C++ linkage, called by engine.cu)
// An SM changes its L1 / shared-memory split only when it is idle, so kernels that ask for different splits cannot
// share an SM: a refresh CTA (5 KiB of shared memory) would wait for every route / policy CTA (1-2 KiB, i.e. the
// smallest split) on that SM to drain.  All kernels of the path therefore ask for the same carveout, large enough
// for their resident CTAs side by side; the rest (>= 128 KiB) stays L1 for the pass-row gathers.
static cudaError_t configure_kernels() {
  static bool done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  static const int kb = []() { const char* v = getenv("CORDUM_SMEM_KB"); return v ? atoi(v) : 32; }();   // tuning knob
  const int pct = (kb * 100 + 227) / 228;
  const void* fns[] = {(const void*)worker_chunk_kernel<128>, (const void*)worker_chunk_kernel<256>, (const void*)worker_merge_kernel,
                       (const void*)policy_kernel<4, 1>, (const void*)policy_kernel<5, 1>, (const void*)policy_kernel<4, 2>,
                       (const void*)policy_kernel<5, 2>, (const void*)policy_kernel<3, 4>, (const void*)policy_kernel<4, 4>,
                       (const void*)route_kernel<true>, (const void*)route_kernel<false>,
                       (const void*)route_kernel<false, 3>, (const void*)route_kernel<false, 5>};
  for (const void* f : fns) {
    e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e != cudaSuccess) return e;
  }
  e = cudaFuncSetAttribute(worker_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CORDUM_POOL_SORT_MAX * 8u));
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64) done[dev] = true;
  return cudaSuccess;
}

cudaError_t launch_worker_pools(const DeviceTables& T, cudaStream_t s, cudaEvent_t loads_read) {
  if (T.n_pools == 0) return loads_read ? cudaEventRecord(loads_read, s) : cudaSuccess;
  if (cudaError_t c = configure_kernels(); c != cudaSuccess) return c;
  static const int nt = []() { const char* v = getenv("CORDUM_CHUNK_NT"); return v ? atoi(v) : 256; }();   // tuning knob
  if (nt == 256) worker_chunk_kernel<256><<<T.n_chunks, 256, 0, s>>>(T);
  else worker_chunk_kernel<128><<<T.n_chunks, 128, 0, s>>>(T);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && loads_read) e = cudaEventRecord(loads_read, s);   // the load table is not read past this point
  if (e != cudaSuccess || T.n_merge == 0) return e;
  worker_merge_kernel<<<T.n_merge, 256, T.merge_smem, s>>>(T);
  return cudaGetLastError();
}

static uint32_t grid_for(uint32_t n_jobs, int sm_count, int resident) {
  static const int waves = []() { const char* v = getenv("CORDUM_WAVES"); return v ? atoi(v) : 4; }();
  const uint32_t tiles = (n_jobs + 31u) / 32u;
  uint32_t blocks = (tiles + 7u) / 8u;                                               // 8 warps (tiles) per 256-thread CTA
  const uint32_t cap = (uint32_t)sm_count * (uint32_t)resident * (uint32_t)waves;   // multiple of SM count x resident CTAs
  return blocks > cap ? cap : blocks;
}

cudaError_t launch_policy(const KParams& P, int sm_count, cudaStream_t s) {
  if (P.n_jobs == 0) return cudaSuccess;
  if (cudaError_t c = configure_kernels(); c != cudaSuccess) return c;
  static const int minb = []() { const char* v = getenv("CORDUM_MINB"); return v ? atoi(v) : 5; }();   // tuning knob
  const uint32_t blocks = grid_for(P.n_jobs, sm_count, minb);
  const uint32_t iu = P.t.item_u4;
  if (iu == 1) { if (minb >= 5) policy_kernel<5, 1><<<blocks, 256, 0, s>>>(P); else policy_kernel<4, 1><<<blocks, 256, 0, s>>>(P); }
  else if (iu == 2) { if (minb >= 5) policy_kernel<5, 2><<<blocks, 256, 0, s>>>(P); else policy_kernel<4, 2><<<blocks, 256, 0, s>>>(P); }
  else if (iu == 4) { if (minb >= 4) policy_kernel<4, 4><<<blocks, 256, 0, s>>>(P); else policy_kernel<3, 4><<<blocks, 256, 0, s>>>(P); }
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_route(const KParams& P, bool route_only, int sm_count, cudaStream_t s) {
  if (P.n_jobs == 0) return cudaSuccess;
  if (cudaError_t c = configure_kernels(); c != cudaSuccess) return c;
  static const int minb = []() { const char* v = getenv("CORDUM_ROUTE_MINB"); return v ? atoi(v) : 4; }();   // tuning knob
  const uint32_t blocks = grid_for(P.n_jobs, sm_count, route_only ? 4 : minb);
  if (route_only) route_kernel<true><<<blocks, 256, 0, s>>>(P);
  else if (minb == 3) route_kernel<false, 3><<<blocks, 256, 0, s>>>(P);
  else if (minb == 5) route_kernel<false, 5><<<blocks, 256, 0, s>>>(P);
  else route_kernel<false><<<blocks, 256, 0, s>>>(P);
  return cudaGetLastError();
}
