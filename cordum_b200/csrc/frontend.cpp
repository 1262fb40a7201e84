// frontend.cpp — micro-batching front-end over the batch ABI (include/cordum_b200.h, cordum_frontend_*).
//
// The reference handles one request at a time: the scheduler's handleJobRequest / processJob
// (core/controlplane/scheduler/engine.go:203-443) calls SafetyChecker.Check and SchedulingStrategy.PickSubject per job,
// and grpc-go runs every SafetyKernel RPC on its own goroutine (kernel.go:106-127).  A GPU round trip per request would
// waste the device, so this front-end turns many concurrent blocking single-request calls into batches:
//     cordum_frontend_submit(request) -> decision            blocking, thread-safe, one request
// A lane thread takes the first waiting request, keeps collecting until the batch holds max_batch requests or max_wait_us
// have passed since that first request, lays the requests out in page-locked envelope buffers, runs ONE cordum_encode +
// cordum_dispatch for all of them, formats the strings the response carries (rule id, reason, subject, snapshot) and wakes
// the callers.  Two lanes alternate, so that packing the next batch overlaps the GPU round trip of the current one.
// Failures are per batch and fail closed: every request of a failed batch gets the engine's status and a DENY record.
#include "../../common/nvtx_range.hpp"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <algorithm>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/cordum_b200.h"

namespace {

struct Ticket {
  const cordum_request* req;
  cordum_response* resp;
  // completion is signalled per ticket: waking every waiter of the front-end for every batch does not scale to
  // thousands of blocked callers
  std::mutex m;
  std::condition_variable cv;
  bool done = false;
};

// ---- decision cache (kernel.go:149-162,250-303; SAFETY_DECISION_CACHE_TTL).  The reference keys a cached response by
// snapshot + SHA-256 of the request's deterministic protobuf bytes with the job id cleared.  The key never leaves the
// process; here it is a 128-bit keyed hash (two multiply-xorshift lanes seeded from std::random_device per front-end, so
// colliding requests cannot be prepared offline) over the same information: every request field in a fixed order,
// repeated fields in their order, labels sorted by key (deterministic marshalling sorts map keys), plus the policy
// generation (which the snapshot string stands for in the reference).  cordum_request has no job id.
struct CacheKey {
  uint64_t a, b;
  bool operator==(const CacheKey& o) const { return a == o.a && b == o.b; }
};
struct CacheKeyHash { size_t operator()(const CacheKey& k) const { return (size_t)(k.a ^ (k.b * 0x9E3779B97F4A7C15ull)); } };
struct CacheEntry {
  cordum_response resp;
  std::chrono::steady_clock::time_point expires;
};
struct CacheShard {
  std::mutex m;
  std::unordered_map<CacheKey, CacheEntry, CacheKeyHash> map;
};
struct KeyHasher {
  uint64_t a, b;
  inline void word(uint64_t x) {
    a = (a ^ x) * 0x9FB21C651E98DF25ull; a ^= a >> 32;
    b = (b ^ (x + 0x632BE59BD9B4E019ull)) * 0xD6E8FEB86659FD93ull; b ^= b >> 29;
  }
  void bytes(const char* p, uint32_t n) {
    word(0x8000000000000000ull | n);   // length first: ("ab","c") and ("a","bc") differ
    while (n >= 8) { uint64_t x; std::memcpy(&x, p, 8); word(x); p += 8; n -= 8; }
    if (n) { uint64_t x = 0; std::memcpy(&x, p, n); word(x); }
  }
  void sv(const cordum_sv& s) { bytes(s.p ? s.p : "", s.p ? s.n : 0); }
};

struct Lane {
  cordum_envelopes* env = nullptr;   // page-locked, owned by the engine
  cordum_batch* batch = nullptr;
  std::thread th;
};

}  // namespace

struct cordum_frontend {
  cordum_engine* eng = nullptr;
  cordum_frontend_opts opts{};
  std::mutex mu;
  std::condition_variable cv_work;
  std::deque<Ticket*> queue;
  bool stop = false;
  std::vector<std::unique_ptr<Lane>> lanes;
  std::atomic<uint64_t> n_batches{0}, n_requests{0}, n_full{0}, n_stale_retries{0};
  uint64_t arena_cap = 0;
  uint32_t max_lists = 0;
  // decision cache (POLICY_ONLY front-ends with cache_ttl_us > 0)
  static constexpr int kShards = 16;
  static constexpr size_t kShardMax = 1u << 16;   // entries per shard before the shard is swept / dropped
  CacheShard cache[kShards];
  uint64_t seed_a = 0, seed_b = 0;
  std::atomic<uint64_t> n_hits{0}, n_misses{0};
  bool cache_on() const { return opts.cache_ttl_us != 0 && opts.mode == CORDUM_MODE_POLICY_ONLY; }
  CacheKey key_of(const cordum_request& q, uint64_t gen) const;
  bool cache_get(const CacheKey& k, cordum_response* out);
  void cache_put(const CacheKey& k, const cordum_response& r);

  void run(Lane& L);
  bool pack(Lane& L, std::vector<Ticket*>& items, std::vector<int32_t>& status);
};

namespace {

inline cordum_str put(uint8_t* arena, uint64_t& at, uint64_t cap, cordum_sv s, bool& overflow) {
  cordum_str r{0, 0};
  if (!s.p || s.n == 0) return r;
  if (at + s.n > cap) { overflow = true; return r; }
  std::memcpy(arena + at, s.p, s.n);
  r.off = (uint32_t)at;
  r.len = s.n;
  at += s.n;
  return r;
}

void fail_closed(cordum_response* r, int32_t status, const char* msg) {
  std::memset(r, 0, sizeof *r);
  r->status = status;
  r->rec.decision = CORDUM_DEC_DENY;
  r->rec.sched_decision = CORDUM_DEC_DENY;
  r->rec.rule_idx = -1;
  r->rec.worker_slot = -1;
  std::snprintf(r->reason, sizeof r->reason, "safety kernel error: %s", msg ? msg : "");   // safety_client.go:98-101
}

}  // namespace

// Lay the requests out as one columnar envelope set.  A request that does not fit (arena or list capacity) is failed
// closed on its own; the rest of the batch goes through.
bool cordum_frontend::pack(Lane& L, std::vector<Ticket*>& items, std::vector<int32_t>& status) {
  cordum_envelopes* e = L.env;
  auto* arena = const_cast<uint8_t*>(e->arena);
  auto col = [](const cordum_str* p) { return const_cast<cordum_str*>(p); };
  auto u8 = [](const uint8_t* p) { return const_cast<uint8_t*>(p); };
  auto u32 = [](const uint32_t* p) { return const_cast<uint32_t*>(p); };
  uint64_t at = 1;   // offset 0 = the empty string
  uint32_t n = 0, n_risk = 0, n_req = 0, n_lab = 0;
  u32(e->risk_off)[0] = 0; u32(e->requires_off)[0] = 0; u32(e->label_off)[0] = 0;
  for (size_t i = 0; i < items.size(); ++i) {
    const cordum_request& q = *items[i]->req;
    const uint64_t at0 = at;
    bool ov = n_risk + q.n_risk_tags > max_lists || n_req + q.n_requires > max_lists || n_lab + q.n_labels > max_lists;
    if (!ov) {
      col(e->topic)[n] = put(arena, at, arena_cap, q.topic, ov);
      col(e->tenant)[n] = put(arena, at, arena_cap, q.tenant, ov);
      col(e->principal_id)[n] = put(arena, at, arena_cap, q.principal_id, ov);
      col(e->effective_config)[n] = put(arena, at, arena_cap, q.effective_config, ov);
      u8(e->has_meta)[n] = q.has_meta ? 1 : 0;
      col(e->meta_tenant_id)[n] = put(arena, at, arena_cap, q.meta_tenant_id, ov);
      col(e->actor_id)[n] = put(arena, at, arena_cap, q.actor_id, ov);
      u8(e->actor_type)[n] = q.actor_type;
      col(e->capability)[n] = put(arena, at, arena_cap, q.capability, ov);
      col(e->pack_id)[n] = put(arena, at, arena_cap, q.pack_id, ov);
      u8(e->approved)[n] = q.approved ? 1 : 0;
      for (uint32_t k = 0; k < q.n_risk_tags && !ov; ++k) col(e->risk_tags)[n_risk + k] = put(arena, at, arena_cap, q.risk_tags[k], ov);
      for (uint32_t k = 0; k < q.n_requires && !ov; ++k) col(e->requires_)[n_req + k] = put(arena, at, arena_cap, q.requires_[k], ov);
      for (uint32_t k = 0; k < q.n_labels && !ov; ++k) {
        col(e->label_keys)[n_lab + k] = put(arena, at, arena_cap, q.labels[k].key, ov);
        col(e->label_vals)[n_lab + k] = put(arena, at, arena_cap, q.labels[k].val, ov);
      }
    }
    if (ov) { at = at0; status[i] = CORDUM_E_CAPACITY; continue; }   // this request alone is refused
    status[i] = (int32_t)n;   // its row in the batch
    n_risk += q.n_risk_tags; n_req += q.n_requires; n_lab += q.n_labels;
    ++n;
    u32(e->risk_off)[n] = n_risk; u32(e->requires_off)[n] = n_req; u32(e->label_off)[n] = n_lab;
  }
  e->n_jobs = n;
  e->arena_len = at;
  return n > 0;
}

void cordum_frontend::run(Lane& L) {
  std::vector<Ticket*> items;
  std::vector<int32_t> row;
  while (true) {
    items.clear();
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_work.wait(lk, [&] { return stop || !queue.empty(); });
      if (stop && queue.empty()) return;
      // the first request opens the batch; collect until it is full or max_wait_us have passed since then
      const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(opts.max_wait_us);
      while (true) {
        while (!queue.empty() && items.size() < opts.max_batch) { items.push_back(queue.front()); queue.pop_front(); }
        if (items.size() >= opts.max_batch || stop || opts.max_wait_us == 0) break;
        if (cv_work.wait_until(lk, deadline, [&] { return stop || !queue.empty(); })) { if (queue.empty()) break; continue; }
        break;   // deadline
      }
      if (!queue.empty()) cv_work.notify_one();   // more work than one batch: wake the other lane
    }
    cordum::NvtxRange nvtx_("cordum:frontend_batch");
    if (items.size() >= opts.max_batch) n_full++;
    row.assign(items.size(), 0);
    int32_t rc = CORDUM_OK;
    const char* msg = "";
    const bool any = pack(L, items, row);
    if (any) {
      // a policy / routing / registry reload between the two calls makes the encoded ids stale: encode again against
      // the new tables (the reference's evaluate reads s.policy once under the lock, kernel.go:140-147 - a request
      // racing a reload is answered under one policy or the other, never refused)
      for (int attempt = 0; attempt < 4; ++attempt) {
        rc = cordum_encode(eng, L.batch, L.env);
        if (rc == CORDUM_OK) rc = cordum_dispatch(eng, L.batch, opts.mode);
        if (rc != CORDUM_E_STALE) break;
        n_stale_retries++;
      }
      if (rc != CORDUM_OK) msg = cordum_last_error();
    }
    const cordum_decision* recs = (any && rc == CORDUM_OK) ? cordum_batch_results(L.batch) : nullptr;
    char snap[sizeof(((cordum_response*)nullptr)->snapshot)] = {0};
    const uint64_t gen = recs ? cordum_batch_policy_gen(L.batch) : 0;
    if (recs) cordum_batch_snapshot(L.batch, snap, sizeof snap);   // the snapshot this batch was evaluated under (kernel.go:141,243), "" if that policy has none
    for (size_t i = 0; i < items.size(); ++i) {
      cordum_response* r = items[i]->resp;
      if (row[i] < 0) { fail_closed(r, row[i], "request exceeds the front-end's envelope capacity"); continue; }
      if (!recs) { fail_closed(r, rc, msg); continue; }
      const uint32_t j = (uint32_t)row[i];
      std::memset(r, 0, sizeof *r);
      r->status = CORDUM_OK;
      r->rec = recs[j];
      r->policy_gen = gen;
      if (r->rec.flags & CORDUM_F_HAS_SNAPSHOT) {   // kernel.go:239-248: the early DENY returns carry neither snapshot nor rule id
        std::memcpy(r->snapshot, snap, sizeof snap);
        cordum_rule_id_at(eng, gen, r->rec.rule_idx, r->rule_id, sizeof r->rule_id);
      }
      // kernel.go:198-215: allow and allow_with_constraints drop the reason
      if (r->rec.reason_code != CORDUM_REASON_NONE && r->rec.decision != CORDUM_DEC_ALLOW && r->rec.decision != CORDUM_DEC_ALLOW_WITH_CONSTRAINTS)
        cordum_reason_flavor(eng, L.batch, j, CORDUM_REASON_FLAVOR_KERNEL, L.env, r->reason, sizeof r->reason);   // the request's own spelling in %q
      else if (r->rec.reason_code == CORDUM_REASON_APPROVAL_GRANTED) cordum_reason_flavor(eng, L.batch, j, CORDUM_REASON_FLAVOR_KERNEL, L.env, r->reason, sizeof r->reason);   // the request's own spelling in %q
      if (r->rec.route_status == CORDUM_ROUTE_OK || r->rec.route_status == CORDUM_ROUTE_OK_PREFERRED)
        cordum_subject(eng, L.batch, j, r->subject, sizeof r->subject);
    }
    n_batches++;
    n_requests += items.size();
    for (Ticket* t : items) {
      std::lock_guard<std::mutex> lk(t->m);   // notify under the lock: the waiter cannot return (and destroy t) before we are done with it
      t->done = true;
      t->cv.notify_one();
    }
  }
}

CacheKey cordum_frontend::key_of(const cordum_request& q, uint64_t gen) const {
  KeyHasher h{seed_a, seed_b};
  h.word(gen);
  h.sv(q.topic); h.sv(q.tenant); h.sv(q.principal_id); h.sv(q.effective_config);
  h.word((uint64_t)(q.has_meta ? 1 : 0) | (uint64_t)q.actor_type << 8 | (uint64_t)(q.approved ? 1 : 0) << 16);
  h.sv(q.meta_tenant_id); h.sv(q.actor_id); h.sv(q.capability); h.sv(q.pack_id);
  h.word(q.n_risk_tags);
  for (uint32_t k = 0; k < q.n_risk_tags; ++k) h.sv(q.risk_tags[k]);
  h.word(q.n_requires);
  for (uint32_t k = 0; k < q.n_requires; ++k) h.sv(q.requires_[k]);
  h.word(q.n_labels);
  if (q.n_labels) {   // a map: order of arrival is not information
    uint32_t idx_small[32];
    std::vector<uint32_t> idx_big;
    uint32_t* idx = idx_small;
    if (q.n_labels > 32) { idx_big.resize(q.n_labels); idx = idx_big.data(); }
    for (uint32_t k = 0; k < q.n_labels; ++k) idx[k] = k;
    std::stable_sort(idx, idx + q.n_labels, [&](uint32_t x, uint32_t y) {
      const cordum_sv &a = q.labels[x].key, &b = q.labels[y].key;
      const int c = std::memcmp(a.p ? a.p : "", b.p ? b.p : "", std::min(a.p ? a.n : 0u, b.p ? b.n : 0u));
      return c != 0 ? c < 0 : (a.p ? a.n : 0u) < (b.p ? b.n : 0u);
    });
    for (uint32_t k = 0; k < q.n_labels; ++k) { h.sv(q.labels[idx[k]].key); h.sv(q.labels[idx[k]].val); }
  }
  return CacheKey{h.a, h.b};
}

bool cordum_frontend::cache_get(const CacheKey& k, cordum_response* out) {
  CacheShard& S = cache[k.a % kShards];
  std::lock_guard<std::mutex> g(S.m);
  auto it = S.map.find(k);
  if (it == S.map.end()) return false;
  if (std::chrono::steady_clock::now() > it->second.expires) { S.map.erase(it); return false; }   // kernel.go:285-288
  *out = it->second.resp;
  return true;
}

void cordum_frontend::cache_put(const CacheKey& k, const cordum_response& r) {
  // kernel.go:250-254 caches what the full evaluation returned; the early topic denials return before that point
  if (r.status != CORDUM_OK || !(r.rec.flags & CORDUM_F_HAS_SNAPSHOT)) return;
  CacheShard& S = cache[k.a % kShards];
  const auto now = std::chrono::steady_clock::now();
  std::lock_guard<std::mutex> g(S.m);
  if (S.map.size() >= kShardMax) {   // the reference's map only shrinks on access; bound it: sweep, then drop
    for (auto it = S.map.begin(); it != S.map.end();) it = now > it->second.expires ? S.map.erase(it) : std::next(it);
    if (S.map.size() >= kShardMax) S.map.clear();
  }
  S.map[k] = CacheEntry{r, now + std::chrono::microseconds(opts.cache_ttl_us)};
}

extern "C" {

int32_t cordum_frontend_create(cordum_engine* e, const cordum_frontend_opts* o, cordum_frontend** out) {
  if (!e || !out) return CORDUM_E_INVALID;
  *out = nullptr;
  auto f = std::make_unique<cordum_frontend>();
  f->eng = e;
  f->opts.max_batch = o && o->max_batch ? o->max_batch : 1024;
  f->opts.max_wait_us = o ? o->max_wait_us : 200;
  f->opts.mode = o && o->mode ? o->mode : CORDUM_MODE_POLICY_AND_ROUTE;
  f->opts.lanes = o && o->lanes ? o->lanes : 2;
  f->opts.arena_bytes_per_request = o && o->arena_bytes_per_request ? o->arena_bytes_per_request : 1024;
  f->opts.cache_ttl_us = o ? o->cache_ttl_us : 0;
  { std::random_device rd; f->seed_a = ((uint64_t)rd() << 32) | rd(); f->seed_b = ((uint64_t)rd() << 32) | rd(); }
  f->arena_cap = (uint64_t)f->opts.max_batch * f->opts.arena_bytes_per_request + 16;
  f->max_lists = f->opts.max_batch * 8;
  for (uint32_t i = 0; i < f->opts.lanes; ++i) {
    auto L = std::make_unique<Lane>();
    cordum_envelope_caps caps{f->opts.max_batch, f->max_lists, f->max_lists, f->max_lists, f->arena_cap};
    int32_t rc = cordum_envelopes_alloc(e, &caps, &L->env);
    if (rc == CORDUM_OK) rc = cordum_batch_alloc(e, f->opts.max_batch, &L->batch);
    if (rc != CORDUM_OK) {
      if (L->batch) cordum_batch_free(L->batch);
      if (L->env) cordum_envelopes_free(e, L->env);
      for (auto& l : f->lanes) { cordum_batch_free(l->batch); cordum_envelopes_free(e, l->env); }
      return rc;
    }
    f->lanes.push_back(std::move(L));
  }
  cordum_frontend* fp = f.get();
  for (auto& L : f->lanes) { Lane* lp = L.get(); lp->th = std::thread([fp, lp] { fp->run(*lp); }); }
  *out = f.release();
  return CORDUM_OK;
}

void cordum_frontend_destroy(cordum_frontend* f) {
  if (!f) return;
  { std::lock_guard<std::mutex> lk(f->mu); f->stop = true; }
  f->cv_work.notify_all();
  for (auto& L : f->lanes) if (L->th.joinable()) L->th.join();
  for (auto& L : f->lanes) { cordum_batch_free(L->batch); cordum_envelopes_free(f->eng, L->env); }
  delete f;
}

int32_t cordum_frontend_submit(cordum_frontend* f, const cordum_request* req, cordum_response* resp) {
  if (!f || !req || !resp) return CORDUM_E_INVALID;
  const bool cached = f->cache_on();
  CacheKey key{0, 0};
  uint64_t key_gen = 0;
  if (cached) {
    key_gen = cordum_policy_generation(f->eng);
    key = f->key_of(*req, key_gen);
    if (f->cache_get(key, resp)) { f->n_hits++; return resp->status; }
    f->n_misses++;
  }
  Ticket t;
  t.req = req; t.resp = resp;
  {
    std::lock_guard<std::mutex> lk(f->mu);
    if (f->stop) { fail_closed(resp, CORDUM_E_STATE, "front-end is shutting down"); return CORDUM_E_STATE; }
    f->queue.push_back(&t);
  }
  f->cv_work.notify_one();
  {
    std::unique_lock<std::mutex> lk(t.m);
    t.cv.wait(lk, [&] { return t.done; });
  }
  // the key carries the generation read before the evaluation: a response evaluated under another policy is not stored
  if (cached && resp->policy_gen == key_gen) f->cache_put(key, *resp);
  return resp->status;
}

/* n requests from one caller (e.g. a Go adapter that has drained a channel): queued together, answered together. */
int32_t cordum_frontend_submit_many(cordum_frontend* f, const cordum_request* reqs, uint32_t n, cordum_response* resps) {
  if (!f || (n && (!reqs || !resps))) return CORDUM_E_INVALID;
  std::vector<std::unique_ptr<Ticket>> ts(n);
  {
    std::lock_guard<std::mutex> lk(f->mu);
    if (f->stop) { for (uint32_t i = 0; i < n; ++i) fail_closed(&resps[i], CORDUM_E_STATE, "front-end is shutting down"); return CORDUM_E_STATE; }
    for (uint32_t i = 0; i < n; ++i) {
      ts[i] = std::make_unique<Ticket>();
      ts[i]->req = &reqs[i]; ts[i]->resp = &resps[i];
      f->queue.push_back(ts[i].get());
    }
  }
  f->cv_work.notify_all();
  int32_t worst = CORDUM_OK;
  for (uint32_t i = 0; i < n; ++i) {
    std::unique_lock<std::mutex> lk(ts[i]->m);
    ts[i]->cv.wait(lk, [&] { return ts[i]->done; });
    if (resps[i].status != CORDUM_OK) worst = resps[i].status;
  }
  return worst;
}

/* Load generator (diagnostics): `threads` native threads submit the given requests round-robin, one blocking call at a
 * time, for `seconds`; per-request latencies in microseconds go to lat_us (up to cap entries, thread-interleaved).
 * Returns the number of requests completed.  Native threads: a Python harness would measure its own interpreter lock. */
uint64_t cordum_frontend_loadgen(cordum_frontend* f, const cordum_request* reqs, uint32_t n_reqs, uint32_t threads, double seconds,
                                 float* lat_us, uint64_t cap) {
  if (!f || !reqs || !n_reqs || !threads) return 0;
  std::atomic<uint64_t> total{0};
  const auto stop_at = std::chrono::steady_clock::now() + std::chrono::duration<double>(seconds);
  std::vector<std::thread> ts;
  for (uint32_t t = 0; t < threads; ++t)
    ts.emplace_back([&, t] {
      cordum_response resp;
      uint64_t i = t, mine = 0;
      while (std::chrono::steady_clock::now() < stop_at) {
        const auto t0 = std::chrono::steady_clock::now();
        cordum_frontend_submit(f, &reqs[i % n_reqs], &resp);
        const float us = std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - t0).count();
        const uint64_t slot = mine * threads + t;
        if (lat_us && slot < cap) lat_us[slot] = us;
        ++mine;
        i += threads;
      }
      total += mine;
    });
  for (auto& th : ts) th.join();
  return total.load();
}

int32_t cordum_frontend_cache_stats(cordum_frontend* f, uint64_t* hits, uint64_t* misses, uint64_t* entries) {
  if (!f) return CORDUM_E_INVALID;
  if (hits) *hits = f->n_hits.load();
  if (misses) *misses = f->n_misses.load();
  if (entries) {
    uint64_t n = 0;
    for (auto& S : f->cache) { std::lock_guard<std::mutex> g(S.m); n += S.map.size(); }
    *entries = n;
  }
  return CORDUM_OK;
}

int32_t cordum_frontend_stats(cordum_frontend* f, uint64_t* batches, uint64_t* requests, uint64_t* full_batches) {
  if (!f) return CORDUM_E_INVALID;
  if (batches) *batches = f->n_batches.load();
  if (requests) *requests = f->n_requests.load();
  if (full_batches) *full_batches = f->n_full.load();
  return CORDUM_OK;
}

}  // extern "C"
