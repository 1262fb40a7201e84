/* Minimal C99 host of libcordum_b200.so: the hello-pack case of BASELINE config 1 (one echo job, one allow rule,
 * a two-worker pool) through the C ABI exactly as a cgo binding would drive it.
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/host_min.c -Lcordum_b200 -lcordum_b200 \
 *       -Wl,-rpath,$PWD/cordum_b200 -o /tmp/host_min && /tmp/host_min
 *
 * Without a GPU cordum_engine_create fails with CORDUM_E_NODEVICE (there is no CPU path): the program says so and
 * exits 0, which is what tests/test_abi.py checks on CPU-only machines. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cordum_b200.h"

static uint8_t arena[1024];
static uint32_t arena_len = 1; /* offset 0 = the empty string */

static cordum_str put(const char* s) {
  cordum_str r;
  r.off = arena_len;
  r.len = (uint32_t)strlen(s);
  memcpy(arena + arena_len, s, r.len);
  arena_len += r.len;
  return r;
}

#define CHECK(call)                                                      \
  do {                                                                   \
    int32_t rc_ = (call);                                                \
    if (rc_ != CORDUM_OK) {                                              \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, cordum_last_error()); \
      return 1;                                                          \
    }                                                                    \
  } while (0)

int main(void) {
  cordum_engine_opts opts;
  cordum_engine* eng = NULL;
  cordum_batch* batch = NULL;
  int32_t rc;
  memset(&opts, 0, sizeof opts);
  rc = cordum_engine_create(&opts, &eng);
  if (rc == CORDUM_E_NODEVICE) {
    printf("no CUDA device: %s (the library has no CPU path)\n", cordum_last_error());
    return 0;
  }
  if (rc != CORDUM_OK) {
    fprintf(stderr, "cordum_engine_create -> %d: %s\n", rc, cordum_last_error());
    return 1;
  }

  {
    static const char policy[] =
        "{\"default_tenant\":\"default\",\"rules\":[{\"id\":\"hello-pack-allow\",\"decision\":\"allow\","
        "\"match\":{\"topics\":[\"job.hello-pack.*\"],\"capabilities\":[\"hello-pack.echo\"]}}]}";
    static const char routing[] =
        "{\"topics\":{\"job.hello-pack.echo\":[\"hello-pack\"]},\"pools\":{\"hello-pack\":{\"requires\":[\"local\"]}}}";
    CHECK(cordum_policy_load(eng, policy, sizeof policy - 1, "cfg:demo", 8));
    CHECK(cordum_routing_load(eng, routing, sizeof routing - 1));
  }
  {
    /* two heartbeats of the hello worker (examples/hello-worker-go/main.go:44-50) */
    cordum_workers w;
    cordum_str ids[2], pools[2];
    int32_t active[2] = {1, 0}, maxp[2] = {4, 4};
    float cpu[2] = {0.f, 0.f}, gpu[2] = {0.f, 0.f};
    uint32_t label_off[3] = {0, 0, 0};
    memset(&w, 0, sizeof w);
    ids[0] = put("hello-worker-a"); ids[1] = put("hello-worker-b");
    pools[0] = put("hello-pack");   pools[1] = pools[0];
    w.n_workers = 2; w.arena = arena; w.arena_len = arena_len;
    w.worker_id = ids; w.pool = pools; w.active_jobs = active; w.max_parallel_jobs = maxp;
    w.cpu_load = cpu; w.gpu_utilization = gpu; w.label_off = label_off;
    CHECK(cordum_workers_load(eng, &w));
  }
  {
    /* one job, columnar */
    cordum_envelopes env;
    cordum_str topic, tenant, empty = {0, 0}, cap, pack, lkeys[3], lvals[3], req[1];
    uint8_t has_meta = 1, actor_type = 0;
    uint32_t off0[2] = {0, 0}, req_off[2] = {0, 1}, label_off[2] = {0, 3};
    const cordum_decision* d;
    char buf[256];
    memset(&env, 0, sizeof env);
    topic = put("job.hello-pack.echo"); tenant = put("default");
    cap = put("hello-pack.echo"); pack = put("hello-pack"); req[0] = put("local");
    lkeys[0] = put("workflow_id"); lvals[0] = put("wf-1");
    lkeys[1] = put("run_id");      lvals[1] = put("run-1");
    lkeys[2] = put("step_id");     lvals[2] = put("step-1");
    env.n_jobs = 1; env.arena = arena; env.arena_len = arena_len;
    env.topic = &topic; env.tenant = &tenant; env.principal_id = &empty; env.effective_config = &empty;
    env.has_meta = &has_meta; env.meta_tenant_id = &empty; env.actor_id = &empty; env.actor_type = &actor_type;
    env.capability = &cap; env.pack_id = &pack;
    env.risk_off = off0; env.risk_tags = &empty;
    env.requires_off = req_off; env.requires_ = req;
    env.label_off = label_off; env.label_keys = lkeys; env.label_vals = lvals;
    CHECK(cordum_batch_alloc(eng, 16, &batch));
    CHECK(cordum_encode(eng, batch, &env));
    CHECK(cordum_dispatch(eng, batch, CORDUM_MODE_POLICY_AND_ROUTE));
    d = cordum_batch_results(batch);
    cordum_rule_id(eng, d[0].rule_idx, buf, sizeof buf);
    printf("decision %u rule %s", (unsigned)d[0].decision, buf);
    cordum_subject(eng, batch, 0, buf, sizeof buf);
    printf(" route_status %u subject %s\n", (unsigned)d[0].route_status, buf);
    if (d[0].decision != CORDUM_DEC_ALLOW || d[0].route_status != CORDUM_ROUTE_OK || d[0].worker_slot != 1) {
      fprintf(stderr, "unexpected result\n");
      return 1;
    }
  }
  cordum_batch_free(batch);
  cordum_engine_destroy(eng);
  return 0;
}
