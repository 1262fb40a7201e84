/* oracle.h — C ABI of the CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A string-level C++ restatement of the reference's Go code for the hot path
 * (see oracle.cpp for the file:line map).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (cordum_b200/) never links, imports or calls it.
 *
 * It consumes the same boundary data formats as the product (the plain-data
 * structs of include/cordum_b200.h) so both can be fed identical inputs.
 */
#ifndef CORDUM_ORACLE_H
#define CORDUM_ORACLE_H
#include "../include/cordum_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_ctx oracle_ctx;

oracle_ctx* oracle_create(void);
void oracle_destroy(oracle_ctx*);
const char* oracle_last_error(void);

/* same documents as cordum_policy_load / cordum_routing_load / cordum_workers_load */
int32_t oracle_policy_load(oracle_ctx*, const char* json, uint64_t len);
int32_t oracle_routing_load(oracle_ctx*, const char* json, uint64_t len);
int32_t oracle_workers_load(oracle_ctx*, const cordum_workers* w);
int32_t oracle_workers_update(oracle_ctx*, uint32_t n, const uint32_t* slots, const cordum_worker_load* loads);

/* Evaluate jobs [first, first+count) of env one by one, exactly as the reference
 * would (linear first-match rule scan per job, full worker scan per job), fanned
 * over `threads` std::threads (the reference: goroutine per RPC).  out[count]. */
int32_t oracle_eval(oracle_ctx*, const cordum_envelopes* env, uint32_t first, uint32_t count,
                    uint32_t mode, uint32_t threads, cordum_decision* out);

/* One job, full string-level response as JSON:
 * {"decision": "...", "reason": "...", "rule_id": "...", "approval_required": b,
 *  "has_snapshot": b, "has_constraints": b, "sched_decision": "...",
 *  "subject": "...", "route_error": "...", "worker_slot": n, "tie": b}
 * Returns the full length (buf is truncated to cap, NUL-terminated). */
int64_t oracle_eval_one_json(oracle_ctx*, const cordum_envelopes* env, uint32_t job, uint32_t mode,
                             char* buf, uint64_t cap);

/* Direct access to the restated string primitives, for differential tests. */
int32_t oracle_path_match(const char* pat, uint64_t plen, const char* name, uint64_t nlen); /* 1 match, 0 no, -1 ErrBadPattern */
int32_t oracle_equal_fold(const char* a, uint64_t alen, const char* b, uint64_t blen);
/* flavor 0: safety kernel (kernel.go:129-257); 1: the gateway's copy of the evaluator used by policy simulation
 * (gateway/policy_bundles.go:1132-1231) - same decisions, effective-config reasons printed with %q */
int64_t oracle_eval_one_json_flavor(oracle_ctx*, const cordum_envelopes* env, uint32_t job, uint32_t mode,
                                    uint32_t flavor, char* buf, uint64_t cap);
/* strconv.Quote */
int64_t oracle_quote(const char* s, uint64_t n, char* buf, uint64_t cap);
int64_t oracle_to_lower(const char* s, uint64_t n, char* buf, uint64_t cap);
/* writes the trimmed span [*off,*off+*len) */
void oracle_trim_space(const char* s, uint64_t n, uint64_t* off, uint64_t* len);
/* normalizeDecision (safety_policy.go:208-223): returns CORDUM_DEC_* of the normalised string;
 * REQUIRE_HUMAN stands for "require_approval". */
int32_t oracle_normalize_decision(const char* s, uint64_t n);
/* ParseEffectiveSafety (effective.go:12-39): 1 ok, 0 not ok. Writes counts for tests. */
int32_t oracle_parse_effective(const char* s, uint64_t n, uint32_t* n_allowed, uint32_t* n_denied);

#ifdef __cplusplus
}
#endif
#endif
