// Package cordumb200 binds libcordum_b200.so (include/cordum_b200.h) and adapts it to the seams of cordum's
// scheduler and safety kernel (SURVEY.md §8b):
//
//	scheduler.SafetyChecker        types.go:29-31     -> Checker           (safety.go)
//	pb.SafetyKernelServer          kernel.go:106-127  -> SafetyKernel      (safety.go)
//	scheduler.SchedulingStrategy   types.go:40-42     -> Strategy          (strategy.go)
//	scheduler.WorkerRegistry       types.go:34-37     -> Registry          (strategy.go)
//
// Every seam is one request per call; the library's micro-batching front-end (cordum_frontend_submit) is exactly that
// call shape, so the adapters below only convert messages: one C allocation per request holds all of its strings, the
// blocking cgo call parks the goroutine while its request rides in a batch, and nothing Go-allocated is visible to C
// (cgo pointer rule).
//
// NOTE: written against the reference @ c7ddbe09 and include/cordum_b200.h; it could not be compiled in the authoring
// image (no Go toolchain, CAP module not vendored).  The same C ABI calls are exercised from Python and C
// (tests/test_frontend.py, tests/test_reference_api.py, examples/host_min.c), which is what CI runs.
package cordumb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../cordum_b200 -lcordum_b200
#include <stdlib.h>
#include <string.h>
#include "cordum_b200.h"
*/
import "C"

import (
	"encoding/json"
	"errors"
	"fmt"
	"os"
	"runtime"
	"sort"
	"strings"
	"sync"
	"time"
	"unsafe"

	pb "github.com/cordum/cordum/core/protocol/pb/v1"
)

// Engine owns one GPU's tables, streams and the front-ends (policy-only for the safety kernel surface, route-only for the
// strategy, policy-and-route for the scheduler's per-job step).
type Engine struct {
	h        *C.cordum_engine
	policyFE *C.cordum_frontend // CORDUM_MODE_POLICY_ONLY
	routeFE  *C.cordum_frontend // CORDUM_MODE_ROUTE_ONLY
	schedFE  *C.cordum_frontend // CORDUM_MODE_POLICY_AND_ROUTE (ProcessJob, process.go)
	mu       sync.Mutex         // serialises table loads and the worker snapshot below

	// worker registry mirror: slot = index in the last cordum_workers_load
	slots   map[string]int
	workers []*pb.Heartbeat
}

// call runs one ABI call and fetches its error text on the same OS thread: cordum_last_error() is thread-local in the
// library and a goroutine may otherwise migrate between the two cgo calls.
func call(f func() C.int32_t) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := f(); rc != 0 {
		return fmt.Errorf("cordum_b200 error %d: %s", int(rc), C.GoString(C.cordum_last_error()))
	}
	return nil
}

// NewEngine fails when no GPU is present (CORDUM_E_NODEVICE): callers must then fail closed (DENY / retryable error).
func NewEngine(device int) (*Engine, error) {
	e := &Engine{slots: map[string]int{}}
	opts := C.cordum_engine_opts{device: C.int32_t(device)}
	if err := call(func() C.int32_t { return C.cordum_engine_create(&opts, &e.h) }); err != nil {
		return nil, err
	}
	// SAFETY_DECISION_CACHE_TTL (kernel.go:46,87,316-327): the safety kernel's decision cache, kept by the policy front-end
	var ttlUs C.uint64_t
	if d, err := time.ParseDuration(strings.TrimSpace(os.Getenv("SAFETY_DECISION_CACHE_TTL"))); err == nil && d > 0 {
		ttlUs = C.uint64_t(d / time.Microsecond)
	}
	mk := func(mode C.uint32_t, out **C.cordum_frontend) error {
		o := C.cordum_frontend_opts{max_batch: 1024, max_wait_us: 200, mode: mode, lanes: 2, arena_bytes_per_request: 2048}
		if mode == C.CORDUM_MODE_POLICY_ONLY {
			o.cache_ttl_us = ttlUs
		}
		return call(func() C.int32_t { return C.cordum_frontend_create(e.h, &o, out) })
	}
	if err := mk(C.CORDUM_MODE_POLICY_ONLY, &e.policyFE); err != nil {
		e.Close()
		return nil, err
	}
	if err := mk(C.CORDUM_MODE_ROUTE_ONLY, &e.routeFE); err != nil {
		e.Close()
		return nil, err
	}
	if err := mk(C.CORDUM_MODE_POLICY_AND_ROUTE, &e.schedFE); err != nil {
		e.Close()
		return nil, err
	}
	return e, nil
}

func (e *Engine) Close() {
	if e.policyFE != nil {
		C.cordum_frontend_destroy(e.policyFE)
	}
	if e.routeFE != nil {
		C.cordum_frontend_destroy(e.routeFE)
	}
	if e.schedFE != nil {
		C.cordum_frontend_destroy(e.schedFE)
	}
	if e.h != nil {
		C.cordum_engine_destroy(e.h)
	}
	e.h, e.policyFE, e.routeFE, e.schedFE = nil, nil, nil, nil
}

// cbytes copies b into C memory (nil for empty); the caller frees it.
func cbytes(b []byte) (*C.char, C.uint64_t) {
	if len(b) == 0 {
		return nil, 0
	}
	return (*C.char)(C.CBytes(b)), C.uint64_t(len(b))
}

// LoadPolicy mirrors (*server).setPolicy (kernel.go:510-521).  policyJSON is the merged *config.SafetyPolicy with its
// yaml tag names as JSON keys (PolicyJSON in safety.go builds it); snapshot is the loader's snapshot id.
func (e *Engine) LoadPolicy(policyJSON []byte, snapshot string) error {
	e.mu.Lock()
	defer e.mu.Unlock()
	p, n := cbytes(policyJSON)
	s, sn := cbytes([]byte(snapshot))
	defer C.free(unsafe.Pointer(p))
	defer C.free(unsafe.Pointer(s))
	return call(func() C.int32_t { return C.cordum_policy_load(e.h, p, n, s, sn) })
}

// LoadRouting mirrors (*LeastLoadedStrategy).UpdateRouting (strategy_least_loaded.go:28-30).
func (e *Engine) LoadRouting(topics map[string][]string, pools map[string][]string) error {
	type pool struct {
		Requires []string `json:"requires"`
	}
	doc := struct {
		Topics map[string][]string `json:"topics"`
		Pools  map[string]pool     `json:"pools"`
	}{Topics: topics, Pools: map[string]pool{}}
	for k, v := range pools {
		doc.Pools[k] = pool{Requires: v}
	}
	b, err := json.Marshal(doc)
	if err != nil {
		return err
	}
	e.mu.Lock()
	defer e.mu.Unlock()
	p, n := cbytes(b)
	defer C.free(unsafe.Pointer(p))
	return call(func() C.int32_t { return C.cordum_routing_load(e.h, p, n) })
}

// Snapshots mirrors ListSnapshots (kernel.go:122-127): newest first, at most 10.
func (e *Engine) Snapshots() []string {
	buf := (*C.char)(C.malloc(1 << 16))
	defer C.free(unsafe.Pointer(buf))
	var n C.uint32_t
	if err := call(func() C.int32_t { return C.cordum_policy_snapshots(e.h, buf, 1<<16, &n) }); err != nil {
		return nil
	}
	out := make([]string, 0, int(n))
	p := unsafe.Pointer(buf)
	for i := 0; i < int(n); i++ {
		s := C.GoString((*C.char)(p))
		out = append(out, s)
		p = unsafe.Add(p, len(s)+1)
	}
	return out
}

// ---- requests ----------------------------------------------------------------------------------------------------

// cArena is one C allocation that holds every string and array of one request.
type cArena struct {
	base unsafe.Pointer
	cap  int
	at   int
}

func (a *cArena) alloc(n, align int) unsafe.Pointer {
	a.at = (a.at + align - 1) &^ (align - 1)
	p := unsafe.Add(a.base, a.at)
	a.at += n
	return p
}

func (a *cArena) sv(s string) C.cordum_sv {
	if s == "" {
		return C.cordum_sv{}
	}
	p := a.alloc(len(s), 1)
	copy(unsafe.Slice((*byte)(p), len(s)), s)
	return C.cordum_sv{p: (*C.char)(p), n: C.uint32_t(len(s))}
}

func actorType(t pb.ActorType) C.uint8_t { // kernel.go:370-379
	switch t {
	case pb.ActorType_ACTOR_TYPE_HUMAN:
		return 1
	case pb.ActorType_ACTOR_TYPE_SERVICE:
		return 2
	}
	return 0
}

// packRequest lays one PolicyCheckRequest (or the PolicyCheckRequest-shaped view of a JobRequest) out in C memory.
// Field set: kernel.go:133-138, 348-368, 381-414; strategy_least_loaded.go:46-62.  free() releases everything.
func packRequest(topic, tenant, principal string, labels map[string]string, meta *pb.JobMetadata, effective []byte, approved bool) (*C.cordum_request, func()) {
	size := int(unsafe.Sizeof(C.cordum_request{})) + len(topic) + len(tenant) + len(principal) + len(effective) + 64
	var tags, reqs []string
	if meta != nil {
		tags, reqs = meta.GetRiskTags(), meta.GetRequires()
		size += len(meta.GetTenantId()) + len(meta.GetActorId()) + len(meta.GetCapability()) + len(meta.GetPackId())
		for _, s := range tags {
			size += len(s)
		}
		for _, s := range reqs {
			size += len(s)
		}
		size += (len(tags) + len(reqs)) * int(unsafe.Sizeof(C.cordum_sv{}))
	}
	for k, v := range labels {
		size += len(k) + len(v)
	}
	size += len(labels)*int(unsafe.Sizeof(C.cordum_kv{})) + 64
	a := &cArena{base: C.calloc(1, C.size_t(size)), cap: size}
	r := (*C.cordum_request)(a.alloc(int(unsafe.Sizeof(C.cordum_request{})), 8))
	r.topic, r.tenant, r.principal_id = a.sv(topic), a.sv(tenant), a.sv(principal)
	r.effective_config = a.sv(string(effective))
	if approved {
		r.approved = 1
	}
	if meta != nil {
		r.has_meta = 1
		r.actor_type = actorType(meta.GetActorType())
		r.meta_tenant_id, r.actor_id = a.sv(meta.GetTenantId()), a.sv(meta.GetActorId())
		r.capability, r.pack_id = a.sv(meta.GetCapability()), a.sv(meta.GetPackId())
		if len(tags) > 0 {
			arr := unsafe.Slice((*C.cordum_sv)(a.alloc(len(tags)*int(unsafe.Sizeof(C.cordum_sv{})), 8)), len(tags))
			for i, s := range tags {
				arr[i] = a.sv(s)
			}
			r.risk_tags, r.n_risk_tags = &arr[0], C.uint32_t(len(tags))
		}
		if len(reqs) > 0 {
			arr := unsafe.Slice((*C.cordum_sv)(a.alloc(len(reqs)*int(unsafe.Sizeof(C.cordum_sv{})), 8)), len(reqs))
			for i, s := range reqs {
				arr[i] = a.sv(s)
			}
			r.requires_, r.n_requires = &arr[0], C.uint32_t(len(reqs))
		}
	}
	if len(labels) > 0 {
		arr := unsafe.Slice((*C.cordum_kv)(a.alloc(len(labels)*int(unsafe.Sizeof(C.cordum_kv{})), 8)), len(labels))
		i := 0
		for k, v := range labels {
			arr[i] = C.cordum_kv{key: a.sv(k), val: a.sv(v)}
			i++
		}
		r.labels, r.n_labels = &arr[0], C.uint32_t(len(labels))
	}
	return r, func() { C.free(a.base) }
}

// submit parks the goroutine in one blocking cgo call while the request rides in a batch.
func submit(fe *C.cordum_frontend, r *C.cordum_request) (*C.cordum_response, func(), error) {
	resp := (*C.cordum_response)(C.calloc(1, C.size_t(unsafe.Sizeof(C.cordum_response{}))))
	free := func() { C.free(unsafe.Pointer(resp)) }
	if rc := C.cordum_frontend_submit(fe, r, resp); rc != 0 {
		msg := C.GoString(&resp.reason[0]) // the front-end writes "safety kernel error: ..." into the response itself
		free()
		return nil, func() {}, errors.New(msg)
	}
	return resp, free, nil
}

// ruleText fetches the pass-through JSON of a rule's constraints / remediations (kernel.go:244,247).
// ruleText resolves a record's rule index against the policy generation the request was evaluated under
// (cordum_response.policy_gen), not the policy in force by the time the response is assembled.
func (e *Engine) ruleText(gen C.uint64_t, idx C.int32_t, f func(*C.cordum_engine, C.uint64_t, C.int32_t, *C.char, C.uint64_t) C.int64_t) []byte {
	const cap = 1 << 16
	buf := (*C.char)(C.malloc(cap))
	defer C.free(unsafe.Pointer(buf))
	n := f(e.h, gen, idx, buf, cap)
	if n <= 0 || n >= cap {
		return nil
	}
	return C.GoBytes(unsafe.Pointer(buf), C.int(n))
}

// ---- workers -----------------------------------------------------------------------------------------------------

// loadWorkersLocked replaces the device registry with e.workers (cordum_workers_load): columnar, C memory.
func (e *Engine) loadWorkersLocked() error {
	n := len(e.workers)
	size := 64
	nl := 0
	for _, hb := range e.workers {
		size += len(hb.GetWorkerId()) + len(hb.GetPool())
		for k, v := range hb.GetLabels() {
			size += len(k) + len(v)
			nl++
		}
	}
	svSz := int(unsafe.Sizeof(C.cordum_str{}))
	size += n*(2*svSz+16) + (n+1)*4 + 2*nl*svSz + 256
	a := &cArena{base: C.calloc(1, C.size_t(size)), cap: size}
	defer C.free(a.base)
	strs := a.alloc(size/2, 1) // string bytes first, offsets are relative to this base
	at := 1                    // offset 0 = the empty string
	put := func(s string) C.cordum_str {
		if s == "" {
			return C.cordum_str{}
		}
		copy(unsafe.Slice((*byte)(unsafe.Add(strs, at)), len(s)), s)
		r := C.cordum_str{off: C.uint32_t(at), len: C.uint32_t(len(s))}
		at += len(s)
		return r
	}
	col := func() []C.cordum_str {
		return unsafe.Slice((*C.cordum_str)(a.alloc(max(n, 1)*svSz, 8)), max(n, 1))
	}
	ids, pools := col(), col()
	active := unsafe.Slice((*C.int32_t)(a.alloc(max(n, 1)*4, 4)), max(n, 1))
	maxp := unsafe.Slice((*C.int32_t)(a.alloc(max(n, 1)*4, 4)), max(n, 1))
	cpu := unsafe.Slice((*C.float)(a.alloc(max(n, 1)*4, 4)), max(n, 1))
	gpu := unsafe.Slice((*C.float)(a.alloc(max(n, 1)*4, 4)), max(n, 1))
	off := unsafe.Slice((*C.uint32_t)(a.alloc((n+1)*4, 4)), n+1)
	lk := unsafe.Slice((*C.cordum_str)(a.alloc(max(nl, 1)*svSz, 8)), max(nl, 1))
	lv := unsafe.Slice((*C.cordum_str)(a.alloc(max(nl, 1)*svSz, 8)), max(nl, 1))
	k := 0
	e.slots = make(map[string]int, n)
	for i, hb := range e.workers {
		ids[i], pools[i] = put(hb.GetWorkerId()), put(hb.GetPool())
		active[i], maxp[i] = C.int32_t(hb.GetActiveJobs()), C.int32_t(hb.GetMaxParallelJobs())
		cpu[i], gpu[i] = C.float(hb.GetCpuLoad()), C.float(hb.GetGpuUtilization())
		for key, val := range hb.GetLabels() {
			lk[k], lv[k] = put(key), put(val)
			k++
		}
		off[i+1] = C.uint32_t(k)
		e.slots[hb.GetWorkerId()] = i
	}
	w := C.cordum_workers{n_workers: C.uint32_t(n), arena: (*C.uint8_t)(strs), arena_len: C.uint64_t(at),
		worker_id: &ids[0], pool: &pools[0], active_jobs: &active[0], max_parallel_jobs: &maxp[0],
		cpu_load: &cpu[0], gpu_utilization: &gpu[0], label_off: &off[0], label_keys: &lk[0], label_vals: &lv[0]}
	// the struct itself must live in C memory too: it holds C pointers only, but cgo checks the argument
	wp := (*C.cordum_workers)(a.alloc(int(unsafe.Sizeof(w)), 8))
	*wp = w
	return call(func() C.int32_t { return C.cordum_workers_load(e.h, wp) })
}

func sameLabels(a, b map[string]string) bool {
	if len(a) != len(b) {
		return false
	}
	for k, v := range a {
		if w, ok := b[k]; !ok || w != v {
			return false
		}
	}
	return true
}

// applyHeartbeat mirrors MemoryRegistry.UpdateHeartbeat (registry_memory.go:43-51) into the device tables: a known
// worker whose pool and labels are unchanged only updates its 16 B load record (the common case, every 10 s per
// worker); a new worker or changed placement data reloads the registry snapshot.
func (e *Engine) applyHeartbeat(hb *pb.Heartbeat) error {
	e.mu.Lock()
	defer e.mu.Unlock()
	if slot, ok := e.slots[hb.GetWorkerId()]; ok {
		old := e.workers[slot]
		if old.GetPool() == hb.GetPool() && sameLabels(old.GetLabels(), hb.GetLabels()) {
			e.workers[slot] = hb
			s := C.uint32_t(slot)
			l := C.cordum_worker_load{active_jobs: C.int32_t(hb.GetActiveJobs()), max_parallel_jobs: C.int32_t(hb.GetMaxParallelJobs()),
				cpu_load: C.float(hb.GetCpuLoad()), gpu_utilization: C.float(hb.GetGpuUtilization())}
			return call(func() C.int32_t { return C.cordum_workers_update(e.h, 1, &s, &l) })
		}
		e.workers[slot] = hb
	} else {
		e.workers = append(e.workers, hb)
	}
	return e.loadWorkersLocked()
}

// expire drops the workers MemoryRegistry has expired (TTL 30 s, registry_memory.go:23,76-81); ids = the survivors.
func (e *Engine) expire(live map[string]*pb.Heartbeat) error {
	e.mu.Lock()
	defer e.mu.Unlock()
	if len(live) == len(e.workers) {
		return nil
	}
	ids := make([]string, 0, len(live))
	for id := range live {
		ids = append(ids, id)
	}
	sort.Strings(ids)
	e.workers = e.workers[:0]
	for _, id := range ids {
		e.workers = append(e.workers, live[id])
	}
	return e.loadWorkersLocked()
}
