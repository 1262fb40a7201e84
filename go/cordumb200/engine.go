// Package cordumb200 binds libcordum_b200.so (include/cordum_b200.h) and adapts it to the seams of
// cordum's scheduler and safety kernel (SURVEY.md §8b).
//
// NOTE: written against the reference @ c7ddbe09; it could not be compiled in the authoring image
// (no Go toolchain, CAP module not vendored).  The identical C ABI is exercised from Python
// (cordum_b200/reference_api.py, tests/test_reference_api.py), which is what CI runs.
package cordumb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../cordum_b200 -lcordum_b200
#include <stdlib.h>
#include "cordum_b200.h"
*/
import "C"

import (
	"encoding/json"
	"errors"
	"fmt"
	"sync"
	"unsafe"

	pb "github.com/cordum/cordum/core/protocol/pb/v1"
)

// Engine owns one GPU's tables and streams.
type Engine struct {
	h  *C.cordum_engine
	mu sync.Mutex // serialises table loads; dispatch takes per-batch locks
}

func lastErr() error { return errors.New(C.GoString(C.cordum_last_error())) }

// NewEngine fails when no GPU is present: callers must then fail closed (DENY / retryable error).
func NewEngine(device int) (*Engine, error) {
	opts := C.cordum_engine_opts{device: C.int32_t(device)}
	var h *C.cordum_engine
	if rc := C.cordum_engine_create(&opts, &h); rc != 0 {
		return nil, fmt.Errorf("cordum_engine_create: %w", lastErr())
	}
	return &Engine{h: h}, nil
}

func (e *Engine) Close() { C.cordum_engine_destroy(e.h) }

// LoadPolicy mirrors (*server).setPolicy (kernel.go:510-521).  policy is the merged
// *config.SafetyPolicy; its yaml tags are used as JSON keys (see policyJSON).
func (e *Engine) LoadPolicy(policyJSON []byte, snapshot string) error {
	e.mu.Lock()
	defer e.mu.Unlock()
	var p *C.char
	if len(policyJSON) > 0 {
		p = (*C.char)(unsafe.Pointer(&policyJSON[0]))
	}
	cs := C.CString(snapshot)
	defer C.free(unsafe.Pointer(cs))
	if rc := C.cordum_policy_load(e.h, p, C.uint64_t(len(policyJSON)), cs, C.uint64_t(len(snapshot))); rc != 0 {
		return lastErr()
	}
	return nil
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
	if rc := C.cordum_routing_load(e.h, (*C.char)(unsafe.Pointer(&b[0])), C.uint64_t(len(b))); rc != 0 {
		return lastErr()
	}
	return nil
}

// ---- multi-GPU: one scheduler process per GPU, jobs sharded by the queue group, registry replicated.
// Rank 0 calls ExchangeID and ships the 128 bytes to the other ranks (e.g. through the config KV the
// schedulers already share); every rank then calls JoinExchange (collective).  Per heartbeat epoch each
// rank hands IngestHeartbeats the loads of ITS slice of worker slots (the heartbeat fan-in is sharded by
// slot): the engine copies them to the GPU, all-gathers the slices over NCCL and refreshes its tables.

// ExchangeID mirrors ncclGetUniqueId.
func ExchangeID() ([]byte, error) {
	id := make([]byte, C.CORDUM_EXCHANGE_ID_BYTES)
	if rc := C.cordum_exchange_unique_id((*C.char)(unsafe.Pointer(&id[0]))); rc != 0 {
		return nil, lastErr()
	}
	return id, nil
}

// JoinExchange blocks until all `world` ranks have joined.
func (e *Engine) JoinExchange(id []byte, rank, world int) error {
	e.mu.Lock()
	defer e.mu.Unlock()
	if rc := C.cordum_exchange_init(e.h, (*C.char)(unsafe.Pointer(&id[0])), C.int32_t(rank), C.int32_t(world)); rc != 0 {
		return lastErr()
	}
	return nil
}

// IngestHeartbeats takes the {active_jobs, max_parallel_jobs, cpu_load, gpu_utilization} records of this
// rank's slots [first, first+len(loads)), in slot order (registry_memory.go:43-51 keeps the latest heartbeat
// per worker).  loads must live in C memory (C.malloc / cudaHostAlloc) until the next call.
func (e *Engine) IngestHeartbeats(loads *C.cordum_worker_load, first, n int) error {
	if rc := C.cordum_workers_ingest(e.h, loads, C.uint32_t(first), C.uint32_t(n)); rc != 0 {
		return lastErr()
	}
	return nil
}

// arena packs Go strings into one C-allocated byte slab + (off,len) spans.  Nothing Go-allocated is
// retained by C after a call returns (cgo pointer rule): the slab is freed by the caller.
type arena struct {
	buf  []byte
	seen map[string]C.cordum_str
}

func (a *arena) add(s string) C.cordum_str {
	if s == "" {
		return C.cordum_str{}
	}
	if sp, ok := a.seen[s]; ok {
		return sp
	}
	sp := C.cordum_str{off: C.uint32_t(len(a.buf)), len: C.uint32_t(len(s))}
	a.buf = append(a.buf, s...)
	a.seen[s] = sp
	return sp
}

// envelopes lays a slice of PolicyCheckRequests out as the columnar cordum_envelopes struct.
// (C memory management elided for brevity: every column is C.malloc'ed and freed after cordum_encode.)
type envelopeBatch struct {
	c     C.cordum_envelopes
	frees []unsafe.Pointer
}

func actorType(t pb.ActorType) uint8 { // kernel.go:370-379
	switch t {
	case pb.ActorType_ACTOR_TYPE_HUMAN:
		return 1
	case pb.ActorType_ACTOR_TYPE_SERVICE:
		return 2
	}
	return 0
}
