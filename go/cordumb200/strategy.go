package cordumb200

/*
#include "cordum_b200.h"
*/
import "C"

import (
	"fmt"
	"sync/atomic"

	"github.com/cordum/cordum/core/controlplane/scheduler"
	pb "github.com/cordum/cordum/core/protocol/pb/v1"
)

// Strategy implements scheduler.SchedulingStrategy (types.go:40-42) plus the two methods cmd/cordum-scheduler calls on
// the concrete type (UpdateRouting / CurrentRouting, strategy_least_loaded.go:28-38; main.go:149, config_overlay.go:141).
type Strategy struct {
	eng     *Engine
	routing atomic.Value // scheduler.PoolRouting
}

func NewStrategy(eng *Engine, routing scheduler.PoolRouting) *Strategy {
	s := &Strategy{eng: eng}
	s.UpdateRouting(routing)
	return s
}

func (s *Strategy) UpdateRouting(r scheduler.PoolRouting) {
	pools := map[string][]string{}
	for name, p := range r.Pools {
		pools[name] = append([]string{}, p.Requires...)
	}
	if err := s.eng.LoadRouting(r.Topics, pools); err == nil { // a refused table (capacity) keeps the previous routing on both sides
		s.routing.Store(r)
	}
}

func (s *Strategy) CurrentRouting() scheduler.PoolRouting {
	if r, ok := s.routing.Load().(scheduler.PoolRouting); ok {
		return r
	}
	return scheduler.PoolRouting{}
}

// PickSubject ignores the `workers` argument: the Registry below mirrors every heartbeat into the device worker table, so
// nothing copies the 64k-entry map per job (the reference does: registry_memory.go:71-84 via engine.go:392).
// Errors wrap the sentinels engine.go:445-472 classifies with errors.Is.
func (s *Strategy) PickSubject(req *pb.JobRequest, _ map[string]*pb.Heartbeat) (string, error) {
	if req == nil || req.GetTopic() == "" {
		return "", fmt.Errorf("missing topic") // strategy_least_loaded.go:41-43
	}
	creq, free := packRequest(req.GetTopic(), "", "", req.GetLabels(), req.GetMeta(), nil, false)
	defer free()
	r, freeResp, err := submit(s.eng.routeFE, creq)
	if err != nil {
		return "", fmt.Errorf("%w: engine: %v", scheduler.ErrNoWorkers, err) // retryable: the job is NAK'd, never mis-routed
	}
	defer freeResp()
	return subjectFromResponse(r, req)
}

// subjectFromResponse maps the routing half of a record to PickSubject's (subject, error) (strategy_least_loaded.go:40-136).
func subjectFromResponse(r *C.cordum_response, req *pb.JobRequest) (string, error) {
	switch r.rec.route_status {
	case C.CORDUM_ROUTE_OK, C.CORDUM_ROUTE_OK_PREFERRED:
		if subject := C.GoString(&r.subject[0]); subject != "" {
			return subject, nil // "worker.<id>.jobs" (bus/nats.go:94-99)
		}
		return req.GetTopic(), nil // empty worker id: fall back to the topic (:131-135)
	case C.CORDUM_ROUTE_MISSING_TOPIC:
		return "", fmt.Errorf("missing topic")
	case C.CORDUM_ROUTE_NO_POOL_PREFERRED:
		return "", fmt.Errorf("%w: preferred pool %q not mapped for topic %q", scheduler.ErrNoPoolMapping, req.GetLabels()["preferred_pool"], req.GetTopic())
	case C.CORDUM_ROUTE_NO_POOL_TOPIC:
		return "", fmt.Errorf("%w: topic %q", scheduler.ErrNoPoolMapping, req.GetTopic())
	case C.CORDUM_ROUTE_NO_POOL_REQUIRES:
		return "", fmt.Errorf("%w: no pool satisfies requires for topic %q", scheduler.ErrNoPoolMapping, req.GetTopic())
	case C.CORDUM_ROUTE_POOL_OVERLOADED:
		return "", fmt.Errorf("%w: topic %q", scheduler.ErrPoolOverloaded, req.GetTopic())
	}
	return "", fmt.Errorf("%w: topic %q", scheduler.ErrNoWorkers, req.GetTopic())
}

// Registry implements scheduler.WorkerRegistry (types.go:34-37): it keeps the reference's MemoryRegistry for Snapshot()
// (main.go:150 publishes it; TTL 30 s, registry_memory.go:23,76-81) and mirrors it into the engine: a heartbeat of a
// known worker with unchanged pool and labels is one cordum_workers_update, anything else reloads the snapshot.
type Registry struct {
	*scheduler.MemoryRegistry
	eng *Engine
}

func NewRegistry(eng *Engine) *Registry {
	return &Registry{MemoryRegistry: scheduler.NewMemoryRegistry(), eng: eng}
}

func (r *Registry) UpdateHeartbeat(hb *pb.Heartbeat) {
	r.MemoryRegistry.UpdateHeartbeat(hb)
	_ = r.eng.applyHeartbeat(hb) // a refused load (capacity) leaves the previous device registry in place
}

// Snapshot also lets the engine drop what the TTL has expired.
func (r *Registry) Snapshot() map[string]*pb.Heartbeat {
	snap := r.MemoryRegistry.Snapshot()
	_ = r.eng.expire(snap)
	return snap
}
