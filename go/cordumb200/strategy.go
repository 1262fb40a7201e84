package cordumb200

import (
	"fmt"
	"sync/atomic"

	"github.com/cordum/cordum/core/controlplane/scheduler"
	pb "github.com/cordum/cordum/core/protocol/pb/v1"
)

// Strategy implements scheduler.SchedulingStrategy (types.go:40-42) plus the two methods
// cmd/cordum-scheduler calls on the concrete type (UpdateRouting / CurrentRouting,
// strategy_least_loaded.go:28-38; main.go:149, config_overlay.go:141).
type Strategy struct {
	eng     *Engine
	routing atomic.Value // scheduler.PoolRouting
	reg     *Registry
}

func NewStrategy(eng *Engine, reg *Registry, routing scheduler.PoolRouting) *Strategy {
	s := &Strategy{eng: eng, reg: reg}
	s.UpdateRouting(routing)
	return s
}

func (s *Strategy) UpdateRouting(r scheduler.PoolRouting) {
	pools := map[string][]string{}
	for name, p := range r.Pools {
		pools[name] = append([]string{}, p.Requires...)
	}
	_ = s.eng.LoadRouting(r.Topics, pools)
	s.routing.Store(r)
}

func (s *Strategy) CurrentRouting() scheduler.PoolRouting {
	if r, ok := s.routing.Load().(scheduler.PoolRouting); ok {
		return r
	}
	return scheduler.PoolRouting{}
}

// PickSubject ignores the `workers` argument: the registry adapter below already mirrors every
// heartbeat into the device worker table, so the engine never copies the 64k-entry map per job
// (the reference does, registry_memory.go:71-84 via engine.go:392).
func (s *Strategy) PickSubject(req *pb.JobRequest, _ map[string]*pb.Heartbeat) (string, error) {
	if req == nil || req.Topic == "" {
		return "", fmt.Errorf("missing topic")
	}
	rec, subject, err := s.eng.route(req) // micro-batched like SafetyKernel.evaluate, MODE_ROUTE_ONLY
	if err != nil {
		return "", fmt.Errorf("%w: engine: %v", scheduler.ErrNoWorkers, err) // retryable: the job is NAK'd, never mis-routed
	}
	switch rec.route_status {
	case 1, 2: // CORDUM_ROUTE_OK, CORDUM_ROUTE_OK_PREFERRED
		return subject, nil
	case 4:
		return "", fmt.Errorf("%w: preferred pool %q not mapped for topic %q", scheduler.ErrNoPoolMapping, req.GetLabels()["preferred_pool"], req.Topic)
	case 5:
		return "", fmt.Errorf("%w: topic %q", scheduler.ErrNoPoolMapping, req.Topic)
	case 6:
		return "", fmt.Errorf("%w: no pool satisfies requires", scheduler.ErrNoPoolMapping)
	case 8:
		return "", fmt.Errorf("%w: pool", scheduler.ErrPoolOverloaded)
	}
	return "", fmt.Errorf("%w: pool", scheduler.ErrNoWorkers)
}

// Registry implements scheduler.WorkerRegistry (types.go:34-37): it keeps the reference's
// MemoryRegistry for Snapshot() (main.go:150 publishes it) and forwards heartbeats to the engine:
// cordum_workers_update for a known worker whose pool/labels are unchanged (the common case, every
// 10 s per worker), cordum_workers_load when the set of workers or their labels changed or a TTL expired.
type Registry struct {
	*scheduler.MemoryRegistry
	eng *Engine
}

func (r *Registry) UpdateHeartbeat(hb *pb.Heartbeat) {
	r.MemoryRegistry.UpdateHeartbeat(hb)
	r.eng.applyHeartbeat(hb)
}
