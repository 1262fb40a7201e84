package cordumb200

import (
	"context"
	"time"

	"github.com/cordum/cordum/core/controlplane/scheduler"
	"github.com/cordum/cordum/core/infra/config"
	pb "github.com/cordum/cordum/core/protocol/pb/v1"
)

// SafetyKernel implements pb.SafetyKernelServer (Check/Evaluate/Explain/Simulate/ListSnapshots,
// kernel.go:106-127) and scheduler.SafetyChecker (types.go:29-31) over one Engine.
//
// grpc-go runs every RPC on its own goroutine; a single GPU round trip per request would waste the
// device, so concurrent calls are micro-batched: a request waits until the batch holds maxBatch
// requests or maxWait has elapsed, then one cordum_encode + cordum_dispatch(POLICY_ONLY) serves all.
type SafetyKernel struct {
	pb.UnimplementedSafetyKernelServer
	eng      *Engine
	in       chan *pending
	maxBatch int
	maxWait  time.Duration
}

type pending struct {
	req  *pb.PolicyCheckRequest
	resp chan *pb.PolicyCheckResponse
}

func NewSafetyKernel(eng *Engine) *SafetyKernel {
	k := &SafetyKernel{eng: eng, in: make(chan *pending, 4096), maxBatch: 1024, maxWait: 200 * time.Microsecond}
	go k.loop()
	return k
}

func (k *SafetyKernel) loop() {
	for first := range k.in {
		batch := []*pending{first}
		timer := time.NewTimer(k.maxWait)
	fill:
		for len(batch) < k.maxBatch {
			select {
			case p := <-k.in:
				batch = append(batch, p)
			case <-timer.C:
				break fill
			}
		}
		timer.Stop()
		resps, err := k.eng.evaluate(batch) // encode + dispatch + materialise strings
		for i, p := range batch {
			if err != nil {
				// fail closed, exactly what SafetyClient does on a transport error (safety_client.go:98-101)
				p.resp <- &pb.PolicyCheckResponse{Decision: pb.DecisionType_DECISION_TYPE_DENY, Reason: "safety kernel error: " + err.Error()}
				continue
			}
			p.resp <- resps[i]
		}
	}
}

func (k *SafetyKernel) evaluate(ctx context.Context, req *pb.PolicyCheckRequest) (*pb.PolicyCheckResponse, error) {
	p := &pending{req: req, resp: make(chan *pb.PolicyCheckResponse, 1)}
	k.in <- p
	select {
	case r := <-p.resp:
		return r, nil
	case <-ctx.Done():
		return nil, ctx.Err()
	}
}

// All four modes are the same function in the reference (kernel.go:129: the mode string is ignored).
func (k *SafetyKernel) Check(ctx context.Context, r *pb.PolicyCheckRequest) (*pb.PolicyCheckResponse, error) {
	return k.evaluate(ctx, r)
}
func (k *SafetyKernel) Evaluate(ctx context.Context, r *pb.PolicyCheckRequest) (*pb.PolicyCheckResponse, error) {
	return k.evaluate(ctx, r)
}
func (k *SafetyKernel) Explain(ctx context.Context, r *pb.PolicyCheckRequest) (*pb.PolicyCheckResponse, error) {
	return k.evaluate(ctx, r)
}
func (k *SafetyKernel) Simulate(ctx context.Context, r *pb.PolicyCheckRequest) (*pb.PolicyCheckResponse, error) {
	return k.evaluate(ctx, r)
}
func (k *SafetyKernel) ListSnapshots(context.Context, *pb.ListSnapshotsRequest) (*pb.ListSnapshotsResponse, error) {
	return &pb.ListSnapshotsResponse{Snapshots: k.eng.snapshots()}, nil
}

// Checker is the in-process scheduler.SafetyChecker: same field mapping as SafetyClient.Check
// (safety_client.go:80-95), no gRPC hop.
type Checker struct{ k *SafetyKernel }

func (c Checker) Check(req *pb.JobRequest) (scheduler.SafetyDecisionRecord, error) {
	creq := &pb.PolicyCheckRequest{
		JobId: req.GetJobId(), Topic: req.GetTopic(), Tenant: scheduler.ExtractTenant(req),
		PrincipalId: req.GetPrincipalId(), Priority: req.GetPriority(), Budget: req.GetBudget(),
		Labels: req.GetLabels(), MemoryId: req.GetMemoryId(), Meta: req.GetMeta(),
	}
	if env := req.GetEnv(); env != nil {
		if eff := env[config.EffectiveConfigEnvVar]; eff != "" {
			creq.EffectiveConfig = []byte(eff)
		}
	}
	resp, err := c.k.evaluate(context.Background(), creq)
	if err != nil {
		return scheduler.SafetyDecisionRecord{Decision: scheduler.SafetyDeny, Reason: "safety kernel error: " + err.Error()}, nil
	}
	return scheduler.SafetyDecisionRecord{
		Decision: decisionFromProto(resp.GetDecision()), Reason: resp.GetReason(), RuleID: resp.GetRuleId(),
		PolicySnapshot: resp.GetPolicySnapshot(), Constraints: resp.GetConstraints(),
		ApprovalRequired: resp.GetApprovalRequired(), ApprovalRef: resp.GetApprovalRef(), Remediations: resp.GetRemediations(),
	}, nil
}

func decisionFromProto(d pb.DecisionType) scheduler.SafetyDecision { // safety_client.go:117-132
	switch d {
	case pb.DecisionType_DECISION_TYPE_ALLOW:
		return scheduler.SafetyAllow
	case pb.DecisionType_DECISION_TYPE_REQUIRE_HUMAN:
		return scheduler.SafetyRequireApproval
	case pb.DecisionType_DECISION_TYPE_THROTTLE:
		return scheduler.SafetyThrottle
	case pb.DecisionType_DECISION_TYPE_ALLOW_WITH_CONSTRAINTS:
		return scheduler.SafetyAllowWithConstraints
	}
	return scheduler.SafetyDeny
}
