package cordumb200

/*
#include "cordum_b200.h"
*/
import "C"

import (
	"context"
	"encoding/json"

	"gopkg.in/yaml.v3"

	"github.com/cordum/cordum/core/controlplane/scheduler"
	"github.com/cordum/cordum/core/infra/config"
	pb "github.com/cordum/cordum/core/protocol/pb/v1"
)

// SafetyKernel implements pb.SafetyKernelServer (Check/Evaluate/Explain/Simulate/ListSnapshots, kernel.go:106-127)
// over one Engine.  grpc-go runs every RPC on its own goroutine; each handler converts its request, blocks in
// cordum_frontend_submit while the request rides in a batch with the other in-flight RPCs, and converts the answer.
type SafetyKernel struct {
	pb.UnimplementedSafetyKernelServer
	eng *Engine
}

func NewSafetyKernel(eng *Engine) *SafetyKernel { return &SafetyKernel{eng: eng} }

// PolicyJSON renders the merged *config.SafetyPolicy with its yaml tag names as JSON keys, which is the document
// cordum_policy_load takes (include/cordum_b200.h).  config's structs carry yaml tags only, so go through a generic map.
func PolicyJSON(p *config.SafetyPolicy) ([]byte, error) {
	if p == nil {
		return nil, nil // nil policy: allow-all (kernel.go:187)
	}
	y, err := yaml.Marshal(p) // gopkg.in/yaml.v3, the module the reference parses its policies with (go.mod)
	if err != nil {
		return nil, err
	}
	var generic map[string]any
	if err := yaml.Unmarshal(y, &generic); err != nil {
		return nil, err
	}
	return json.Marshal(generic)
}

var decisionToProto = map[C.uint8_t]pb.DecisionType{ // include/cordum_b200.h CORDUM_DEC_* -> pb.go:61-66
	C.CORDUM_DEC_ALLOW:                  pb.DecisionType_DECISION_TYPE_ALLOW,
	C.CORDUM_DEC_DENY:                   pb.DecisionType_DECISION_TYPE_DENY,
	C.CORDUM_DEC_REQUIRE_HUMAN:          pb.DecisionType_DECISION_TYPE_REQUIRE_HUMAN,
	C.CORDUM_DEC_THROTTLE:               pb.DecisionType_DECISION_TYPE_THROTTLE,
	C.CORDUM_DEC_ALLOW_WITH_CONSTRAINTS: pb.DecisionType_DECISION_TYPE_ALLOW_WITH_CONSTRAINTS,
}

func toProtoConstraints(c config.PolicyConstraints) *pb.PolicyConstraints { // kernel.go:416-445 (emptiness was decided by the engine)
	return &pb.PolicyConstraints{
		Budgets: &pb.BudgetConstraints{MaxRuntimeMs: c.Budgets.MaxRuntimeMs, MaxRetries: c.Budgets.MaxRetries,
			MaxArtifactBytes: c.Budgets.MaxArtifactBytes, MaxConcurrentJobs: c.Budgets.MaxConcurrentJobs},
		Sandbox: &pb.SandboxProfile{Isolated: c.Sandbox.Isolated, NetworkAllowlist: c.Sandbox.NetworkAllowlist,
			FsReadOnly: c.Sandbox.FsReadOnly, FsReadWrite: c.Sandbox.FsReadWrite},
		Toolchain:      &pb.ToolchainConstraints{AllowedTools: c.Toolchain.AllowedTools, AllowedCommands: c.Toolchain.AllowedCommands},
		Diff:           &pb.DiffConstraints{MaxFiles: c.Diff.MaxFiles, MaxLines: c.Diff.MaxLines, DenyPathGlobs: c.Diff.DenyPathGlobs},
		RedactionLevel: c.RedactionLevel,
	}
}

// evaluate is the batched form of (*server).evaluate (kernel.go:129-257) for one request.
func (k *SafetyKernel) evaluate(ctx context.Context, req *pb.PolicyCheckRequest) (*pb.PolicyCheckResponse, error) {
	if err := ctx.Err(); err != nil {
		return nil, err
	}
	creq, free := packRequest(req.GetTopic(), req.GetTenant(), req.GetPrincipalId(), req.GetLabels(), req.GetMeta(), req.GetEffectiveConfig(), false)
	defer free()
	r, freeResp, err := submit(k.eng.policyFE, creq)
	if err != nil {
		// Evaluate never returns a Go error for policy outcomes (kernel.go:172,175); an engine failure fails closed,
		// exactly what SafetyClient does on a transport error (safety_client.go:98-101)
		return &pb.PolicyCheckResponse{Decision: pb.DecisionType_DECISION_TYPE_DENY, Reason: err.Error()}, nil
	}
	defer freeResp()
	return k.eng.policyResponse(r, req.GetJobId()), nil
}

// policyResponse assembles the PolicyCheckResponse of one front-end response (kernel.go:233-248): strings and pass-through
// objects come from the policy generation the request was evaluated under (r.policy_gen), not from the policy in force now.
func (e *Engine) policyResponse(r *C.cordum_response, jobID string) *pb.PolicyCheckResponse {
	out := &pb.PolicyCheckResponse{
		Decision: decisionToProto[r.rec.decision], Reason: C.GoString(&r.reason[0]),
		PolicySnapshot: C.GoString(&r.snapshot[0]), RuleId: C.GoString(&r.rule_id[0]),
		ApprovalRequired: r.rec.flags&C.CORDUM_F_APPROVAL_REQUIRED != 0,
	}
	if out.ApprovalRequired {
		out.ApprovalRef = jobID // kernel.go:233-237
	}
	if r.rec.flags&C.CORDUM_F_CONSTRAINTS != 0 {
		if js := e.ruleText(r.policy_gen, r.rec.rule_idx, func(h *C.cordum_engine, g C.uint64_t, i C.int32_t, b *C.char, n C.uint64_t) C.int64_t {
			return C.cordum_rule_constraints_json_at(h, g, i, b, n)
		}); js != nil {
			var c config.PolicyConstraints
			if json.Unmarshal(js, &c) == nil { // the engine passes the rule's YAML-tagged object through as JSON
				out.Constraints = toProtoConstraints(c)
			}
		}
	}
	if r.rec.rule_idx >= 0 {
		if js := e.ruleText(r.policy_gen, r.rec.rule_idx, func(h *C.cordum_engine, g C.uint64_t, i C.int32_t, b *C.char, n C.uint64_t) C.int64_t {
			return C.cordum_rule_remediations_json_at(h, g, i, b, n)
		}); js != nil {
			var rems []config.PolicyRemediation
			if json.Unmarshal(js, &rems) == nil {
				for _, rem := range rems { // kernel.go:326-346
					out.Remediations = append(out.Remediations, &pb.PolicyRemediation{Id: rem.ID, Title: rem.Title, Summary: rem.Summary,
						ReplacementTopic: rem.ReplacementTopic, ReplacementCapability: rem.ReplacementCapability,
						AddLabels: rem.AddLabels, RemoveLabels: append([]string{}, rem.RemoveLabels...)})
				}
			}
		}
	}
	return out
}

// All four modes are the same function in the reference (kernel.go:106-120,129: the mode string is ignored).
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
	return &pb.ListSnapshotsResponse{Snapshots: k.eng.Snapshots()}, nil
}

// Checker is the in-process scheduler.SafetyChecker (types.go:29-31): same field mapping as SafetyClient.Check
// (safety_client.go:80-95), no gRPC hop.
type Checker struct{ k *SafetyKernel }

func NewChecker(k *SafetyKernel) Checker { return Checker{k: k} }

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
	return recordFromResponse(resp), nil
}

func recordFromResponse(resp *pb.PolicyCheckResponse) scheduler.SafetyDecisionRecord { // safety_client.go:103-114
	return scheduler.SafetyDecisionRecord{
		Decision: decisionFromProto(resp.GetDecision()), Reason: resp.GetReason(), RuleID: resp.GetRuleId(),
		PolicySnapshot: resp.GetPolicySnapshot(), Constraints: resp.GetConstraints(),
		ApprovalRequired: resp.GetApprovalRequired(), ApprovalRef: resp.GetApprovalRef(), Remediations: resp.GetRemediations(),
	}
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
