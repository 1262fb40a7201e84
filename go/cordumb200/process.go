package cordumb200

/*
#include "cordum_b200.h"
*/
import "C"

import (
	"github.com/cordum/cordum/core/controlplane/scheduler"
	"github.com/cordum/cordum/core/infra/config"
	pb "github.com/cordum/cordum/core/protocol/pb/v1"
)

// Scheduled is what Engine.processJob (engine.go:294-347,393) needs from the two calls it makes per job today -
// checkSafetyDecision (:484-531, which calls SafetyChecker.Check) and, for jobs that may dispatch,
// SchedulingStrategy.PickSubject (:392-393) - delivered by ONE blocking submit on the policy-and-route front-end.  The
// handler keeps its state machine, locks, job-store writes and publishes; only where the decision and the subject come
// from changes.  Concurrent handlers (JetStream delivers up to MaxAckPending messages, bus/nats.go:174) are batched by
// the library.
type Scheduled struct {
	Record   scheduler.SafetyDecisionRecord
	Subject  string // set when the decision lets the job dispatch and a worker was picked
	RouteErr error  // PickSubject's error for a dispatchable job (wraps ErrNoPoolMapping / ErrNoWorkers / ErrPoolOverloaded)
}

// ProcessJob evaluates and routes one job.  approved = the caller verified the stored approval record and its job hash
// (engine.go:484-522: label approval_granted, prev.JobHash == HashJobRequest(req)); the engine then returns
// {ALLOW, "approval granted"} without consulting the policy, as the reference does, and routes the job.
func (e *Engine) ProcessJob(req *pb.JobRequest, approved bool) (Scheduled, error) {
	var eff []byte
	if env := req.GetEnv(); env != nil {
		if v := env[config.EffectiveConfigEnvVar]; v != "" {
			eff = []byte(v)
		}
	}
	creq, free := packRequest(req.GetTopic(), scheduler.ExtractTenant(req), req.GetPrincipalId(), req.GetLabels(), req.GetMeta(), eff, approved)
	defer free()
	r, freeResp, err := submit(e.schedFE, creq)
	if err != nil { // fail closed, as SafetyClient does on a transport error (safety_client.go:98-101)
		return Scheduled{Record: scheduler.SafetyDecisionRecord{Decision: scheduler.SafetyDeny, Reason: "safety kernel error: " + err.Error()}}, nil
	}
	defer freeResp()
	out := Scheduled{Record: recordFromResponse(e.policyResponse(r, req.GetJobId()))}
	// engine.go:528-530: an ALLOW that still needs approval is held; the record's sched_decision carries that post-step
	if r.rec.sched_decision == C.CORDUM_DEC_REQUIRE_HUMAN {
		out.Record.Decision = scheduler.SafetyRequireApproval
	}
	if r.rec.route_status != C.CORDUM_ROUTE_NOT_ATTEMPTED {
		out.Subject, out.RouteErr = subjectFromResponse(r, req)
	}
	return out, nil
}
