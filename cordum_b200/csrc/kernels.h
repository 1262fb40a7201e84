// kernels.h — kernel parameter block and launchers (kernels.cu <-> engine.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/cordum_b200.h"
#include "tables.h"

struct KParams {
  JobRecords recs;        // device pointers to the encoded (topic-sorted) records of this batch
  DeviceTables t;         // device pointers to the compiled tables
  cordum_decision* out;   // device, n_jobs records
  uint32_t n_jobs;
  uint32_t* route_count;     // device: number of jobs policy_kernel found dispatchable (null: no compaction)
  uint2* route_list;         // device: {slot in the sorted records, head word policy_kernel wrote} per dispatchable job
  uint32_t honor_approved;   // POLICY_AND_ROUTE: jobs flagged JF_APPROVED bypass the policy (engine.go:484-522)
};

// loads_read (optional) is recorded once T.loads has been consumed (between the chunk and the merge kernel)
cudaError_t launch_worker_pools(const DeviceTables& T, cudaStream_t s, cudaEvent_t loads_read);
cudaError_t launch_policy(const KParams& P, int sm_count, cudaStream_t s);
cudaError_t launch_route(const KParams& P, bool route_only, int sm_count, cudaStream_t s);
