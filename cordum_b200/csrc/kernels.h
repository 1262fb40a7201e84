// kernels.h — kernel parameter block and launchers (kernels.cu <-> engine.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/cordum_b200.h"
#include "tables.h"

struct KParams {
  JobColumns cols;        // device pointers to the encoded columns of this batch
  DeviceTables t;         // device pointers to the compiled tables
  cordum_decision* out;   // device, n_jobs records
  uint32_t n_jobs;
};

cudaError_t launch_worker_pools(const DeviceTables& T, cudaStream_t s);
cudaError_t launch_dispatch(const KParams& P, uint32_t mode, int sm_count, cudaStream_t s);
