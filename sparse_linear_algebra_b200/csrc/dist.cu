// dist.cu — multi-GPU plumbing (one process per GPU; NCCL over NVLink).  Filled in by the row-partitioned path.
#include "common.cuh"

sla_status sla_dist_attach(sla_ctx* c, const void* nccl_id128) {
  (void)nccl_id128;
  return sla_fail(c, SLA_ERR_COMM, "sla_init_dist: multi-GPU support is not built into this library yet");
}

void sla_dist_detach(sla_ctx* c) { (void)c; }

extern "C" sla_status sla_nccl_unique_id(void* out128) {
  (void)out128;
  return SLA_ERR_COMM;
}
