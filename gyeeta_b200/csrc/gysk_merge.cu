// gysk_merge.cu — multi-GPU merge step. Filled in after the single-GPU path is parity-green.
#include "gysk_kernels.cuh"

extern "C" {

int gysk_set_logical_map(gysk_engine *, const uint64_t *, const uint64_t *, uint32_t) { return GYSK_ERR_NOTSUP; }
int gysk_merge_prepare(gysk_engine *) { return GYSK_ERR_NOTSUP; }
int gysk_merge_buffers(gysk_engine *, gysk_buffer_desc *, uint32_t, uint32_t *) { return GYSK_ERR_NOTSUP; }
int gysk_merge_tdigest_slab(gysk_engine *, void **, uint64_t *) { return GYSK_ERR_NOTSUP; }
int gysk_merge_finish(gysk_engine *, const void *, uint32_t) { return GYSK_ERR_NOTSUP; }
int gysk_query_logical(gysk_engine *, const uint64_t *, uint32_t, gysk_svc_summary *) { return GYSK_ERR_NOTSUP; }

}
