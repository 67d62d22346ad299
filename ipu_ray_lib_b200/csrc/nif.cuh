// Interface of the NIF (neural image field) environment-light evaluator, see nif.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200rt.h"

namespace rt {

struct NifModel;

NifModel* nif_create(const b200rt_nif_desc& desc, int device);
void nif_destroy(NifModel* m);
const char* nif_last_error();

// n (u,v) pairs -> n bgr triples. Device pointers.
int nif_eval_uv(NifModel* m, const float* dUv, uint32_t n, float* dBgrOut, cudaStream_t stream, int* launches);

// Wavefront form used by the path tracer: queue[0..*dCount) holds slot indices of escaped samples;
// (u,v) is read from slotEscape[5*slot+3..4], bgr is written to slotEnv[3*slot..]. The count stays on
// the device (no host round trip); maxCount bounds it. maxBatch > 0 evaluates the queue in serial launches of at most
// that many rays (IpuScene::setMaxNifBatchSize).
int nif_eval_queue(NifModel* m, const float* slotEscape, const uint32_t* queue, const uint32_t* dCount,
                   uint32_t maxCount, uint32_t maxBatch, float* slotEnv, cudaStream_t stream, int* launches);

}  // namespace rt
