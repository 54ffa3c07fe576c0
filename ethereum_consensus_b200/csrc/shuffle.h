// Device-side building blocks of shuffle.cu used by the resident-state entry point in capi_ssz.cu.
#pragma once
#include "engine.h"

namespace b200 {
int32_t shuffle_on_device(Engine& e, const uint64_t* idx_dev, uint64_t n, const uint8_t seed[32], uint32_t rounds, uint64_t* out_dev);
int32_t active_indices_on_device(Engine& e, const uint8_t* recs_dev, uint64_t n, uint64_t epoch, uint64_t* out_dev, uint64_t* out_n);
int32_t shuffle_scratch(Engine& e, uint64_t n, uint64_t** a, uint64_t** b);
}  // namespace b200
