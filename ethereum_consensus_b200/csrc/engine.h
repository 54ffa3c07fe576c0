// Process-global engine context: one CUDA device, one stream, grow-only device/pinned buffers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/b200_consensus.h"

namespace b200 {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    // grow-only; contents are NOT preserved
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 4096;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Engine {
    bool ready = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_copy[17] = {nullptr};  // H2D slice k done (copy stream) -> compute stream may hash slice k
    std::mutex mu;
    std::string last_error;
    uint64_t launches = 0;
    uint64_t collectives = 0;  // NCCL collectives issued by the library (comm.cu)
    float last_kernel_ms = 0.f;
    // SSZ scratch (one-shot calls)
    DevBuf arena, fields, planbuf;
    PinnedBuf staging;
    DevBuf xch_dev;        // b200_comm_all_gather_bytes scratch
    PinnedBuf xch_host;
    uint32_t* d_zero = nullptr;  // 65 zero-subtree hashes, word form
    // BLS scratch lives in bls_engine (opaque here)
    void* bls = nullptr;
};

Engine& engine();

#define B200_CUDA_TRY(expr)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            engine().last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);        \
            return B200_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

}  // namespace b200
