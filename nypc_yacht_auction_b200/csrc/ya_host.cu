// Yacht-Auction B200 engine -- host-buffer entry points: the same fused ply, called with HOST
// pointers (pinned or pageable).  Host->device and device->host copies are part of the call, so
// this is what an external (non-torch) caller of the plug-in pays end to end.
#include <cuda_runtime.h>
#include <cstdint>
#include <new>
#include "../../include/yacht_b200.h"

namespace {
struct HostCtx {
    int64_t n = 0;
    int with_masks = 0;
    cudaStream_t stream = nullptr;
    uint32_t* states = nullptr;
    int8_t* players = nullptr;
    int32_t* ply = nullptr;
    uint32_t* episode = nullptr;
    int32_t* actions = nullptr;
    float* outcome = nullptr;
    uint8_t* masks = nullptr;
    int32_t* err = nullptr;
};

template <typename T>
cudaError_t dev_alloc(T** p, size_t count) { return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)); }
}  // namespace

extern "C" {

int ya_host_create(int64_t n, int with_masks, void** handle) {
    if (n <= 0 || !handle) return (int)cudaErrorInvalidValue;
    HostCtx* c = new (std::nothrow) HostCtx();
    if (!c) return (int)cudaErrorMemoryAllocation;
    c->n = n;
    c->with_masks = with_masks;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = dev_alloc(&c->states, (size_t)n * 8);
    if (e == cudaSuccess) e = dev_alloc(&c->players, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(&c->ply, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(&c->episode, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(&c->actions, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(&c->outcome, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(&c->err, 1);
    if (e == cudaSuccess && with_masks) e = dev_alloc(&c->masks, (size_t)n * YA_ACTION_SIZE);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->err, 0, sizeof(int32_t), c->stream);
    if (e != cudaSuccess) { ya_host_destroy(c); return (int)e; }
    *handle = c;
    return 0;
}

int ya_host_destroy(void* handle) {
    HostCtx* c = static_cast<HostCtx*>(handle);
    if (!c) return 0;
    cudaFree(c->states); cudaFree(c->players); cudaFree(c->ply); cudaFree(c->episode);
    cudaFree(c->actions); cudaFree(c->outcome); cudaFree(c->masks); cudaFree(c->err);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int ya_host_play_ply(void* handle, uint32_t* states, int8_t* players, int32_t* ply, uint32_t* episode,
                     int32_t* actions, float* outcome, uint8_t* masks, int32_t* err_flag,
                     uint64_t seed, uint64_t game_base, int auto_reset) {
    HostCtx* c = static_cast<HostCtx*>(handle);
    if (!c) return (int)cudaErrorInvalidValue;
    const size_t n = (size_t)c->n;
    cudaStream_t s = c->stream;
    cudaError_t e;
#define YA_TRY(x) do { e = (x); if (e != cudaSuccess) return (int)e; } while (0)
    YA_TRY(cudaMemcpyAsync(c->states, states, n * 32, cudaMemcpyHostToDevice, s));
    YA_TRY(cudaMemcpyAsync(c->players, players, n, cudaMemcpyHostToDevice, s));
    YA_TRY(cudaMemcpyAsync(c->ply, ply, n * 4, cudaMemcpyHostToDevice, s));
    YA_TRY(cudaMemcpyAsync(c->episode, episode, n * 4, cudaMemcpyHostToDevice, s));
    uint8_t* dmasks = (masks && c->with_masks) ? c->masks : nullptr;
    int rc = ya_play_ply(c->states, c->n, c->players, c->ply, c->episode, c->actions, c->outcome, dmasks, c->err,
                         c->n, seed, game_base, auto_reset, s);
    if (rc) return rc;
    YA_TRY(cudaMemcpyAsync(states, c->states, n * 32, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaMemcpyAsync(players, c->players, n, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaMemcpyAsync(ply, c->ply, n * 4, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaMemcpyAsync(episode, c->episode, n * 4, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaMemcpyAsync(actions, c->actions, n * 4, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaMemcpyAsync(outcome, c->outcome, n * 4, cudaMemcpyDeviceToHost, s));
    if (dmasks) YA_TRY(cudaMemcpyAsync(masks, dmasks, n * YA_ACTION_SIZE, cudaMemcpyDeviceToHost, s));
    if (err_flag) YA_TRY(cudaMemcpyAsync(err_flag, c->err, 4, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaStreamSynchronize(s));
#undef YA_TRY
    return 0;
}

}  // extern "C"
