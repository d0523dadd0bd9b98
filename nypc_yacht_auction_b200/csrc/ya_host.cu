// Yacht-Auction B200 engine -- host-buffer entry points: the same fused ply, called with HOST
// pointers (pinned or pageable).  Host->device and device->host copies are part of the call, so
// this is what an external (non-torch) caller of the plug-in pays end to end.
#include <cuda_runtime.h>
#include <cstdint>
#include <new>
#include "../../include/yacht_b200.h"

namespace {
constexpr int kStreams = 4;         // copy/compute pipeline depth of the record path (full-duplex PCIe)

struct HostCtx {
    int64_t n = 0;
    int with_masks = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t lanes[kStreams] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t* records = nullptr;
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
    for (int i = 0; i < kStreams && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&c->lanes[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = dev_alloc(&c->states, (size_t)n * 8);
    if (e == cudaSuccess) e = dev_alloc(&c->players, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(&c->ply, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(&c->episode, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(&c->actions, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(&c->outcome, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(&c->err, 1);
    if (e == cudaSuccess) e = dev_alloc(&c->records, (size_t)n * 16);
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
    cudaFree(c->actions); cudaFree(c->outcome); cudaFree(c->masks); cudaFree(c->err); cudaFree(c->records);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (int i = 0; i < kStreams; ++i) if (c->lanes[i]) cudaStreamDestroy(c->lanes[i]);
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
    // masks == NULL: the mask is still materialised in HBM when the context owns a mask buffer
    int rc = ya_play_ply(c->states, c->n, c->players, c->ply, c->episode, c->actions, c->outcome,
                         c->with_masks ? c->masks : nullptr, c->err, c->n, seed, game_base, auto_reset, s);
    if (rc) return rc;
    YA_TRY(cudaMemcpyAsync(states, c->states, n * 32, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaMemcpyAsync(players, c->players, n, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaMemcpyAsync(ply, c->ply, n * 4, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaMemcpyAsync(episode, c->episode, n * 4, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaMemcpyAsync(actions, c->actions, n * 4, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaMemcpyAsync(outcome, c->outcome, n * 4, cudaMemcpyDeviceToHost, s));
    if (masks && c->with_masks) YA_TRY(cudaMemcpyAsync(masks, c->masks, n * YA_ACTION_SIZE, cudaMemcpyDeviceToHost, s));
    if (err_flag) YA_TRY(cudaMemcpyAsync(err_flag, c->err, 4, cudaMemcpyDeviceToHost, s));
    YA_TRY(cudaStreamSynchronize(s));
#undef YA_TRY
    return 0;
}

int ya_host_play_plies_records(void* handle, uint32_t* records, int plies, uint8_t* masks, int32_t* err_flag,
                               uint64_t seed, uint64_t game_base, int auto_reset) {
    HostCtx* c = static_cast<HostCtx*>(handle);
    if (!c || !records || plies < 1) return (int)cudaErrorInvalidValue;
    const int64_t n = c->n;
    cudaError_t e;
#define YA_TRY(x) do { e = (x); if (e != cudaSuccess) return (int)e; } while (0)
    // Games are independent: the batch is cut into kStreams slices (multiples of 8 games keep the mask
    // rows 16-byte aligned); each slice is one H2D copy, `plies` kernels, one D2H copy on its own stream,
    // so one slice's results travel back over the full-duplex link while the next slice travels in.
    const int64_t per = ((n + kStreams - 1) / kStreams + 7) & ~int64_t(7);
    uint8_t* dmasks = c->with_masks ? c->masks : nullptr;
    for (int k = 0; k < kStreams; ++k) {
        const int64_t g0 = k * per;
        if (g0 >= n) break;
        const int64_t m = (g0 + per <= n) ? per : n - g0;
        cudaStream_t s = c->lanes[k];
        YA_TRY(cudaMemcpyAsync(c->records + 16 * g0, records + 16 * g0, (size_t)m * 64, cudaMemcpyHostToDevice, s));
        for (int p = 0; p < plies; ++p) {
            int rc = ya_play_ply_records(c->records + 16 * g0, dmasks ? dmasks + g0 * YA_ACTION_SIZE : nullptr, c->err, m, seed,
                                         game_base + (uint64_t)g0, auto_reset, s);
            if (rc) return rc;
        }
        YA_TRY(cudaMemcpyAsync(records + 16 * g0, c->records + 16 * g0, (size_t)m * 64, cudaMemcpyDeviceToHost, s));
        if (masks && dmasks) YA_TRY(cudaMemcpyAsync(masks + g0 * YA_ACTION_SIZE, dmasks + g0 * YA_ACTION_SIZE,
                                                    (size_t)m * YA_ACTION_SIZE, cudaMemcpyDeviceToHost, s));
    }
    for (int k = 0; k < kStreams; ++k) YA_TRY(cudaStreamSynchronize(c->lanes[k]));
    if (err_flag) {
        YA_TRY(cudaMemcpyAsync(err_flag, c->err, 4, cudaMemcpyDeviceToHost, c->stream));
        YA_TRY(cudaStreamSynchronize(c->stream));
    }
#undef YA_TRY
    return 0;
}

int ya_host_play_ply_records(void* handle, uint32_t* records, uint8_t* masks, int32_t* err_flag,
                             uint64_t seed, uint64_t game_base, int auto_reset) {
    return ya_host_play_plies_records(handle, records, 1, masks, err_flag, seed, game_base, auto_reset);
}

}  // extern "C"
