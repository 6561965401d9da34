// evk_host.cu -- host-buffer pipeline: voxel build for events that live in HOST memory.
//
// The reference's consumers hand over host arrays (numpy / CPU tensors: data loaders
// lib/data_loaders/base_dataset.py:446-453).  Moving 16 B/event over PCIe costs far more than the
// scatter itself, so the copy is chunked and double-buffered: while chunk k is scattered on the
// compute stream, chunk k+1 is in flight on the copy stream.  The accumulation grid stays on the
// device (L2 resident) for the whole call and is read back once.
#include <stdlib.h>

#include "evk_common.cuh"

struct evk_pipeline {
    int64_t chunk;            // events per chunk
    float *stage[2][4];       // device staging: x,y,t,p per slot
    cudaStream_t copy, compute;
    cudaEvent_t copied[2], consumed[2];
    float *grid;              // device output grid
    size_t grid_bytes;
    void *ws;                 // voxel workspace
    size_t ws_bytes;
    unsigned long long *oob_dev;
    unsigned long long *oob_pinned;
};

extern "C" {

int evk_pipeline_create(evk_pipeline_t **out, int64_t chunk_events)
{
    using namespace evk;
    if (!out) { set_error("evk_pipeline_create: null out"); return EVK_E_ARG; }
    if (chunk_events <= 0) chunk_events = (int64_t)4 << 20;
    chunk_events = (chunk_events + 3) & ~(int64_t)3;
    evk_pipeline *p = (evk_pipeline *)calloc(1, sizeof(evk_pipeline));
    if (!p) { set_error("evk_pipeline_create: out of host memory"); return EVK_E_ARG; }
    p->chunk = chunk_events;
    for (int s = 0; s < 2; ++s)
        for (int a = 0; a < 4; ++a)   // slot 2 (timestamps) is 8 bytes wide so that it can also stage float64 stamps
            EVK_CUDA(cudaMalloc(&p->stage[s][a], (size_t)chunk_events * (a == 2 ? sizeof(double) : sizeof(float))));
    EVK_CUDA(cudaStreamCreateWithFlags(&p->copy, cudaStreamNonBlocking));
    EVK_CUDA(cudaStreamCreateWithFlags(&p->compute, cudaStreamNonBlocking));
    for (int s = 0; s < 2; ++s) {
        EVK_CUDA(cudaEventCreateWithFlags(&p->copied[s], cudaEventDisableTiming));
        EVK_CUDA(cudaEventCreateWithFlags(&p->consumed[s], cudaEventDisableTiming));
    }
    EVK_CUDA(cudaMalloc(&p->oob_dev, sizeof(unsigned long long)));
    EVK_CUDA(cudaMallocHost(&p->oob_pinned, sizeof(unsigned long long)));
    *out = p;
    return EVK_OK;
}

void evk_pipeline_destroy(evk_pipeline_t *p)
{
    if (!p) return;
    for (int s = 0; s < 2; ++s) {
        for (int a = 0; a < 4; ++a) cudaFree(p->stage[s][a]);
        if (p->copied[s]) cudaEventDestroy(p->copied[s]);
        if (p->consumed[s]) cudaEventDestroy(p->consumed[s]);
    }
    if (p->copy) cudaStreamDestroy(p->copy);
    if (p->compute) cudaStreamDestroy(p->compute);
    cudaFree(p->grid);
    cudaFree(p->ws);
    cudaFree(p->oob_dev);
    cudaFreeHost(p->oob_pinned);
    free(p);
}

int evk_voxel_host_f32(evk_pipeline_t *p, const float *x, const float *y, const float *t, const float *pol,
                       int64_t n, float t0, float dt, int B, int H, int W, unsigned flags, float *out_host,
                       unsigned long long *oob_host)
{
    using namespace evk;
    if (!p || !out_host || n < 0 || B < 1 || H < 1 || W < 1 || (n > 0 && (!x || !y || !t || !pol))) {
        set_error("evk_voxel_host_f32: bad arguments");
        return EVK_E_ARG;
    }
    const size_t grid_bytes = (size_t)B * H * W * sizeof(float);
    if (p->grid_bytes < grid_bytes) {
        cudaFree(p->grid);
        p->grid = nullptr; p->grid_bytes = 0;
        EVK_CUDA(cudaMalloc(&p->grid, grid_bytes));
        p->grid_bytes = grid_bytes;
    }
    const size_t ws_bytes = evk_voxel_workspace_bytes(B, H, W, flags);
    if (p->ws_bytes < ws_bytes) {
        cudaFree(p->ws);
        p->ws = nullptr; p->ws_bytes = 0;
        EVK_CUDA(cudaMalloc(&p->ws, ws_bytes));
        p->ws_bytes = ws_bytes;
    }
    EVK_CUDA(cudaMemsetAsync(p->oob_dev, 0, sizeof(unsigned long long), p->compute));
    EVK_CUDA(cudaMemsetAsync(p->grid, 0, grid_bytes, p->compute));
    const float *src[4] = {x, y, t, pol};
    int64_t done = 0;
    int k = 0;
    // every chunk accumulates into the same grid; chunk 0 has already been zeroed above
    const unsigned cflags = (flags & ~EVK_VARIANT_MASK) | EVK_ACCUMULATE | EVK_VARIANT_GLOBAL_RED;
    while (done < n) {
        const int s = k & 1;
        const int64_t m = (n - done < p->chunk) ? (n - done) : p->chunk;
        if (k >= 2) EVK_CUDA(cudaStreamWaitEvent(p->copy, p->consumed[s], 0));
        for (int a = 0; a < 4; ++a)
            EVK_CUDA(cudaMemcpyAsync(p->stage[s][a], src[a] + done, (size_t)m * sizeof(float), cudaMemcpyHostToDevice, p->copy));
        EVK_CUDA(cudaEventRecord(p->copied[s], p->copy));
        EVK_CUDA(cudaStreamWaitEvent(p->compute, p->copied[s], 0));
        int rc = evk_voxel_f32(p->stage[s][0], p->stage[s][1], p->stage[s][2], p->stage[s][3], m, t0, dt, B, H, W,
                               cflags, p->grid, nullptr, 0, p->oob_dev, p->compute);
        if (rc) return rc;
        EVK_CUDA(cudaEventRecord(p->consumed[s], p->compute));
        done += m;
        ++k;
    }
    EVK_CUDA(cudaMemcpyAsync(out_host, p->grid, grid_bytes, cudaMemcpyDefault, p->compute));  // out may be host or device
    EVK_CUDA(cudaMemcpyAsync(p->oob_pinned, p->oob_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->compute));
    EVK_CUDA(cudaStreamSynchronize(p->compute));
    if (oob_host) *oob_host = *p->oob_pinned;
    return EVK_OK;
}

// Same pipeline for the storage layout (int16 x, int16 y, float64 t, uint8 p): 13 B/event cross PCIe
// instead of 16, and the casts happen on the device (evk_voxel_packed_f32).
int evk_voxel_host_packed_f32(evk_pipeline_t *p, const int16_t *x, const int16_t *y, const double *t, const uint8_t *pol,
                              int64_t n, double t_first, double t_last, int B, int H, int W, unsigned flags, float *out_host,
                              unsigned long long *oob_host)
{
    using namespace evk;
    if (!p || !out_host || n < 0 || B < 1 || H < 1 || W < 1 || (n > 0 && (!x || !y || !t || !pol))) {
        set_error("evk_voxel_host_packed_f32: bad arguments");
        return EVK_E_ARG;
    }
    const size_t grid_bytes = (size_t)B * H * W * sizeof(float);
    if (p->grid_bytes < grid_bytes) {
        cudaFree(p->grid);
        p->grid = nullptr; p->grid_bytes = 0;
        EVK_CUDA(cudaMalloc(&p->grid, grid_bytes));
        p->grid_bytes = grid_bytes;
    }
    EVK_CUDA(cudaMemsetAsync(p->oob_dev, 0, sizeof(unsigned long long), p->compute));
    EVK_CUDA(cudaMemsetAsync(p->grid, 0, grid_bytes, p->compute));
    const unsigned cflags = (flags & ~(EVK_VARIANT_MASK | EVK_AUTO_SPAN)) | EVK_ACCUMULATE | EVK_VARIANT_GLOBAL_RED;
    int64_t done = 0;
    int k = 0;
    while (done < n) {
        const int s = k & 1;
        const int64_t m = (n - done < p->chunk) ? (n - done) : p->chunk;
        if (k >= 2) EVK_CUDA(cudaStreamWaitEvent(p->copy, p->consumed[s], 0));
        EVK_CUDA(cudaMemcpyAsync(p->stage[s][0], x + done, (size_t)m * sizeof(int16_t), cudaMemcpyHostToDevice, p->copy));
        EVK_CUDA(cudaMemcpyAsync(p->stage[s][1], y + done, (size_t)m * sizeof(int16_t), cudaMemcpyHostToDevice, p->copy));
        EVK_CUDA(cudaMemcpyAsync(p->stage[s][2], t + done, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, p->copy));
        EVK_CUDA(cudaMemcpyAsync(p->stage[s][3], pol + done, (size_t)m * sizeof(uint8_t), cudaMemcpyHostToDevice, p->copy));
        EVK_CUDA(cudaEventRecord(p->copied[s], p->copy));
        EVK_CUDA(cudaStreamWaitEvent(p->compute, p->copied[s], 0));
        int rc = evk_voxel_packed_f32((const int16_t *)p->stage[s][0], (const int16_t *)p->stage[s][1], (const double *)p->stage[s][2],
                                      (const uint8_t *)p->stage[s][3], m, t_first, t_last, B, H, W, cflags, p->grid, nullptr, 0,
                                      p->oob_dev, p->compute);
        if (rc) return rc;
        EVK_CUDA(cudaEventRecord(p->consumed[s], p->compute));
        done += m;
        ++k;
    }
    EVK_CUDA(cudaMemcpyAsync(out_host, p->grid, grid_bytes, cudaMemcpyDefault, p->compute));
    EVK_CUDA(cudaMemcpyAsync(p->oob_pinned, p->oob_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->compute));
    EVK_CUDA(cudaStreamSynchronize(p->compute));
    if (oob_host) *oob_host = *p->oob_pinned;
    return EVK_OK;
}

}  // extern "C"
