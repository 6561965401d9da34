// evk_host.cu -- host-buffer pipeline: voxel build for events that live in HOST memory.
//
// The reference's consumers hand over host arrays (numpy / CPU tensors: data loaders
// lib/data_loaders/base_dataset.py:446-453).  Moving 16 B/event over PCIe costs far more than the
// scatter itself, so the copy is chunked and double-buffered: while chunk k is scattered on the
// compute stream, chunk k+1 is in flight on the copy stream.  The accumulation grid stays on the
// device (L2 resident) for the whole call and is read back once.
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "evk_common.cuh"

// ---- content hash of a host buffer (identity of a cached event set) ---------------------------------
// Four independent multiply-rotate lanes over 32-byte stripes (the structure of XXH64), merged and
// avalanched; the tail is zero-padded into one more stripe and the length is mixed in.  Every byte of
// the buffer enters the result.  Large buffers are cut into equal pieces hashed on separate threads and
// the piece hashes are chained in order, so the value does not depend on the thread count.
namespace {
constexpr uint64_t kP1 = 0x9E3779B185EBCA87ull, kP2 = 0xC2B2AE3D27D4EB4Full, kP3 = 0x165667B19E3779F9ull,
                   kP4 = 0x85EBCA77C2B2AE63ull, kP5 = 0x27D4EB2F165667C5ull;
inline uint64_t rotl64(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }
inline uint64_t lane_round(uint64_t acc, uint64_t in) { return rotl64(acc + in * kP2, 31) * kP1; }
inline uint64_t lane_merge(uint64_t h, uint64_t v) { return (h ^ lane_round(0, v)) * kP1 + kP4; }

uint64_t hash_piece(const unsigned char *p, size_t n, uint64_t seed)
{
    uint64_t v0 = seed + kP1 + kP2, v1 = seed + kP2, v2 = seed, v3 = seed - kP1;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        uint64_t w[4];
        memcpy(w, p + i, 32);
        v0 = lane_round(v0, w[0]); v1 = lane_round(v1, w[1]); v2 = lane_round(v2, w[2]); v3 = lane_round(v3, w[3]);
    }
    if (i < n) {
        uint64_t w[4] = {0, 0, 0, 0};
        memcpy(w, p + i, n - i);
        v0 = lane_round(v0, w[0]); v1 = lane_round(v1, w[1]); v2 = lane_round(v2, w[2]); v3 = lane_round(v3, w[3]);
    }
    uint64_t h = rotl64(v0, 1) + rotl64(v1, 7) + rotl64(v2, 12) + rotl64(v3, 18);
    h = lane_merge(h, v0); h = lane_merge(h, v1); h = lane_merge(h, v2); h = lane_merge(h, v3);
    h += (uint64_t)n * kP5;
    h ^= h >> 33; h *= kP2; h ^= h >> 29; h *= kP3; h ^= h >> 32;
    return h;
}
}  // namespace

// hash k buffers at once: all 1 MiB pieces of all buffers are spread over one set of threads
static void hash_many(const void *const *ptrs, const size_t *nbytes, int k, const uint64_t *seeds, uint64_t *out)
{
    constexpr size_t kPiece = (size_t)1 << 20;               // fixed piece size: the result is thread-count independent
    struct Job { const unsigned char *p; size_t len; uint64_t seed; uint64_t *dst; };
    std::vector<Job> jobs;
    std::vector<std::vector<uint64_t>> piece_hash(k);
    for (int a = 0; a < k; ++a) {
        const unsigned char *p = (const unsigned char *)ptrs[a];
        const size_t n = p ? nbytes[a] : 0;
        const size_t pieces = n ? (n + kPiece - 1) / kPiece : 1;
        piece_hash[a].resize(pieces);
        for (size_t j = 0; j < pieces; ++j) {
            const size_t off = j * kPiece, len = n ? ((n - off < kPiece) ? n - off : kPiece) : 0;
            jobs.push_back({n ? p + off : (const unsigned char *)"", len, seeds[a] + j, &piece_hash[a][j]});
        }
    }
    unsigned hw = std::thread::hardware_concurrency();
    size_t nthreads = hw ? hw : 4;
    if (nthreads > 8) nthreads = 8;
    if (nthreads > jobs.size() / 2) nthreads = jobs.size() / 2 ? jobs.size() / 2 : 1;
    auto work = [&](size_t first, size_t step) {
        for (size_t j = first; j < jobs.size(); j += step) *jobs[j].dst = hash_piece(jobs[j].p, jobs[j].len, jobs[j].seed);
    };
    if (nthreads <= 1) work(0, 1);
    else {
        std::vector<std::thread> th;
        for (size_t t = 1; t < nthreads; ++t) th.emplace_back(work, t, nthreads);
        work(0, nthreads);
        for (auto &t : th) t.join();
    }
    for (int a = 0; a < k; ++a) {
        const std::vector<uint64_t> &h = piece_hash[a];
        out[a] = (h.size() == 1) ? h[0] : hash_piece((const unsigned char *)h.data(), h.size() * sizeof(uint64_t), seeds[a] ^ (uint64_t)nbytes[a]);
    }
}

extern "C" uint64_t evk_host_hash64(const void *data, size_t nbytes, uint64_t seed)
{
    uint64_t out = 0;
    hash_many(&data, &nbytes, 1, &seed, &out);
    return out;
}

extern "C" void evk_host_hash64_multi(const void *const *ptrs, const size_t *nbytes, int k, uint64_t seed, uint64_t *out)
{
    if (k <= 0 || !ptrs || !nbytes || !out) return;
    std::vector<uint64_t> seeds(k);
    for (int a = 0; a < k; ++a) seeds[a] = seed + 0x9E3779B97F4A7C15ull * (uint64_t)a;
    hash_many(ptrs, nbytes, k, seeds.data(), out);
}

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
    // pageable sources: pinned bounce buffers filled by host threads (cudaMemcpyAsync from pageable memory is a
    // single-threaded staged copy inside the driver, far below PCIe speed)
    float *bounce[2][4];
    int64_t bounce_events;    // capacity of one bounce slot
};

namespace {
// is `ptr` ordinary pageable host memory (not pinned / registered / managed / device)?
bool is_pageable(const void *ptr)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// copy k arrays of `bytes` each with a few host threads (a single memcpy stream tops out well below PCIe 5 x16)
void parallel_copy(void *const *dst, const void *const *src, int k, size_t bytes)
{
    // a quarter of the hardware threads, at most 24 (a memcpy thread moves ~5-8 GB/s; PCIe 5 x16 takes ~55 GB/s; the process
    // may be bound to one NUMA node); EVK_HOST_COPY_THREADS overrides
    static int nthreads = 0;
    if (nthreads == 0) {
        const char *e = getenv("EVK_HOST_COPY_THREADS");
        unsigned hw = std::thread::hardware_concurrency();
        int v = (e && *e) ? atoi(e) : (hw ? (int)(hw / 4) : 4);
        nthreads = v < 1 ? 1 : (v > 24 ? 24 : v);
    }
    const size_t total = bytes * k;
    if (total < ((size_t)4 << 20) || nthreads <= 1) {
        for (int a = 0; a < k; ++a) memcpy(dst[a], src[a], bytes);
        return;
    }
    constexpr size_t kBlock = (size_t)1 << 20;
    const size_t blocks_per = (bytes + kBlock - 1) / kBlock, nblocks = blocks_per * k;
    auto work = [&](size_t first) {
        for (size_t b = first; b < nblocks; b += (size_t)nthreads) {
            const int a = (int)(b / blocks_per);
            const size_t off = (b % blocks_per) * kBlock, len = (bytes - off < kBlock) ? bytes - off : kBlock;
            memcpy((char *)dst[a] + off, (const char *)src[a] + off, len);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, (size_t)t);
    work(0);
    for (auto &t : th) t.join();
}
}  // namespace

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
    for (int s = 0; s < 2; ++s)
        for (int a = 0; a < 4; ++a) cudaFreeHost(p->bounce[s][a]);
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
    // ordinary (pageable) host arrays -- what the reference's callers hand over -- go through pinned bounce buffers that
    // a few host threads fill while the previous chunk is on the wire
    const bool bounce = n > 0 && (is_pageable(x) || is_pageable(y) || is_pageable(t) || is_pageable(pol));
    if (bounce && p->bounce_events < p->chunk) {
        for (int s = 0; s < 2; ++s)
            for (int a = 0; a < 4; ++a) {
                cudaFreeHost(p->bounce[s][a]);
                p->bounce[s][a] = nullptr;
                EVK_CUDA(cudaMallocHost(&p->bounce[s][a], (size_t)p->chunk * sizeof(float)));
            }
        p->bounce_events = p->chunk;
    }
    // every chunk accumulates into the same grid; chunk 0 has already been zeroed above
    const unsigned cflags = (flags & ~EVK_VARIANT_MASK) | EVK_ACCUMULATE | EVK_VARIANT_GLOBAL_RED;
    while (done < n) {
        const int s = k & 1;
        const int64_t m = (n - done < p->chunk) ? (n - done) : p->chunk;
        const float *from[4] = {src[0] + done, src[1] + done, src[2] + done, src[3] + done};
        if (bounce) {
            if (k >= 2) EVK_CUDA(cudaEventSynchronize(p->copied[s]));      // the bounce slot's previous H2D has left it
            void *dst4[4] = {p->bounce[s][0], p->bounce[s][1], p->bounce[s][2], p->bounce[s][3]};
            const void *src4[4] = {from[0], from[1], from[2], from[3]};
            parallel_copy(dst4, src4, 4, (size_t)m * sizeof(float));
            for (int a = 0; a < 4; ++a) from[a] = p->bounce[s][a];
        }
        if (k >= 2) EVK_CUDA(cudaStreamWaitEvent(p->copy, p->consumed[s], 0));
        for (int a = 0; a < 4; ++a)
            EVK_CUDA(cudaMemcpyAsync(p->stage[s][a], from[a], (size_t)m * sizeof(float), cudaMemcpyHostToDevice, p->copy));
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
