// Experiment: throughput of per-event 16-byte TMA bulk reductions (cp.reduce.async.bulk ... add.f32) from
// shared memory to pseudo-random global addresses, alone and mixed with LSU reductions (red.global.add.v4.f32).
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void tma_red16(float *gdst, const float *ssrc)
{
    unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 16;" ::"l"(gdst), "r"(saddr) : "memory");
}

// mode: number of TMA reductions out of every 8 operations (0 = all LSU, 8 = all TMA)
__global__ void __launch_bounds__(256) k(float *base, unsigned nslots, int per_thread, int tma_of_8, unsigned seed)
{
    __shared__ __align__(16) float stage[256 * 4 * 4];   // 4 staging slots per thread
    unsigned s = seed ^ (blockIdx.x * 2654435761u) ^ (threadIdx.x * 40503u);
    int slot = 0;
    for (int i = 0; i < per_thread; ++i) {
        s = s * 1664525u + 1013904223u;
        float *dst = base + (size_t)((s >> 4) % nslots) * 4;
        const bool use_tma = (i & 7) < tma_of_8;
        if (use_tma) {
            float *st = stage + (threadIdx.x * 4 + slot) * 4;
            st[0] = 1.0f; st[1] = 0.5f; st[2] = 0.0f; st[3] = 0.0f;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tma_red16(dst, st);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            slot = (slot + 1) & 3;
            if (slot == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging slots free again
        } else {
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(1.0f), "f"(0.5f), "f"(0.0f), "f"(0.0f));
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
    const unsigned nslots = 9830400 / 16;   // 9.8 MB of 16-byte slots
    float *buf;
    cudaMalloc(&buf, (size_t)nslots * 16);
    cudaMemset(buf, 0, (size_t)nslots * 16);
    const int grid = 148 * 8, per = 50000000 / (grid * 256) + 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep)
        for (int m : {0, 1, 2, 4, 8}) {
            cudaEventRecord(e0);
            k<<<grid, 256>>>(buf, nslots, per, m, 77u + rep);
            cudaEventRecord(e1);
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("TMA %d/8 of the reductions: %.3f ms for %.1f M 16-byte reductions (%.1f G/s)\n", m, ms, (double)grid * 256 * per / 1e6,
                   (double)grid * 256 * per / ms / 1e6);
        }
    // sanity: total mass
    return 0;
}
