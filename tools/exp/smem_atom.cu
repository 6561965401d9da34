// Experiment (round 2): can on-chip (shared / distributed shared memory) accumulators beat the L2 vector
// reductions?  Measures, on the real access pattern (one accumulator update per event, uniformly random
// cell, events streamed from HBM with 16-byte evict-first loads):
//   A. local shared-memory atomics: u32 / u64 / f32 (CAS) at spread addresses, 1 CTA per SM, ~200 KB table
//   B. remote DSMEM atomics (red.shared::cluster) for cluster sizes 2/4/8/16
//   C. hybrid: a fraction of the events goes to the cluster's DSMEM table, the rest to L2 red.v4.f32
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_atom smem_atom.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_random(uint4 *a, size_t n4, unsigned seed)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        unsigned s = seed ^ (unsigned)(i * 2654435761ull);
        uint4 v;
        s = s * 1664525u + 1013904223u; s ^= s >> 15; s *= 2246822519u; s ^= s >> 13; v.x = s;
        s = s * 1664525u + 1013904223u; s ^= s >> 15; s *= 2246822519u; s ^= s >> 13; v.y = s;
        s = s * 1664525u + 1013904223u; s ^= s >> 15; s *= 2246822519u; s ^= s >> 13; v.z = s;
        s = s * 1664525u + 1013904223u; s ^= s >> 15; s *= 2246822519u; s ^= s >> 13; v.w = s;
        a[i] = v;
    }
}

enum { OP_U32 = 0, OP_U64 = 1, OP_F32 = 2, OP_U64_RET = 3, OP_NONE = 4, OP_U32X2 = 5, OP_U32X4 = 6, OP_F32X4 = 7 };

__device__ __forceinline__ void red_add4(float *a, float4 v)
{
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ unsigned mapa(unsigned saddr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}

template <int OP, bool LOCAL>
__device__ __forceinline__ void smem_update(unsigned caddr, unsigned r, unsigned long long &sink)
{
    if (LOCAL) {   // shared::cta window: ATOMS.*
        if (OP == OP_U32) asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0], %1;" ::"r"(caddr), "r"(r & 0xffu));
        else if (OP == OP_U32X2) {
            asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0], %1;" ::"r"(caddr), "r"(r & 0xffu));
            asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0+4], %1;" ::"r"(caddr), "r"((r >> 8) & 0xffu));
        }
        else if (OP == OP_U32X4) {   // bilinear footprint in a 241-wide image
            asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0], %1;" ::"r"(caddr), "r"(r & 0xffu));
            asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0+4], %1;" ::"r"(caddr), "r"((r >> 8) & 0xffu));
            asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0+964], %1;" ::"r"(caddr), "r"((r >> 16) & 0xffu));
            asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0+968], %1;" ::"r"(caddr), "r"((r >> 24) & 0xffu));
        }
        else if (OP == OP_F32X4) {
            const float w = __uint_as_float((r & 0x007fffffu) | 0x3f000000u);
            asm volatile("red.relaxed.cta.shared::cta.add.f32 [%0], %1;" ::"r"(caddr), "f"(w));
            asm volatile("red.relaxed.cta.shared::cta.add.f32 [%0+4], %1;" ::"r"(caddr), "f"(1.f - w));
            asm volatile("red.relaxed.cta.shared::cta.add.f32 [%0+964], %1;" ::"r"(caddr), "f"(w * 0.5f));
            asm volatile("red.relaxed.cta.shared::cta.add.f32 [%0+968], %1;" ::"r"(caddr), "f"(0.5f - w * 0.5f));
        }
        else if (OP == OP_U64) asm volatile("red.relaxed.cta.shared::cta.add.u64 [%0], %1;" ::"r"(caddr), "l"((unsigned long long)(r & 0xffffu) | ((unsigned long long)(r >> 20) << 32)));
        else if (OP == OP_F32) asm volatile("red.relaxed.cta.shared::cta.add.f32 [%0], %1;" ::"r"(caddr), "f"(__uint_as_float((r & 0x007fffffu) | 0x3f000000u)));
        else if (OP == OP_U64_RET) {
            unsigned old;
            asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(caddr), "r"(r & 0xffffu));
            sink += old;
        }
        return;
    }
    if (OP == OP_U32X2) {
        asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(caddr), "r"(r & 0xffu));
        asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0+4], %1;" ::"r"(caddr), "r"((r >> 8) & 0xffu));
    }
    if (OP == OP_U32) asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(caddr), "r"(r & 0xffu));
    else if (OP == OP_U64) asm volatile("red.relaxed.cluster.shared::cluster.add.u64 [%0], %1;" ::"r"(caddr), "l"((unsigned long long)(r & 0xffffu) | ((unsigned long long)(r >> 20) << 32)));
    else if (OP == OP_F32) asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" ::"r"(caddr), "f"(__uint_as_float((r & 0x007fffffu) | 0x3f000000u)));
    else if (OP == OP_U64_RET) {
        unsigned long long old;
        asm volatile("atom.relaxed.cluster.shared::cluster.add.u64 %0, [%1], %2;" : "=l"(old) : "r"(caddr), "l"((unsigned long long)(r & 0xffffu)));
        sink += old;
    }
}

// Every event: r = random 32 bits read from HBM (4 arrays x 16-byte loads = 16 B/event like the voxel kernel).
// pix = r % npix.  pix < covered -> DSMEM table of the cluster (rank = pix / per_cta), else L2 v4 RED.
// covered == npix : everything on chip; covered == 0 : everything through L2.
template <int OP, int CLUSTER>
__global__ void __launch_bounds__(1024, 1)
hybrid(const uint4 *__restrict__ a0, const uint4 *__restrict__ a1, const uint4 *__restrict__ a2, const uint4 *__restrict__ a3,
       size_t n4, unsigned npix, unsigned covered, unsigned per_cta, int stride_bytes, float *l2ws, unsigned long long *out)
{
    extern __shared__ __align__(16) unsigned char table[];
    for (unsigned i = threadIdx.x; i < per_cta * (unsigned)stride_bytes / 4; i += blockDim.x) reinterpret_cast<unsigned *>(table)[i] = 0;
    unsigned crank = 0;
    if (CLUSTER > 1) { cg::this_cluster().sync(); crank = cg::this_cluster().block_rank(); } else __syncthreads();
    (void)crank;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(table);
    unsigned long long sink = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = __ldcs(a0 + i), y = __ldcs(a1 + i), t = __ldcs(a2 + i), p = __ldcs(a3 + i);
        const unsigned rr[4] = {x.x ^ y.x ^ t.x ^ p.x, x.y ^ y.y ^ t.y ^ p.y, x.z ^ y.z ^ t.z ^ p.z, x.w ^ y.w ^ t.w ^ p.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned r = rr[k];
            const unsigned pix = (unsigned)(((unsigned long long)r * npix) >> 32);
            if (pix < covered) {
                if (OP != OP_NONE) {
                    const unsigned rank = pix / per_cta, off = pix - rank * per_cta;
                    unsigned addr = sbase + off * (unsigned)stride_bytes;
                    if (CLUSTER > 1) addr = mapa(addr, rank);
                    smem_update<OP, CLUSTER == 1>(addr, r, sink);
                }
            } else if (l2ws) {
                const float w = __uint_as_float((r & 0x007fffffu) | 0x3f000000u);
                red_add4(l2ws + ((size_t)pix * 2 + (r & 1)) * 4, make_float4(w, 1.0f - w, 0.f, 0.f));
            }
        }
    }
    if (CLUSTER > 1) cg::this_cluster().sync(); else __syncthreads();
    // flush (cost included: it is part of the design) -- here just a checksum so the table is live
    unsigned long long s = sink;
    for (unsigned i = threadIdx.x; i < per_cta * (unsigned)stride_bytes / 4; i += blockDim.x) s += reinterpret_cast<unsigned *>(table)[i];
    if (s == 0x123456789abcdefull) out[0] = s;
}

template <int OP, int CLUSTER>
static float run(const uint4 *a, size_t n4, unsigned npix, double cover_frac, int stride_bytes, size_t smem, float *l2ws, unsigned long long *out, int *grid_out)
{
    auto kern = hybrid<OP, CLUSTER>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CLUSTER > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    cfg.gridDim = dim3(CLUSTER);
    if (CLUSTER > 1) { CK(cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg)); } else nclusters = 148;
    if (nclusters < 1) { *grid_out = 0; return -1.f; }
    cfg.gridDim = dim3(nclusters * CLUSTER);
    *grid_out = nclusters * CLUSTER;
    unsigned per_cta = (unsigned)(smem / stride_bytes);
    unsigned covered = (unsigned)(cover_frac * npix);
    if (covered > per_cta * CLUSTER) covered = per_cta * CLUSTER;
    const uint4 *a0 = a, *a1 = a + n4, *a2 = a + 2 * n4, *a3 = a + 3 * n4;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, kern, a0, a1, a2, a3, n4, npix, covered, per_cta, stride_bytes, l2ws, out));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    return best;
}

template <int OP, int CLUSTER>
static void sweep(const char *name, const uint4 *a, size_t n4, unsigned npix, int stride, float *l2ws, unsigned long long *out)
{
    const size_t smem = 200 * 1024;
    const double max_cover = (double)(smem / stride) * CLUSTER / npix;
    for (double f : {0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.65, 0.8, 1.0}) {
        if (f > max_cover + 1e-9 && f != 0.0) { f = max_cover; }
        int grid;
        float ms = run<OP, CLUSTER>(a, n4, npix, f, stride, smem, l2ws, out, &grid);
        printf("%-10s cluster=%2d grid=%3d stride=%dB on-chip share=%.3f : %.3f ms  (%.1f Gev/s)\n", name, CLUSTER, grid, stride, f, ms,
               n4 * 4 / ms / 1e6);
        if (f >= max_cover) break;
    }
}

int main(int argc, char **argv)
{
    const size_t n = 50000000 / 4 * 4, n4 = n / 4;
    const unsigned npix = 480 * 640;
    uint4 *a; float *l2ws; unsigned long long *out;
    CK(cudaMalloc(&a, n * 16));
    CK(cudaMalloc(&l2ws, (size_t)npix * 2 * 16));
    CK(cudaMalloc(&out, 64));
    CK(cudaMemset(l2ws, 0, (size_t)npix * 2 * 16));
    fill_random<<<148 * 8, 256>>>(a, n, 1234u);
    CK(cudaDeviceSynchronize());
    int g;
    printf("== read-only floor (no updates at all) ==\n");
    printf("read only: %.3f ms\n", run<OP_NONE, 1>(a, n4, npix, 1.0, 8, 200 * 1024, nullptr, out, &g));
    printf("== all L2 (covered = 0) ==\n");
    printf("L2 v4 RED only: %.3f ms\n", run<OP_U64, 1>(a, n4, npix, 0.0, 8, 200 * 1024, l2ws, out, &g));
    printf("== A. local shared atomics only, table = 200 KB, every event hits the local table (npix := table size) ==\n");
    printf("local u32      : %.3f ms\n", run<OP_U32, 1>(a, n4, 200 * 1024 / 4, 1.0, 4, 200 * 1024, nullptr, out, &g));
    printf("local u32 x2   : %.3f ms\n", run<OP_U32X2, 1>(a, n4, 200 * 1024 / 8, 1.0, 8, 200 * 1024, nullptr, out, &g));
    printf("local u32 x4 (181x241 image, bilinear footprint): %.3f ms\n", run<OP_U32X4, 1>(a, n4, 180 * 241 - 2, 1.0, 4, 200 * 1024, nullptr, out, &g));
    printf("local f32 x4 (181x241 image, bilinear footprint): %.3f ms\n", run<OP_F32X4, 1>(a, n4, 180 * 241 - 2, 1.0, 4, 200 * 1024, nullptr, out, &g));
    printf("local u64      : %.3f ms\n", run<OP_U64, 1>(a, n4, 200 * 1024 / 8, 1.0, 8, 200 * 1024, nullptr, out, &g));
    printf("local u64 ret  : %.3f ms\n", run<OP_U64_RET, 1>(a, n4, 200 * 1024 / 8, 1.0, 8, 200 * 1024, nullptr, out, &g));
    printf("local f32 (CAS): %.3f ms\n", run<OP_F32, 1>(a, n4, 200 * 1024 / 4, 1.0, 4, 200 * 1024, nullptr, out, &g));
    printf("== B. DSMEM only: every event hits the cluster's table (npix := cluster table size) ==\n");
#define DS(OPN, C, STR) { float ms = run<OPN, C>(a, n4, (unsigned)(200 * 1024 / STR) * C, 1.0, STR, 200 * 1024, nullptr, out, &g); \
        printf("dsmem %-10s cluster=%2d grid=%3d: %.3f ms (%.1f G/s)\n", #OPN, C, g, ms, n / ms / 1e6); }
    DS(OP_U32, 2, 4) DS(OP_U32, 4, 4) DS(OP_U32, 8, 4) DS(OP_U32, 16, 4)
    DS(OP_U64, 2, 8) DS(OP_U64, 4, 8) DS(OP_U64, 8, 8) DS(OP_U64, 16, 8)
    DS(OP_F32, 4, 4) DS(OP_F32, 8, 4)
    DS(OP_U32X2, 4, 8) DS(OP_U32X2, 8, 8)
    DS(OP_U64_RET, 8, 8)
    printf("== C. hybrid voxel-like: 480x640 pixels, 8-byte packed cell per pixel on chip, rest via L2 v4 RED ==\n");
    sweep<OP_U64, 1>("u64", a, n4, npix, 8, l2ws, out);
    sweep<OP_U64, 2>("u64", a, n4, npix, 8, l2ws, out);
    sweep<OP_U64, 4>("u64", a, n4, npix, 8, l2ws, out);
    sweep<OP_U64, 8>("u64", a, n4, npix, 8, l2ws, out);
    sweep<OP_U64, 16>("u64", a, n4, npix, 8, l2ws, out);
    sweep<OP_U32, 4>("u32", a, n4, npix, 4, l2ws, out);
    sweep<OP_U32, 8>("u32", a, n4, npix, 4, l2ws, out);
    sweep<OP_U32X2, 8>("u32x2", a, n4, npix, 8, l2ws, out);
    sweep<OP_U32X2, 16>("u32x2", a, n4, npix, 8, l2ws, out);
    sweep<OP_F32, 8>("f32", a, n4, npix, 4, l2ws, out);
    return 0;
}
