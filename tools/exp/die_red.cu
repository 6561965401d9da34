// Experiment: do die-local reductions (each SM only RED-ing into L2 lines homed on its own die) run faster?
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <algorithm>
#include <vector>

#define CHUNK_WORDS 512   // 2 KB

__global__ void probe(const unsigned *base, int nchunks, int reps, unsigned short *lat, unsigned *smids)
{
    if (threadIdx.x != 0) return;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    smids[blockIdx.x] = smid;
    unsigned sink = 0;
    for (int c = 0; c < nchunks; ++c) {
        const unsigned *p = base + (size_t)c * CHUNK_WORDS;
        unsigned best = 0xffffffffu, off = 0;
        for (int r = 0; r < reps; ++r) {
            unsigned v;
            long long t0, t1 = 0;
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mov.u64 %0, %%clock64;\n\t"
                         "ld.global.cv.u32 %2, [%3];\n\t"
                         "setp.lt.u32 p, %2, 0x7fffffff;\n\t"
                         "@p mov.u64 %1, %%clock64;\n\t}"
                         : "=l"(t0), "+l"(t1), "=r"(v) : "l"(p + off) : "memory");
            off = (off + 8 + v) & (CHUNK_WORDS - 1);
            sink += v;
            unsigned d = (unsigned)(t1 - t0);
            if (r > 0 && d < best) best = d;
        }
        lat[(size_t)blockIdx.x * nchunks + c] = (unsigned short)min(best, 65535u);
    }
    if (sink == 0x12345) smids[blockIdx.x] = 0;
}

// every thread issues `per_thread` vector reductions to pseudo-random 16-byte slots of chunks taken from the list
// of its SM's group (list[g], count[g]); sm_group[smid] gives the group
__global__ void red_kernel(float *base, const int *listA, int nA, const int *listB, int nB, const unsigned char *sm_group,
                           int per_thread, unsigned seed)
{
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const int *list = sm_group[smid] ? listB : listA;
    const int n = sm_group[smid] ? nB : nA;
    unsigned s = seed ^ (blockIdx.x * 2654435761u) ^ (threadIdx.x * 40503u);
    for (int i = 0; i < per_thread; ++i) {
        s = s * 1664525u + 1013904223u;
        const unsigned c = list[(s >> 8) % (unsigned)n];
        s = s * 1664525u + 1013904223u;
        const unsigned slot = (s >> 10) & 127u;   // 128 16-byte slots per chunk
        float *a = base + (size_t)c * CHUNK_WORDS + slot * 4;
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(1.0f), "f"(0.5f), "f"(0.0f), "f"(0.0f));
    }
}

int main()
{
    const int P = 16384, reps = 5, grid = 148 * 4, want = 4800;
    unsigned *buf; unsigned short *lat; unsigned *smids;
    cudaMalloc(&buf, (size_t)P * CHUNK_WORDS * 4);
    cudaMemset(buf, 0, (size_t)P * CHUNK_WORDS * 4);
    cudaMalloc(&lat, (size_t)grid * P * 2);
    cudaMalloc(&smids, grid * 4);
    probe<<<grid, 32>>>(buf, P, reps, lat, smids);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<unsigned short> hl((size_t)grid * P);
    std::vector<unsigned> hs(grid);
    cudaMemcpy(hl.data(), lat, hl.size() * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(hs.data(), smids, grid * 4, cudaMemcpyDeviceToHost);
    std::vector<int> row(256, -1);
    for (int b = 0; b < grid; ++b) if (row[hs[b]] < 0) row[hs[b]] = b;
    std::vector<int> sms;
    for (int s = 0; s < 256; ++s) if (row[s] >= 0) sms.push_back(s);
    const int S = (int)sms.size();
    // centre: subtract per-SM mean, then per-chunk mean
    std::vector<float> M((size_t)S * P);
    for (int i = 0; i < S; ++i) {
        double m = 0; for (int c = 0; c < P; ++c) m += hl[(size_t)row[sms[i]] * P + c]; m /= P;
        for (int c = 0; c < P; ++c) M[(size_t)i * P + c] = hl[(size_t)row[sms[i]] * P + c] - (float)m;
    }
    for (int c = 0; c < P; ++c) { double m = 0; for (int i = 0; i < S; ++i) m += M[(size_t)i * P + c]; m /= S; for (int i = 0; i < S; ++i) M[(size_t)i * P + c] -= (float)m; }
    // group SMs by the sign of their correlation with SM 0 (two refinement rounds against the group mean)
    std::vector<float> ref(P); for (int c = 0; c < P; ++c) ref[c] = M[c];
    std::vector<int> grp(S, 0);
    for (int it = 0; it < 3; ++it) {
        for (int i = 0; i < S; ++i) { double d = 0; for (int c = 0; c < P; ++c) d += (double)M[(size_t)i * P + c] * ref[c]; grp[i] = d < 0; }
        std::fill(ref.begin(), ref.end(), 0.f);
        for (int i = 0; i < S; ++i) for (int c = 0; c < P; ++c) ref[c] += (grp[i] ? -1.f : 1.f) * M[(size_t)i * P + c];
    }
    int nB = 0; for (int i = 0; i < S; ++i) nB += grp[i];
    // chunk home: ref[c] < 0 means group-0 SMs see it faster
    std::vector<int> LA, LB; double sep = 0;
    for (int c = 0; c < P; ++c) { (ref[c] < 0 ? LA : LB).push_back(c); sep += fabs(ref[c]) / S; }
    printf("SMs %d: group0 %d group1 %d; chunks near group0 %zu near group1 %zu; mean |latency gap| %.1f cycles\n", S, S - nB, nB, LA.size(), LB.size(), 2 * sep / P);
    if ((int)LA.size() < want || (int)LB.size() < want) { printf("not enough chunks per die\n"); return 1; }
    // lists: local (A for group 0, B for group 1), swapped (all remote), mixed (the same 2*want chunks for everybody)
    std::vector<int> a(LA.begin(), LA.begin() + want), b(LB.begin(), LB.begin() + want), mix;
    for (int i = 0; i < want; ++i) mix.push_back(i % 2 ? a[i] : b[i]);
    std::vector<unsigned char> g256(256, 0);
    for (int i = 0; i < S; ++i) g256[sms[i]] = (unsigned char)grp[i];
    int *dA, *dB, *dM; unsigned char *dG;
    cudaMalloc(&dA, want * 4); cudaMalloc(&dB, want * 4); cudaMalloc(&dM, want * 4); cudaMalloc(&dG, 256);
    cudaMemcpy(dA, a.data(), want * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, b.data(), want * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dM, mix.data(), want * 4, cudaMemcpyHostToDevice); cudaMemcpy(dG, g256.data(), 256, cudaMemcpyHostToDevice);
    const int rgrid = 148 * 8, per = 50000000 / (rgrid * 256) + 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[3] = {"die-local (each SM -> chunks homed on its die)", "all remote (swapped lists)", "mixed (same chunks for all SMs)"};
    for (int rep = 0; rep < 2; ++rep)
        for (int mode = 0; mode < 3; ++mode) {
            const int *l0 = mode == 0 ? dA : mode == 1 ? dB : dM, *l1 = mode == 0 ? dB : mode == 1 ? dA : dM;
            cudaEventRecord(e0);
            red_kernel<<<rgrid, 256>>>((float *)buf, l0, want, l1, want, dG, per, 1234u + rep);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%-50s %.3f ms for %.1f M v4 REDs (%.1f G/s)\n", names[mode], ms, (double)rgrid * 256 * per / 1e6, (double)rgrid * 256 * per / ms / 1e6);
        }
    return 0;
}
