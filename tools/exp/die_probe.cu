// Experiment: can the L2 "home die" of a 2 KB chunk, and the die of an SM, be told apart by L2-hit latency?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o die_probe die_probe.cu && ./die_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

__global__ void probe(const unsigned *base, int nchunks, int chunk_words, int reps, unsigned *lat, unsigned *smids)
{
    if (threadIdx.x != 0) return;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    smids[blockIdx.x] = smid;
    unsigned sink = 0;
    for (int c = 0; c < nchunks; ++c) {
        const unsigned *p = base + (size_t)c * chunk_words;
        unsigned best = 0xffffffffu;
        unsigned off = 0;
        for (int r = 0; r < reps; ++r) {
            unsigned v;
            long long t0, t1 = 0;
            // the second clock read is predicated on the loaded value, so it cannot issue before the load returns
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mov.u64 %0, %%clock64;\n\t"
                         "ld.global.cv.u32 %2, [%3];\n\t"
                         "setp.lt.u32 p, %2, 0x7fffffff;\n\t"
                         "@p mov.u64 %1, %%clock64;\n\t}"
                         : "=l"(t0), "+l"(t1), "=r"(v) : "l"(p + off) : "memory");
            off = (off + 8 + v) & (chunk_words - 1);   // v == 0: next sector
            sink += v;
            unsigned d = (unsigned)(t1 - t0);
            if (r > 0 && d < best) best = d;
        }
        lat[(size_t)blockIdx.x * nchunks + c] = best;
    }
    if (sink == 0x12345) smids[blockIdx.x] = 0;
}

int main()
{
    const int nchunks = 512, chunk_bytes = 2048, chunk_words = chunk_bytes / 4, reps = 12;
    const int grid = 148 * 6;
    unsigned *buf, *lat, *smids;
    cudaMalloc(&buf, (size_t)nchunks * chunk_bytes);
    cudaMemset(buf, 0, (size_t)nchunks * chunk_bytes);
    cudaMalloc(&lat, (size_t)grid * nchunks * 4);
    cudaMalloc(&smids, grid * 4);
    probe<<<grid, 32>>>(buf, nchunks, chunk_words, reps, lat, smids);
    probe<<<grid, 32>>>(buf, nchunks, chunk_words, reps, lat, smids);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<unsigned> hl((size_t)grid * nchunks), hs(grid);
    cudaMemcpy(hl.data(), lat, hl.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hs.data(), smids, grid * 4, cudaMemcpyDeviceToHost);
    // one row per SM (first CTA seen)
    std::vector<int> row_of_sm(256, -1);
    for (int b = 0; b < grid; ++b) if (row_of_sm[hs[b]] < 0) row_of_sm[hs[b]] = b;
    // global latency histogram
    std::vector<unsigned> all;
    for (int s = 0; s < 256; ++s) if (row_of_sm[s] >= 0) for (int c = 0; c < nchunks; ++c) all.push_back(hl[(size_t)row_of_sm[s] * nchunks + c]);
    std::sort(all.begin(), all.end());
    printf("latency percentiles: min %u p10 %u p25 %u p50 %u p75 %u p90 %u max %u\n", all[0], all[all.size() / 10], all[all.size() / 4],
           all[all.size() / 2], all[all.size() * 3 / 4], all[all.size() * 9 / 10], all.back());
    unsigned thr = (all[all.size() / 4] + all[all.size() * 3 / 4]) / 2;
    // histogram
    int hist[40] = {0};
    for (unsigned v : all) { int b = (int)(v - all[0]) / 8; if (b > 39) b = 39; hist[b]++; }
    printf("histogram (bin 8 cycles from %u):", all[0]);
    for (int i = 0; i < 40; ++i) printf(" %d", hist[i]);
    printf("\nthreshold %u\n", thr);
    // signature of SM 0's row vs others: fraction of chunks classified equal
    int ref = -1;
    for (int s = 0; s < 256; ++s) if (row_of_sm[s] >= 0) { ref = s; break; }
    int nsm = 0, same = 0, opp = 0, unclear = 0;
    for (int s = 0; s < 256; ++s) {
        if (row_of_sm[s] < 0) continue;
        ++nsm;
        int eq = 0;
        for (int c = 0; c < nchunks; ++c) {
            bool a = hl[(size_t)row_of_sm[ref] * nchunks + c] > thr, b = hl[(size_t)row_of_sm[s] * nchunks + c] > thr;
            eq += (a == b);
        }
        double f = (double)eq / nchunks;
        if (f > 0.9) ++same; else if (f < 0.1) ++opp; else ++unclear;
        if (s < 12 || f > 0.1 && f < 0.9) printf("sm %3d agreement with sm %d: %.3f\n", s, ref, f);
    }
    int far = 0;
    for (int c = 0; c < nchunks; ++c) far += hl[(size_t)row_of_sm[ref] * nchunks + c] > thr;
    printf("SMs %d: same-die-as-ref %d, other-die %d, unclear %d; chunks far from ref SM: %d / %d\n", nsm, same, opp, unclear, far, nchunks);
    return 0;
}
