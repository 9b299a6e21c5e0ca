// Does prefetch.global.L1 / .L2 (CCTL.E.PF1 / PF2) hide the latency of a later load on B200?  One warp per SM walks random
// cold lines of a 2 GB buffer; per step: [prefetch of the next address], ~600 cycles of dependent FMAs, then the load.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_prefetch tools/ubench_prefetch.cu && tools/ubench_prefetch
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void Walk(const float* __restrict__ buf, size_t nLines, float* out, int steps, int work)
{
    unsigned h = (blockIdx.x * 9781u + threadIdx.x * 6271u) | 1u;
    float acc = 0.0f, x = 1.0001f;
    for (int s = 0; s < steps; s++)
    {
        h = h * 1664525u + 1013904223u;
        const float* p = buf + (size_t)(h % nLines) * 32; // one 128-byte line per lane
        if (MODE == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        if (MODE == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        for (int i = 0; i < work; i++) x = fmaf(x, 1.0000001f, 1e-9f); // dependent chain: ~4 cycles each
        acc += __ldg(p) + x;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main()
{
    const size_t bytes = 2ull << 30, nLines = bytes / 128;
    float *buf, *out;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    cudaMalloc(&out, 148 * 32 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int steps = 2000;
    for (int work : {0, 150, 300})
        for (int mode = 0; mode < 3; mode++)
        {
            float best = 1e9f;
            for (int rep = 0; rep < 3; rep++)
            {
                cudaEventRecord(e0);
                if (mode == 0) Walk<0><<<148, 32>>>(buf, nLines, out, steps, work);
                if (mode == 1) Walk<1><<<148, 32>>>(buf, nLines, out, steps, work);
                if (mode == 2) Walk<2><<<148, 32>>>(buf, nLines, out, steps, work);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("work %3d fma  %-12s  %.1f ns per step\n", work, mode == 0 ? "no prefetch" : (mode == 1 ? "prefetch.L1" : "prefetch.L2"), best * 1e6f / steps);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
