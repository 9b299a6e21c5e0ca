// Micro-benchmark (B200): issue / pipe rates that decide how the REBLUR tap loops are written.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_issue tools/ubench_issue.cu && tools/ubench_issue
// Prints warp-instructions per clock per SM for: FFMA, FFMA2 (packed fp32x2, sm_100), FFMA+IADD3 interleaved, MUFU.RCP, F2I.FLOOR,
// FADD.RM (the XU-free floor), HADD2.F32 (fp16 -> fp32 unpack), LDG.128 (L1 hits).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048;

template <int MODE> __global__ void __launch_bounds__(256) K(float* out, const float4* in, float seed)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = seed + i + threadIdx.x;
    float2 b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = make_float2(seed + i, seed - i + threadIdx.x);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(seed, -seed);
    int acc = threadIdx.x;
    const float4* p = in + (threadIdx.x & 31);
    for (int it = 0; it < ITER; it++)
    {
        if (MODE == 0)
        {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], 1.0001f, seed);
        }
        else if (MODE == 1)
        {
#pragma unroll
            for (int i = 0; i < 8; i++) b[i] = __ffma2_rn(b[i], m, c);
        }
        else if (MODE == 2)
        {
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                a[i] = fmaf(a[i], 1.0001f, seed);
                acc = acc * 3 + (acc >> 3) + i; // IMAD / SHF / IADD mix
            }
        }
        else if (MODE == 3)
        {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __frcp_rn(a[i]) == 0.f ? a[i] : __fdividef(1.0f, a[i]);
        }
        else if (MODE == 4)
        {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = (float)__float2int_rd(a[i]) + 0.37f;
        }
        else if (MODE == 5)
        {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fadd_rd(a[i], 12582912.0f) - 12582911.63f;
        }
        else if (MODE == 6)
        {
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                b[i] = __ffma2_rn(b[i], m, c);
                acc = acc * 3 + (acc >> 3) + i;
            }
        }
        else if (MODE == 7)
        {
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                float4 v = __ldg(p + ((acc + i * 37) & 1023));
                a[i] += v.x + v.y + v.z + v.w;
                acc += __float_as_int(v.x) & 7;
            }
        }
        else if (MODE == 8)
        {
            // 8 MUFU + 8 FFMA2 + 8 IADD: does the XU pipe overlap with packed math?
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                b[i] = __ffma2_rn(b[i], m, c);
                a[i] = __fdividef(1.0f, a[i]);
                acc += i ^ (acc >> 2);
            }
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; i++) s += b[i].x + b[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
}

template <int MODE> void Run(const char* name, double instrPerIter, float* out, const float4* in)
{
    const int blocks = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    K<MODE><<<blocks, 256>>>(out, in, 1.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    K<MODE><<<blocks, 256>>>(out, in, 1.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warpInstr = (double)blocks * 8 * ITER * instrPerIter;
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-34s %8.3f ms  %6.2f warp-instr/clk/SM (nominal clock %d MHz; counts the named instructions only)\n", name, ms, warpInstr / cycles / 148.0, clk / 1000);
}

int main()
{
    float* out;
    float4* in;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaMalloc(&in, 2048 * sizeof(float4));
    cudaMemset(in, 0, 2048 * sizeof(float4));
    Run<0>("FFMA x16", 16, out, in);
    Run<1>("FFMA2 x8 (=16 fp32 fma)", 8, out, in);
    Run<2>("FFMA x8 + int x~24 (counts 8)", 8, out, in);
    Run<6>("FFMA2 x8 + int x~24 (counts 8)", 8, out, in);
    Run<3>("MUFU.RCP x16 (+select)", 16, out, in);
    Run<4>("F2I.FLOOR + I2F x16 (counts 16)", 16, out, in);
    Run<5>("FADD.RM + FADD x16 (counts 16)", 16, out, in);
    Run<7>("LDG.128 x4 L1-resident", 4, out, in);
    Run<8>("FFMA2 + MUFU + IADD x8 (counts 8)", 8, out, in);
    return 0;
}
