// fp32x2_latency.cu -- latency / throughput of scalar FADD/FMUL vs packed FADD2/FMUL2 on sm_100a.
// One warp per SM subpartition for latency (dependent chain), 4..16 warps for throughput.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{
    unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmul(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

// MODE 0: dependent scalar FADD chain; 1: dependent FADD2 chain; 2: 8 independent scalar chains; 3: 8 independent packed chains
// 4: dependent FMUL; 5: dependent FMUL2
template <int MODE>
__global__ void k(float *out, int iters, long long *cyc)
{
    float a[8]; unsigned long long p[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
    const float c = 1.0000001f; const unsigned long long pc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0) a[0] = fadd(a[0], c);
            if (MODE == 1) p[0] = add2(p[0], pc);
            if (MODE == 4) a[0] = fmul(a[0], c);
            if (MODE == 5) p[0] = mul2(p[0], pc);
            if (MODE == 2) { for (int i = 0; i < 8; i++) a[i] = fadd(a[i], c); }
            if (MODE == 3) { for (int i = 0; i < 8; i++) p[i] = add2(p[i], pc); }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; i++) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE> void run(const char *name, int threads, int ops_per_iter)
{
    float *out; long long *cyc, h; cudaMalloc(&out, 4 * 148 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<MODE><<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * 16.0 * ops_per_iter);
    printf("%-34s threads/SM %4d: %.2f clk per instruction per warp; SM-wide %.2f warp-instr/clk\n", name, threads, per,
           (threads / 32) / per);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("dependent FADD (latency)", 32, 1);
    run<1>("dependent FADD2 (latency)", 32, 1);
    run<4>("dependent FMUL (latency)", 32, 1);
    run<5>("dependent FMUL2 (latency)", 32, 1);
    for (int t : {128, 256, 384, 512, 1024}) {
        run<2>("8 independent FADD chains", t, 8);
        run<3>("8 independent FADD2 chains", t, 8);
    }
    return 0;
}
