// ffma_rate.cu — per-SM FP32 FMA issue rates on B200 for the operand forms the fused
// separable-filter kernel can choose between.  Prints FMA/clk/SM for each variant.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int NACC = 16;

__constant__ float cw[32];

template <int V>
__global__ void __launch_bounds__(512) k(float* out, const float* in, long long* clk)
{
    float acc[NACC];
    float a = in[threadIdx.x], b = in[threadIdx.x + 512];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = in[i] + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (V == 0) acc[i] = fmaf(a, b, acc[i]);                  // reg, reg, reg
            if (V == 1) acc[i] = fmaf(a, cw[i], acc[i]);              // reg, const bank, reg
            if (V == 2) acc[i] = fmaf(a, 1.0009765625f, acc[i]);      // reg, imm, reg
            if (V == 5) acc[i] = fmaf(acc[(i + 1) % NACC], cw[i], acc[i]);  // 2 distinct regs + const
        }
        if (V == 3 || V == 4) {
#pragma unroll
            for (int i = 0; i < NACC; i += 2) {
                // packed: (acc[i],acc[i+1]) += (a,b) * (w,w)
                unsigned long long d, x, w;
                asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(acc[i]), "f"(acc[i + 1]));
                asm("mov.b64 %0, {%1,%2};" : "=l"(x) : "f"(a), "f"(b));
                if (V == 3) asm("mov.b64 %0, {%1,%2};" : "=l"(w) : "f"(b), "f"(b));
                else asm("mov.b64 %0, {%1,%2};" : "=l"(w) : "f"(cw[i]), "f"(cw[i]));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(x), "l"(w));
                asm("mov.b64 {%0,%1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(d));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name, int threads)
{
    float *out, *in;
    long long* clk;
    int blocks = 148;
    cudaMalloc(&out, blocks * 1024 * 4);
    cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0, 4096 * 4);
    cudaMalloc(&clk, blocks * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<V><<<blocks, threads>>>(out, in, clk);
    cudaEventRecord(e0);
    k<V><<<blocks, threads>>>(out, in, clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < blocks; ++i) c += h[i];
    c /= blocks;
    double fma = (double)ITERS * NACC * threads;
    printf("%-34s threads=%4d  %.1f FMA/clk/SM   (%.3f ms, %.2f TFMA/s chip, clk %.0f MHz)\n", name, threads,
           fma / c, ms, fma * blocks / (ms * 1e-3) / 1e12, c / (ms * 1e-3) / 1e6);
    cudaFree(out); cudaFree(in); cudaFree(clk);
}

int main()
{
    float w[32];
    for (int i = 0; i < 32; ++i) w[i] = 1.0f + i * 1e-3f;
    cudaMemcpyToSymbol(cw, w, sizeof w);
    for (int threads : {128, 256, 512, 1024}) {
        run<0>("FFMA reg*reg+reg", threads);
        run<1>("FFMA reg*const+reg", threads);
        run<2>("FFMA reg*imm+reg", threads);
        run<5>("FFMA reg2*const+reg", threads);
        run<3>("FFMA2 packed reg*reg", threads);
        run<4>("FFMA2 packed reg*{c,c}", threads);
    }
    return 0;
}
