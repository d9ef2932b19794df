// tools/ubench.cu -- instruction-throughput probes used to size the fused kernel (DESIGN.md "issue budget").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define ITERS 4096
#define UNROLL 8

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, float a, float b)
{
    float r[UNROLL];
    float2 r2[UNROLL];
    #pragma unroll
    for (int i = 0; i < UNROLL; i++) { r[i] = a + threadIdx.x * 1e-6f + i; r2[i] = make_float2(r[i], r[i] + 1.f); }
    for (int it = 0; it < ITERS; it++) {
        #pragma unroll
        for (int i = 0; i < UNROLL; i++) {
            if (OP == 0) r[i] = __fmaf_rn(r[i], a, b);
            else if (OP == 1) {
                unsigned long long x, aa, bb;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(x) : "f"(r2[i].x), "f"(r2[i].y));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(aa) : "f"(a));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(bb) : "f"(b));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(aa), "l"(bb));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(r2[i].x), "=f"(r2[i].y) : "l"(x));
            }
            else if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
            else if (OP == 3) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
            else if (OP == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
            else if (OP == 5) r[i] = (float)__float2int_rd(r[i]) + a;          // F2I + I2F + FADD
            else if (OP == 6) r[i] = __fmul_rn(r[i], a);
            else if (OP == 7) r[i] = __fadd_rn(r[i], a);
            else if (OP == 8) { int v = __float_as_int(r[i]); v = (v >> 3) + (v & 0x7fff); r[i] = __int_as_float(v); }  // ALU pipe
            else if (OP == 9) r[i] = __fdiv_rn(r[i], b);
            else if (OP == 10) { __half2 h = __floats2half2_rn(r[i], a); r[i] = __low2float(h) + __high2float(h); }
            else if (OP == 11) r[i] = __fadd_rd(r[i], 8388608.0f) - 8388608.0f;
        }
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < UNROLL; i++) s += r[i] + r2[i].x + r2[i].y;
    if (s == 123.456f) out[0] = s;
}

template <int OP>
void run(const char* name, double ops_per_iter)
{
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 2, threads = 1024;
    k<OP><<<blocks, threads>>>(d, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) k<OP><<<blocks, threads>>>(d, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double total = (double)blocks * threads * ITERS * UNROLL * ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f Gop/s (thread-ops)  %6.2f op/clk/SM @1.9GHz\n", name, ms, total / ms / 1e6, total / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    run<0>("FFMA", 1);
    run<1>("FFMA2 (fma.rn.f32x2) [2 fma]", 2);
    run<6>("FMUL", 1);
    run<7>("FADD", 1);
    run<8>("ALU (shf+lop+iadd) [3 ops]", 3);
    run<2>("MUFU.EX2", 1);
    run<3>("MUFU.LG2", 1);
    run<4>("MUFU.RCP", 1);
    run<5>("F2I.FLOOR+I2F+FADD [3 ops]", 3);
    run<11>("FADD.RM + FADD [2 ops]", 2);
    run<9>("__fdiv_rn", 1);
    run<10>("F2F.f16x2 pack + 2 unpack+add", 1);
    return 0;
}
