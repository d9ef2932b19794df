// tools/latency.cu -- dependent-issue latency and few-chain throughput of the instructions the frame kernel lives on (sm_100a).
// One warp per SM sub-partition unless stated; cycles from clock64 around an unrolled dependent chain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/latency tools/latency.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define N 2048

template <int OP, int CHAINS>
__global__ void k(float* out, long long* cyc, float a, float b)
{
    float2 r[CHAINS]; float s[CHAINS];
    #pragma unroll
    for (int i = 0; i < CHAINS; i++) { r[i] = make_float2(a + i + threadIdx.x * 1e-6f, a - i); s[i] = a + i * 0.5f + threadIdx.x * 1e-6f; }
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    __syncthreads();
    const long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < N / 16; it++) {
        #pragma unroll
        for (int u = 0; u < 16; u++) {
            #pragma unroll
            for (int i = 0; i < CHAINS; i++) {
                if (OP == 0) r[i] = __ffma2_rn(r[i], A, B);                               // FFMA2, register operands
                else if (OP == 1) s[i] = __fmaf_rn(s[i], a, b);                           // FFMA
                else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(s[i]));   // MUFU.RCP
                else if (OP == 3) { s[i] = __int_as_float((__float_as_int(s[i]) + 0x3f3504f3) & 0x7fffffff); }   // IADD + LOP
                else if (OP == 4) s[i] = (float)__float_as_int(s[i]) * 1e-9f;            // I2FP + FMUL
                else if (OP == 5) r[i] = __fadd2_rn(r[i], B);                             // FADD2
                else if (OP == 6) r[i] = __ffma2_rn(r[i], make_float2(0.999f, 0.999f), make_float2(0.001f, 0.001f));   // FFMA2 with immediates
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0;
    #pragma unroll
    for (int i = 0; i < CHAINS; i++) acc += r[i].x + r[i].y + s[i];
    if (acc == 123.456f) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP, int CHAINS>
void run(const char* name, int warps)
{
    float* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 8);
    k<OP, CHAINS><<<1, 32 * warps>>>(d, c, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    k<OP, CHAINS><<<1, 32 * warps>>>(d, c, 1.0001f, 0.5f);
    long long h = 0; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-34s chains=%d warps/SM=%2d : %7.2f cycles per instruction per warp, %6.3f warp-instr/clk/SMSP\n", name, CHAINS, warps,
           (double)h / N / CHAINS, (double)N * CHAINS * (warps / 4.0 < 1 ? 1 : warps / 4.0) / h);
    cudaFree(d); cudaFree(c);
}

int main()
{
    run<0, 1>("FFMA2 dependent", 4); run<0, 2>("FFMA2", 4); run<0, 3>("FFMA2", 4); run<0, 4>("FFMA2", 4); run<0, 3>("FFMA2", 12); run<0, 3>("FFMA2", 28);
    run<6, 1>("FFMA2 imm dependent", 4); run<6, 3>("FFMA2 imm", 4); run<6, 3>("FFMA2 imm", 12);
    run<1, 1>("FFMA dependent", 4); run<1, 3>("FFMA", 4); run<1, 3>("FFMA", 12);
    run<5, 1>("FADD2 dependent", 4); run<5, 3>("FADD2", 4);
    run<2, 1>("MUFU.RCP dependent", 4); run<2, 3>("MUFU.RCP", 4); run<2, 3>("MUFU.RCP", 12);
    run<3, 1>("IADD+LOP dependent (2 instr)", 4); run<3, 3>("IADD+LOP", 4);
    run<4, 1>("I2FP+FMUL dependent (2 instr)", 4); run<4, 3>("I2FP+FMUL", 4);
    return 0;
}
