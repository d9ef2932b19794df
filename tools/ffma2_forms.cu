// tools/ffma2_forms.cu -- throughput of FFMA2 by operand form (sm_100a): packed registers, a scalar register broadcast to both
// lanes (.F32), a 32-bit immediate, a uniform register / constant-bank operand.  Independent chains, W warps on one SM,
// cycles per FFMA2 per scheduler (2.0 = the FMA pipe's packed rate).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o tools/ffma2_forms tools/ffma2_forms.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 256
#define CH 6

struct Prm { float a, b, c[8]; };

template <int FORM>
__global__ void __launch_bounds__(1024, 1) k(const __grid_constant__ Prm prm, float* out, long long* cyc, const float* gsrc)
{
    float2 r[CH];
    #pragma unroll
    for (int i = 0; i < CH; i++) r[i] = make_float2(prm.a + i + threadIdx.x * 1e-6f, prm.a - i);
    // "register" operands the compiler cannot fold: loaded from global memory
    const float sa = gsrc[threadIdx.x & 7], sb = gsrc[8 + (threadIdx.x & 7)];
    const float2 A = make_float2(sa, sa * 1.0001f), B = make_float2(sb, sb * 0.9999f);
    __syncthreads();
    const long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        #pragma unroll
        for (int u = 0; u < 8; u++) {
            #pragma unroll
            for (int i = 0; i < CH; i++) {
                if (FORM == 0) r[i] = __ffma2_rn(r[i], A, B);                                                        // packed regs
                if (FORM == 1) r[i] = __ffma2_rn(r[i], make_float2(sa, sa), B);                                      // scalar reg broadcast multiplicand
                if (FORM == 2) r[i] = __ffma2_rn(r[i], A, make_float2(0.001f + u * 0.125f, 0.001f + u * 0.125f));    // immediate addend
                if (FORM == 3) r[i] = __ffma2_rn(r[i], make_float2(sa, sa), make_float2(0.001f + u * 0.125f, 0.001f + u * 0.125f));   // broadcast reg x + imm
                if (FORM == 4) r[i] = __ffma2_rn(r[i], make_float2(prm.c[u], prm.c[u]), B);                          // constant bank / uniform register multiplicand
                if (FORM == 5) r[i] = __ffma2_rn(r[i], A, make_float2(prm.c[u], prm.c[u]));                          // constant bank / uniform register addend
                if (FORM == 6) r[i] = __ffma2_rn(r[i], A, make_float2(sb, sb));                                      // scalar reg broadcast addend
                if (FORM == 7) r[i] = __fadd2_rn(r[i], make_float2(-1.0f, -1.0f));                                   // FADD2 immediate
                if (FORM == 8) r[i] = __fmul2_rn(r[i], make_float2(prm.c[u], prm.c[u]));                             // FMUL2 uniform
                if (FORM == 9) r[i] = __ffma2_rn(r[i], r[(i + 1) % CH], make_float2(0.001f + u * 0.125f, 0.001f + u * 0.125f));   // Horner form: packed x packed + imm
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0;
    #pragma unroll
    for (int i = 0; i < CH; i++) acc += r[i].x + r[i].y;
    if (acc == 123.456f) out[0] = acc;
    atomicMax((unsigned long long*)cyc, (unsigned long long)(t1 - t0));
}

template <int FORM>
void run(const char* name)
{
    float* d; long long* c; float* g; cudaMalloc(&d, 4); cudaMalloc(&c, 8); cudaMalloc(&g, 64);
    float h[16]; for (int i = 0; i < 16; i++) h[i] = 0.999f + i * 1e-4f; cudaMemcpy(g, h, 64, cudaMemcpyHostToDevice);
    Prm p; p.a = 1.0001f; p.b = 0.5f; for (int i = 0; i < 8; i++) p.c[i] = 0.998f + i * 1e-4f;
    printf("%-52s", name);
    for (int warps : {4, 8, 16}) {
        cudaMemset(c, 0, 8); k<FORM><<<1, 32 * warps>>>(p, d, c, g);
        cudaMemset(c, 0, 8); k<FORM><<<1, 32 * warps>>>(p, d, c, g);
        cudaDeviceSynchronize();
        long long hh = 0; cudaMemcpy(&hh, c, 8, cudaMemcpyDeviceToHost);
        printf("  w%-2d %5.2f", warps, (double)hh / ((double)ITERS * 8 * CH * (warps / 4.0)));
    }
    printf("   cycles per instruction per scheduler\n");
}

int main()
{
    if (cudaSetDevice(0)) { printf("no device\n"); return 1; }
    run<0>("FFMA2 R, R2, R2, R2 (packed registers)");
    run<1>("FFMA2 R, R2, R.F32 (broadcast), R2");
    run<2>("FFMA2 R, R2, R2, imm");
    run<3>("FFMA2 R, R2, R.F32, imm");
    run<9>("FFMA2 R, R2, R2', imm (Horner step)");
    run<4>("FFMA2 R, R2, const/uniform, R2");
    run<5>("FFMA2 R, R2, R2, const/uniform");
    run<6>("FFMA2 R, R2, R2, R.F32 (broadcast addend)");
    run<7>("FADD2 R, R2, imm");
    run<8>("FMUL2 R, R2, const/uniform");
    return 0;
}
