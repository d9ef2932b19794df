// tools/issue_probe.cu -- issue-slot model of sm_100a for the instruction mixes the frame kernel lives on:
// does a packed FFMA2 (two FMA-pipe cycles) leave its second cycle's issue slot to another pipe?  Each pattern runs
// independent chains on W warps of one SM; cycles from clock64 around the unrolled loop.  Reported: cycles per bundle
// per scheduler (SM sub-partition), to be compared with the bundle's instruction count (1 IPC) and FMA-pipe cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_probe tools/issue_probe.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 512
#define CH 4

#define F2(r)   asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r) : "l"(A2), "l"(B2))
#define F1(s)   asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s) : "f"(a), "f"(b))
#define IA(i)   asm volatile("add.s32 %0, %0, %1;" : "+r"(i) : "r"(ib))
#define AL(i, u) do { if ((u) & 1) LO(i); else IA(i); } while (0)     /* one ALU op per unit; alternating add / xor so ptxas cannot merge a chain */
#define LO(i)   asm volatile("xor.b32 %0, %0, %1;" : "+r"(i) : "r"(ic))
#define SH(i)   asm volatile("shf.l.wrap.b32 %0, %0, %0, %1;" : "+r"(i) : "r"(ic))
#define LEA_(i) asm volatile("mad.lo.s32 %0, %0, 8, %1;" : "+r"(i) : "r"(ib))
#define MU(s)   asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s))
#define I2F_(i, s) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(s) : "r"(i)); asm volatile("mov.b32 %0, %1;" : "=r"(i) : "f"(s))
#define F2I_(i, s) asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(i) : "f"(s)); asm volatile("mov.b32 %0, %1;" : "=f"(s) : "r"(i))
#define PR(i)   asm volatile("prmt.b32 %0, %0, %1, 0x0040;" : "+r"(i) : "r"(ic))
#define FS(s)   asm volatile("{.reg .pred p; setp.ge.f32 p, %0, %1; selp.f32 %0, %0, %2, p;}" : "+f"(s) : "f"(a), "f"(b))
#define FMX(s)  asm volatile("max.f32 %0, %0, %1;" : "+f"(s) : "f"(b))
#define LDS_(i) asm volatile("ld.shared.b32 %0, [%0];" : "+r"(i))
#define H2F(i, r) asm volatile("{.reg .f16 lo, hi; .reg .f32 x, y; mov.b32 {lo, hi}, %1; cvt.f32.f16 x, lo; cvt.f32.f16 y, hi; mov.b64 %0, {x, y};}" : "=l"(r) : "r"(i)); \
                  asm volatile("{.reg .f32 x, y; mov.b64 {x, y}, %1; mov.b32 %0, x;}" : "=r"(i) : "l"(r))

template <int PAT>
__global__ void k(float* out, long long* cyc, float a, float b, int ib, int ic)
{
    __shared__ int sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (int)__cvta_generic_to_shared(&sm[(i * 33 + 7) & 1023]);
    unsigned long long r[CH]; float s[CH]; int q[CH];
    unsigned long long A2, B2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(A2) : "f"(a)); asm("mov.b64 %0, {%1, %1};" : "=l"(B2) : "f"(b));
    #pragma unroll
    for (int i = 0; i < CH; i++) { float x = a + i + threadIdx.x * 1e-6f; asm("mov.b64 %0, {%1, %1};" : "=l"(r[i]) : "f"(x)); s[i] = x; q[i] = (int)__cvta_generic_to_shared(&sm[(threadIdx.x * 17 + i * 5) & 1023]); }
    __syncthreads();
    const long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        #pragma unroll
        for (int u = 0; u < 4; u++) {
            #pragma unroll
            for (int i = 0; i < CH; i++) {
                if (PAT == 0)  { F2(r[i]); }
                if (PAT == 1)  { F2(r[i]); AL(q[i], u); }
                if (PAT == 2)  { F2(r[i]); IA(q[i]); LO(q[i]); }
                if (PAT == 3)  { F1(s[i]); AL(q[i], u); }
                if (PAT == 4)  { IA(q[i]); LO(q[i]); }
                if (PAT == 5)  { F2(r[i]); F1(s[i]); }
                if (PAT == 6)  { F1(s[i]); }
                if (PAT == 7)  { F2(r[i]); MU(s[i]); }
                if (PAT == 8)  { MU(s[i]); }
                if (PAT == 9)  { I2F_(q[i], s[i]); }
                if (PAT == 10) { F2I_(q[i], s[i]); }
                if (PAT == 11) { LEA_(q[i]); }
                if (PAT == 12) { PR(q[i]); }
                if (PAT == 13) { FS(s[i]); }
                if (PAT == 14) { FMX(s[i]); }
                if (PAT == 15) { LDS_(q[i]); }
                if (PAT == 16) { F2(r[i]); LDS_(q[i]); }
                if (PAT == 17) { H2F(q[i], r[i]); }
                if (PAT == 18) { SH(q[i]); }
                if (PAT == 19) { F2(r[i]); IA(q[i]); LO(q[i]); F1(s[i]); }
                if (PAT == 20) { F2(r[i]); F2(r[(i + 1) % CH]); AL(q[i], u); }
                if (PAT == 21) { F1(s[i]); F1(s[(i + 1) % CH]); AL(q[i], u); }
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0;
    #pragma unroll
    for (int i = 0; i < CH; i++) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(r[i])); acc += x + y + s[i] + (float)q[i]; }
    if (acc == 123.456f) out[0] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int PAT>
void run(const char* name, int instr_per_unit)
{
    float* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 8);
    printf("%-44s", name);
    for (int warps : {4, 8, 16, 28}) {
        k<PAT><<<1, 32 * warps>>>(d, c, 1.0001f, 0.5f, 3, 5);
        cudaDeviceSynchronize();
        k<PAT><<<1, 32 * warps>>>(d, c, 1.0001f, 0.5f, 3, 5);
        long long h = 0; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        const double units = (double)ITERS * 4 * CH * (warps / 4.0);       // units issued per scheduler
        printf("  w%-2d %6.2f cyc/unit", warps, (double)h / units);
    }
    printf("   (%d instr/unit)\n", instr_per_unit);
    cudaFree(d); cudaFree(c);
}

int main()
{
    cudaError_t e = cudaSetDevice(0); if (e) { printf("no device\n"); return 1; }
    printf("cycles per unit per scheduler; columns: warps per SM (4 = one per scheduler)\n");
    run<0>("FFMA2", 1);
    run<6>("FFMA", 1);
    run<1>("FFMA2 + IADD", 2);
    run<2>("FFMA2 + IADD + LOP", 3);
    run<19>("FFMA2 + IADD + LOP + FFMA", 4);
    run<20>("FFMA2 + FFMA2 + IADD", 3);
    run<21>("FFMA + FFMA + IADD", 3);
    run<3>("FFMA + IADD", 2);
    run<4>("IADD + LOP", 2);
    run<5>("FFMA2 + FFMA", 2);
    run<7>("FFMA2 + MUFU.EX2", 2);
    run<8>("MUFU.EX2", 1);
    run<9>("I2F (+mov)", 1);
    run<10>("F2I (+mov)", 1);
    run<11>("IMAD/LEA (mad.lo x8)", 1);
    run<12>("PRMT", 1);
    run<13>("FSETP + FSEL", 2);
    run<14>("FMNMX", 1);
    run<15>("LDS.32 dependent (pointer chase)", 1);
    run<16>("FFMA2 + LDS.32", 2);
    run<17>("half2 -> float2 (2 cvt)", 2);
    run<18>("SHF", 1);
    return 0;
}
