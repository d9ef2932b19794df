// tools/tma_probe.cu -- which 2-D TMA loads (cp.async.bulk.tensor.2d) does this GPU accept?  One variant per process (argv[1]),
// because an illegal instruction poisons the context.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Params { alignas(64) CUtensorMap tm; int x, y, bytes; unsigned char* out; int pad[5]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_nested(const __grid_constant__ Params prm)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t bar = smem_u32(smem), dst = smem_u32(smem + 1024);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(prm.bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(dst), "l"(&prm.tm), "r"(prm.x), "r"(prm.y), "r"(bar) : "memory");
    }
    uint32_t ok = 0; long long t0 = clock64();
    while (!ok && clock64() - t0 < 200000000ll)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar) : "memory");
    for (int i = threadIdx.x; i < prm.bytes; i += blockDim.x) prm.out[i] = ok ? smem[1024 + i] : 0xEE;
}

__global__ void k_direct(const __grid_constant__ CUtensorMap tm, int x, int y, int bytes, unsigned char* out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t bar = smem_u32(smem), dst = smem_u32(smem + 1024);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(dst), "l"(&tm), "r"(x), "r"(y), "r"(bar) : "memory");
    }
    uint32_t ok = 0; long long t0 = clock64();
    while (!ok && clock64() - t0 < 200000000ll)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = ok ? smem[1024 + i] : 0xEE;
}

int main(int argc, char** argv)
{
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    struct V { const char* name; CUtensorMapDataType dt; int es; int W, H, pitch; int bw, bh; int x, y; int nested; int l2; } vs[] = {
        {"u32 72x6 direct",      CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 960, 540, 4096, 72, 6, 8, 8, 0, 2},
        {"u16 136x8 direct",     CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 1920, 1080, 4096, 136, 8, 8, 8, 0, 2},
        {"u16 136x8 nested",     CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 1920, 1080, 4096, 136, 8, 8, 8, 1, 2},
        {"u16 128x8 direct",     CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 1920, 1080, 4096, 128, 8, 8, 8, 0, 2},
        {"u16 64x8 direct",      CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 1920, 1080, 4096, 64, 8, 8, 8, 0, 2},
        {"u8 144x8 direct",      CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 1920, 1080, 2048, 144, 8, 16, 8, 0, 2},
        {"u16 136x8 neg coords", CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 1920, 1080, 4096, 136, 8, -3, -3, 0, 2},
        {"u16 136x8 small 150x70", CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 150, 70, 512, 136, 8, -3, 0, 0, 2},
        {"u16 136x8 l2 none",    CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 1920, 1080, 4096, 136, 8, 8, 8, 0, 0},
        {"f32 64x8 direct",      CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 960, 540, 4096, 64, 8, 8, 8, 0, 2},
        {"u16 136x8 odd x",      CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 1920, 1080, 4096, 136, 8, 5, 8, 0, 2},
        {"u16 144x8 x=-8 y=-3",  CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 1920, 1080, 4096, 144, 8, -8, -3, 0, 2},
        {"u16 144x8 small x=-8", CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 150, 70, 512, 144, 8, -8, 66, 0, 2},
        {"u32 76x6 x=-4 y=-1",   CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 75, 35, 512, 76, 6, -4, -1, 0, 2},
        {"u32 76x6 x=2 (8B)",    CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 960, 540, 4096, 76, 6, 2, 0, 0, 2},
        {"u8 160x8 x=-16",       CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 1920, 1080, 2048, 160, 8, -16, 1075, 0, 2},
    };
    const int nv = sizeof(vs) / sizeof(vs[0]);
    if (v < 0 || v >= nv) { printf("variants 0..%d\n", nv - 1); return 0; }
    const V& c = vs[v];
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess || !fnp) { printf("no encode fn\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    std::vector<unsigned char> h((size_t)c.pitch * c.H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)(i * 7 + (i >> 8));
    unsigned char *d, *out;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    const int bytes = c.bw * c.bh * c.es;
    cudaMalloc(&out, bytes);
    CUtensorMap tm;
    const cuuint64_t dims[2] = { (cuuint64_t)c.W, (cuuint64_t)c.H }; const cuuint64_t strides[1] = { (cuuint64_t)c.pitch };
    const cuuint32_t box[2] = { (cuuint32_t)c.bw, (cuuint32_t)c.bh }; const cuuint32_t es[2] = { 1, 1 };
    CUresult r = enc(&tm, c.dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)c.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("variant %d (%s): encode failed %d\n", v, c.name, (int)r); return 1; }
    const int smem = 1024 + bytes + 1024;
    cudaError_t e;
    if (c.nested) {
        Params p; memset(&p, 0, sizeof(p)); p.tm = tm; p.x = c.x; p.y = c.y; p.bytes = bytes; p.out = out;
        cudaFuncSetAttribute(k_nested, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_nested<<<1, 128, smem>>>(p);
    } else {
        cudaFuncSetAttribute(k_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_direct<<<1, 128, smem>>>(tm, c.x, c.y, bytes, out);
    }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d (%s): %s\n", v, c.name, cudaGetErrorString(e)); return 2; }
    std::vector<unsigned char> o(bytes); cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
    int bad = 0, timeout = 0;
    for (int r2 = 0; r2 < c.bh; r2++) for (int i = 0; i < c.bw * c.es; i++) {
        const int gy = c.y + r2, gxb = c.x * c.es + i;
        unsigned char want = 0;
        if (gy >= 0 && gy < c.H && gxb >= 0 && gxb < c.W * c.es) want = h[(size_t)gy * c.pitch + gxb];
        const unsigned char got = o[r2 * c.bw * c.es + i];
        if (got != want) { bad++; if (got == 0xEE) timeout++; }
    }
    printf("variant %d (%s): ok, %d mismatching bytes of %d%s\n", v, c.name, bad, bytes, timeout ? " (wait timed out)" : "");
    return 0;
}
