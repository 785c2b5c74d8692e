// membench.cu — read-side ceilings of one B200 for the glt_stream design (context for the roofline):
//   ldg_read : grid-stride 16-byte loads, 8 in flight per thread, summed
//   tma_read : persistent CTAs, ring of NST stages of CH bytes filled by 1-D bulk copies, consumer
//              only acknowledges (no compute): the ceiling of the producer/ring structure itself
//   copy     : 16-byte load + store (what MEASURED_PEAKS.json's hbm_gbs measures)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o membench membench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void ldg_read(const float4* __restrict__ p, size_t n4, float* out) {
    float acc = 0.f;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcs(p + i + j * stride);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
    }
    for (; i < n4; i += stride) { float4 v = __ldcs(p + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void copy16(const float4* __restrict__ p, float4* __restrict__ q, size_t n4) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldcs(p + i + j * stride);
#pragma unroll
        for (int j = 0; j < 4; ++j) __stcs(q + i + j * stride, v[j]);
    }
    for (; i < n4; i += stride) __stcs(q + i, __ldcs(p + i));
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t tx) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(tx) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}

// chunk = CH bytes; each stage gets `split` copies of CH/split bytes
__global__ void __launch_bounds__(64, 1) tma_read(const unsigned char* __restrict__ p, size_t nchunks, int ch, int nst, int split) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm);
    uint64_t* empty = full + 32;
    unsigned char* st = sm + 512;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t u = 0;
            for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
                mbar_wait(&empty[s], (u & 1) ^ 1);
                mbar_expect(&full[s], ch);
                const int part = ch / split;
                for (int j = 0; j < split; ++j) bulk(st + (size_t)s * ch + j * part, p + c * (size_t)ch + j * part, part, &full[s]);
                if (++s == nst) { s = 0; ++u; }
            }
        }
    } else {
        int s = 0; uint32_t u = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            mbar_wait(&full[s], u & 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == nst) { s = 0; ++u; }
        }
    }
}

int main() {
    const size_t bytes = 3200ull << 20;
    unsigned char *a, *b; float* o;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&o, 4));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 0, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, double moved, auto fn) {
        fn(); fn(); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) fn();
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, moved / ms / 1e6);
        fflush(stdout);
    };
    const size_t n4 = bytes / 16;
    for (int bpsm : {8, 16, 32}) {
        char nm[64]; snprintf(nm, 64, "ldg_read  (%d x 256 thr / SM)", bpsm);
        timeit(nm, (double)bytes, [&] { ldg_read<<<148 * bpsm, 256>>>((const float4*)a, n4, o); });
    }
    timeit("copy16    (16 x 256 thr / SM, R+W bytes)", 2.0 * bytes, [&] { copy16<<<148 * 16, 256>>>((const float4*)a, (float4*)b, n4); });
    CK(cudaFuncSetAttribute(tma_read, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    struct Cfg { int ch, nst, split; } cfgs[] = {{36864, 5, 1}, {36864, 5, 4}, {36864, 5, 16}, {36864, 5, 32}, {36864, 3, 1}, {18432, 10, 1}, {18432, 10, 16},
                                                {9216, 20, 1}, {73728, 2, 1}, {73728, 2, 64}};
    for (auto c : cfgs) {
        char nm[64]; snprintf(nm, 64, "tma_read  chunk %d x %d stages, %d copies", c.ch, c.nst, c.split);
        const size_t nchunks = bytes / c.ch;
        timeit(nm, (double)nchunks * c.ch, [&] { tma_read<<<148, 64, 512 + c.ch * c.nst>>>(a, nchunks, c.ch, c.nst, c.split); });
    }
    return 0;
}
