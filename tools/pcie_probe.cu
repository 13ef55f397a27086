// Development aid: how fast can SMs read page-locked host memory over PCIe?  (a) 16-byte volatile loads, lanes of many warps;
// (b) TMA bulk copies (cp.async.bulk host -> shared memory) of 3 KB, one thread per warp, two in flight.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pcie_probe tools/pcie_probe.cu && ./build/pcie_probe
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void ldg_kernel(const float4* __restrict__ src, float4* __restrict__ dst, unsigned n4) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
        float4 a, b;
        asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(src + i));
        const bool two = i + stride < n4;
        if (two) asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(src + i + stride));
        __stcg(dst + i, a);
        if (two) __stcg(dst + i + stride, b);
    }
}
constexpr unsigned kTile = 3072;
__global__ void tma_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, unsigned ntiles, int depth) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ unsigned long long bar[8][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    if (lane != 0) return;
    unsigned char* buf = sm + warp * depth * kTile;
    for (int s = 0; s < depth; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[warp][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned first = blockIdx.x * W + warp, stride = gridDim.x * W;
    unsigned issued = 0, done = 0;
    for (unsigned t = first; ; t += stride) {
        if (t < ntiles) {
            const int s = issued % depth;
            if (issued >= (unsigned)depth) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the slot's store has read it
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[warp][s])), "r"(kTile) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + s * kTile)),
                         "l"(src + (size_t)t * kTile), "r"(kTile), "r"(smem_u32(&bar[warp][s])) : "memory");
            ++issued;
        }
        // retire the oldest outstanding tile once `depth` are in flight (or at the end)
        if (issued - done == (unsigned)depth || (t >= ntiles && done < issued)) {
            const int s = done % depth;
            const unsigned par = (done / depth) & 1;
            asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar[warp][s])), "r"(par) : "memory");
            const unsigned tt = first + done * stride;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)tt * kTile), "r"(smem_u32(buf + s * kTile)), "r"(kTile) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            ++done;
        }
        if (t >= ntiles && done == issued) break;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
    const size_t bytes = 3u << 20;   // 3 MB, like cfg2
    unsigned char *h, *hd, *d;
    CK(cudaHostAlloc((void**)&h, bytes, cudaHostAllocMapped));
    for (size_t i = 0; i < bytes; ++i) h[i] = (unsigned char)(i * 131u >> 3);
    CK(cudaHostGetDevicePointer((void**)&hd, h, 0));
    CK(cudaMalloc((void**)&d, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int blocks : {37, 74, 148, 296}) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            ldg_kernel<<<blocks, 256>>>((const float4*)hd, (float4*)d, (unsigned)(bytes / 16));
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("ldg  %3d blocks x 256 threads: %7.1f us  %5.1f GB/s\n", blocks, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
    }
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int depth : {2, 4}) for (int warps : {2, 4, 8}) for (int blocks : {37, 148}) {
        const size_t smem = (size_t)warps * depth * kTile;
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemset(d, 0, bytes);
            cudaEventRecord(e0);
            tma_kernel<<<blocks, warps * 32, smem>>>(hd, d, (unsigned)(bytes / kTile), depth);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        unsigned char* chk = (unsigned char*)malloc(bytes);
        cudaMemcpy(chk, d, bytes, cudaMemcpyDeviceToHost);
        size_t bad = 0; for (size_t i = 0; i < bytes; ++i) bad += chk[i] != h[i];
        free(chk);
        printf("tma  %3d blocks x %d warps, depth %d: %7.1f us  %5.1f GB/s  (%zu bytes wrong)\n", blocks, warps, depth, ms * 1e3, bytes / (ms * 1e-3) / 1e9, bad);
    }
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("cudaMemcpyAsync H2D: %7.1f us  %5.1f GB/s\n", ms * 1e3, bytes / (ms * 1e-3) / 1e9);
    return 0;
}
