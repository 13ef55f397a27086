// tc_probe3.cu — development microbenchmark: tcgen05.mma issue rate vs N, operand kind and A source (smem / TMEM)
//   (A) tcgen05.mma kind::tf32 issue rate alone, for several K-major operand layouts (K = 16 as two K = 8 steps):
//         L0  64-byte rows, SWIZZLE_64B, K-step = +32 B inside the row       (what chamfer_tc.cu uses)
//         L1  32-byte rows, SWIZZLE_32B, one dense 4 KB tile per K-step
//         L2  128-byte rows, SWIZZLE_128B (half of every row unused)         (knn_tc.cu layout)
//       and for N = 128 (two row tiles) / N = 256 (one row tile)
//   (B) tcgen05.ld read-out rate alone (x16 / x32, 8 or 16 warps, with / without the FMNMX3 tree)
//   (C) both at once, with no dependency between them: do they interfere?
// One CTA per SM; cycles per "tile step" (256 rows x 128 candidates = 32768 filter values) from clock64.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tc_probe2 tools/tc_probe2.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc(unsigned* slot_in_smem, int cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned addr, int cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"l"((unsigned long long)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait32(unsigned (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                   "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
                   "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),
                   "+r"(r[31])
                 :
                 : "memory");
}
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float min32(const unsigned (&r)[32]) {
    float t[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) t[i] = min3(__uint_as_float(r[3 * i]), __uint_as_float(r[3 * i + 1]), __uint_as_float(r[3 * i + 2]));
    t[10] = fminf(__uint_as_float(r[30]), __uint_as_float(r[31]));
    const float u0 = min3(t[0], t[1], t[2]), u1 = min3(t[3], t[4], t[5]), u2 = min3(t[6], t[7], t[8]), u3 = fminf(t[9], t[10]);
    return fminf(min3(u0, u1, u2), u3);
}
// layout 0: SW64 rows of 64 B; 1: SW32 rows of 32 B; 2: SW128 rows of 128 B
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr, int layout) {
    const unsigned long long sbo = layout == 0 ? 32ull : layout == 1 ? 16ull : 64ull;
    const unsigned long long lt = layout == 0 ? 4ull : layout == 1 ? 6ull : 2ull;
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (lt << 61);
}


__device__ __forceinline__ void umma_f16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate), "r"(0u)
        : "memory");
}
// A operand in TMEM ("TS" form)
__device__ __forceinline__ void umma_tf32_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc),
        "r"(accumulate), "r"(0u)
        : "memory");
}

struct P3 {
    int iters, N, kind, ats, order, depth;   // kind 0 tf32 / 1 bf16; ats: A from TMEM; order 0: (r0k0,r0k1,r1k0,r1k1), 1: (r0k0,r1k0,r0k1,r1k1)
    long long* out;
};

__global__ void __launch_bounds__(64, 1) probe3_kernel(P3 p) {
    extern __shared__ unsigned char smem_raw_[];
    unsigned char* smem = smem_raw_ + ((1024u - (smem_u32(smem_raw_) & 1023u)) & 1023u);
    __shared__ unsigned long long bar[8];
    __shared__ unsigned s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (float)((i * 2654435761u) >> 22);
    if (tid == 0) {
        for (int s = 0; s < 8; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = s_tmem;
    if (warp == 1 && lane == 0) {
        const int N = p.N;
        // two "row tiles" (independent accumulators) x two K steps per step, accumulators double-buffered by step parity:
        // D columns: ((r*2 + parity) * N) — needs 4N <= 448 when A lives in TMEM columns 448..511
        const unsigned a0 = smem_u32(smem), b0 = a0 + 48 * 1024;
        const unsigned fmt = p.kind == 0 ? 2u : 1u;  // tf32 : bf16
        const unsigned idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        const long long t0 = clock64();
        for (int i = 0; i < p.iters; ++i) {
            if (i >= p.depth) mbar_wait(&bar[(i - p.depth) & 7], ((i - p.depth) >> 3) & 1);
            const unsigned stage = (unsigned)(i & 1) * (unsigned)(N * 64);
            for (int j = 0; j < 4; ++j) {
                const int r = p.order ? (j & 1) : (j >> 1), ks = p.order ? (j >> 1) : (j & 1);
                const unsigned d = tmem + (unsigned)((r * 2 + (i & 1)) * N);
                const unsigned long long bd = umma_desc(b0 + stage + ks * 32, 0);
                if (p.ats) umma_tf32_ts(d, tmem + 448u + (unsigned)(r * 16 + ks * 8), bd, idesc, ks > 0);
                else if (p.kind == 0) umma_tf32(d, umma_desc(a0 + r * 128 * 64 + ks * 32, 0), bd, idesc, ks > 0);
                else umma_f16(d, umma_desc(a0 + r * 128 * 64 + ks * 32, 0), bd, idesc, ks > 0);
            }
            umma_commit(&bar[i & 7]);
        }
        for (int i = p.iters > p.depth ? p.iters - p.depth : 0; i < p.iters; ++i) mbar_wait(&bar[i & 7], (i >> 3) & 1);
        p.out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

static void run(const char* name, int N, int kind, int ats, int order, int depth, int iters = 4000) {
    long long* d_out;
    const int grid = 148;
    CK(cudaMalloc(&d_out, grid * sizeof(long long)));
    CK(cudaMemset(d_out, 0, grid * sizeof(long long)));
    const size_t smem = 97 * 1024 + 1024;
    CK(cudaFuncSetAttribute(probe3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    P3 p{iters, N, kind, ats, order, depth, d_out};
    probe3_kernel<<<grid, 64, smem>>>(p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double m = 0;
    for (int i = 0; i < grid; ++i) m += (double)h[i];
    m /= grid * (double)iters;
    const double kel = kind == 0 ? 8 : 16;
    printf("%-44s N=%3d  %7.1f clk / 4 MMAs = %6.1f clk per MMA  -> %7.1f MAC/clk/SM, %6.1f accumulator values/clk/SM (K=2 steps)\n", name, N, m, m / 4,
           4 * 128.0 * N * kel / m, 2 * 128.0 * N / m);
    CK(cudaFree(d_out));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d\n", prop.name, prop.multiProcessorCount);
    for (int N : {32, 64, 96, 112, 128, 192, 256}) if (4 * N <= 512) run("tf32 SS, order r-major", N, 0, 0, 0, 2);
    for (int N : {64, 128}) run("tf32 SS, order k-major (independent pairs)", N, 0, 0, 1, 2);
    for (int N : {32, 64, 96, 112}) run("tf32 TS (A in TMEM), order r-major", N, 0, 1, 0, 2);
    for (int N : {64, 112}) run("tf32 TS (A in TMEM), order k-major", N, 0, 1, 1, 2);
    for (int N : {64, 128}) run("bf16 SS (K=16 per MMA), order r-major", N, 1, 0, 0, 2);
    return 0;
}
