// tc_probe2.cu — development microbenchmark: what bounds the tensor-core chamfer sweep?
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

struct P2 {
    int iters, mode;    // mode bit 0: MMA loop, bit 1: LD loop, bit 2: LD loop with the FMNMX3 tree
    int layout, n256;   // operand layout; N = 256 (one row tile) instead of N = 128 (two row tiles)
    int depth;          // tile steps the MMA issuer may run ahead of their completion
    long long* out;     // [grid][2] cycles of the MMA loop / of the slowest read-out warp
    float* sink;
};

template <int EPI>
__global__ void __launch_bounds__((EPI + 1) * 32, 1) probe2_kernel(P2 p) {
    extern __shared__ unsigned char smem_raw_[];
    unsigned char* smem = smem_raw_ + ((1024u - (smem_u32(smem_raw_) & 1023u)) & 1023u);
    __shared__ unsigned long long bar[8];
    __shared__ unsigned s_tmem;
    __shared__ long long s_cyc[EPI];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // operands: any finite data will do
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
    if (warp == EPI) {
        if (lane == 0 && (p.mode & 1)) {
            const int rowb = p.layout == 0 ? 64 : p.layout == 1 ? 32 : 128;
            const unsigned a0 = smem_u32(smem), b0 = a0 + 48 * 1024;
            const int N = p.n256 ? 256 : 128, RT = p.n256 ? 1 : 2;
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
            const long long t0 = clock64();
            for (int i = 0; i < p.iters; ++i) {
                if (i >= p.depth) mbar_wait(&bar[(i - p.depth) & 7], ((i - p.depth) >> 3) & 1);
                const unsigned stage = (unsigned)(i & 1) * (unsigned)(N * rowb * (p.layout == 1 ? 2 : 1));
                for (int r = 0; r < RT; ++r)
                    for (int ks = 0; ks < 2; ++ks) {
                        // layout 1: the two K-steps are separate dense tiles; otherwise +32 B inside the row
                        const unsigned ao = p.layout == 1 ? (unsigned)((r * 2 + ks) * 128 * 32) : (unsigned)(r * 128 * rowb + ks * 32);
                        const unsigned bo = p.layout == 1 ? stage + (unsigned)(ks * N * 32) : stage + (unsigned)(ks * 32);
                        umma_tf32(tmem + (unsigned)((r * 2 + (i & 1)) * 128) * (p.n256 ? 0u : 1u) + (p.n256 ? (unsigned)((i & 1) * 256) : 0u),
                                  umma_desc(a0 + ao, p.layout), umma_desc(b0 + bo, p.layout), idesc, ks > 0);
                    }
                umma_commit(&bar[i & 7]);
            }
            for (int i = p.iters > p.depth ? p.iters - p.depth : 0; i < p.iters; ++i) mbar_wait(&bar[i & 7], (i >> 3) & 1);
            p.out[blockIdx.x * 2] = clock64() - t0;
        }
    } else if (p.mode & 6) {
        const int quad = warp & 3, part = warp >> 2;     // EPI / 4 warps share a lane quadrant: each reads its share of the 512 columns
        constexpr int PARTS = EPI / 4;
        float acc = 0.f;
        unsigned keep = 0;
        const long long t0 = clock64();
        for (int i = 0; i < p.iters; ++i) {
            // one tile step = 256 columns of this lane quadrant (two row tiles x 128), split over the PARTS warps of the quadrant
            const unsigned base = tmem + ((unsigned)(quad * 32) << 16) + (unsigned)((i & 1) * 256 + part * (256 / PARTS));
            unsigned v[2][32];
            tmem_ld32(base, v[0]);
#pragma unroll
            for (int q = 0; q < 256 / PARTS / 32; ++q) {
                tmem_wait32(v[q & 1]);
                if (q + 1 < 256 / PARTS / 32) tmem_ld32(base + (q + 1) * 32, v[(q + 1) & 1]);
                if (p.mode & 4) acc += min32(v[q & 1]);
                else keep ^= v[q & 1][0] ^ v[q & 1][31];
            }
        }
        const long long dt = clock64() - t0;
        if (lane == 0) s_cyc[warp] = dt;
        if (acc == 123.456f || keep == 0x12345u) p.sink[0] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0 && (p.mode & 6)) {
        long long m = 0;
        for (int w = 0; w < EPI; ++w) m = s_cyc[w] > m ? s_cyc[w] : m;
        p.out[blockIdx.x * 2 + 1] = m;
    }
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int EPI>
static void run(const char* name, int mode, int layout, int n256, int depth, int iters = 4000) {
    long long* d_out;
    float* d_sink;
    const int grid = 148;
    CK(cudaMalloc(&d_out, grid * 2 * sizeof(long long)));
    CK(cudaMalloc(&d_sink, 4));
    CK(cudaMemset(d_out, 0, grid * 2 * sizeof(long long)));
    auto k = probe2_kernel<EPI>;
    const size_t smem = 97 * 1024 + 1024;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    P2 p{iters, mode, layout, n256, depth, d_out, d_sink};
    k<<<grid, (EPI + 1) * 32, smem>>>(p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(grid * 2);
    CK(cudaMemcpy(h.data(), d_out, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
    double m = 0, l = 0;
    for (int i = 0; i < grid; ++i) { m += (double)h[2 * i]; l += (double)h[2 * i + 1]; }
    m /= grid * (double)iters; l /= grid * (double)iters;
    printf("%-58s EPI=%2d  MMA %7.1f clk/step (%6.1f values/clk/SM)   read-out %7.1f clk/step (%6.1f values/clk/SM)\n", name, EPI, m, m > 0 ? 32768.0 / m : 0.0, l,
           l > 0 ? 32768.0 / l : 0.0);
    CK(cudaFree(d_out)); CK(cudaFree(d_sink));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d   (one tile step = 4 MMAs M128 N128 K8 or 2 MMAs M128 N256 K8 = 32768 accumulator values)\n", prop.name, prop.multiProcessorCount);
    run<8>("MMA only, SW64 64B rows, N=128, depth 2", 1, 0, 0, 2);
    run<8>("MMA only, SW64 64B rows, N=128, depth 6", 1, 0, 0, 6);
    run<8>("MMA only, SW32 dense K-step tiles, N=128, depth 2", 1, 1, 0, 2);
    run<8>("MMA only, SW128 128B rows, N=128, depth 2", 1, 2, 0, 2);
    run<8>("MMA only, SW64 64B rows, N=256, depth 2", 1, 0, 1, 2);
    run<8>("MMA only, SW32 dense K-step tiles, N=256, depth 2", 1, 1, 1, 2);
    run<8>("read-out only (no ALU)", 2, 0, 0, 2);
    run<16>("read-out only (no ALU)", 2, 0, 0, 2);
    run<8>("read-out only + FMNMX3 tree", 4, 0, 0, 2);
    run<16>("read-out only + FMNMX3 tree", 4, 0, 0, 2);
    run<8>("MMA (SW64, N=128) + read-out (no ALU) concurrently", 3, 0, 0, 2);
    run<16>("MMA (SW64, N=128) + read-out (no ALU) concurrently", 3, 0, 0, 2);
    run<8>("MMA (SW64, N=128) + read-out + FMNMX3 concurrently", 5, 0, 0, 2);
    run<16>("MMA (SW64, N=128) + read-out + FMNMX3 concurrently", 5, 0, 0, 2);
    run<16>("MMA (SW32, N=128) + read-out + FMNMX3 concurrently", 5, 1, 0, 2);
    run<16>("MMA (SW32, N=256) + read-out + FMNMX3 concurrently", 5, 1, 1, 2);
    return 0;
}
