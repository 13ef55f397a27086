// tc_probe.cu — development probe (not part of the library): can the chamfer filter sweep run on tcgen05?
//
// One direction of the sweep as a K = 16 TF32 GEMM with the accumulator in TMEM:
//     f_ij = |q_i|² + |c_j|² - 2 q_i.c_j = sum_k R[i][k] C[j][k]
//     R[i] = {xh,xh,xl,xl, yh,yh,yl,yl | zh,zh,zl,zl, nh,nl,1,1}          (x = xh + xl: two TF32 pieces of an FP32 value)
//     C[j] = {Xh,Xl,Xh,Xl, Yh,Yl,Yh,Yl | Zh,Zl,Zh,Zl, 1,1,Nh,Nl}          (X = -2 x of the candidate)
// The probe answers three questions on a real B200:
//   1. how fast can the row minima (with the 32-column chunk locator b1 / c1 / b2 of chamfer.cu) be pulled out of TMEM
//      (tcgen05.ld + FMNMX3), against the MMA rate and the smem refill (cp.async.bulk) — modes 0 / 1 / 2;
//   2. how accurate is the tensor core's FP32 accumulation of the split products (error in units of u (|q|² + |c|²));
//   3. does the K-major SWIZZLE_64B operand image (64-byte rows) work like the SWIZZLE_128B one used by knn_tc.cu.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o build/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
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
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
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
// issue only: the registers are valid after tmem_ld_wait(r) (which carries them as in/out operands so that no consumer
// can be scheduled ahead of the wait)
__device__ __forceinline__ void tmem_ld32_issue(unsigned taddr, unsigned (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld_wait(unsigned (&r)[32]) {
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

// K-major operand descriptors.  SW = 1: SWIZZLE_64B, rows of 64 B, 8-row groups 512 B apart; SW = 0: SWIZZLE_128B, rows of
// 128 B (only the first 64 B hold data), 8-row groups 1024 B apart (the layout knn_tc.cu uses).
template <int SW>
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((SW ? 32ull : 64ull) << 32) | (1ull << 46) | ((SW ? 4ull : 2ull) << 61);
}

struct ProbeParams {
    const float* rowimg;  // [items][RT*128][ROWB/4]   query operand image (already in the shared-memory layout)
    const float* colimg;  // [B][M][ROWB/4]            candidate operand image
    int M;                // candidates per batch element (multiple of TN)
    int items_per_batch;
    float4* out;          // [items*RT*128] {b1, b2, c1, -}
    float* dump;          // null or [RT*128][M]: the whole filter matrix of item 0
    int mode;             // 0 full, 1 tcgen05.ld only (no minima), 2 no read-out at all (MMA + refill only)
};

template <int RT, int TN, int EW, int STAGES, int SW>
__global__ void __launch_bounds__((RT * EW + 2) * 32, 1) tcmin_kernel(ProbeParams p) {
    constexpr int ROWB = SW ? 64 : 128;
    constexpr int NEPI = RT * EW;            // read-out warps
    constexpr int CW = TN / (EW / 4);        // columns of a buffer per read-out warp
    constexpr int TCOLS = RT * 2 * TN;       // TMEM columns: two accumulators per row tile
    static_assert(TCOLS == 256 || TCOLS == 512 || TCOLS == 128, "TMEM allocation must be a power of two");
    extern __shared__ unsigned char smem_raw_[];
    unsigned char* smem = smem_raw_ + ((1024u - (smem_u32(smem_raw_) & 1023u)) & 1023u);
    unsigned char* s_a = smem;                              // [RT][128][ROWB]
    unsigned char* s_b = s_a + RT * 128 * ROWB;             // [STAGES][TN][ROWB]
    __shared__ unsigned long long full_b[STAGES], empty_b[STAGES], tfull[2], tempty[2], a_bar;
    __shared__ unsigned s_tmem;
    __shared__ float4 s_merge[NEPI * 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int item = blockIdx.x, b = item / p.items_per_batch;
    const int ntiles = p.M / TN;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], NEPI); }
        mbar_init(&a_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&s_tmem, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = s_tmem;

    if (warp == NEPI) {
        // ---- producer: the query tiles once, then the candidate tiles through the ring -----------------------------
        if (lane == 0) {
            mbar_expect_tx(&a_bar, RT * 128 * ROWB);
            tma_bulk_g2s(s_a, p.rowimg + (size_t)item * RT * 128 * (ROWB / 4), RT * 128 * ROWB, &a_bar);
            const float* src = p.colimg + (size_t)b * p.M * (ROWB / 4);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % STAGES, n = t / STAGES;
                mbar_wait(&empty_b[s], (n & 1) ^ 1);
                mbar_expect_tx(&full_b[s], TN * ROWB);
                tma_bulk_g2s(s_b + (size_t)s * TN * ROWB, src + (size_t)t * TN * (ROWB / 4), TN * ROWB, &full_b[s]);
            }
        }
    } else if (warp == NEPI + 1) {
        // ---- MMA issuer -------------------------------------------------------------------------------------------------
        if (lane == 0) {
            constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(TN >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
            mbar_wait(&a_bar, 0);
            const unsigned a0 = smem_u32(s_a), b0 = smem_u32(s_b);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % STAGES, n = t / STAGES, acc = t & 1, na = t >> 1;
                mbar_wait(&full_b[s], n & 1);
                mbar_wait(&tempty[acc], (na & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int r = 0; r < RT; ++r)
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        umma_tf32(tmem + (unsigned)((r * 2 + acc) * TN), umma_desc<SW>(a0 + r * 128 * ROWB + k * 32),
                                  umma_desc<SW>(b0 + s * TN * ROWB + k * 32), idesc, k > 0);
                umma_commit(&empty_b[s]);
                umma_commit(&tfull[acc]);
            }
        }
    } else {
        // ---- read-out: thread <-> query row (TMEM lane), CW columns of every accumulator --------------------------------
        const int quad = warp & 3, r = (warp >> 2) % RT, part = warp / (4 * RT);
        float b1 = INFINITY, b2 = INFINITY;
        int c1 = 0;
        unsigned keep = 0u;
        for (int t = 0; t < ntiles; ++t) {
            const int acc = t & 1, na = t >> 1;
            mbar_wait(&tfull[acc], na & 1);
            tc_fence_after();
            if (p.mode != 2) {
                const unsigned base = tmem + ((unsigned)(quad * 32) << 16) + (unsigned)((r * 2 + acc) * TN + part * CW);
                unsigned v[2][32];
                tmem_ld32_issue(base, v[0]);
#pragma unroll
                for (int k = 0; k < CW / 32; ++k) {
                    tmem_ld_wait(v[k & 1]);
                    if (k + 1 < CW / 32) tmem_ld32_issue(base + (k + 1) * 32, v[(k + 1) & 1]);
                    if (p.mode == 0) {
                        const float m = min32(v[k & 1]);
                        const int chunk = (t * TN + part * CW) / 32 + k;
                        b2 = fminf(b2, fmaxf(b1, m));
                        c1 = m < b1 ? chunk : c1;
                        b1 = fminf(b1, m);
                        if (p.dump && item == 0) {
                            float* d = p.dump + (size_t)(r * 128 + quad * 32 + lane) * p.M + (size_t)chunk * 32;
#pragma unroll
                            for (int i = 0; i < 32; ++i) d[i] = __uint_as_float(v[k & 1][i]);
                        }
                    } else {
                        keep ^= v[k & 1][0] ^ v[k & 1][31];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (p.mode == 1) c1 = (int)keep;
        s_merge[warp * 32 + lane] = make_float4(b1, b2, __int_as_float(c1), 0.f);
    }
    tc_fence_before();
    __syncthreads();
    if (warp < 4 * RT) {
        // merge the column parts of a row (ascending column order: part 0 holds the lower chunk ids of every tile — the
        // probe only checks b1 / b2 and that c1 points at a chunk holding b1)
        float4 best = s_merge[warp * 32 + lane];
#pragma unroll
        for (int q = 1; q < EW / 4; ++q) {
            const float4 o = s_merge[(warp + q * 4 * RT) * 32 + lane];
            if (o.x < best.x) { best.y = fminf(best.x, o.y); best.x = o.x; best.z = o.z; }
            else best.y = fminf(best.y, o.x);
        }
        const int quad = warp & 3, r = warp >> 2;
        p.out[((size_t)item * RT + r) * 128 + quad * 32 + lane] = best;
    }
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

// ---------------------------------------------------------------------------------------------------------------------
static float rn_tf32(float x) {  // round to 11 significant bits (ties away: still a valid TF32 value, same error bound)
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xffffe000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}
static void split(float x, float& h, float& l) { h = rn_tf32(x); l = rn_tf32(x - h); }

// operand row -> 16 floats; query role or candidate role
static void make_row(const float* pt, bool candidate, float out[16]) {
    const float n = fmaf(pt[2], pt[2], fmaf(pt[1], pt[1], pt[0] * pt[0]));
    float nh, nl;
    split(n, nh, nl);
    for (int d = 0; d < 3; ++d) {
        float h, l;
        split(candidate ? -2.0f * pt[d] : pt[d], h, l);
        float* o = out + 4 * d;
        if (candidate) { o[0] = h; o[1] = l; o[2] = h; o[3] = l; }
        else { o[0] = h; o[1] = h; o[2] = l; o[3] = l; }
    }
    if (candidate) { out[12] = 1.f; out[13] = 1.f; out[14] = nh; out[15] = nl; }
    else { out[12] = nh; out[13] = nl; out[14] = 1.f; out[15] = 1.f; }
}
// place 16 floats of row r (index within its tile-independent 8-row group pattern) into the image
static void put_row(float* img, size_t row, const float v[16], int SW) {
    const int rowf = SW ? 16 : 32;
    float* base = img + row * rowf;
    for (int c = 0; c < 4; ++c) {
        const int cc = SW ? (c ^ (int)((row >> 1) & 3)) : (c ^ (int)(row & 7));
        memcpy(base + 4 * cc, v + 4 * c, 16);
    }
}

template <int RT, int TN, int EW, int STAGES, int SW>
static void run_config(const char* name, int B, int N, int M, const std::vector<float>& A, const std::vector<float>& C, bool check) {
    constexpr int ROWB = SW ? 64 : 128;
    const int rowf = ROWB / 4;
    const int items_per_batch = N / (RT * 128), items = B * items_per_batch;
    std::vector<float> rimg((size_t)B * N * rowf, 0.f), cimg((size_t)B * M * rowf, 0.f);
    for (size_t i = 0; i < (size_t)B * N; ++i) { float v[16]; make_row(&A[3 * i], false, v); put_row(rimg.data(), i, v, SW); }
    for (size_t j = 0; j < (size_t)B * M; ++j) { float v[16]; make_row(&C[3 * j], true, v); put_row(cimg.data(), j, v, SW); }
    float *d_r, *d_c, *d_dump = nullptr;
    float4* d_out;
    CK(cudaMalloc(&d_r, rimg.size() * 4)); CK(cudaMalloc(&d_c, cimg.size() * 4)); CK(cudaMalloc(&d_out, (size_t)B * N * sizeof(float4)));
    CK(cudaMemcpy(d_r, rimg.data(), rimg.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_c, cimg.data(), cimg.size() * 4, cudaMemcpyHostToDevice));
    if (check) CK(cudaMalloc(&d_dump, (size_t)RT * 128 * M * 4));
    auto kern = tcmin_kernel<RT, TN, EW, STAGES, SW>;
    const size_t smem = (size_t)RT * 128 * ROWB + (size_t)STAGES * TN * ROWB + 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = (RT * EW + 2) * 32;
    ProbeParams p{d_r, d_c, M, items_per_batch, d_out, d_dump, 0};
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("== %s: RT=%d TN=%d EW=%d STAGES=%d SW=%d  items=%d threads=%d smem=%zu occupancy(blocks/SM by smem+regs)=%d\n", name, RT, TN, EW, STAGES, SW, items,
           threads, smem, occ);
    for (int mode = 0; mode < 3; ++mode) {
        p.mode = mode;
        p.dump = (mode == 0) ? d_dump : nullptr;
        kern<<<items, threads, smem>>>(p);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        p.dump = nullptr;
        for (int i = 0; i < 3; ++i) kern<<<items, threads, smem>>>(p);
        CK(cudaEventRecord(e0));
        const int reps = 20;
        for (int i = 0; i < reps; ++i) kern<<<items, threads, smem>>>(p);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / reps, pd = (double)B * N * M;
        printf("   mode %d (%s): %8.2f us  %.3e pair-dirs/s  = %.1f pair-dirs/clk/SM @1.965GHz (both directions: %.3e pairs/s)\n", mode,
               mode == 0 ? "full" : mode == 1 ? "ld only" : "mma+refill only", us, pd / (us * 1e-6), pd / (us * 1e-6) / 148 / 1.965e9, pd / (us * 1e-6) / 2);
    }
    if (check) {
        // accuracy of the filter matrix of item 0 against the exact squared distance of the same FP32 points
        std::vector<float> D((size_t)RT * 128 * M);
        std::vector<float4> out((size_t)B * N);
        p.mode = 0; p.dump = d_dump;
        kern<<<items, threads, smem>>>(p);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), d_dump, D.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(float4), cudaMemcpyDeviceToHost));
        const double u = ldexp(1.0, -24);
        double maxe = 0, sume = 0, maxe_split = 0;
        size_t bad = 0;
        for (int i = 0; i < RT * 128; ++i)
            for (int j = 0; j < M; ++j) {
                const float* a = &A[3 * (size_t)i];
                const float* c = &C[3 * (size_t)j];
                double d = 0, na = 0, nc = 0;
                for (int k = 0; k < 3; ++k) { d += ((double)a[k] - c[k]) * ((double)a[k] - c[k]); na += (double)a[k] * a[k]; nc += (double)c[k] * c[k]; }
                // what the 16 products sum to in exact arithmetic (isolates the tensor core's own accumulation error)
                float rv[16], cv[16];
                make_row(a, false, rv); make_row(c, true, cv);
                double ex = 0;
                for (int k = 0; k < 16; ++k) ex += (double)rv[k] * cv[k];
                const double e = fabs((double)D[(size_t)i * M + j] - d) / (u * (na + nc));
                const double es = fabs((double)D[(size_t)i * M + j] - ex) / (u * (na + nc));
                if (!(e < 1e6)) ++bad;
                maxe = std::max(maxe, e); sume += e; maxe_split = std::max(maxe_split, es);
            }
        printf("   accuracy (item 0, %d x %d): max |f - d| = %.2f u(na+nc), mean %.3f; against the exact sum of the 16 split products: max %.2f u(na+nc); garbage entries %zu\n",
               RT * 128, M, maxe, sume / ((double)RT * 128 * M), maxe_split, bad);
        // (b1, c1, b2) of item 0 against the dumped matrix
        size_t wrong = 0;
        for (int i = 0; i < RT * 128; ++i) {
            float b1 = INFINITY, b2 = INFINITY;
            int c1 = 0;
            for (int ch = 0; ch < M / 32; ++ch) {
                float m = INFINITY;
                for (int j = 0; j < 32; ++j) m = fminf(m, D[(size_t)i * M + ch * 32 + j]);
                b2 = fminf(b2, fmaxf(b1, m));
                if (m < b1) c1 = ch;
                b1 = fminf(b1, m);
            }
            const float4 o = out[i];
            int oc;
            memcpy(&oc, &o.z, 4);
            if (o.x != b1 || o.y != b2 || (EW == 4 && oc != c1)) { if (wrong < 4) printf("   row %d: got (%g, %g, %d) want (%g, %g, %d)\n", i, o.x, o.y, oc, b1, b2, c1); ++wrong; }
        }
        printf("   locator triples of item 0: %zu of %d rows differ from the dumped matrix\n", wrong, RT * 128);
        // every row of every item: b1 against the exact minimum
        double maxrow = 0;
        for (int bb = 0; bb < B; ++bb)
            for (int i = 0; i < N; i += 37) {
                const float* a = &A[3 * ((size_t)bb * N + i)];
                double best = 1e300, na = 0, ncm = 0;
                for (int k = 0; k < 3; ++k) na += (double)a[k] * a[k];
                for (int j = 0; j < M; ++j) {
                    const float* c = &C[3 * ((size_t)bb * M + j)];
                    double d = 0, nc = 0;
                    for (int k = 0; k < 3; ++k) { d += ((double)a[k] - c[k]) * ((double)a[k] - c[k]); nc += (double)c[k] * c[k]; }
                    best = std::min(best, d); ncm = std::max(ncm, nc);
                }
                maxrow = std::max(maxrow, fabs((double)out[(size_t)bb * N + i].x - best) / (u * (na + ncm)));
            }
        printf("   row minima over all items (every 37th row): max |b1 - min d| = %.2f u(na + max nc)\n", maxrow);
    }
    CK(cudaFree(d_r)); CK(cudaFree(d_c)); CK(cudaFree(d_out));
    if (d_dump) CK(cudaFree(d_dump));
}

int main(int argc, char** argv) {
    int B = argc > 1 ? atoi(argv[1]) : 37, N = 4096, M = 4096;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d\n", prop.name, prop.multiProcessorCount);
    std::mt19937 rng(12345);
    std::uniform_real_distribution<float> uni(-0.5f, 0.5f);  // centred unit cube
    std::vector<float> A((size_t)B * N * 3), C((size_t)B * M * 3);
    for (auto& x : A) x = uni(rng);
    for (auto& x : C) x = uni(rng);
    // known-good operand layout first (SWIZZLE_128B as in knn_tc.cu), then the 64-byte rows
    run_config<1, 128, 4, 4, 0>("sw128 1x128", B, N, M, A, C, true);
    run_config<1, 128, 4, 4, 1>("sw64  1x128", B, N, M, A, C, true);
    run_config<1, 256, 4, 4, 1>("sw64  1x256 (4 read-out warps)", B, N, M, A, C, true);
    run_config<1, 256, 8, 4, 1>("sw64  1x256 (8 read-out warps)", B, N, M, A, C, true);
    run_config<2, 128, 4, 4, 1>("sw64  2x128 (8 read-out warps)", B, N, M, A, C, true);
    run_config<2, 128, 8, 4, 1>("sw64  2x128 (16 read-out warps)", B, N, M, A, C, true);
    return 0;
}
