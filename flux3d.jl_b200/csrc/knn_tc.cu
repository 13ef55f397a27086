// knn_tc.cu — kNN graph with the Gram matrix on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as knn.cu (src/models/dgcnn.jl:3-9, :32-45; bit-exact (distance, index) order, first hit dropped by
// position), different shape of work.  The N x F . F x N contraction of a cloud against itself is the one genuinely
// dense product on this hot path (F = 64 in EdgeConv2): 128 queries x 128 candidates x F per tcgen05.mma batch,
// accumulators in TMEM, operands staged in shared memory in the canonical K-major SWIZZLE_128B layout.
//
// The tensor cores see TF32 (10-bit mantissa), so the Gram matrix is only a FILTER — exactly the certify-and-
// re-evaluate scheme of the chamfer sweep:
//   d~_ij = |x_i|² + |x_j|² - 2 G~_ij,   |d~ - d| <= E_i := 2^-7.5 |x_i| max_j|x_j| + 2 (F+4) u (|x_i|² + max_j|x_j|²)
//   (truncation to TF32 costs at most 2^-10 relative per operand; FP32 accumulation of <= 64 exact products adds < 2^-17;
//   the second term covers the FP32 roundings of the norms and of the final fma, which do not vanish with |x_i|)
// FOUR THREADS own one query (its TMEM lane; thread = (lane quadrant warp%4, column quarter warp/4) of a 16-warp CTA):
//   pass 1  stream the row of d~ out of TMEM, 32 columns of every 128-column tile per thread; keep the two smallest of
//           every 32-column group; the 64 local minima of a query are exchanged through shared memory and a
//           64-input bitonic network gives T~ = the (K+1)-th smallest of them — an upper bound of the (K+1)-th
//           smallest d~, because they are a subset;
//   pass 2  recompute the Gram tiles (the MMA is ~free) and collect every candidate with d~ <= T~ + 2 E_i: a provable
//           superset of the exact (K+1)-nearest set (typically K+5 .. K+15 candidates);
//   exact   re-evaluate those candidates in the reference arithmetic (s = s + (a_d - b_d)², every op rounded); every
//           candidate then finds its own rank among the 64-bit (distance bits << 32 | index) keys by counting and
//           writes itself to its output slot.
// All 32 lanes of a warp run the same data-independent instruction stream on 32 different queries (no shuffles, no
// divergence in passes 1-2); a query whose candidate set overflows its 64 slots (heavy ties) falls back to an exact
// scan of the whole cloud.  Clouds with N > 1024, F > 64 or K > 31 use the CUDA-core kernel in knn.cu.
#include <algorithm>

#include "f3d_common.cuh"

namespace f3d {
namespace {

constexpr int kTQ = 128;          // queries per CTA = UMMA M = TMEM lanes
constexpr int kTN = 128;          // candidates per MMA tile = UMMA N = TMEM columns of the accumulator
constexpr int kTcThreads = 544;   // 16 read-out warps: lane quadrant (warp % 4) x column quarter (warp / 4); warp 16 issues the MMAs
constexpr int kCap = 128;         // candidate slots per query
constexpr int kMaxTiles = 8;      // N <= 1024
constexpr int kGroups = kMaxTiles * (kTN / 32);  // 32 groups of 32 columns -> 64 local minima (two per group) per query
constexpr float kErrRel = 0.005524272f;          // 2^-7.5

struct KnnTcParams {
    const float* X;   // [B][N][F]
    int N, F, Kp, K;  // Kp: F rounded up to a multiple of 32 (one SWIZZLE_128B atom = 32 floats of K)
    int32_t* idx;     // [B][N][K]
    float* dist;      // [B][N][K] or null
    unsigned* stats;  // optional [2]: queries that overflowed to the exact scan, candidates re-evaluated exactly
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void tmem_alloc(unsigned* slot_in_smem, int cols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned addr, int cols) {  // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {  // arrives on bar when all prior MMAs are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"l"((unsigned long long)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B canonical layout (cute::UMMA::make_umma_desc<Major::K>): a region is [rows][128 B]; the 16-byte
// chunk c of row r sits at r*128 + ((c ^ (r & 7)) << 4); 8-row groups are 1024 B apart (SBO); LBO = 1 (unused with swizzle)
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
constexpr unsigned kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(kTN >> 3) << 17) | ((unsigned)(kTQ >> 4) << 24);

// exact (reference-arithmetic) squared distance between a query held in the swizzled A tile (own row) and a global row
__device__ __forceinline__ float exact_dist(const unsigned char* s_a, int row, const float* __restrict__ xj, int F) {
    float s = 0.0f;
    for (int d0 = 0; d0 < F; d0 += 4) {
        const int h = d0 >> 5, c = (d0 >> 2) & 7;
        const float4 q = *reinterpret_cast<const float4*>(s_a + (size_t)h * (kTQ * 128) + row * 128 + ((c ^ (row & 7)) << 4));
        float t;
        t = __fsub_rn(q.x, __ldg(xj + d0)); s = __fadd_rn(s, __fmul_rn(t, t));
        if (d0 + 1 < F) { t = __fsub_rn(q.y, __ldg(xj + d0 + 1)); s = __fadd_rn(s, __fmul_rn(t, t)); }
        if (d0 + 2 < F) { t = __fsub_rn(q.z, __ldg(xj + d0 + 2)); s = __fadd_rn(s, __fmul_rn(t, t)); }
        if (d0 + 3 < F) { t = __fsub_rn(q.w, __ldg(xj + d0 + 3)); s = __fadd_rn(s, __fmul_rn(t, t)); }
    }
    return s;
}
// in-register bitonic sort of 64 floats (ascending), fully unrolled: data-independent, identical in every lane
__device__ __forceinline__ void sort64(float (&v)[64]) {
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const float a = v[i], b = v[l];
                    const bool up = (i & k) == 0;
                    v[i] = up ? fminf(a, b) : fmaxf(a, b);
                    v[l] = up ? fmaxf(a, b) : fminf(a, b);
                }
            }
        }
    }
}

__device__ __forceinline__ void cpa4(void* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpa16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}

// Shared memory of one CTA:  A tile | 3 B tiles (later: exact distances) | candidate ids | norms x2 | group minima | counters
__global__ void __launch_bounds__(kTcThreads, 1) knn_tc_kernel(KnnTcParams p) {
    extern __shared__ unsigned char smem_raw_[];
    unsigned char* smem = smem_raw_ + ((1024u - (smem_u32(smem_raw_) & 1023u)) & 1023u);  // SWIZZLE_128B atoms need 1024 B alignment (pointer arithmetic keeps LDS/STS)
    const int halves = p.Kp >> 5;                                   // K atoms of 32 floats
    const size_t b_tile = (size_t)halves * kTN * 128;               // one staged candidate tile
    unsigned char* s_a = smem;                                      // [halves][128 rows][128 B]
    unsigned char* s_b = s_a + (size_t)halves * kTQ * 128;          // 3 x [halves][128 rows][128 B]
    const size_t b_region = max(3 * b_tile, sizeof(float) * kCap * kTQ);
    float* s_dex = reinterpret_cast<float*>(s_b);                   // [kCap][128] exact distances, reuses the B region after pass 2
    unsigned short* s_cand = reinterpret_cast<unsigned short*>(s_b + b_region);  // [kCap][128] candidate ids (N <= 1024)
    float* s_nc = reinterpret_cast<float*>(s_cand + kCap * kTQ);    // [1024] squared norms of all candidates of the cloud (+inf past N)
    float* s_gm = s_nc + kMaxTiles * kTN;                           // [2 kGroups][128] two smallest d~ of every group, per query
    int* s_cnt = reinterpret_cast<int*>(s_gm + 2 * kGroups * kTQ);  // [128] candidates collected per query
    __shared__ unsigned long long s_bar[2];
    __shared__ unsigned s_tmem, s_maxnc;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_mma = warp == 16;                  // the 17th warp only issues tcgen05.mma (one elected lane)
    const int quad = warp & 3, qtr = (warp >> 2) & 3;  // TMEM lane quadrant / column quarter of every tile
    const int row = quad * 32 + lane;                // this thread's query (TMEM lane)
    const int b = blockIdx.y, q0 = blockIdx.x * kTQ;
    const float* Xb = p.X + (size_t)b * p.N * p.F;
    const int ntiles = (p.N + kTN - 1) / kTN;
    const int ksteps = (p.F + 7) >> 3;             // MMA k-steps of 8 floats that hold data
    const int c4_used = 2 * ksteps;                // 16-byte chunks per row that the MMAs read
    const bool vec16 = (p.F & 3) == 0 && (reinterpret_cast<uintptr_t>(p.X) & 15) == 0;

    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&s_tmem, 2 * kTN);   // two accumulators: MMA(s+1) runs while tile s is being read
    if (tid < kTQ) s_cnt[tid] = 0;
    if (tid == 0) s_maxnc = 0u;
    // zero the chunks the MMAs read once: cp.async below only ever writes the valid floats, so the K padding stays 0
    for (int e = tid; e < (kTQ + 3 * kTN) * c4_used; e += kTcThreads) {
        const int r = e / c4_used, c4 = e - r * c4_used, h = c4 >> 3, c = c4 & 7;
        unsigned char* base = r < kTQ ? s_a + (size_t)h * (kTQ * 128) + r * 128
                                      : s_b + (size_t)((r - kTQ) / kTN) * b_tile + (size_t)h * (kTN * 128) + ((r - kTQ) % kTN) * 128;
        *reinterpret_cast<float4*>(base + ((c ^ (r & 7)) << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    // asynchronous staging of `rows` rows starting at global row g0 into a swizzled region (rows past N: row N-1 for A)
    auto stage = [&](unsigned char* region, int region_rows, int g0, bool clamp) {
        if (vec16) {
            const int c4n = p.F >> 2;
            const int sh = (c4n & (c4n - 1)) == 0 ? __ffs(c4n) - 1 : -1;   // power-of-two row length: shifts instead of a division
            for (int e = tid; e < region_rows * c4n; e += kTcThreads) {
                const int r = sh >= 0 ? (e >> sh) : e / c4n, c4 = e - r * c4n;
                int g = g0 + r;
                if (g >= p.N) { if (!clamp) continue; g = p.N - 1; }
                cpa16(region + (size_t)(c4 >> 3) * (region_rows * 128) + r * 128 + (((c4 & 7) ^ (r & 7)) << 4), Xb + (size_t)g * p.F + 4 * c4);
            }
        } else {
            for (int e = tid; e < region_rows * p.F; e += kTcThreads) {
                const int r = e / p.F, d = e - r * p.F, c4 = d >> 2;
                int g = g0 + r;
                if (g >= p.N) { if (!clamp) continue; g = p.N - 1; }
                cpa4(region + (size_t)(c4 >> 3) * (region_rows * 128) + r * 128 + (((c4 & 7) ^ (r & 7)) << 4) + ((d & 3) << 2), Xb + (size_t)g * p.F + d);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // squared norm of a staged row (exact FP32 bits are in the tile; any rounding will do — it only enters the filter)
    auto row_norm = [&](const unsigned char* region, int region_rows, int r) {
        float n = 0.0f;
        for (int c4 = 0; c4 < c4_used; ++c4) {
            const float4 q = *reinterpret_cast<const float4*>(region + (size_t)(c4 >> 3) * (region_rows * 128) + r * 128 + (((c4 & 7) ^ (r & 7)) << 4));
            n += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
        }
        return n;
    };

    // squared norms of all candidates of the cloud, once (any rounding will do — they only enter the filter)
    {
        float mymax = 0.0f;
        for (int j = tid; j < kMaxTiles * kTN; j += kTcThreads) {
            float n = INFINITY;
            if (j < p.N) {
                const float* x = Xb + (size_t)j * p.F;
                n = 0.0f;
                if (vec16) {
#pragma unroll 4
                    for (int d = 0; d < p.F; d += 4) { const float4 q = __ldg(reinterpret_cast<const float4*>(x + d)); n += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w; }
                } else {
                    for (int d = 0; d < p.F; ++d) { const float q = __ldg(x + d); n += q * q; }
                }
                mymax = fmaxf(mymax, n);
            }
            s_nc[j] = n;
        }
        atomicMax(&s_maxnc, __float_as_uint(mymax));  // norms are >= 0: bit order == value order
    }
    // The tile sequence is pass 1 (tiles 0..ntiles-1) followed by pass 2 (the same tiles again): L steps, software-
    // pipelined — step s: [cp.async tile s+2] | norms + MMA of tile s | read-out (TMEM) of tile s-1.
    const int L = 2 * ntiles;
    stage(s_a, kTQ, q0, true);                      // group 0: A (it is waited for together with step 0)
    for (int u = 0; u < 3; ++u) {                   // groups 1..3: sequence steps 0, 1, 2 (all three B buffers are free)
        if (u < L) stage(s_b + (size_t)u * b_tile, kTN, (u % ntiles) * kTN, false);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }

    float thr = 0.0f, nq = 0.0f;
    unsigned tmem = 0;
    unsigned phbits = 0u;  // phase parity of the two mbarriers (bit u & 1)
    // read-out of sequence step u (its MMA batch has been committed to s_bar[u & 1])
    auto read_out = [&](int u) {
        mbar_wait(&s_bar[u & 1], (phbits >> (u & 1)) & 1u);
        phbits ^= 1u << (u & 1);
        tc_fence_after();
        // the B buffer of step u is free again: start staging sequence step u + 3
        if (u + 3 < L) stage(s_b + (size_t)((u + 3) % 3) * b_tile, kTN, ((u + 3) % ntiles) * kTN, false);
        else asm volatile("cp.async.commit_group;" ::: "memory");   // keep the group count uniform
        if (is_mma) return;
        const int t = u % ntiles;
        const float* nc = s_nc + t * kTN;
        float v[32];
        tmem_ld32(tmem + ((unsigned)(quad * 32) << 16) + (unsigned)((u & 1) * kTN + qtr * 32), v);
        if (u < ntiles) {  // pass 1: the two smallest d~ of this thread's 32-column group
            float m1 = INFINITY, m2 = INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float n = nc[qtr * 32 + i];
                const float d = fmaf(-2.0f, v[i], nq + n);
                m2 = fminf(m2, fmaxf(m1, d));
                m1 = fminf(m1, d);
            }
            s_gm[(2 * (t * 4 + qtr)) * kTQ + row] = m1;
            s_gm[(2 * (t * 4 + qtr) + 1) * kTQ + row] = m2;
        } else {           // pass 2: collect the candidate superset
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float n = nc[qtr * 32 + i];
                const float d = fmaf(-2.0f, v[i], nq + n);
                if (n < INFINITY && !(d > thr)) {
                    const int slot = atomicAdd(&s_cnt[row], 1);
                    if (slot < kCap) s_cand[slot * kTQ + row] = (unsigned short)(t * kTN + qtr * 32 + i);
                }
            }
        }
    };

    // MMA descriptors of the A tile (constant) — K = 8 floats = 32 B per step, 4 steps per 128-byte atom
    const unsigned a0 = smem_u32(s_a), b00 = smem_u32(s_b);
#pragma unroll 1
    for (int sidx = 0; sidx < L; ++sidx) {
        // step `sidx` has landed.  Groups in commit order: A, steps 0, 1, 2, then one per read_out (step u + 3), so when
        // step sidx is needed at most steps sidx+1 (and, at sidx = 0, sidx+2) are younger
        if (sidx == 0) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else asm volatile("cp.async.wait_group 1;" ::: "memory");
        proxy_fence_async();  // cp.async smem writes -> visible to the tensor core (async proxy)
        tc_fence_before();    // every earlier tcgen05.ld of this thread is ordered before the barrier
        __syncthreads();      // ... so the accumulator that MMA(sidx) overwrites has been read out (step sidx-2)
        if (sidx == 0) { tc_fence_after(); tmem = s_tmem; if (!is_mma) nq = row_norm(s_a, kTQ, row); }
        if (is_mma) {
            if (lane == 0) {
                tc_fence_after();
                const unsigned b0 = b00 + (unsigned)((sidx % 3) * b_tile), acc = tmem + (unsigned)((sidx & 1) * kTN);
                for (int k = 0; k < ksteps; ++k)
                    umma_tf32(acc, umma_desc(a0 + (k >> 2) * (kTQ * 128) + (k & 3) * 32), umma_desc(b0 + (k >> 2) * (kTN * 128) + (k & 3) * 32), kIdesc, k > 0);
                umma_commit(&s_bar[sidx & 1]);
            }
            __syncwarp();
            // the MMA warp also helps staging: it takes part in read_out's cp.async schedule below (same group counts)
        }
        if (sidx >= 1) read_out(sidx - 1);
        if (sidx == ntiles) {
            // pass 1 is complete (its last tile was just read out): exchange the local minima and fix the threshold
            __syncthreads();
            if (!is_mma) {
                float gm[2 * kGroups];
#pragma unroll
                for (int g = 0; g < 2 * kGroups; ++g) gm[g] = (g < 8 * ntiles) ? s_gm[g * kTQ + row] : INFINITY;
                sort64(gm);
                float Tsel = gm[0];
#pragma unroll
                for (int i = 1; i < 64; ++i) Tsel = (i == p.K) ? gm[i] : Tsel;  // (K+1)-th smallest, K < 64
                // collect threshold: T~ + 2 E (E bounds |d~ - d|); not finite => collect everything (and overflow to the exact scan)
                // ... plus an ABSOLUTE term for the filter's own FP32 roundings, which do not shrink with |x_i|: the norms are
                // summed in another order than the exact distance (<= F u each, relative), the add nq + n and the fma round once
                // more — <= (F + 3) u (nq + max n), doubled for slack.  Without it a query of zero / tiny norm (zero-padded
                // points, dead ReLU features) got thr == Tsel and a near-tied true neighbour could be filtered out.
                const float maxnc = __uint_as_float(s_maxnc);
                thr = Tsel + 2.0f * (kErrRel * sqrtf(nq) * sqrtf(maxnc) + 2.0f * (float)(p.F + 4) * 5.9604645e-8f * (nq + maxnc));
            }
        }
    }
    read_out(L - 1);
    tc_fence_before();
    __syncthreads();  // all TMEM reads and MMAs are done; the B region may be reused
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 2 * kTN); }

    // ---- exact re-evaluation + ordered selection ----------------------------------------------------------------
    const int qi = q0 + row;
    const int cnt = s_cnt[row];
    const bool live = qi < p.N && !is_mma;
    const size_t obase = ((size_t)b * p.N + min(qi, p.N - 1)) * p.K;
    if (live && qtr == 0 && p.stats) { if (cnt > kCap) atomicAdd(p.stats, 1u); atomicAdd(p.stats + 1, (unsigned)min(cnt, kCap)); }
    int ncand = cnt;  // entries of (s_dex, s_cand) that take part in the ranking
    if (live && cnt <= kCap) {
        // exact distances: this query's candidates are split over its 4 threads, 2 in flight per thread
        for (int s = qtr; s < cnt; s += 8) {
            int j[2];
            float acc[2] = {0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 2; ++u) j[u] = s_cand[min(s + 4 * u, cnt - 1) * kTQ + row];
#pragma unroll 8
            for (int d0 = 0; d0 < p.F; d0 += 4) {
                const int h = d0 >> 5, c = (d0 >> 2) & 7;
                const float4 q = *reinterpret_cast<const float4*>(s_a + (size_t)h * (kTQ * 128) + row * 128 + ((c ^ (row & 7)) << 4));
                float x[2][4];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float* xj = Xb + (size_t)j[u] * p.F + d0;
                    if (vec16) { const float4 w = __ldg(reinterpret_cast<const float4*>(xj)); x[u][0] = w.x; x[u][1] = w.y; x[u][2] = w.z; x[u][3] = w.w; }
                    else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) x[u][e] = (d0 + e < p.F) ? __ldg(xj + e) : 0.0f;
                    }
                }
                const float qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (d0 + e < p.F) { const float tt = __fsub_rn(qq[e], x[u][e]); acc[u] = __fadd_rn(acc[u], __fmul_rn(tt, tt)); }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (s + 4 * u < cnt) s_dex[(s + 4 * u) * kTQ + row] = acc[u];
        }
    } else if (live) {
        // candidate set overflowed (heavy ties): every thread of the query scans a quarter of the cloud exactly and keeps
        // its K+1 best in a sorted (distance, index) list of its own 32 slots; the ranking below merges the four lists
        const int base = qtr * 32, cap = p.K + 1;
        int have = 0;
        for (int j = qtr; j < p.N; j += 4) {
            const float d = exact_dist(s_a, row, Xb + (size_t)j * p.F, p.F);
            if (have == cap) {
                const float dl = s_dex[(base + cap - 1) * kTQ + row];
                if (d > dl || (d == dl && j > (int)s_cand[(base + cap - 1) * kTQ + row])) continue;
            }
            int pos = have < cap ? have : cap - 1;
            while (pos > 0) {
                const float dp = s_dex[(base + pos - 1) * kTQ + row];
                const int jp = s_cand[(base + pos - 1) * kTQ + row];
                if (!(dp > d || (dp == d && jp > j))) break;
                s_dex[(base + pos) * kTQ + row] = dp;
                s_cand[(base + pos) * kTQ + row] = (unsigned short)jp;
                --pos;
            }
            s_dex[(base + pos) * kTQ + row] = d;
            s_cand[(base + pos) * kTQ + row] = (unsigned short)j;
            if (have < cap) ++have;
        }
        for (int s = have; s < 32; ++s) { s_dex[(base + s) * kTQ + row] = INFINITY; s_cand[(base + s) * kTQ + row] = 0xffff; }
        ncand = 128;
    }
    __syncthreads();
    if (!live) return;
    // every candidate finds its rank in the ascending (distance, index) order by counting; ranks 1..K are the
    // neighbours, rank 0 is dropped by position (dgcnn.jl:6)
    for (int s = qtr; s < ncand; s += 4) {
        const float d = s_dex[s * kTQ + row];
        const int j = s_cand[s * kTQ + row];
        if (!(d < INFINITY)) continue;
        int rank = 0;
#pragma unroll 4
        for (int o = 0; o < ncand; ++o) {  // no early exit: independent loads, 4 in flight
            const float dd = s_dex[o * kTQ + row];
            const int jj = s_cand[o * kTQ + row];
            rank += (dd < d || (dd == d && jj < j)) ? 1 : 0;
        }
        if (rank >= 1 && rank <= p.K) {
            p.idx[obase + rank - 1] = j;
            if (p.dist) p.dist[obase + rank - 1] = d;
        }
    }
}

// gathered (F,K,N,B) and edge features (2F,K,N,B) from the neighbour indices: pure data movement.  One CTA per point
// (b, i); its K*W output floats are contiguous, threads walk them in units of V floats (V = 4 when F % 4 == 0 and
// everything is 16-byte aligned), so stores are coalesced for every F and all index arithmetic is 32-bit.
template <int V>
__global__ void __launch_bounds__(256) knn_emit_kernel(const float* __restrict__ X, const int32_t* __restrict__ idx, int N, int F, int K,
                                                       float* __restrict__ gathered, float* __restrict__ edge) {
    const unsigned bi = blockIdx.x;                 // b*N + i
    const unsigned b = bi / (unsigned)N;
    const float* xi = X + (size_t)bi * F;
    const int32_t* nn = idx + (size_t)bi * K;
    const unsigned W = edge ? 2 * F : F, Wv = W / V, Fv = F / V;
    for (unsigned t = threadIdx.x; t < (unsigned)K * Wv; t += 256) {
        const unsigned k = t / Wv, cv = t - k * Wv;  // neighbour, V-float column
        const float* xj = X + ((size_t)b * N + __ldg(nn + k)) * F;
        float vi[V], vj[V], o[V];
        const bool second = edge && cv >= Fv;        // the (x_j - x_i) half of an edge row
        const unsigned c = (second ? cv - Fv : cv) * V;
        if (V == 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(xi + c)), w = __ldg(reinterpret_cast<const float4*>(xj + c));
            vi[0] = a.x; vi[1 % V] = a.y; vi[2 % V] = a.z; vi[3 % V] = a.w;
            vj[0] = w.x; vj[1 % V] = w.y; vj[2 % V] = w.z; vj[3 % V] = w.w;
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e) { vi[e] = __ldg(xi + c + e); vj[e] = __ldg(xj + c + e); }
        }
#pragma unroll
        for (int e = 0; e < V; ++e) o[e] = edge ? (second ? __fsub_rn(vj[e], vi[e]) : vi[e]) : vj[e];   // cat(X, KNNGraph - X; dims=1)  dgcnn.jl:45
        float* dst = (edge ? edge : gathered) + ((size_t)bi * K + k) * W + cv * V;
        if (V == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1 % V], o[2 % V], o[3 % V]);
        else {
#pragma unroll
            for (int e = 0; e < V; ++e) dst[e] = o[e];
        }
        if (edge && gathered && second) {
            float* g = gathered + ((size_t)bi * K + k) * F + c;
            if (V == 4) *reinterpret_cast<float4*>(g) = make_float4(vj[0], vj[1 % V], vj[2 % V], vj[3 % V]);
            else {
#pragma unroll
                for (int e = 0; e < V; ++e) g[e] = vj[e];
            }
        }
    }
}

}  // namespace

// called from f3d_knn_graph (knn.cu); returns false if the shape is outside this path
bool knn_tc_supported(int N, int F, int K) { return N <= kMaxTiles * kTN && F <= 64 && K + 1 <= 32 && N >= 2; }

// gathered (F,K,N,B) / edge (2F,K,N,B) tensors from the neighbour indices (also used behind knn_gram.cu)
int32_t knn_emit_launch(const float* X, int B, int N, int F, int K, const int32_t* idx, float* gathered, float* edge, cudaStream_t stream) {
    if (gathered || edge) {
        const bool v4 = (F & 3) == 0 && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(gathered) | reinterpret_cast<uintptr_t>(edge)) & 15) == 0;
        if (v4) knn_emit_kernel<4><<<(unsigned)(B * N), 256, 0, stream>>>(X, idx, N, F, K, gathered, edge);
        else knn_emit_kernel<1><<<(unsigned)(B * N), 256, 0, stream>>>(X, idx, N, F, K, gathered, edge);
        F3D_CHECK_LAUNCH("knn_emit_kernel");
    }
    return F3D_OK;
}

int32_t knn_tc_launch(const float* X, int B, int N, int F, int K, int32_t* idx, float* dist, float* gathered, float* edge, unsigned* stats, cudaStream_t stream) {
    KnnTcParams p;
    p.X = X; p.N = N; p.F = F; p.Kp = (F + 31) / 32 * 32; p.K = K; p.idx = idx; p.dist = dist; p.stats = stats;
    if (stats) F3D_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(unsigned), stream));
    const int halves = p.Kp / 32;
    const size_t b_region = std::max((size_t)3 * halves * kTN * 128, sizeof(float) * kCap * kTQ);
    const size_t smem = (size_t)halves * kTQ * 128 + b_region + sizeof(unsigned short) * kCap * kTQ + sizeof(float) * (kMaxTiles * kTN + 2 * kGroups * kTQ) +
                        sizeof(int) * kTQ + 1024;
    F3D_CUDA(cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((N + kTQ - 1) / kTQ, B);
    knn_tc_kernel<<<grid, kTcThreads, smem, stream>>>(p);
    F3D_CHECK_LAUNCH("knn_tc_kernel");
    return knn_emit_launch(X, B, N, F, K, idx, gathered, edge, stream);
}

}  // namespace f3d
