// knn_gram.cu — kNN graph of a point / feature cloud for sm_100a: Gram-matrix filter on tcgen05 + TMEM, operands fed by TMA
// bulk copies, exact re-evaluation out of the operand tiles themselves.
//
// Replaces CreateSingleKNNGraph / the EdgeConv prologue of Flux3D.jl src/models/dgcnn.jl:3-9 (NearestNeighbors.knn with K+1
// neighbours, sorted, the first one dropped by position) with the same bits as knn.cu / knn_tc.cu: every
// reported index and distance is evaluated in the reference arithmetic (sequential s = s + (a_d - b_d)²); the tensor cores
// only decide WHICH ~35 of the N candidates of a query are worth evaluating.
//
// knn_gram_prepare_{split,plain}_kernel write the cloud once as an OPERAND IMAGE: 256-point tiles, rows of 128 bytes in the canonical
// K-major SWIZZLE_128B layout — byte for byte what a tcgen05.mma wants to find in shared memory, so that a tile (32 KB per
// 32 features) travels with ONE cp.async.bulk — plus the squared norms.  Two row formats:
//   plain  (5 <= F <= 64)  the FP32 features; kind::tf32 reads their upper 19 bits: |g~ - g| <= 2^-8 |q||c|
//   split  (F <= 4)        A part [xh | xh | xl | . | x] and B part [xh | xl | xh | 0] of the same 128-byte row (x = xh + xl in
//                          TF32 pieces): g~ = xh.ch + xh.cl + xl.ch, |g~ - g| <= 2^-19 |q||c| — low-dimensional clouds have
//                          neighbours much closer than their norms, a 2^-8 filter would pass half the cloud
// knn_gram_kernel: CTA = 128 queries of one cloud against all its candidates (256-candidate tiles), warp-specialised:
//   warp 16, one thread   TMA producer and MMA issuer: bulk copies into a 2-stage ring, ceil(F/8) tcgen05.mma.kind::tf32
//                         (M = 128, N = 256) per tile into one of two 256-column TMEM accumulators, tcgen05.commit
//   warps 0-15            read-out, thread <-> (query row = TMEM lane, 64-column quarter): two tcgen05.ld of 32 columns,
//                         d' = |c|² - 2 g~ (the query's own norm does not change its ranking)
//     pass 1   the minimum of every 32-candidate chunk (FMNMX3 tree); the (K+1)-th smallest of a row's chunk minima T~ bounds
//              its (K+1)-th smallest distance from above (bitonic network, thread <-> row)
//     pass 2   the Gram tiles again (the MMA is nearly free); every candidate with d' <= T~ + 2E — a provable superset of the
//              answer, ~35 of 1024 — is re-evaluated in the reference arithmetic RIGHT THERE, out of the operand tile that
//              is still in shared memory (the image holds the exact FP32 values): no gather from L2, balanced over the
//              lanes of the warp that found the hits
//   ranking  every candidate finds its output slot by counting the (distance, index) keys below it
// Rows whose candidate set overflows (> 64: heavy ties, degenerate clouds) are redone at the end of the same CTA by an exact scan
// of the whole cloud.  Two launches per call (prepare, search), no memset.
#include <algorithm>

#include "f3d_common.cuh"

namespace f3d {
namespace {

constexpr int kGQ = 128;            // queries per CTA = UMMA M = TMEM lanes
constexpr int kGN = 256;            // candidates per tile = UMMA N = TMEM columns of an accumulator
constexpr int kGReadWarps = 16;
constexpr int kGThreads = (kGReadWarps + 1) * 32;
constexpr int kGCap = 64;           // candidate slots per query
constexpr int kGSeg = 256;          // hits per warp and tile (~75 on uniform clouds)
constexpr int kGMaxTiles = 8;       // N <= 2048
constexpr int kGHalf = kGN * 128;   // bytes of a tile per 32 features
constexpr float kGErrPlain = 0.005524272f;        // 2^-7.5 (2^-8 + slack), as knn_tc.cu
constexpr float kGErrSplit = 7.62939453125e-06f;  // 2^-17  (2^-19 + slack)
constexpr unsigned kGIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(kGN >> 3) << 17) | ((unsigned)(kGQ >> 4) << 24);

struct KnnGramParams {
    const float* X;            // [B][N][F]
    int B, N, F, K;
    int Np, ntiles, halves, ksteps, split, fold;
    unsigned char* img;        // [B][ntiles][halves][256][128 B]
    float* nrm;                // [B][Np]   squared norms (+inf past N)
    int32_t* idx;              // [B][N][K]
    float* dist;               // [B][N][K] or null
    unsigned* stats;           // [0] rows redone by the exact scan, [1] candidates re-evaluated exactly, [3..5] why, [7] = 2: this path served the call
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void tmem_alloc(unsigned* slot_in_smem, int cols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned addr, int cols) {  // the same warp
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
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {  // arrives on bar when all prior MMAs of this thread are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"l"((unsigned long long)__cvta_generic_to_shared(bar)) : "memory");
}
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
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// K-major SWIZZLE_128B canonical layout (cute::UMMA::make_umma_desc<Major::K>): a region is [rows][128 B]; the 16-byte
// chunk c of row r sits at r*128 + ((c ^ (r & 7)) << 4); 8-row groups are 1024 B apart (SBO); LBO unused with swizzle
__device__ __forceinline__ unsigned long long umma_desc128(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }
// x rounded to 11 significant bits (Veltkamp, 2^13 + 1) on the FMA pipe; no contraction: the library is built with -fmad=false
__device__ __forceinline__ float rn_tf32(float x) {
    const float c = __fmul_rn(x, 8193.0f);
    return __fsub_rn(c, __fsub_rn(c, x));
}
__device__ __forceinline__ unsigned swz(int r, int c) { return (unsigned)r * 128u + (unsigned)((c ^ (r & 7)) << 4); }

// A split row (F <= 4).  A part: floats [0, F) xh, [F, 2F) xh, [2F, 3F) xl, [12, 12 + F) x;  B part: [16, 16 + F) xh, [16 + F, ..) xl,
// [16 + 2F, ..) xh.  fold (F <= 3): two spare columns carry -|c|²/2 (two TF32 pieces) against 1, 1 in the A part — the accumulator
// then holds g' = g - |c|²/2 = -d'/2 and the read-out needs neither the norms nor an FMA per value.  Returns |x|².
template <int F>
__device__ __forceinline__ float split_row(const float* x, bool live, bool fold, float (&row)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) row[i] = 0.0f;
    float n = 0.0f;
    if (live) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const float v = __ldg(x + f), h = rn_tf32(v), l = rn_tf32(__fsub_rn(v, h));
            n = __fadd_rn(n, __fmul_rn(v, v));
            row[f] = h; row[F + f] = h; row[2 * F + f] = l; row[12 + f] = v;
            row[16 + f] = h; row[16 + F + f] = l; row[16 + 2 * F + f] = h;
        }
    }
    if (fold && F <= 3) {
        const float m = live ? -0.5f * n : -1.0e30f, mh = rn_tf32(m), ml = live ? rn_tf32(__fsub_rn(m, mh)) : 0.0f;
        row[(3 * F) & 31] = 1.0f; row[(3 * F + 1) & 31] = 1.0f;
        row[(16 + 3 * F) & 31] = mh; row[(16 + 3 * F + 1) & 31] = ml;
    }
    return n;
}

// ---- prepare: the operand image + norms -----------------------------------------------------------------------------------
// split rows (F <= 4): grid (ntiles, B), thread <-> row of the tile (three loads, eight 16-byte stores)
__global__ void __launch_bounds__(kGN) knn_gram_prepare_split_kernel(KnnGramParams p) {
    const int t = blockIdx.x, b = blockIdx.y, r = threadIdx.x, j = t * kGN + r;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the search grid may set itself up (it waits for this grid's end)
    unsigned char* tile = p.img + (size_t)(b * p.ntiles + t) * kGHalf;
    const float* x = p.X + ((size_t)b * p.N + min(j, p.N - 1)) * p.F;
    const bool live = j < p.N;
    float n = 0.0f;
    float row[32];
    switch (p.F) {
        case 1: n = split_row<1>(x, live, p.fold != 0, row); break;
        case 2: n = split_row<2>(x, live, p.fold != 0, row); break;
        case 3: n = split_row<3>(x, live, p.fold != 0, row); break;
        default: n = split_row<4>(x, live, p.fold != 0, row); break;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(tile + swz(r, c)) = make_float4(row[4 * c], row[4 * c + 1], row[4 * c + 2], row[4 * c + 3]);
    if (t == 0 && b == 0 && r < 8) p.stats[r] = r == 7 ? 2u : 0u;   // diagnostics (the search grid only counts after this grid has ended); [7]: this path served the call
    p.nrm[(size_t)b * p.Np + j] = live ? n : INFINITY;
}
// plain rows (5 <= F <= 64): thread <-> (row, 16-byte chunk) — 8 or 16 threads per row, 32 or 16 rows per block: every load and
// every store is one 16-byte access of a coalesced run, and a cfg3-sized call is 4096 blocks instead of 128
__global__ void __launch_bounds__(256) knn_gram_prepare_plain_kernel(KnnGramParams p) {
    const int cpr = 8 * p.halves, rpb = 256 / cpr, bpt = kGN / rpb;     // chunks per row, rows per block, blocks per tile
    const int t = blockIdx.x / bpt, rb = blockIdx.x - t * bpt, b = blockIdx.y;
    const int c = threadIdx.x % cpr, r = rb * rpb + threadIdx.x / cpr, j = t * kGN + r;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    unsigned char* tile = p.img + ((size_t)(b * p.ntiles + t) * p.halves) * kGHalf;
    const float* x = p.X + ((size_t)b * p.N + min(j, p.N - 1)) * p.F;
    const bool live = j < p.N, vec = (p.F & 3) == 0 && (reinterpret_cast<uintptr_t>(p.X) & 15) == 0;
    const int d0 = c * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live && d0 < p.F) {
        if (vec) v = __ldg(reinterpret_cast<const float4*>(x + d0));
        else {
            v.x = __ldg(x + d0);
            if (d0 + 1 < p.F) v.y = __ldg(x + d0 + 1);
            if (d0 + 2 < p.F) v.z = __ldg(x + d0 + 2);
            if (d0 + 3 < p.F) v.w = __ldg(x + d0 + 3);
        }
    }
    *reinterpret_cast<float4*>(tile + (size_t)(c >> 3) * kGHalf + swz(r, c & 7)) = v;
    float n = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;   // any rounding will do: the norm only enters the filter
    for (int o = cpr >> 1; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);   // (the threads of a row are cpr consecutive lanes)
    if (blockIdx.x == 0 && b == 0 && threadIdx.x < 8) p.stats[threadIdx.x] = threadIdx.x == 7 ? 2u : 0u;
    if (c == 0) p.nrm[(size_t)b * p.Np + j] = live ? n : INFINITY;
}

// exact (reference-arithmetic) squared distance between row r of the query tile and row c of a candidate tile, both in
// shared memory as operand images (zero padding adds +0.0: the sum over the real features is unchanged)
__device__ __forceinline__ float exact_pair(const unsigned char* s_a, int r, const unsigned char* s_b, int c, int nchunks, bool split) {
    float s = 0.0f;
    if (split) {
        const float4 q = *reinterpret_cast<const float4*>(s_a + swz(r, 3));
        const float4 w = *reinterpret_cast<const float4*>(s_b + swz(c, 3));
        float t;
        t = __fsub_rn(q.x, w.x); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(q.y, w.y); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(q.z, w.z); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(q.w, w.w); s = __fadd_rn(s, __fmul_rn(t, t));
        return s;
    }
#pragma unroll 4
    for (int c4 = 0; c4 < nchunks; ++c4) {
        const int h = c4 >> 3, cc = c4 & 7;
        const float4 q = *reinterpret_cast<const float4*>(s_a + (size_t)h * (kGQ * 128) + swz(r, cc));
        const float4 w = *reinterpret_cast<const float4*>(s_b + (size_t)h * kGHalf + swz(c, cc));
        float t;
        t = __fsub_rn(q.x, w.x); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(q.y, w.y); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(q.z, w.z); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(q.w, w.w); s = __fadd_rn(s, __fmul_rn(t, t));
    }
    return s;
}

// in-register bitonic sort (ascending), fully unrolled: data-independent, identical in every lane
template <int NS>
__device__ __forceinline__ void sort_regs(float (&v)[NS]) {
#pragma unroll
    for (int k = 2; k <= NS; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
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
__device__ __forceinline__ float min16f(const float* d) {
    const float a = min3f(d[0], d[1], d[2]), b = min3f(d[3], d[4], d[5]), c = min3f(d[6], d[7], d[8]), e = min3f(d[9], d[10], d[11]), f = min3f(d[12], d[13], d[14]);
    return fminf(min3f(a, b, c), min3f(e, f, d[15]));
}
__device__ __forceinline__ float max16u(const unsigned* v) {
    float d[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = -__uint_as_float(v[i]);
    return -min16f(d);
}
template <int S>
__device__ __forceinline__ void rank_row(const float* s_dex, const unsigned short* s_cand, int row, int part, int cnt, int K, int32_t* idx, float* dist) {
    unsigned long long key[S];
    int rank[S];
#pragma unroll
    for (int e = 0; e < S; ++e) {
        const int sidx = part + 4 * e;
        key[e] = sidx < cnt ? ((unsigned long long)__float_as_uint(s_dex[sidx * kGQ + row]) << 32) | s_cand[sidx * kGQ + row] : ~0ull;
        rank[e] = 0;
    }
#pragma unroll 2
    for (int o = 0; o < cnt; ++o) {
        const unsigned okh = __float_as_uint(s_dex[o * kGQ + row]), okl = s_cand[o * kGQ + row];
        // counts the keys NOT below the own key as the (absent) borrow of a 64-bit subtraction other - own — three instructions of one
        // subtract-with-borrow chain per pair, no compare / select: n += 1 - borrow  (n - (-1 + borrow))
#pragma unroll
        for (int e = 0; e < S; ++e)
            asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %3;\n\tsubc.cc.u32 t, %2, %4;\n\tsubc.u32 %0, %0, 0xffffffff;\n\t}"
                : "+r"(rank[e])
                : "r"(okl), "r"(okh), "r"((unsigned)(key[e] & 0xffffffffu)), "r"((unsigned)(key[e] >> 32)));
    }
#pragma unroll
    for (int e = 0; e < S; ++e) rank[e] = cnt - rank[e];   // keys below the own key
#pragma unroll
    for (int e = 0; e < S; ++e)
        if (part + 4 * e < cnt && rank[e] >= 1 && rank[e] <= K) {
            idx[rank[e] - 1] = (int32_t)(key[e] & 0xffffffffu);
            if (dist) dist[rank[e] - 1] = __uint_as_float((unsigned)(key[e] >> 32));
        }
}

#ifdef F3D_KNN_PROF
#define GPROF(k) do { if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) p.stats[8 + (k)] = (unsigned)(clock64() - gt0_); } while (0)
#else
#define GPROF(k)
#endif
__global__ void __launch_bounds__(kGThreads, 1) knn_gram_kernel(KnnGramParams p) {
#ifdef F3D_KNN_PROF
    const long long gt0_ = clock64();
#endif
    extern __shared__ unsigned char smem_raw_[];
    unsigned char* smem = smem_raw_ + ((1024u - (smem_u32(smem_raw_) & 1023u)) & 1023u);   // SWIZZLE_128B atoms: 1024-byte alignment
    const int halves = p.halves;
    const size_t b_tile = (size_t)halves * kGHalf;
    unsigned char* s_a = smem;                                        // [halves][128][128 B]   the queries' operand rows
    unsigned char* s_b = s_a + (size_t)halves * kGQ * 128;            // 2 x [halves][256][128 B] candidate tiles
    float* s_dex = reinterpret_cast<float*>(s_b + 2 * b_tile);        // [kGCap][128] exact distances (pass 1: the chunk minima, [<= 64][128])
    unsigned short* s_cand = reinterpret_cast<unsigned short*>(s_dex + kGCap * kGQ);   // [kGCap][128] candidate ids
    unsigned short* s_seg = s_cand + kGCap * kGQ;                     // [16][kGSeg] hits of a warp in the current tile: lane | column << 5
    float* s_nc = reinterpret_cast<float*>(s_seg + kGReadWarps * kGSeg);   // [Np] squared norms of the cloud
    float* s_thr = s_nc + kGMaxTiles * kGN;                           // [128]
    int* s_cnt = reinterpret_cast<int*>(s_thr + kGQ);                 // [128] candidates per query
    int* s_ovf = s_cnt + kGQ;                                         // [128] the query's candidate set is incomplete
    int* s_segcnt = s_ovf + kGQ;                                      // [16]
    __shared__ unsigned long long a_full, full_b[2], empty_b[2], tfull[2], tempty[2], s_key[2][8];
    __shared__ unsigned s_tmem;
    __shared__ int s_nfix;
    __shared__ unsigned s_wmax[kGReadWarps + 1];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, q0 = blockIdx.x * kGQ;
    const int ntiles = p.ntiles, L = 2 * ntiles;
    const unsigned char* cloud = p.img + (size_t)b * ntiles * b_tile;
    if (tid == 0) {
        mbar_init(&a_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], kGReadWarps); mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kGReadWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&s_tmem, 2 * kGN);
    for (int i = tid; i < kGQ; i += kGThreads) { s_cnt[i] = 0; s_ovf[i] = 0; }
    if (tid < kGReadWarps) s_segcnt[tid] = 0;
    if (tid == 0) s_nfix = 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");   // launched programmatically behind the prepare grid: the image is complete
    {   // the cloud's norms, and their maximum (padding is +inf) for the filter's error bound
        float lm = 0.0f;
        for (int i = tid; i < p.Np; i += kGThreads) {
            const float v = __ldcg(p.nrm + (size_t)b * p.Np + i);
            s_nc[i] = v;
            if (v < INFINITY) lm = fmaxf(lm, v);
        }
        const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(lm));   // norms are >= 0: bit order == value order
        if (lane == 0) s_wmax[warp] = wm;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = s_tmem;
    GPROF(0);

    if (warp == kGReadWarps) {
        // ---- producer + MMA issuer (one thread) -------------------------------------------------------------------------------
        if (lane == 0) {
            mbar_expect_tx(&a_full, (unsigned)(halves * kGQ * 128));
            for (int h = 0; h < halves; ++h)
                tma_bulk_g2s(s_a + (size_t)h * kGQ * 128, cloud + (size_t)(q0 / kGN) * b_tile + (size_t)h * kGHalf + (size_t)(q0 % kGN) * 128, kGQ * 128, &a_full);
            for (int u = 0; u < 2 && u < L; ++u) {
                mbar_expect_tx(&full_b[u], (unsigned)b_tile);
                tma_bulk_g2s(s_b + (size_t)u * b_tile, cloud + (size_t)(u % ntiles) * b_tile, (unsigned)b_tile, &full_b[u]);
            }
            const unsigned a0 = smem_u32(s_a), b0 = smem_u32(s_b);
            const unsigned boff = p.split ? 64u : 0u;   // split rows: the B part is the second half of the 128-byte row
            mbar_wait(&a_full, 0);
            for (int u = 0; u < L; ++u) {
                const unsigned s = u & 1, ph = (u >> 1) & 1;
                mbar_wait(&full_b[s], ph);
                mbar_wait(&tempty[s], ph ^ 1);     // accumulator s has been pulled out of TMEM (tile u - 2)
                tc_fence_after();
                for (int k = 0; k < p.ksteps; ++k)
                    umma_tf32(tmem + s * kGN, umma_desc128(a0 + (unsigned)(k >> 2) * (kGQ * 128) + (k & 3) * 32),
                              umma_desc128(b0 + s * (unsigned)b_tile + (unsigned)(k >> 2) * kGHalf + boff + (k & 3) * 32), kGIdesc, k > 0);
                umma_commit(&tfull[s]);
                if (u >= 1 && u + 1 < L) {         // refill the other stage (tile u - 1 has been read out and re-evaluated) with tile u + 1
                    const unsigned so = s ^ 1, pho = ((u - 1) >> 1) & 1;
                    mbar_wait(&empty_b[so], pho);
                    mbar_expect_tx(&full_b[so], (unsigned)b_tile);
                    tma_bulk_g2s(s_b + (size_t)so * b_tile, cloud + (size_t)((u + 1) % ntiles) * b_tile, (unsigned)b_tile, &full_b[so]);
                }
            }
        }
    } else {
        // ---- read-out -----------------------------------------------------------------------------------------------------------
        const int quad = warp & 3, part = warp >> 2;
        const int row = quad * 32 + lane;
        const int nchunks4 = (p.F + 3) >> 2;
        const bool fine = ntiles <= 4;
        float thr = 0.0f;
        for (int u = 0; u < L; ++u) {
            const unsigned s = u & 1, ph = (u >> 1) & 1;
            const int t = u < ntiles ? u : u - ntiles;
            if (u == ntiles) {
                // pass 1 is complete: the (K+1)-th smallest chunk minimum of every row, widened by the filter's error bound
                named_bar_sync(1, kGReadWarps * 32);
                GPROF(1);
                const int nch = fine ? ntiles * 16 : ntiles * 8;   // <= 64
                if (part < 2 && (part == 0 || nch > 32)) {
                    float sv[32];
#pragma unroll
                    for (int g = 0; g < 32; ++g) sv[g] = part * 32 + g < nch ? s_dex[(part * 32 + g) * kGQ + row] : INFINITY;
                    sort_regs<32>(sv);
#pragma unroll
                    for (int g = 0; g < 32; ++g) s_dex[(part * 32 + g) * kGQ + row] = sv[g];
                }
                named_bar_sync(1, kGReadWarps * 32);
                if (part == 0) {
                    // (K+1)-th smallest of the union of the two sorted halves: K + 1 steps of a merge
                    int ia = 0, ib = 32;
                    const int ibend = nch > 32 ? 64 : 32;
                    float Tsel = INFINITY;
                    for (int st = 0; st <= p.K; ++st) {
                        const float a = ia < 32 ? s_dex[ia * kGQ + row] : INFINITY, bb = ib < ibend ? s_dex[ib * kGQ + row] : INFINITY;
                        if (a <= bb) { Tsel = a; ++ia; } else { Tsel = bb; ++ib; }
                    }
                    const float nq = s_nc[min(q0 + row, p.Np - 1)];
                    unsigned mb = 0u;
#pragma unroll
                    for (int w = 0; w <= kGReadWarps; ++w) mb = max(mb, s_wmax[w]);
                    const float maxnc = __uint_as_float(mb);
                    // E bounds |d' - (d - nq)|: the Gram entry's relative error on |q||c| plus the FP32 roundings of the norms and of the
                    // fma, which do not shrink with |q| (<= (F + 3) u (nq + max nc), doubled for slack)
                    const float E = (p.split ? kGErrSplit : kGErrPlain) * sqrtf(nq) * sqrtf(maxnc) + 2.0f * (float)(p.F + 4) * 5.9604645e-8f * (nq + maxnc);
                    float th = Tsel + 2.0f * E;
                    if (!(th < INFINITY)) { th = -INFINITY; s_ovf[row] = 1; }   // NaN / inf: certify nothing, the exact scan takes the row
                    s_thr[row] = th;
                }
                named_bar_sync(1, kGReadWarps * 32);
                GPROF(2);
                thr = s_thr[row];
            }
            mbar_wait(&tfull[s], ph);
            tc_fence_after();
            const unsigned base = tmem + ((unsigned)(quad * 32) << 16) + (unsigned)(s * kGN + part * 64);
            unsigned v[2][32];
            tmem_ld32_issue(base, v[0]);
            tmem_ld32_issue(base + 32, v[1]);
            tmem_ld_wait(v[0]);
            tmem_ld_wait(v[1]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[s]);     // the accumulator is in registers: the MMAs of tile u + 2 may start
            const float* nc = s_nc + t * kGN + part * 64;
            if (u < ntiles) {
                // chunk minima of d' (fold: d' = -2 g', the smallest d' is -2 x the largest accumulator value).  Clouds of up to 1024
                // points use 16-candidate chunks (64 minima per row: the threshold passes ~25 candidates instead of ~34)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float lo, hi;
                    if (p.fold) {
                        lo = -2.0f * max16u(&v[q][0]);
                        hi = -2.0f * max16u(&v[q][16]);
                    } else {
                        float d[32];
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 n4 = *reinterpret_cast<const float4*>(nc + q * 32 + i);
                            d[i] = fmaf(-2.0f, __uint_as_float(v[q][i]), n4.x);
                            d[i + 1] = fmaf(-2.0f, __uint_as_float(v[q][i + 1]), n4.y);
                            d[i + 2] = fmaf(-2.0f, __uint_as_float(v[q][i + 2]), n4.z);
                            d[i + 3] = fmaf(-2.0f, __uint_as_float(v[q][i + 3]), n4.w);
                        }
                        lo = min16f(&d[0]);
                        hi = min16f(&d[16]);
                    }
                    const int ch = t * 8 + part * 2 + q;
                    if (fine) { s_dex[(2 * ch) * kGQ + row] = lo; s_dex[(2 * ch + 1) * kGQ + row] = hi; }
                    else s_dex[ch * kGQ + row] = fminf(lo, hi);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_b[s]);   // pass 1 never reads the candidate tile itself
            } else {
                unsigned short* seg = s_seg + warp * kGSeg;
                // hit masks without branches or atomics (one predicated OR per value), then one warp scan places every lane's hits
                unsigned hm[2] = {0u, 0u};
                if (p.fold) {
                    const float tg = -0.5f * thr;   // d' = -2 g' exactly: d' > thr  <=>  g' < -thr / 2
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (!(__uint_as_float(v[q][i]) < tg)) hm[q] |= 1u << i;
                } else
#pragma unroll
                for (int q = 0; q < 2; ++q) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 n4 = *reinterpret_cast<const float4*>(nc + q * 32 + i);
                        if (!(fmaf(-2.0f, __uint_as_float(v[q][i]), n4.x) > thr)) hm[q] |= 1u << i;
                        if (!(fmaf(-2.0f, __uint_as_float(v[q][i + 1]), n4.y) > thr)) hm[q] |= 1u << (i + 1);
                        if (!(fmaf(-2.0f, __uint_as_float(v[q][i + 2]), n4.z) > thr)) hm[q] |= 1u << (i + 2);
                        if (!(fmaf(-2.0f, __uint_as_float(v[q][i + 3]), n4.w) > thr)) hm[q] |= 1u << (i + 3);
                    }
                }
                const int mine = __popc(hm[0]) + __popc(hm[1]);
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                const int nh_all = __shfl_sync(0xffffffffu, incl, 31);
                int pos = incl - mine;
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    for (unsigned m = hm[q]; m; m &= m - 1) {
                        const int i = __ffs(m) - 1;
                        if (pos < kGSeg) seg[pos] = (unsigned short)(lane | ((q * 32 + i) << 5));   // 5 + 6 bits
                        else s_ovf[row] = 1;
                        ++pos;
                    }
                __syncwarp();
                // the warp's hits, balanced over its lanes, out of the operand tiles in shared memory
                const int nh = min(nh_all, kGSeg);
                const unsigned char* bt = s_b + (size_t)s * b_tile;
                for (int e = lane; e < nh; e += 32) {
                    const unsigned ent = seg[e];
                    const int r = quad * 32 + (int)(ent & 31u), c = part * 64 + (int)(ent >> 5);
                    const float dd = exact_pair(s_a, r, bt, c, nchunks4, p.split != 0);
                    const int sl = atomicAdd(&s_cnt[r], 1);
                    if (sl < kGCap) { s_dex[sl * kGQ + r] = dd; s_cand[sl * kGQ + r] = (unsigned short)(t * kGN + c); }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_b[s]);
            }
        }
        // ---- ranking: every candidate finds its rank in the ascending (distance, index) order by counting; ranks 1..K are the
        // neighbours, rank 0 is dropped by position (dgcnn.jl:6) ------------------------------------------------------------------
        named_bar_sync(1, kGReadWarps * 32);
        GPROF(4);
        const int qi = q0 + row;
        if (part == 0) {   // diagnostics: candidates re-evaluated exactly (one atomic per warp)
            unsigned c = qi < p.N ? (unsigned)min(s_cnt[row], kGCap) : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) atomicAdd(p.stats + 1, c);
        }
        if (qi < p.N) {
            const int cnt = s_cnt[row];
            const bool ovf = cnt > kGCap || s_ovf[row] != 0 || cnt < p.K + 1;
            const int cmax = __reduce_max_sync(__activemask(), ovf ? 0 : cnt);   // (warp-uniform choice of the ranking's register budget)
            if (ovf) {
                if (part == 0) {
                    atomicAdd(p.stats, 1u);
                    atomicAdd(p.stats + (cnt > kGCap ? 3 : (s_ovf[row] ? 4 : 5)), 1u);
                    reinterpret_cast<int*>(s_seg)[atomicAdd(&s_nfix, 1)] = row;   // (the hit segments are no longer needed)
                }
            } else {
                // the thread's own candidates (every fourth) live in registers as 64-bit keys (distance bits : index — distances
                // are >= 0, bit order == value order); one pass over all the row's candidates counts the keys below each of them
                const size_t obase = ((size_t)b * p.N + qi) * p.K;
                if (cmax <= 32) rank_row<8>(s_dex, s_cand, row, part, cnt, p.K, p.idx + obase, p.dist ? p.dist + obase : nullptr);
                else if (cmax <= 40) rank_row<10>(s_dex, s_cand, row, part, cnt, p.K, p.idx + obase, p.dist ? p.dist + obase : nullptr);
                else if (cmax <= 48) rank_row<12>(s_dex, s_cand, row, part, cnt, p.K, p.idx + obase, p.dist ? p.dist + obase : nullptr);
                else rank_row<16>(s_dex, s_cand, row, part, cnt, p.K, p.idx + obase, p.dist ? p.dist + obase : nullptr);
            }
        }
        // ---- rows whose candidate set is incomplete (> 64 candidates: heavy ties, degenerate clouds; a threshold that is not finite):
        // the exact scan of the whole cloud, here — two groups of eight warps take a row each: all its distances into shared memory
        // (the candidate-tile ring is free now), then K + 1 selection rounds over 64-bit (distance bits : index) keys -------------------
        named_bar_sync(1, kGReadWarps * 32);
        const int nfix = s_nfix;
        if (nfix) {
            const int grp = warp >> 3, gt = tid & 255, gw = warp & 7;
            float* sd = reinterpret_cast<float*>(s_b + (size_t)grp * b_tile);   // [N] (N <= 2048: 8 KB of a >= 32 KB stage)
            const float* Xb = p.X + (size_t)b * p.N * p.F;
            for (int e = grp; e < nfix; e += 2) {
                const int qf = q0 + reinterpret_cast<const int*>(s_seg)[e];
                const float* xq = Xb + (size_t)qf * p.F;
                for (int j = gt; j < p.N; j += 256) {
                    const float* xj = Xb + (size_t)j * p.F;
                    float sum = 0.0f;
                    for (int d = 0; d < p.F; ++d) { const float t = __fsub_rn(__ldg(xq + d), __ldg(xj + d)); sum = __fadd_rn(sum, __fmul_rn(t, t)); }
                    sd[j] = sum;
                }
                named_bar_sync(2 + grp, 256);
                unsigned long long last = 0ull;   // keys are (distance bits << 32 | index) + 1: strictly increasing from round to round
                for (int r = 0; r <= p.K; ++r) {
                    unsigned long long best = ~0ull;
                    for (int j = gt; j < p.N; j += 256) {
                        const unsigned long long key = (((unsigned long long)__float_as_uint(sd[j]) << 32) | (unsigned)j) + 1ull;   // d >= 0 (or NaN: sorts last)
                        if (key > last && key < best) best = key;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { const unsigned long long ot = __shfl_xor_sync(0xffffffffu, best, o); best = ot < best ? ot : best; }
                    if (lane == 0) s_key[grp][gw] = best;
                    named_bar_sync(2 + grp, 256);
                    best = s_key[grp][0];
#pragma unroll
                    for (int w = 1; w < 8; ++w) best = s_key[grp][w] < best ? s_key[grp][w] : best;
                    named_bar_sync(2 + grp, 256);
                    last = best;
                    if (gt == 0 && r >= 1 && best != ~0ull) {
                        const unsigned long long key = best - 1ull;
                        const size_t o = ((size_t)b * p.N + qf) * p.K + r - 1;
                        p.idx[o] = (int32_t)(key & 0xffffffffu);
                        if (p.dist) p.dist[o] = __uint_as_float((unsigned)(key >> 32));
                    }
                }
            }
        }
    }
    GPROF(5);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 2 * kGN); }
}

struct GramPlan {
    int Np, ntiles, halves, ksteps, split, fold;
    size_t off_stats, off_nrm, off_img, total;
};
GramPlan make_gram_plan(int B, int N, int F) {
    GramPlan pl;
    pl.Np = (int)align_up((size_t)N, kGN);
    pl.ntiles = pl.Np / kGN;
    pl.split = F <= 4 ? 1 : 0;
    pl.fold = F <= 3 ? 1 : 0;
    pl.halves = pl.split ? 1 : (F + 31) / 32;
    pl.ksteps = pl.split ? 2 : (F + 7) / 8;
    size_t o = 0;
    pl.off_stats = o; o = align_up(o + 64, 256);
    pl.off_nrm = o;   o = align_up(o + sizeof(float) * (size_t)B * pl.Np, 1024);
    pl.off_img = o;   o = align_up(o + (size_t)B * pl.ntiles * pl.halves * kGHalf, 256);
    pl.total = o;
    return pl;
}
size_t gram_smem_bytes(int halves) {
    return (size_t)halves * kGQ * 128 + 2 * (size_t)halves * kGHalf + sizeof(float) * kGCap * kGQ + sizeof(unsigned short) * kGCap * kGQ +
           sizeof(unsigned short) * kGReadWarps * kGSeg + sizeof(float) * (kGMaxTiles * kGN + kGQ) + sizeof(int) * (2 * kGQ + kGReadWarps) + 1024;
}

}  // namespace

// called from f3d_knn_graph (knn.cu)
// (the threshold is the (K+1)-th smallest of N / 32 chunk minima: with fewer than 1.5 (K+1) chunks it would pass most of the cloud)
bool knn_gram_supported(int N, int F, int K) {
    const int chunks = N <= 4 * kGN ? (N + 15) / 16 : (N + 31) / 32;
    return F <= 64 && K + 1 <= 32 && N <= kGMaxTiles * kGN && 2 * chunks >= 3 * (K + 1);
}
size_t knn_gram_workspace_bytes(int B, int N, int F) { return make_gram_plan(B, N, F).total; }

int32_t knn_gram_launch(const float* X, int B, int N, int F, int K, int32_t* idx, float* dist, void* ws, size_t ws_bytes, cudaStream_t stream) {
    const GramPlan pl = make_gram_plan(B, N, F);
    if (!ws || ws_bytes < pl.total) return fail(F3D_ERR_WORKSPACE, "f3d_knn_graph: workspace %zu < required %zu bytes", ws_bytes, pl.total);
    unsigned char* w = static_cast<unsigned char*>(ws);
    KnnGramParams p;
    p.X = X; p.B = B; p.N = N; p.F = F; p.K = K;
    p.Np = pl.Np; p.ntiles = pl.ntiles; p.halves = pl.halves; p.ksteps = pl.ksteps; p.split = pl.split; p.fold = pl.fold;
    p.img = w + pl.off_img;
    p.nrm = reinterpret_cast<float*>(w + pl.off_nrm);
    p.idx = idx; p.dist = dist;
    p.stats = reinterpret_cast<unsigned*>(w + pl.off_stats);
    const size_t smem = gram_smem_bytes(pl.halves);
    static bool attr_done[2];
    if (!attr_done[pl.halves - 1]) {
        F3D_CUDA(cudaFuncSetAttribute(knn_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gram_smem_bytes(2)));
        attr_done[0] = attr_done[1] = true;
    }
    if (pl.split) knn_gram_prepare_split_kernel<<<dim3(pl.ntiles, B), kGN, 0, stream>>>(p);
    else knn_gram_prepare_plain_kernel<<<dim3(pl.ntiles * pl.halves * 8, B), 256, 0, stream>>>(p);   // blocks per tile = 256 rows / (256 / (8 halves)) rows per block
    F3D_CHECK_LAUNCH("knn_gram_prepare_kernel");
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((N + kGQ - 1) / kGQ, B); cfg.blockDim = dim3(kGThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        F3D_CUDA(cudaLaunchKernelEx(&cfg, knn_gram_kernel, p));
    }
    F3D_CHECK_LAUNCH("knn_gram_kernel");
    return F3D_OK;
}

}  // namespace f3d
