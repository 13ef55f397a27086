// chamfer_tc.cu — the chamfer filter sweep on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Same contract and the same bits as chamfer.cu (src/metrics/pcloud.jl:39-52, :54-70 semantics): every reported index and
// distance is evaluated in the reference arithmetic; the tensor cores only decide WHERE to look.
//
// The N x 3 . 3 x M contraction of the squared distance is the dense part of the path.  One direction of the search
// ("for every query row, the nearest candidate column") is a K = 16 TF32 GEMM whose accumulator never leaves the SM:
//     f_ij = |q_i|² + |c_j|² - 2 q_i.c_j = sum_k R[i][k] C[j][k]
//     R[i] = {xh,xh,xl,xl, yh,yh,yl,yl | zh,zh,zl,zl, nh,nl,1,1}        (x = xh + xl: an FP32 value as two TF32 pieces)
//     C[j] = {Xh,Xl,Xh,Xl, Yh,Yl,Yh,Yl | Zh,Zl,Zh,Zl, 1,1,Nh,Nl}        (X = -2x of the candidate; points are centred)
// Both directions run the same kernel with the clouds swapped (the MMA is nearly free; what costs is reading the
// accumulator: in TMEM a thread owns one query ROW, so a row minimum is an in-thread FMNMX3 tree, a column minimum would
// need one warp reduction per column).  Measured error of f against the exact distance: <= 6.4 u (|q|² + |c|²), u = 2^-24
// (tools/tc_probe.cu; 5.3 u of it is the tensor core's own accumulation) — tighter than the FP32 expanded form of
// chamfer.cu; the certificate below assumes 35 u.
//
// Work item = 256 query rows of one batch element and direction (two 128-row UMMA tiles) against ALL candidates of the
// other cloud, streamed as 256-candidate tiles.  One persistent CTA per SM, 32 warps with fixed roles (setmaxnreg moves the
// registers to where they are needed: 88 per read-out thread, 40 for everybody else — the whole register file):
//     producer (1 thread)    TMA bulk copies (cp.async.bulk -> UBLKCP) of compact operands {x', y', z', |p'|²} (16 B / point,
//                            written once by chamfer_tc_prepare_kernel) into a 4-slot ring
//     converters (4 warps)   thread <-> point: split the four values into TF32 pieces and write the 64-byte K-major
//                            SWIZZLE_64B operand row (the 64 B/point image is never in HBM or L2: at 256 rows per candidate
//                            tile a pre-expanded image would need ~9 TB/s of L2 -> SM traffic at the rate the MMAs run)
//     MMA issuer (1 thread)  four tcgen05.mma.kind::tf32 (2 row tiles x 2 K-steps, M = 128, N = 256) per candidate tile into
//                            one 256-column TMEM accumulator per row tile (all 512 columns; the pair is the double buffer)
//     read-out (16 warps)    thread <-> query row (TMEM lane) and column quarter: two tcgen05.ld of 32 columns in flight, an
//                            FMNMX3 tree per 32-candidate chunk, and the locator record — the two smallest chunk minima with
//                            their chunk numbers (carried in the low mantissa bits) and the third smallest value — plus one
//                            minimum per 2048-candidate supertile
//     certifiers (8 warps)   the item that has just left the read-out, while the tensor cores work on the next one:
//                            lane <-> row merges the four column quarters' records and decides — b2 > b1 + window proves that
//                            the exact argmin (lowest index on ties) lies in chunk c1; else b3 > b1 + window: in c1 or c2; else
//                            the row is ambiguous — then fetches the located 32 candidates (384 contiguous bytes of the
//                            ORIGINAL cloud) with one TMA bulk copy per row into its own shared-memory row and re-evaluates
//                            them in the reference arithmetic: index, distance, and the item's partial sum of the loss.
// No locator record, flag or completion counter goes through global memory; the step is three launches (prepare, sweep,
// cleanup).  chamfer_tc_cleanup_kernel (launched programmatically; its blocks take the SMs as the sweep's CTAs leave) rescans
// the supertiles within the window for the rare ambiguous rows and reduces the loss in a fixed order (bitwise repeatable);
// for a sharded batch the sum over ranks is fused into its last block (peer mailboxes over NVLink, as in chamfer.cu).
#include <algorithm>
#include <cstdlib>

#include "f3d_common.cuh"

namespace f3d {
namespace {

constexpr int kTQ = 128;                    // UMMA M = TMEM lanes = query rows per row tile
constexpr int kRT = 2;                      // row tiles per work item
constexpr int kItemRows = kTQ * kRT;        // 256
constexpr int kTNc = 256;                   // UMMA N = candidates per tile.  A tcgen05.mma costs ~245 cycles however small it is
                                            // (tools/tc_probe3.cu: N = 32 ... 128 all take 245; N = 256 takes 293 = 87 % of the TF32 rate)
constexpr int kTcChunk = 32;                // locator chunk (candidates re-evaluated per certified row)
constexpr int kSuper = 2048;                // supertile: one stored minimum per row, column quarter and 2048 candidates
constexpr int kTilesPerSuper = kSuper / kTNc;
constexpr int kRowB = 64;                   // operand row: 16 TF32 = 64 bytes
constexpr int kStages = 3;                  // operand ring (16 KB per stage)
// Tunables (development: tools/build_variants.sh)
#ifndef F3D_TC_CSTAGES
#define F3D_TC_CSTAGES 4
#endif
#ifndef F3D_TC_EPI_REGS
#define F3D_TC_EPI_REGS 88
#endif
constexpr int kCStages = F3D_TC_CSTAGES;    // compact ring (4 KB per slot)
constexpr int kParts = 4;                   // a row's 256 accumulator columns are read by four warps, 64 columns each
constexpr int kEpiWarps = 4 * kParts;       // read-out warps: warp w reads lane quadrant w & 3, column quarter w >> 2 of BOTH row tiles' accumulators
constexpr int kConvWarps = 4;               // converters: thread <-> two points of a 256-point tile
constexpr int kWarpProducer = kEpiWarps + kConvWarps, kWarpMma = kWarpProducer + 1;
constexpr int kCertWarps = 8;               // certifiers: warp <-> 32 rows of the item that has just left the read-out
constexpr int kWarpCert = kWarpProducer + 4;
constexpr int kTcThreads = (kWarpCert + kCertWarps) * 32;   // 1024: eight warpgroups (4 read-out, converters, {producer, MMA issuer, two idle warps}, 2 certifier)
// Registers: the CTA is launched with kLaunchRegs per thread; the four read-out warpgroups then grow to kEpiRegs — both
// tcgen05.ld of an accumulator quarter are in flight at once, 64 registers of data — and the other four shrink to kAuxRegs
// (setmaxnreg).  1024 x 64 = 512 x 88 + 512 x 40: the whole register file.
#ifndef F3D_TC_LAUNCH_REGS
#define F3D_TC_LAUNCH_REGS 64
#endif
#ifndef F3D_TC_AUX_REGS
#define F3D_TC_AUX_REGS 40
#endif
constexpr int kLaunchRegs = F3D_TC_LAUNCH_REGS, kEpiRegs = F3D_TC_EPI_REGS, kAuxRegs = F3D_TC_AUX_REGS;
static_assert(kTcThreads * kLaunchRegs >= kEpiWarps * 32 * kEpiRegs + (kTcThreads - kEpiWarps * 32) * kAuxRegs, "register budget");
static_assert(kTcThreads * kLaunchRegs <= 65536 && kCertWarps * 32 == kItemRows, "one CTA per SM; certifier lane <-> row of the item");
static_assert(kItemRows == kTNc, "the query rows of an item travel as one 256-point compact tile");
constexpr int kChunkPitch = 400;            // bytes per staged chunk of the certifiers (384 B of data; 25 x 16 B: conflict-free 16-byte rows)
constexpr float kPadN = 1.0e30f;            // |p'|² of padded points: never a minimum for in-contract inputs
constexpr float kTcNormLimit = 1.0e29f, kTcNormFloor = 1.0e-30f;  // outside: certify nothing (overflow / underflow of the pieces)
// Certificate: |f - d| <= kErrAbs (nq + nc) + 5.001 u d  (u = 2^-24; d = the reference-arithmetic distance).  35 u =
// 12 u (two-piece TF32 representation of the coordinates and norms) + 7.02 u (FP32 norms, centring — as in chamfer.cu) +
// 16 u allowed for the tensor core's accumulation (measured <= 5.3 u).  Hence "b2 > b1 + 70.1 u (nq + max nc) + 10.01 u b1"
// proves that the exact argmin lies in chunk c1.  Every certified row is also CHECKED: the exact minimum found in the
// chunk must lie within kErrAbs (nq + nc) + kErrRel d of b1, else the row is rescanned like an ambiguous one and counted
// (header word kHdrViol; tests assert it stays 0).
constexpr float kTcWinAbs = 4.2e-6f;        // >= 70.1 u = 4.178e-6
constexpr float kTcWinRel = 6.2e-7f;        // >= 10.01 u
constexpr float kTcErrAbs = 2.1e-6f, kTcErrRel = 3.1e-7f;
// workspace header (ints); the first three words are diagnostics a caller may read after the call
constexpr int kHdrDone = 0, kHdrAmb = 1, kHdrViol = 2, kHdrInts = 64;

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
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {  // arrives on bar when all prior MMAs of this thread are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"l"((unsigned long long)__cvta_generic_to_shared(bar)) : "memory");
}
// issue only; the registers are valid after tmem_ld_wait(r), which carries them as in/out operands so that no consumer can
// be scheduled ahead of the wait
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
__device__ __forceinline__ int ld_acquire_i32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
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
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// K-major SWIZZLE_64B operand descriptor (cute::UMMA::make_umma_desc<Major::K>, LayoutType::B64): rows of 64 B, the
// 16-byte chunk c of row r at r*64 + ((c ^ ((r >> 1) & 3)) << 4), 8-row groups 512 B apart (SBO = 32), version 1
__device__ __forceinline__ unsigned long long umma_desc64(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr unsigned kIdescTc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(kTNc >> 3) << 17) | ((unsigned)(kTQ >> 4) << 24);

// x = h + l + e with h, l representable in TF32 (11 significant bits) and |e| <= 2^-22 |x|
// Veltkamp splitting with 2^13 + 1: three FP32 operations on the (idle) FMA pipe give x rounded to 11 significant bits — the
// ALU pipe, which the read-out's FMNMX work saturates (ncu: 79 % active), is left alone.  No contraction: -fmad=false.
__device__ __forceinline__ float rn_tf32(float x) {
    const float c = __fmul_rn(x, 8193.0f);
    return __fsub_rn(c, __fsub_rn(c, x));
}
__device__ __forceinline__ void split_tf32(float x, float& h, float& l) {
    h = rn_tf32(x);
    l = rn_tf32(__fsub_rn(x, h));   // x - h is exact
}

// The locator record of a query row: the two smallest chunk minima with their chunk ids and the third smallest value.
//   b2 > b1 + window              => the exact argmin lies in chunk c1
//   else b3 > b1 + window         => it lies in chunk c1 or c2 (two rescans instead of one; ~0.4 % of rows on uniform clouds)
//   else                          => ambiguous: every supertile within the window is scanned (cleanup kernel; ties, degenerate input)
struct Loc3 {
    float b1, b2, b3;
    int c1, c2;
};
// Inside the sweep the chunk id travels IN the value: the low `idbits` mantissa bits of a chunk minimum are replaced by the
// thread's running chunk number (2 x tile + chunk of its quarter), so that three FMNMX keep the two smallest chunk minima
// together with where they came from — no compares, no selects (5 ALU instructions less per chunk on the pipe that binds).
// The values move by at most 2^(idbits-23) relative; the certificate's relative term carries that (tc_win_rel below).
struct Loc3v {
    float b1, b2, b3;   // b1 <= b2 <= b3, ids in the low bits
};
__device__ __forceinline__ void loc3v_insert(Loc3v& l, float m) {
    l.b3 = fminf(l.b3, fmaxf(l.b2, m));
    l.b2 = fminf(l.b2, fmaxf(l.b1, m));
    l.b1 = fminf(l.b1, m);
}
__host__ __device__ __forceinline__ int tc_idbits(int ntiles) {   // bits for 2 * ntiles chunk numbers
    int b = 1;
    while ((1 << b) < 2 * ntiles) ++b;
    return b;
}
__device__ __forceinline__ void loc3_insert(Loc3& l, float m, int chunk) {   // strict '<': the EARLIEST chunk reaching a value keeps it
    const bool lt1 = m < l.b1, lt2 = m < l.b2;
    l.b3 = fminf(l.b3, fmaxf(l.b2, m));
    l.c2 = lt1 ? l.c1 : (lt2 ? chunk : l.c2);
    l.b2 = fminf(l.b2, fmaxf(l.b1, m));
    l.c1 = lt1 ? chunk : l.c1;
    l.b1 = fminf(l.b1, m);
}
// ---- prepare: compact operands {x', y', z', |p'|²} of both clouds, once per call ---------------------------------------
struct TcPrepParams {
    const float* A;    // [B][N][3]
    const float* Bp;   // [B][M][3]
    int N, M, NpA, NpB;
    float4* PA;        // [B][NpA]
    float4* PB;        // [B][NpB]
    unsigned* maxn;    // [2][B][nblk]  max |p'|² per cloud, batch element and block of this grid (float bits): no atomics, nothing to zero
    int nblk;          // blocks per cloud and element (gridDim.x)
    int* hdr;          // the workspace header: zeroed by this grid's first block (the sweep counts only after this grid has ended)
};
constexpr int kPrepT = 256, kPrepPts = 4;   // thread <-> four consecutive points: 48 bytes in, 64 bytes out
__global__ void __launch_bounds__(kPrepT) chamfer_tc_prepare_kernel(TcPrepParams p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the sweep's set-up overlaps this grid; it waits before it reads
    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.y;
    const bool isA = blockIdx.z == 0;
    const int n = isA ? p.N : p.M, np = isA ? p.NpA : p.NpB;
    const int i0 = ((int)blockIdx.x * kPrepT + tid) * kPrepPts;
    if ((int)blockIdx.x * kPrepT * kPrepPts >= np) {   // (the grid is sized for the larger cloud)
        if (tid == 0) p.maxn[((size_t)(isA ? 0 : 1) * gridDim.y + b) * p.nblk + blockIdx.x] = 0u;
        return;
    }
    const float* gA = p.A + (size_t)b * p.N * 3;
    const float* gB = p.Bp + (size_t)b * p.M * 3;
    // the centre of the batch element: the mean of 32 + 32 strided sample points (any point works — the certificate uses the
    // norms actually obtained); every warp computes it with the same operations => the same bits everywhere
    const int ia = (int)(((long)lane * p.N) >> 5), ib = (int)(((long)lane * p.M) >> 5);
    float cx = __ldg(gA + 3 * ia) + __ldg(gB + 3 * ib);
    float cy = __ldg(gA + 3 * ia + 1) + __ldg(gB + 3 * ib + 1);
    float cz = __ldg(gA + 3 * ia + 2) + __ldg(gB + 3 * ib + 2);
    const float* src = (isA ? gA : gB) + 3 * (size_t)i0;
    float r[3 * kPrepPts];
    if (i0 + kPrepPts <= n && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src) + 1), v2 = __ldg(reinterpret_cast<const float4*>(src) + 2);
        r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w; r[8] = v2.x; r[9] = v2.y; r[10] = v2.z; r[11] = v2.w;
    } else {
#pragma unroll
        for (int u = 0; u < kPrepPts; ++u) {
            const float* s1 = (isA ? gA : gB) + 3 * (size_t)min(i0 + u, n - 1);
            r[3 * u] = __ldg(s1); r[3 * u + 1] = __ldg(s1 + 1); r[3 * u + 2] = __ldg(s1 + 2);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cx += __shfl_xor_sync(0xffffffffu, cx, o);
        cy += __shfl_xor_sync(0xffffffffu, cy, o);
        cz += __shfl_xor_sync(0xffffffffu, cz, o);
    }
    cx *= (1.0f / 64.0f); cy *= (1.0f / 64.0f); cz *= (1.0f / 64.0f);
    float mx = 0.f;
    float4* dst = (isA ? p.PA : p.PB) + (size_t)b * np;
#pragma unroll
    for (int u = 0; u < kPrepPts; ++u) {
        const int i = i0 + u;
        float x = 0.f, y = 0.f, z = 0.f, nrm = kPadN;
        if (i < n) {
            x = r[3 * u] - cx; y = r[3 * u + 1] - cy; z = r[3 * u + 2] - cz;
            nrm = fmaf(z, z, fmaf(y, y, x * x));
            mx = fmaxf(mx, nrm);
        }
        if (i < np) dst[i] = make_float4(x, y, z, nrm);
    }
    __shared__ unsigned s_mx[kPrepT / 32];
    const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));  // norms are >= 0 (NaN bits order above everything: caught by the limit test)
    if (lane == 0) s_mx[tid >> 5] = wm;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid < kHdrInts) p.hdr[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        unsigned m = s_mx[0];
#pragma unroll
        for (int w = 1; w < kPrepT / 32; ++w) m = max(m, s_mx[w]);
        p.maxn[((size_t)(isA ? 0 : 1) * gridDim.y + b) * p.nblk + blockIdx.x] = m;
    }
}

// ---- host arrays: PCIe reads by the sweep CTAs' own spare warps (f3d_chamfer_pipe_run) --------------------------------------
__device__ __forceinline__ float4 ld_host_f4(const float* p) {
    float4 v;  // volatile: never served from a stale cache line of a previous call's bytes at the same host address
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
// One batch element of both clouds (whole 16-byte units: N, M are multiples of 4 on this path) by warp w of W cooperating warps:
// the loads of both clouds are issued before the first store — four 16-byte PCIe reads per lane in flight
__device__ __forceinline__ void upload_pair_warp(const float4* __restrict__ sa, float4* __restrict__ da, unsigned na4, const float4* __restrict__ sb,
                                                 float4* __restrict__ db, unsigned nb4, int w, int W, int lane) {
    const unsigned stride = (unsigned)W * 32u, n = max(na4, nb4);
    for (unsigned i = (unsigned)w * 32u + lane; i < n; i += 2 * stride) {
        const unsigned j = i + stride;
        float4 a0, a1, b0, b1;
        if (i < na4) a0 = ld_host_f4(reinterpret_cast<const float*>(sa + i));
        if (i < nb4) b0 = ld_host_f4(reinterpret_cast<const float*>(sb + i));
        if (j < na4) a1 = ld_host_f4(reinterpret_cast<const float*>(sa + j));
        if (j < nb4) b1 = ld_host_f4(reinterpret_cast<const float*>(sb + j));
        if (i < na4) __stcg(da + i, a0);
        if (i < nb4) __stcg(db + i, b0);
        if (j < na4) __stcg(da + j, a1);
        if (j < nb4) __stcg(db + j, b1);
    }
}
// batch elements in flight: element b is moved by the uploader warps w with w % G == b % G
__device__ __forceinline__ int upload_group_size(int W, int b, int G) { return (W - b % G + G - 1) / G; }

// ---- sweep -------------------------------------------------------------------------------------------------------------
#ifdef F3D_TC_PROF
// development: cycles every role spends waiting / working, per CTA (read back with f3d_debug_read_tc)
__device__ long long g_tcprof[148 * 16];
__device__ long long g_certprof[148 * 8];
#define PROF_DECL long long pf_[6] = {0, 0, 0, 0, 0, 0}; long long pt_ = clock64()
#define PROF(k) do { const long long n_ = clock64(); pf_[k] += n_ - pt_; pt_ = n_; } while (0)
#define PROF_OUT(base, n) do { if (blockIdx.x < 148) for (int i_ = 0; i_ < (n); ++i_) g_tcprof[blockIdx.x * 16 + (base) + i_] = pf_[i_]; } while (0)
#else
#define PROF_DECL
#define PROF(k)
#define PROF_OUT(base, n)
#endif
// What the certifier warps of the sweep and the cleanup kernel share.  "Slot" = item * 256 + position: the ambiguous rows of an
// item are listed in row order at fixed positions, so that the loss stays bitwise repeatable whatever order they are handled in.
struct TcFinParams {
    const float* A;
    const float* Bp;
    const float4* PA;
    const float4* PB;
    int B, N, M, NpA, NpB, rbA, rbB, nstA, nstB;
    const float* tilemin;
    const unsigned* maxn;   // [2][B][nblk] (resident inputs: chamfer_tc_prepare_kernel's block maxima)
    int nblk;
    int32_t* nnA;
    int32_t* nnB;
    double* partial;        // [nitems]       sum of the certified rows' distances of every item
    int* ambcnt;            // [nitems]       ambiguous rows of every item ...
    int* ambq;              // [nitems][256]  ... their row numbers, in row order ...
    float* amblim;          // [nitems][256]  ... and their windows' upper ends
    int* amblist;           // [rows]         slots of all ambiguous rows, in arrival order: the cleanup's work list
    float* ambd;            // [nitems][256]  the exact distances the cleanup finds, at the rows' slots
    int nitems;
    int* hdr;               // kHdr*
    float w1, w2;
    double denomA, denomB;
    float* loss;
    float* terms;
    ChamferPeerSum peer;
};

struct TcSweepParams {
    const float4* PA;
    const float4* PB;
    int B, NpA, NpB;       // padded cloud sizes (multiples of 256)
    int rbA, rbB;          // 256-row blocks per batch element: NpA / 256, NpB / 256
    float* tilemin;        // [B][ (nstB*NpA + nstA*NpB) * kParts ]  per supertile of the searched cloud, read-out group and row
    int nstA, nstB;        // supertiles of cloud A / B
    int wait_prepare;      // launched programmatically behind the prepare grid
    // Host arrays (f3d_chamfer_pipe_run): no prepare grid, no copy.  The two spare warps of EVERY sweep CTA pull the clouds out of
    // page-locked host memory over PCIe, batch element by batch element, into staging copies in HBM (f.A / f.Bp) and count every
    // element as it lands; the producer waits per element and streams the raw points (12 B each) as the tiles; the converters
    // centre them and take the norms on the fly.
    unsigned* arrived;     // [B] uploader warps that have delivered their share of an element (target: the element's group); null: resident inputs
    int up_groups;         // batch elements the uploader warps move concurrently
    const float* hA;       // the host arrays as the device addresses them (page-locked, mapped)
    const float* hB;
    TcFinParams f;         // the certifier warps' inputs and outputs
};

// kRaw: host arrays (uploader warps, raw tiles, centre / norms on the fly); the resident-input instantiation carries none of that code
template <bool kRaw>
__global__ void __maxnreg__(kLaunchRegs) chamfer_tc_sweep_kernel(TcSweepParams p) {
    extern __shared__ unsigned char smem_raw_[];
    unsigned char* smem = smem_raw_ + ((1024u - (smem_u32(smem_raw_) & 1023u)) & 1023u);
    unsigned char* s_a = smem;                                            // [2][kRT][128][64 B]   query operand tiles, double-buffered per item
    unsigned char* s_b = s_a + 2 * kRT * kTQ * kRowB;                     // [kStages][256][64 B]  candidate operand ring
    float4* s_c = reinterpret_cast<float4*>(s_b + kStages * kTNc * kRowB);  // [kCStages][256]      compact ring
    float4* s_pub = s_c + kCStages * kTNc;                                // [kParts][256]        locator triples per column quarter
    unsigned char* s_chunk = reinterpret_cast<unsigned char*>(s_pub + kParts * kItemRows);   // [256][400 B] the certifiers' located chunks
    __shared__ unsigned long long cert_bar[kCertWarps], cfull[kCStages], cempty[kCStages], full_b[kStages], empty_b[kStages], tfull[kRT], tempty[kRT], a_full[2],
        a_empty[2], pub_full, pub_empty;
    __shared__ unsigned s_tmem;

    __shared__ float s_ctr[4][4];               // host arrays: the batch element's centre, per item
    __shared__ unsigned s_maxw[4][kConvWarps];  // ... and the largest candidate norm each converter warp has seen in the item
    __shared__ double s_part[2][kCertWarps];
    __shared__ int s_cnt[2][kCertWarps];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef F3D_TC_PROF
    if (tid == 0 && blockIdx.x < 148) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); g_certprof[blockIdx.x * 8 + 5] = (long long)g_; }  // CTA start
#endif
    // the cleanup grid may be set up right away: its blocks take the SMs as the CTAs of this grid leave and wait for the grid's end
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        for (int s = 0; s < kCStages; ++s) { mbar_init(&cfull[s], 1); mbar_init(&cempty[s], kConvWarps); }
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_b[s], kConvWarps); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < kRT; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kEpiWarps); }
        for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], kConvWarps); mbar_init(&a_empty[s], 1); }
        mbar_init(&pub_full, kEpiWarps); mbar_init(&pub_empty, kCertWarps);
        for (int s = 0; s < kCertWarps; ++s) mbar_init(&cert_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&s_tmem, kRT * kTNc);  // all 512 columns: one 256-column accumulator per row tile
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = s_tmem;

    const int ipe = p.rbA + p.rbB, nitems = p.B * ipe;   // items per batch element: row blocks of A, then of B
    // every role walks the same item sequence; tile / slot counters run on across items so the pipelines never drain
    if (warp >= kEpiWarps) {
      // the four auxiliary warpgroups (converters; producer / MMA issuer / two idle warps; certifiers) hand registers back
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kAuxRegs));
      if (warp == kWarpProducer) {
        if (lane == 0) {
            if (p.wait_prepare) asm volatile("griddepcontrol.wait;" ::: "memory");  // the compact operands are complete and visible
            unsigned h = 0;  // compact tiles streamed so far
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int b = item / ipe, k = item - b * ipe;
                const bool dir = k >= p.rbA;
                const int rb = dir ? k - p.rbA : k;
                const int npq = dir ? p.NpB : p.NpA, npc = dir ? p.NpA : p.NpB;
                const float4* Pq = (dir ? p.PB : p.PA) + (size_t)b * npq + (size_t)rb * kItemRows;
                const float4* Pc = (dir ? p.PA : p.PB) + (size_t)b * npc;
                const int ntiles = npc / kTNc;
                if (kRaw) {   // host arrays: this batch element may still be crossing PCIe
                    while (ld_acquire_i32(reinterpret_cast<const int*>(p.arrived) + b) < upload_group_size(2 * (int)gridDim.x, b, p.up_groups)) __nanosleep(100);
                    asm volatile("fence.proxy.async;" ::: "memory");   // written with generic stores while this kernel runs; the TMA reads through the async proxy
                    const int nq = dir ? p.f.M : p.f.N, nc = dir ? p.f.N : p.f.M;   // (multiples of 4: every tile is whole 16-byte units)
                    const float* Rq = (dir ? p.f.Bp : p.f.A) + ((size_t)b * nq + (size_t)rb * kItemRows) * 3;
                    const float* Rc = (dir ? p.f.A : p.f.Bp) + (size_t)b * nc * 3;
                    for (int t = -1; t < ntiles; ++t, ++h) {
                        const unsigned s = h % kCStages, n = h / kCStages;
                        const int left = t < 0 ? nq - rb * kItemRows : nc - t * kTNc;
                        const unsigned bytes = (unsigned)max(0, min(left, kTNc)) * 12u;
                        mbar_wait(&cempty[s], (n & 1) ^ 1);
                        if (bytes) {
                            mbar_expect_tx(&cfull[s], bytes);
                            tma_bulk_g2s(s_c + s * kTNc, t < 0 ? Rq : Rc + (size_t)t * kTNc * 3, bytes, &cfull[s]);
                        } else {
                            mbar_arrive(&cfull[s]);   // a tile of padding only
                        }
                    }
                } else {
                    for (int t = -1; t < ntiles; ++t, ++h) {   // t = -1: the item's 256 query rows
                        const unsigned s = h % kCStages, n = h / kCStages;
                        mbar_wait(&cempty[s], (n & 1) ^ 1);
                        mbar_expect_tx(&cfull[s], kTNc * 16);
                        tma_bulk_g2s(s_c + s * kTNc, t < 0 ? Pq : Pc + (size_t)t * kTNc, kTNc * 16, &cfull[s]);
                    }
                }
            }
        }
      } else if (warp == kWarpMma) {
        if (lane == 0) {
            // The two row tiles' accumulators alternate: while the read-out warps of row tile r pull tile t out of TMEM, the
            // MMAs of row tile 1-r run — each accumulator is single-buffered, the pair is the double buffer.
            unsigned g = 0, it = 0;
            const unsigned a0 = smem_u32(s_a), b0 = smem_u32(s_b);
            PROF_DECL;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
                const int b = item / ipe, k = item - b * ipe;
                const int ntiles = (k >= p.rbA ? p.NpA : p.NpB) / kTNc;
                const unsigned ab = it & 1;
                mbar_wait(&a_full[ab], (it >> 1) & 1);
                PROF(0);
                for (int t = 0; t < ntiles; ++t, ++g) {
                    const unsigned s = g % kStages, n = g / kStages;
                    mbar_wait(&full_b[s], n & 1);
                    PROF(1);
#pragma unroll
                    for (int r = 0; r < kRT; ++r) {
                        mbar_wait(&tempty[r], (g & 1) ^ 1);
                        PROF(2 + r);
                        tc_fence_after();
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            umma_tf32(tmem + (unsigned)(r * kTNc), umma_desc64(a0 + (ab * kRT + r) * kTQ * kRowB + ks * 32),
                                      umma_desc64(b0 + s * kTNc * kRowB + ks * 32), kIdescTc, ks > 0);
                        umma_commit(&tfull[r]);
                    }
                    umma_commit(&empty_b[s]);
                    PROF(4);
                }
                umma_commit(&a_empty[ab]);  // the item's MMAs have read its query tiles
            }
            PROF_OUT(0, 5);
        }
      } else if (warp > kWarpMma && warp < kWarpCert) {
        // ---- uploaders (host arrays only): the two spare warps of every CTA; every warp moves its share of every element and counts
        // it on its own, so the grid's warps spread over two or three elements and keep the PCIe read queue full ----------------
        if (kRaw) {
            const unsigned na4 = (unsigned)p.f.N * 3u / 4u, nb4 = (unsigned)p.f.M * 3u / 4u;   // float4 units per element (N, M multiples of 4)
            const int W = 2 * (int)gridDim.x, w = 2 * (int)blockIdx.x + (warp - kWarpMma - 1);
            const int G = p.up_groups, g = w % G, wg = w / G, Wg = (W - g + G - 1) / G;
            for (int b = g; b < p.B; b += G) {
                upload_pair_warp(reinterpret_cast<const float4*>(p.hA) + (size_t)b * na4, reinterpret_cast<float4*>(const_cast<float*>(p.f.A)) + (size_t)b * na4, na4,
                                 reinterpret_cast<const float4*>(p.hB) + (size_t)b * nb4, reinterpret_cast<float4*>(const_cast<float*>(p.f.Bp)) + (size_t)b * nb4, nb4,
                                 wg, Wg, lane);
                __threadfence();   // this lane's stores are visible device-wide ...
                __syncwarp();      // ... for every lane of the warp ...
                if (lane == 0) atomicAdd(p.arrived + b, 1u);  // ... before the element counts as delivered by this warp
            }
        }
      } else if (warp < kWarpProducer) {
        // ---- converters: thread <-> points pt and pt + 128 of the 256-point tile; compact {x, y, z, n} -> 16 TF32 pieces in the
        // operand row ----------------------------------------------------------------------------------------------------------
        const int pt = (warp - kEpiWarps) * 32 + lane;
        const unsigned sw = (unsigned)((pt >> 1) & 3);   // (pt + 128) has the same swizzle phase
        unsigned h = 0, g = 0, it = 0;
        PROF_DECL;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int b = item / ipe, k = item - b * ipe;
            const int ntiles = (k >= p.rbA ? p.NpA : p.NpB) / kTNc;
            const unsigned ab = it & 1;
            const bool dirc = k >= p.rbA;
            float cx = 0.f, cy = 0.f, cz = 0.f, mxn = 0.f;
            if (kRaw) {
                // the centre of the batch element, with the operations of chamfer_tc_prepare_kernel (the same bits in every CTA and
                // in the certifiers): converter warp 0 fetches the 32 + 32 sample points once the element has arrived
                if (warp == kEpiWarps) {
                    if (lane == 0) while (ld_acquire_i32(reinterpret_cast<const int*>(p.arrived) + b) < upload_group_size(2 * (int)gridDim.x, b, p.up_groups)) __nanosleep(100);
                    __syncwarp();
                    const float* gA = p.f.A + (size_t)b * p.f.N * 3;
                    const float* gB = p.f.Bp + (size_t)b * p.f.M * 3;
                    const int ia = (int)(((long)lane * p.f.N) >> 5), ib = (int)(((long)lane * p.f.M) >> 5);
                    float sx = __ldcg(gA + 3 * ia) + __ldcg(gB + 3 * ib);
                    float sy = __ldcg(gA + 3 * ia + 1) + __ldcg(gB + 3 * ib + 1);
                    float sz = __ldcg(gA + 3 * ia + 2) + __ldcg(gB + 3 * ib + 2);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        sx += __shfl_xor_sync(0xffffffffu, sx, o);
                        sy += __shfl_xor_sync(0xffffffffu, sy, o);
                        sz += __shfl_xor_sync(0xffffffffu, sz, o);
                    }
                    if (lane == 0) { s_ctr[it & 3][0] = sx * (1.0f / 64.0f); s_ctr[it & 3][1] = sy * (1.0f / 64.0f); s_ctr[it & 3][2] = sz * (1.0f / 64.0f); }
                }
                named_bar_sync(2, kConvWarps * 32);
                cx = s_ctr[it & 3][0]; cy = s_ctr[it & 3][1]; cz = s_ctr[it & 3][2];
            }
            mbar_wait(&a_empty[ab], ((it >> 1) & 1) ^ 1);
            for (int t = -1; t < ntiles; ++t, ++h) {
                const unsigned cs = h % kCStages;
                PROF(2);
                mbar_wait(&cfull[cs], (h / kCStages) & 1);
                PROF(0);
                float4 v[2];
                if (kRaw) {
                    // raw points (12 B each): centre, norm — what chamfer_tc_prepare_kernel writes for resident inputs
                    const float* raw = reinterpret_cast<const float*>(s_c + cs * kTNc);
                    const int n = t < 0 ? (dirc ? p.f.M : p.f.N) - (k - (dirc ? p.rbA : 0)) * kItemRows : (dirc ? p.f.N : p.f.M) - t * kTNc;   // points left from this tile on
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int r = pt + 128 * u;
                        float x = 0.f, y = 0.f, z = 0.f, nrm = kPadN;
                        if (r < n) {
                            x = raw[3 * r] - cx; y = raw[3 * r + 1] - cy; z = raw[3 * r + 2] - cz;
                            nrm = fmaf(z, z, fmaf(y, y, x * x));
                            if (t >= 0) mxn = fmaxf(mxn, nrm);
                        }
                        v[u] = make_float4(x, y, z, nrm);
                    }
                    if (t == ntiles - 1) {   // the item's largest candidate norm, published before the last tile is handed over
                        const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(mxn));   // norms are >= 0: bit order == value order
                        if (lane == 0) s_maxw[it & 3][warp - kEpiWarps] = m;
                    }
                } else {
                    v[0] = s_c[cs * kTNc + pt];
                    v[1] = s_c[cs * kTNc + pt + 128];
                }
                // The compact slot is released only AFTER the stores below, which consume v: an mbarrier.arrive does not
                // wait for an LDS in flight, and a 16-byte warp load executes in four quarter-warp passes — released right
                // after the load, the slot was now and then refilled by the TMA before the last pass had read it (rows
                // 24-31 of a tile then held the points of the tile one ring turn later).
                unsigned char* dst;
                unsigned s = 0;
                if (t < 0) {
                    // query role: rows 0..127 are row tile 0, rows 128..255 row tile 1 (contiguous 8 KB each)
                    dst = s_a + ab * kRT * kTQ * kRowB + (unsigned)pt * kRowB;
                } else {
                    s = g % kStages;
                    mbar_wait(&empty_b[s], ((g / kStages) & 1) ^ 1);
                    PROF(1);
                    dst = s_b + s * kTNc * kRowB + (unsigned)pt * kRowB;
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    float xh, xl, yh, yl, zh, zl, nh, nl;
                    unsigned char* d = dst + u * 128 * kRowB;
                    if (t < 0) {
                        split_tf32(v[u].x, xh, xl); split_tf32(v[u].y, yh, yl); split_tf32(v[u].z, zh, zl); split_tf32(v[u].w, nh, nl);
                        *reinterpret_cast<float4*>(d + ((0u ^ sw) << 4)) = make_float4(xh, xh, xl, xl);
                        *reinterpret_cast<float4*>(d + ((1u ^ sw) << 4)) = make_float4(yh, yh, yl, yl);
                        *reinterpret_cast<float4*>(d + ((2u ^ sw) << 4)) = make_float4(zh, zh, zl, zl);
                        *reinterpret_cast<float4*>(d + ((3u ^ sw) << 4)) = make_float4(nh, nl, 1.0f, 1.0f);
                    } else {
                        split_tf32(-2.0f * v[u].x, xh, xl); split_tf32(-2.0f * v[u].y, yh, yl); split_tf32(-2.0f * v[u].z, zh, zl); split_tf32(v[u].w, nh, nl);
                        *reinterpret_cast<float4*>(d + ((0u ^ sw) << 4)) = make_float4(xh, xl, xh, xl);
                        *reinterpret_cast<float4*>(d + ((1u ^ sw) << 4)) = make_float4(yh, yl, yh, yl);
                        *reinterpret_cast<float4*>(d + ((2u ^ sw) << 4)) = make_float4(zh, zl, zh, zl);
                        *reinterpret_cast<float4*>(d + ((3u ^ sw) << 4)) = make_float4(1.0f, 1.0f, nh, nl);
                    }
                }
                proxy_fence_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&cempty[cs]);
                    mbar_arrive(t < 0 ? &a_full[ab] : &full_b[s]);
                }
                if (t >= 0) ++g;
            }
        }
        if (warp == kEpiWarps && lane == 0) PROF_OUT(5, 3);
      } else if (warp >= kWarpCert) {
        // ---- certifiers: the item that has just left the read-out is certified and re-evaluated in the reference arithmetic
        // while the tensor cores work on the next one.  Phase 1, lane <-> row: merge the four column quarters' locator records
        // and decide — b2 > b1 + window: the exact argmin lies in chunk c1; else b3 > b1 + window: in c1 or c2; else ambiguous
        // (listed for the cleanup kernel).  Phase 2, warp <-> row, lane <-> candidate of the located 32-candidate chunk: three
        // coalesced loads, one distance, REDUX.MIN + ballot for (minimum, lowest index) — four rows in flight.
        const TcFinParams& f = p.f;
        const int cw = warp - kWarpCert;
        const unsigned full = 0xffffffffu;
        if (p.wait_prepare) asm volatile("griddepcontrol.wait;" ::: "memory");   // norms and norm maxima come from the prepare grid
        unsigned it = 0, cphase = 0;
        // this warp's barrier, as an address of its own: folded into another shared-memory base as "[R + -0xc0]" the operand is
        // mis-tracked by compute-sanitizer's synccheck ("missing init" on an initialised barrier)
        unsigned long long* cbar = &cert_bar[cw];
        asm volatile("" : "+l"(cbar));
        PROF_DECL;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int b = item / ipe, k = item - b * ipe;
            const bool dir = k >= p.rbA;                       // false: rows of A search B; true: rows of B search A
            const int rb = dir ? k - p.rbA : k;
            const int Q = dir ? f.M : f.N, R = dir ? f.N : f.M;        // queries / searched points per element
            const float* gQ = (dir ? f.Bp : f.A) + (size_t)b * Q * 3;
            const float* gP = (dir ? f.A : f.Bp) + (size_t)b * R * 3;
            const unsigned pb = it & 1;
            const int idbits = tc_idbits((dir ? p.NpA : p.NpB) / kTNc);
            const unsigned idmask = (1u << idbits) - 1u;
            const int rin = cw * 32 + lane;
            const int q = rb * kItemRows + rin;
            const bool valid = q < Q;
            // what does not depend on the sweep is fetched before the wait (device-resident inputs; in upload mode the element
            // may still be crossing PCIe — its points are only known to have arrived once the item's records are here)
            float qx = 0.f, qy = 0.f, qz = 0.f, nq = 0.f, other = 0.f;
            const bool early = !kRaw;
            if (valid && early) {
                nq = __ldcg(&((dir ? p.PB : p.PA) + (size_t)b * (dir ? p.NpB : p.NpA) + q)->w);
                unsigned om = 0u;
                for (int kb = 0; kb < f.nblk; ++kb) om = max(om, __ldcg(f.maxn + ((size_t)(dir ? 0 : 1) * p.B + b) * f.nblk + kb));   // (same address in every lane)
                other = __uint_as_float(om);
                qx = __ldcg(gQ + 3 * (size_t)q); qy = __ldcg(gQ + 3 * (size_t)q + 1); qz = __ldcg(gQ + 3 * (size_t)q + 2);
            }
            PROF(4);
            mbar_wait(&pub_full, it & 1);
            PROF(0);
            Loc3 e;
            e.b1 = INFINITY; e.b2 = INFINITY; e.b3 = INFINITY; e.c1 = 0; e.c2 = 0;
            {
                // quarter q4's record: values with the quarter's running chunk number (2 * tile + chunk) in their low bits ->
                // global chunk id = tile * 8 + q4 * 2 + chunk.  Quarters hold disjoint chunk sets: insert the two located minima;
                // the third value can only be third or later
                const float4* src = s_pub + rin;
#pragma unroll
                for (int q4 = 0; q4 < kParts; ++q4) {
                    const float4 o = src[q4 * kItemRows];
                    const unsigned i1 = __float_as_uint(o.x) & idmask, i2 = __float_as_uint(o.y) & idmask;
                    loc3_insert(e, o.x, (int)((i1 >> 1) * (kTNc / kTcChunk) + q4 * (kTNc / kParts / kTcChunk) + (i1 & 1u)));
                    loc3_insert(e, o.y, (int)((i2 >> 1) * (kTNc / kTcChunk) + q4 * (kTNc / kParts / kTcChunk) + (i2 & 1u)));
                    e.b3 = fminf(e.b3, o.z);
                }
            }
            // (released after the merge has consumed the loads: an arrive does not wait for an LDS in flight)
            __syncwarp();
            if (lane == 0) mbar_arrive(&pub_empty);
            // ---- phase 1 ----
            float best = 0.f, win = 0.f, errlim = 0.f;
            int loc1 = -1, loc2 = -1;      // chunks to re-evaluate (-1: none)
            bool amb = false;
            if (valid) {
                if (!early) {   // host arrays (the element had arrived before this item's first tile; L2 loads): the norms as the converters took them
                    qx = __ldcg(gQ + 3 * (size_t)q); qy = __ldcg(gQ + 3 * (size_t)q + 1); qz = __ldcg(gQ + 3 * (size_t)q + 2);
                    const float x = qx - s_ctr[it & 3][0], y = qy - s_ctr[it & 3][1], z = qz - s_ctr[it & 3][2];
                    nq = fmaf(z, z, fmaf(y, y, x * x));
                    other = __uint_as_float(max(max(s_maxw[it & 3][0], s_maxw[it & 3][1]), max(s_maxw[it & 3][2], s_maxw[it & 3][3])));
                }
                best = fmaxf(e.b1, 0.0f);
                // the chunk numbers embedded in the located values moved them by up to 2^(idbits-23) relative
                const float idrel = ldexpf(1.0f, idbits - 23);
                win = fmaf(kTcWinRel + 2.0f * idrel, best, kTcWinAbs * (nq + other));
                errlim = fmaf(idrel, best, kTcErrAbs * (nq + other));
                // written so that NaN / inf / out-of-range norms can only make the row ambiguous, never certified
                const bool sane = nq <= kTcNormLimit && other <= kTcNormLimit && nq + other >= kTcNormFloor;
                const bool one = sane && fmaxf(e.b2, 0.0f) > best + win;
                const bool two = sane && !one && fmaxf(e.b3, 0.0f) > best + win;
                amb = !(one || two);
                if (!amb) loc1 = e.c1;
                if (two) loc2 = e.c2;
            }
            PROF(1);
            // ---- phase 2: the located chunk (both located chunks) in the reference arithmetic.  Every row fetches its chunk —
            // 32 candidates = 384 contiguous bytes — with ONE bulk copy (TMA) into its own shared-memory row: a gather of 256
            // chunks per item that costs the SM eight instructions; the scan then runs thread <-> row out of shared memory ----
            float d = INFINITY;
            int j = 0x7fffffff;
            unsigned char* myrow = s_chunk + rin * kChunkPitch;
            if (kRaw) asm volatile("fence.proxy.async;" ::: "memory");   // host arrays: the points were stored by uploader warps while this kernel runs
            for (int pass = 0; pass < 2; ++pass) {
                const int c = pass == 0 ? loc1 : loc2;
                if (pass == 1 && !__any_sync(full, c >= 0)) break;   // (about one warp in nine holds a two-chunk row)
                const int j0 = c * kTcChunk;
                const float* pp = gP + 3 * (size_t)(c < 0 ? 0 : j0);
                const bool staged = c >= 0 && j0 + kTcChunk <= R && (reinterpret_cast<uintptr_t>(pp) & 15u) == 0;
                // one arrival per warp and pass, carrying the bytes of all its rows' copies
                const unsigned nstaged = __popc(__ballot_sync(full, staged));
                if (lane == 0) mbar_expect_tx(cbar, nstaged * (kTcChunk * 12));
                __syncwarp();
                if (staged) {
                    tma_bulk_g2s(myrow, pp, kTcChunk * 12, cbar);
                } else {
                    if (c >= 0) {   // ragged end of the cloud, or a cloud that is not 16-byte aligned there: direct loads
                        const int j1 = min(j0 + kTcChunk, R);
#pragma unroll 4
                        for (int jj = j0; jj < j1; ++jj) {
                            const float dd = sqdist3<false>(qx, qy, qz, __ldcg(gP + 3 * (size_t)jj), __ldcg(gP + 3 * (size_t)jj + 1), __ldcg(gP + 3 * (size_t)jj + 2));
                            if (dd < d || (dd == d && jj < j)) { d = dd; j = jj; }
                        }
                    }
                }
                mbar_wait(cbar, cphase);
                cphase ^= 1u;
                if (staged) {
                    // (a - b)² is the same bits as (b - a)²: one operand order serves both directions of the reference's A .- B
#pragma unroll 2
                    for (int h = 0; h < kTcChunk / 4; ++h) {  // 4 candidates = three 16-byte shared-memory loads per trip
                        const float4 v0 = *reinterpret_cast<const float4*>(myrow + h * 48);
                        const float4 v1 = *reinterpret_cast<const float4*>(myrow + h * 48 + 16);
                        const float4 v2 = *reinterpret_cast<const float4*>(myrow + h * 48 + 32);
                        const float cc[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float dd = sqdist3<false>(qx, qy, qz, cc[3 * u], cc[3 * u + 1], cc[3 * u + 2]);
                            const int jj = j0 + h * 4 + u;
                            if (dd < d || (dd == d && jj < j)) { d = dd; j = jj; }   // lowest index on ties, also across the two chunks
                        }
                    }
                }
                __syncwarp();
                PROF(2 + pass);
            }
            double mine = 0.0;
            if (valid && !amb) {
                // self-check of the error bound at the located minimum
                if (!(fabsf(d - best) <= fmaf(kTcErrRel, d, errlim))) {
                    amb = true;
                    atomicAdd(f.hdr + kHdrViol, 1);
                } else {
                    int32_t* nn = dir ? f.nnB : f.nnA;
                    if (nn) nn[(size_t)b * Q + q] = j;
                    mine = (double)d;
                }
            }
            // ---- the item's partial sum (fixed order) and its ambiguous rows, listed in row order at fixed slots ----
            const unsigned ambmask = __ballot_sync(full, valid && amb);
            mine = warp_sum(mine);
            if (lane == 0) { s_part[pb][cw] = mine; s_cnt[pb][cw] = __popc(ambmask); }
            named_bar_sync(1, kCertWarps * 32);   // (s_part / s_cnt are double-buffered by item parity: one barrier per item is enough)
            if (ambmask) {
                int lbase = 0;
                if (lane == 0) lbase = atomicAdd(f.hdr + kHdrAmb, __popc(ambmask));   // the work list's order is arbitrary; results go to fixed slots
                lbase = __shfl_sync(full, lbase, 0);
                if (valid && amb) {
                    const int wpos = __popc(ambmask & ((1u << lane) - 1u));
                    int pos = wpos;
                    for (int w = 0; w < cw; ++w) pos += s_cnt[pb][w];
                    const size_t slot = (size_t)item * kItemRows + pos;
                    f.ambq[slot] = q;
                    f.amblim[slot] = best + win;
                    f.amblist[lbase + wpos] = (int)slot;
                }
            }
            if (cw == 0 && lane == 0) {
                double s2 = 0.0;
                int c = 0;
#pragma unroll
                for (int w = 0; w < kCertWarps; ++w) { s2 += s_part[pb][w]; c += s_cnt[pb][w]; }
                f.partial[item] = s2;
                f.ambcnt[item] = c;
            }
        }
#ifdef F3D_TC_PROF
        if (cw == 0 && lane == 0 && blockIdx.x < 148) {
            for (int i_ = 0; i_ < 5; ++i_) g_certprof[blockIdx.x * 8 + i_] = pf_[i_];
            unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); g_certprof[blockIdx.x * 8 + 6] = (long long)g_;   // certifier end
        }
#endif
      }
    } else {
        // the four read-out warpgroups take them: both tcgen05.ld of an accumulator quarter in flight = 64 registers of data
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
        // ---- read-out: thread <-> (query row of BOTH row tiles, column quarter): TMEM lane quad*32 + lane, columns part*64 .. +63
        // of whichever accumulator has been filled — all 16 warps pull a finished accumulator out together while the MMAs of
        // the other row tile run ------------------------------------------------------------------------------------------------
        const int quad = warp & 3, part = warp >> 2;
        unsigned g = 0, it = 0;
        PROF_DECL;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int b = item / ipe, k = item - b * ipe;
            const bool dir = k >= p.rbA;
            const int rb = dir ? k - p.rbA : k;
            const int npq = dir ? p.NpB : p.NpA, npc = dir ? p.NpA : p.NpB;
            const int ntiles = npc / kTNc;
            const int rin = quad * 32 + lane;               // row within its row tile
            float* tm = p.tilemin + (size_t)kParts * ((size_t)b * ((size_t)p.nstB * p.NpA + (size_t)p.nstA * p.NpB) + (dir ? (size_t)p.nstB * p.NpA : 0)) +
                        (size_t)rb * kItemRows + rin;
            Loc3v loc[kRT];
            float stmin[kRT];
#pragma unroll
            for (int r = 0; r < kRT; ++r) { loc[r].b1 = INFINITY; loc[r].b2 = INFINITY; loc[r].b3 = INFINITY; stmin[r] = INFINITY; }
            const unsigned idmask = (1u << tc_idbits(ntiles)) - 1u;
            for (int t = 0; t < ntiles; ++t, ++g) {
#pragma unroll
                for (int r = 0; r < kRT; ++r) {
                    PROF(2);
                    mbar_wait(&tfull[r], g & 1);
                    PROF(0);
                    tc_fence_after();
                    const unsigned base = tmem + ((unsigned)(quad * 32) << 16) + (unsigned)(r * kTNc + part * (kTNc / kParts));
                    // both 32-column loads of this quarter are issued before the (single) wait: one TMEM round trip per accumulator
                    unsigned v[2][32];
                    tmem_ld32_issue(base, v[0]);
                    tmem_ld32_issue(base + 32, v[1]);
                    tmem_ld_wait(v[0]);
                    tmem_ld_wait(v[1]);
                    // the accumulator quarter is in registers: hand the accumulator back NOW — the next MMAs into it run while
                    // the minima below are taken (the accumulator is busy for one TMEM round trip, not for the whole read-out)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[r]);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float m = min32(v[q]);
                        loc3v_insert(loc[r], __uint_as_float((__float_as_uint(m) & ~idmask) | (unsigned)(2 * t + q)));
                        stmin[r] = fminf(stmin[r], m);
                    }
                    PROF(1);
                }
                if ((t & (kTilesPerSuper - 1)) == kTilesPerSuper - 1 || t == ntiles - 1) {
#pragma unroll
                    for (int r = 0; r < kRT; ++r) {
                        tm[((size_t)(t / kTilesPerSuper) * kParts + part) * npq + r * kTQ] = stmin[r];
                        stmin[r] = INFINITY;
                    }
                }
            }
            // hand the two rows' triples to the certifier warps (double-buffered by item parity)
            // (single buffer: the certifiers merge an item's records as soon as they are complete, an item's time before the next write)
            mbar_wait(&pub_empty, (it & 1) ^ 1);
#pragma unroll
            for (int r = 0; r < kRT; ++r)
                s_pub[part * kItemRows + r * kTQ + rin] = make_float4(loc[r].b1, loc[r].b2, loc[r].b3, 0.f);
            __syncwarp();
            if (lane == 0) mbar_arrive(&pub_full);
        }
        if ((warp == 0 || warp == 8) && lane == 0) PROF_OUT(8 + (warp >> 3) * 3, 3);
#ifdef F3D_TC_PROF
        if (tid == 0 && blockIdx.x < 148) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); g_tcprof[blockIdx.x * 16 + 14] = (long long)g_; }  // the read-out's end
#endif
    }
    tc_fence_before();
    __syncthreads();
#ifdef F3D_TC_PROF
    if (tid == 0 && blockIdx.x < 148) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); g_tcprof[blockIdx.x * 16 + 15] = (long long)g_; }
#endif
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, kRT * kTNc); }
}

// ---- finalize: certify, re-evaluate exactly, reduce the loss -----------------------------------------------------------
#ifdef F3D_TC_PROF
__device__ unsigned long long g_cleanprof[4096 * 4];
#define CPROF(k) do { if (threadIdx.x == 0 && blockIdx.x < 4096) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); g_cleanprof[blockIdx.x * 4 + (k)] = g_; } } while (0)
#else
#define CPROF(k)
#endif

// Cleanup (after the sweep): the rows the certifier warps could not certify (a handful on uniform clouds; every row of
// tie-heavy inputs) are rescanned by a whole block each, taken from the work list: every supertile whose filter minimum lies
// within the row's window is scanned in the reference arithmetic; the distance found goes to the row's FIXED slot.  The last
// block then adds up the items' partial sums and those slots in a fixed order (bitwise repeatable).
constexpr int kCleanT = 1024;   // wide blocks: the last one adds up thousands of partial sums in one round trip
__global__ void __launch_bounds__(kCleanT, 1) chamfer_tc_cleanup_kernel(TcFinParams p) {
    constexpr int kMaxSel = 32;
    __shared__ int s_sel[kMaxSel], s_nsel;
    __shared__ unsigned s_wd[kCleanT / 32];
    __shared__ int s_wj[kCleanT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ipe = p.rbA + p.rbB;
    asm volatile("griddepcontrol.wait;" ::: "memory");   // launched programmatically: the sweep grid has ended, its stores are visible
    CPROF(0);
    const int total_amb = __ldcg(p.hdr + kHdrAmb);
    for (int e = blockIdx.x; e < total_amb; e += gridDim.x) {
        const int slot = __ldcg(p.amblist + e);
        const int item = slot / kItemRows;
        const int b = item / ipe, k = item - b * ipe;
        const bool dir = k >= p.rbA;
        const int Q = dir ? p.M : p.N, R = dir ? p.N : p.M;
        const int npq = dir ? p.NpB : p.NpA;
        const int nst = dir ? p.nstA : p.nstB;                      // supertiles of the searched cloud
        const float* gQ = (dir ? p.Bp : p.A) + (size_t)b * Q * 3;
        const float* gP = (dir ? p.A : p.Bp) + (size_t)b * R * 3;
        const float* tmb = p.tilemin + (size_t)kParts * ((size_t)b * ((size_t)p.nstB * p.NpA + (size_t)p.nstA * p.NpB) + (dir ? (size_t)p.nstB * p.NpA : 0));
        const int qq = __ldcg(p.ambq + slot);
        const float lim = __ldcg(p.amblim + slot);  // NaN / inf -> scan everything (comparison below)
        const float ax = __ldg(gQ + 3 * (size_t)qq), ay = __ldg(gQ + 3 * (size_t)qq + 1), az = __ldg(gQ + 3 * (size_t)qq + 2);
        float d = INFINITY;
        int j = 0x7fffffff;
        if (tid == 0) s_nsel = 0;
        __syncthreads();
        for (int t0 = 0; t0 < nst; t0 += kCleanT) {
            const int t = t0 + tid;
            if (t < nst) {
                float e1 = __ldcg(tmb + (size_t)t * kParts * npq + qq);   // the supertile's minimum: one partial per column quarter
#pragma unroll
                for (int q4 = 1; q4 < kParts; ++q4) e1 = fminf(e1, __ldcg(tmb + ((size_t)t * kParts + q4) * npq + qq));
                if (!(fmaxf(e1, 0.0f) > lim)) {
                    const int pos = atomicAdd(&s_nsel, 1);
                    if (pos < kMaxSel) s_sel[pos] = t;
                }
            }
        }
        __syncthreads();
        const int nsel = s_nsel;
        const bool all = nsel > kMaxSel;   // tie-heavy input: more supertiles within the window than the list holds
        const int total = all ? ((R + 3) & ~3) : nsel * kSuper;
        for (int idx = 4 * tid; idx < total; idx += 4 * kCleanT) {
            const int sidx = idx / kSuper, jb = all ? idx : s_sel[sidx] * kSuper + (idx - sidx * kSuper);
            if (jb >= R) continue;
            const float* pp = gP + 3 * (size_t)jb;
            float c[12];
            if (jb + 4 <= R && (reinterpret_cast<uintptr_t>(pp) & 15u) == 0) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(pp));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(pp) + 1);
                const float4 v2 = __ldg(reinterpret_cast<const float4*>(pp) + 2);
                c[0] = v0.x; c[1] = v0.y; c[2] = v0.z; c[3] = v0.w; c[4] = v1.x; c[5] = v1.y;
                c[6] = v1.z; c[7] = v1.w; c[8] = v2.x; c[9] = v2.y; c[10] = v2.z; c[11] = v2.w;
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int jj = min(jb + u, R - 1);  // clamped duplicates are harmless (same value, same index)
                    c[3 * u] = __ldg(gP + 3 * (size_t)jj); c[3 * u + 1] = __ldg(gP + 3 * (size_t)jj + 1); c[3 * u + 2] = __ldg(gP + 3 * (size_t)jj + 2);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dd = dir ? sqdist3<false>(c[3 * u], c[3 * u + 1], c[3 * u + 2], ax, ay, az)
                                     : sqdist3<false>(ax, ay, az, c[3 * u], c[3 * u + 1], c[3 * u + 2]);
                const int jj = min(jb + u, R - 1);
                if (dd < d || (dd == d && jj < j)) { d = dd; j = jj; }
            }
        }
        // (d, j) minimum over the block, lowest j on ties (d >= 0 or +inf: bit order == value order)
        const unsigned mb = __reduce_min_sync(0xffffffffu, __float_as_uint(d));
        const int jm = __reduce_min_sync(0xffffffffu, (__float_as_uint(d) == mb) ? j : 0x7fffffff);
        if (lane == 0) { s_wd[warp] = mb; s_wj[warp] = jm; }
        __syncthreads();
        if (tid == 0) {
            unsigned bd = s_wd[0];
            int bj = s_wj[0];
#pragma unroll
            for (int u = 1; u < kCleanT / 32; ++u) {
                const unsigned od = s_wd[u];
                const int oj = s_wj[u];
                if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
            }
            int32_t* nn = dir ? p.nnB : p.nnA;
            if (nn) nn[(size_t)b * Q + qq] = bj;
            p.ambd[slot] = __uint_as_float(bd);
        }
        __syncthreads();  // s_wd / s_wj / s_nsel are reused by the next row
    }
    // ---- the grid's last block adds up every partial sum in a fixed order (bitwise repeatable).  It is the designated
    // reducer (the grid is at most one block per SM: all blocks are resident together): the items' sums are in its registers
    // by the time the other blocks have finished their rows ---------------------------------------------------------------
    CPROF(1);
    if (blockIdx.x != gridDim.x - 1) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(p.hdr + kHdrDone, 1);
        }
        return;
    }
    double sa = 0.0, sb = 0.0;
    for (int k0 = 0; k0 < p.nitems; k0 += 4 * kCleanT) {   // per item: four of them in flight per thread
        double v[4];
        int cn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int kk = k0 + u * kCleanT + tid;
            v[u] = kk < p.nitems ? __ldcg(p.partial + kk) : 0.0;
            cn[u] = kk < p.nitems ? __ldcg(p.ambcnt + kk) : 0;
        }
        if (k0 == 0) {   // the other blocks' distances (ambd) are complete and visible from here on
            if (tid == 0)
                while (ld_acquire_i32(p.hdr + kHdrDone) < (int)gridDim.x - 1) __nanosleep(64);
            __syncthreads();
            CPROF(2);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int kk = k0 + u * kCleanT + tid;
            if (kk < p.nitems) {
                double t = v[u];
                for (int e = 0; e < cn[u]; ++e) t += (double)__ldcg(p.ambd + (size_t)kk * kItemRows + e);   // the item's ambiguous rows, in row order
                if (kk % ipe < p.rbA) sa += t; else sb += t;
            }
        }
    }
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    __shared__ double s_a2[kCleanT / 32], s_b2[kCleanT / 32];
    if (lane == 0) { s_a2[warp] = sa; s_b2[warp] = sb; }
    __syncthreads();
    if (tid == 0) {
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < kCleanT / 32; ++w) { ta += s_a2[w]; tb += s_b2[w]; }
        const float dAB = (float)(ta / p.denomA), dBA = (float)(tb / p.denomB);  // pcloud.jl:47-48
        if (p.terms) { p.terms[0] = dAB; p.terms[1] = dBA; }
        const float l = __fadd_rn(__fmul_rn(p.w1, dAB), __fmul_rn(p.w2, dBA));   // pcloud.jl:50
        if (p.peer.nranks <= 1) p.loss[0] = l;
        s_a2[0] = (double)l;
    }
    CPROF(3);
    if (p.peer.nranks <= 1) return;
    // the one exchange of the sharded path, fused (see chamfer.cu: chamfer_filter_finalize_kernel)
    __shared__ float s_v[kMaxPeerRanks];
    __syncthreads();
    if (tid < p.peer.nranks) s_v[tid] = peer_exchange(p.peer, tid, (float)s_a2[0]);
    __syncthreads();
    if (tid == 0) {
        float sum = 0.0f;
        for (int rr = 0; rr < p.peer.nranks; ++rr) sum = __fadd_rn(sum, s_v[rr]);
        p.loss[0] = sum;
    }
}

struct TcPlan {
    int NpA, NpB, rbA, rbB, nstA, nstB, nitems, nblk;
    size_t off_PA, off_PB, off_tilemin, off_partial, off_ambcnt, off_ambq, off_amblim, off_amblist, off_ambd, zero_from, off_hdr, off_maxn, off_arrived, zero_bytes, total;
};

TcPlan make_tc_plan(int B, int N, int M) {
    TcPlan pl;
    pl.NpA = (int)align_up((size_t)N, kItemRows);
    pl.NpB = (int)align_up((size_t)M, kItemRows);
    pl.rbA = pl.NpA / kItemRows;
    pl.rbB = pl.NpB / kItemRows;
    pl.nstA = (pl.NpA + kSuper - 1) / kSuper;
    pl.nstB = (pl.NpB + kSuper - 1) / kSuper;
    pl.nitems = B * (pl.rbA + pl.rbB);
    size_t o = 0;
    pl.off_hdr = o;     o = align_up(o + sizeof(int) * kHdrInts, 256);   // diagnostics first: a caller can find them
    pl.off_arrived = o; o = align_up(o + sizeof(int) * (size_t)B, 256);   // host arrays: arrived [B]
    pl.zero_from = 0;
    pl.zero_bytes = o;  // host arrays only: header and arrival counters are zeroed by ONE memset per call (resident inputs: by the prepare grid)
    pl.nblk = (std::max(pl.NpA, pl.NpB) + kPrepT * kPrepPts - 1) / (kPrepT * kPrepPts);
    pl.off_maxn = o;    o = align_up(o + sizeof(unsigned) * 2 * (size_t)B * pl.nblk, 256);
    pl.off_PA = o;      o = align_up(o + sizeof(float4) * (size_t)B * pl.NpA, 256);
    pl.off_PB = o;      o = align_up(o + sizeof(float4) * (size_t)B * pl.NpB, 256);
    pl.off_tilemin = o; o = align_up(o + sizeof(float) * kParts * (size_t)B * ((size_t)pl.nstB * pl.NpA + (size_t)pl.nstA * pl.NpB), 256);
    pl.off_partial = o; o = align_up(o + sizeof(double) * (size_t)pl.nitems, 256);
    pl.off_ambcnt = o;  o = align_up(o + sizeof(int) * (size_t)pl.nitems, 256);
    pl.off_ambq = o;    o = align_up(o + sizeof(int) * (size_t)pl.nitems * kItemRows, 256);
    pl.off_amblim = o;  o = align_up(o + sizeof(float) * (size_t)pl.nitems * kItemRows, 256);
    pl.off_amblist = o; o = align_up(o + sizeof(int) * (size_t)pl.nitems * kItemRows, 256);
    pl.off_ambd = o;    o = align_up(o + sizeof(float) * (size_t)pl.nitems * kItemRows, 256);
    pl.total = o;
    return pl;
}

constexpr size_t kTcSmem = 2 * kRT * kTQ * kRowB + kStages * kTNc * kRowB + kCStages * kTNc * 16 + kParts * kItemRows * 16 + kItemRows * kChunkPitch + 1024;  // 213 KB

}  // namespace

size_t chamfer_tc_workspace_bytes(int B, int N, int M) { return make_tc_plan(B, N, M).total; }
#ifdef F3D_TC_PROF
extern "C" __attribute__((visibility("default"))) int f3d_debug_read_tc(void* host, size_t nbytes) { return (int)cudaMemcpyFromSymbol(host, g_tcprof, nbytes); }
extern "C" __attribute__((visibility("default"))) int f3d_debug_read_cert(void* host, size_t nbytes) { return (int)cudaMemcpyFromSymbol(host, g_certprof, nbytes); }
extern "C" __attribute__((visibility("default"))) int f3d_debug_read_clean(void* host, size_t nbytes) { return (int)cudaMemcpyFromSymbol(host, g_cleanprof, nbytes); }
#endif

// The tensor-core sweep is the default whenever both clouds have at least 512 points (measured against the CUDA-core sweep of
// chamfer.cu, step time cold: 1 x 512² 20.5 vs 24.5 us, 1 x 1024² 21 vs 31, 8 x 1000² 23 vs 31, 4 x 2048² 25 vs 31, 64 x 1024² 47 vs 48;
// below that the per-item set-up no longer pays: 4 x 300 x 517 33 vs 27, 1 x 2048 x 64 22.5 vs 16.5).
bool chamfer_tc_possible(int B, int N, int M) {
    if (B <= 0 || N < 1 || M < 1) return false;
    const TcPlan pl = make_tc_plan(B, N, M);
    if ((long long)B * (pl.rbA + pl.rbB) * kItemRows > 0x3fffffffLL) return false;   // slots are ints
    return std::max(pl.NpA, pl.NpB) <= 131072;   // running chunk numbers travel in <= 10 mantissa bits, global chunk ids as 16 bits
}
// host arrays through the sweep's own uploader warps: every 256-point tile of the raw clouds must be whole 16-byte units (TMA)
bool chamfer_tc_upload_possible(int N, int M) { return (N & 3) == 0 && (M & 3) == 0; }
bool chamfer_tc_supported(int B, int N, int M) {
    return chamfer_tc_possible(B, N, M) && std::min(N, M) >= 512;
}

int32_t chamfer_tc_launch(const float* A, const float* Bp, int32_t B, int32_t N, int32_t M, float w1, float w2, int32_t B_total,
                          float* loss_dev, float* terms_dev, int32_t* nnA_dev, int32_t* nnB_dev, void* ws, size_t ws_bytes, int32_t flags,
                          cudaStream_t stream, const ChamferUpload* upload, const ChamferPeerSum* peer) {
    const TcPlan pl = make_tc_plan(B, N, M);
    if (!ws || ws_bytes < pl.total) return fail(F3D_ERR_WORKSPACE, "f3d_chamfer_fwd: workspace %zu < required %zu bytes", ws_bytes, pl.total);
    unsigned char* w = static_cast<unsigned char*>(ws);
    int dev = 0;
    F3D_CUDA(cudaGetDevice(&dev));
    static unsigned char attr_done[256];
    static int sm_count[256];
    if (dev < 0 || dev >= 256 || !attr_done[dev]) {
        F3D_CUDA(cudaFuncSetAttribute(chamfer_tc_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
        F3D_CUDA(cudaFuncSetAttribute(chamfer_tc_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
        // the grids that run BEFORE / BESIDE the sweep must leave the SMs in the sweep's shared-memory configuration: an SM that
        // an upload CTA has configured for a large L1 cannot take a sweep CTA (130 KB of shared memory) until it drains
        F3D_CUDA(cudaFuncSetAttribute(chamfer_tc_sweep_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        F3D_CUDA(cudaFuncSetAttribute(chamfer_tc_sweep_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        F3D_CUDA(cudaFuncSetAttribute(chamfer_tc_prepare_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int sms = 0;
        F3D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 256) { sm_count[dev] = sms; attr_done[dev] = 1; }
    }
    const int sms = (dev >= 0 && dev < 256 && sm_count[dev] > 0) ? sm_count[dev] : 148;
    if (upload) F3D_CUDA(cudaMemsetAsync(w + pl.zero_from, 0, pl.zero_bytes, stream));   // (resident inputs: the prepare grid zeroes the header — no memset node in the step)

    TcPrepParams pp;
    pp.A = A; pp.Bp = Bp; pp.N = N; pp.M = M; pp.NpA = pl.NpA; pp.NpB = pl.NpB;
    pp.PA = reinterpret_cast<float4*>(w + pl.off_PA);
    pp.PB = reinterpret_cast<float4*>(w + pl.off_PB);
    pp.maxn = reinterpret_cast<unsigned*>(w + pl.off_maxn);
    pp.nblk = pl.nblk;
    pp.hdr = reinterpret_cast<int*>(w + pl.off_hdr);
    // Host arrays (upload): A / Bp are staging buffers; the sweep's own spare warps fill them over PCIe — no prepare grid, the
    // converters centre the raw points and take the norms on the fly
    if (upload && ((long long)B * N * 3 >= 0x7fffffffLL || (long long)B * M * 3 >= 0x7fffffffLL))
        return fail(F3D_ERR_INVALID, "chamfer_tc_launch: batch too large for the in-grid upload");
    if (!upload) {
        chamfer_tc_prepare_kernel<<<dim3(pl.nblk, B, 2), kPrepT, 0, stream>>>(pp);
        F3D_CHECK_LAUNCH("chamfer_tc_prepare_kernel");
    }

    TcFinParams fp;
    fp.A = A; fp.Bp = Bp; fp.PA = pp.PA; fp.PB = pp.PB;
    fp.B = B; fp.N = N; fp.M = M; fp.NpA = pl.NpA; fp.NpB = pl.NpB; fp.rbA = pl.rbA; fp.rbB = pl.rbB; fp.nstA = pl.nstA; fp.nstB = pl.nstB;
    fp.tilemin = reinterpret_cast<float*>(w + pl.off_tilemin); fp.maxn = pp.maxn; fp.nblk = pl.nblk;
    fp.nnA = nnA_dev; fp.nnB = nnB_dev;
    fp.partial = reinterpret_cast<double*>(w + pl.off_partial);
    fp.ambcnt = reinterpret_cast<int*>(w + pl.off_ambcnt);
    fp.ambq = reinterpret_cast<int*>(w + pl.off_ambq);
    fp.amblim = reinterpret_cast<float*>(w + pl.off_amblim);
    fp.amblist = reinterpret_cast<int*>(w + pl.off_amblist);
    fp.ambd = reinterpret_cast<float*>(w + pl.off_ambd);
    fp.nitems = pl.nitems;
    fp.hdr = reinterpret_cast<int*>(w + pl.off_hdr);
    fp.w1 = w1; fp.w2 = w2;
    fp.denomA = (double)N * (double)B_total;
    fp.denomB = (double)M * (double)B_total;
    fp.loss = loss_dev; fp.terms = terms_dev;
    if (peer) fp.peer = *peer;
    else { fp.peer.mailboxes = nullptr; fp.peer.nranks = 0; fp.peer.rank = 0; fp.peer.seq = 0; fp.peer.timeout_ns = 0; fp.peer.fault = nullptr; }

    TcSweepParams sp;
    sp.PA = pp.PA; sp.PB = pp.PB; sp.B = B; sp.NpA = pl.NpA; sp.NpB = pl.NpB; sp.rbA = pl.rbA; sp.rbB = pl.rbB;
    sp.tilemin = reinterpret_cast<float*>(w + pl.off_tilemin);
    sp.nstA = pl.nstA; sp.nstB = pl.nstB;
    sp.wait_prepare = upload ? 0 : 1;
    sp.arrived = upload ? reinterpret_cast<unsigned*>(w + pl.off_arrived) : nullptr;
    sp.up_groups = 2;   // (1, 2: 171 us per cfg2 call; 4: 176-185; 8: 181 — same box)
#ifdef F3D_TC_EXP_ENV
    if (const char* e = getenv("F3D_UP_GROUPS")) sp.up_groups = std::max(1, std::min(atoi(e), 16));   // development
#endif
    sp.hA = upload ? upload->A_host_dev : nullptr;
    sp.hB = upload ? upload->B_host_dev : nullptr;
    sp.f = fp;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(std::min(pl.nitems, sms)); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = kTcSmem; cfg.stream = stream;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (upload) F3D_CUDA(cudaLaunchKernelEx(&cfg, chamfer_tc_sweep_kernel<true>, sp));
        else F3D_CUDA(cudaLaunchKernelEx(&cfg, chamfer_tc_sweep_kernel<false>, sp));
    }
    F3D_CHECK_LAUNCH("chamfer_tc_sweep_kernel");
    if (flags & F3D_FLAG_SWEEP_ONLY) return F3D_OK;
    {
        // programmatic launch: the blocks are set up while the sweep runs and take the SMs as its CTAs leave (griddepcontrol.wait
        // holds them until the whole sweep grid has ended)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(std::min(pl.nitems, sms));   // one wave: the reducer block waits for the others
        cfg.blockDim = dim3(kCleanT); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
#ifdef F3D_TC_EXP_ENV
        if (getenv("F3D_TC_NOPDL")) attr[0].val.programmaticStreamSerializationAllowed = 0;  // development: cleanup launched strictly after the sweep
#endif
        cfg.attrs = attr; cfg.numAttrs = 1;
        F3D_CUDA(cudaLaunchKernelEx(&cfg, chamfer_tc_cleanup_kernel, fp));
    }
    F3D_CHECK_LAUNCH("chamfer_tc_cleanup_kernel");
    return F3D_OK;
}

}  // namespace f3d
