// chamfer.cu — brute-force bidirectional nearest neighbour + Chamfer loss for sm_100a.
//
// Replaces Flux3D.jl src/metrics/pcloud.jl:39-52 (_chamfer_distance) and :72-86
// (_nearest_neighbors(::CuArray,::CuArray): cuBLAS batched GEMM + N×M×B matrix + two argmins),
// with the semantics of the CPU method :54-70 (exact 1-NN under the direct-difference Euclidean
// metric).  Nothing of size N×M is ever stored.
//
// Roofline: 0.006 algorithmic bytes per pair — the FP32 pipe binds, not HBM.  The design lever is
// FP32-pipe cycles and issue slots per pair:
//   * every pair distance is computed ONCE and feeds both directions (row min and column min);
//   * distances are evaluated two columns at a time with sm_100 packed FP32 (FADD2/FMUL2[/FFMA2]),
//     halving the issue slots of the subtract/multiply part; the additions stay scalar in the exact
//     mode because ptxas (12.9) contracts mul.rn.f32x2+add.rn.f32x2 into FFMA2 even with .rn;
//   * minima use 3-input FMNMX3; argmin INDICES are not tracked per pair at all: the sweep keeps only
//     the minimum value plus a coarse locator (32-column chunk id per row; 8-row lane ballot per
//     column) and a cheap finalize pass re-evaluates ≤32 / ≤8 candidates to recover the exact index
//     with the lowest-index tie rule;
//   * the column direction needs a cross-lane min per column: one REDUX per column per warp,
//     amortised over the 8 rows each lane holds in registers.
//
// Tile: 256 rows (8 per lane, lane owns rows 8*lane..8*lane+7 of the row block, held in registers
// by every warp of the CTA) × 4 warps × cols_per_warp columns staged in shared memory as packed
// column pairs {x0,x1,y0,y1},{z0,z1} so that one broadcast LDS.128 + LDS.64 feeds 16 pair distances
// per lane.
#include <algorithm>
#include <cstdlib>

#include "f3d_common.cuh"

namespace f3d {
namespace {

constexpr int kRowsPerLane = 8;
constexpr int kTileRows = 32 * kRowsPerLane;  // 256
constexpr int kWarps = 4;
constexpr int kThreads = 32 * kWarps;
#ifndef F3D_CHUNK
#define F3D_CHUNK 32
#endif
constexpr int kChunk = F3D_CHUNK;  // columns per argmin-locator chunk
#ifndef F3D_MAXCPW
#define F3D_MAXCPW 256
#endif
constexpr int kMaxColsPerWarp = F3D_MAXCPW;
// header of the counter region (ints): every word that many CTAs hit at once has its own 128-byte line
constexpr int kHdrStarted = 32, kHdrTimeout = 128, kHdrInts = 256;
constexpr float kPadA = 1.0e18f;   // padded rows / columns sit ~1e18 apart from everything:
constexpr float kPadB = -1.0e18f;  // d ≈ 1e37 (finite), never a minimum for in-contract inputs

struct SweepParams {
    const float* A;   // [B][N][3]
    const float* Bp;  // [B][M][3]
    int N, M;
    int cols_per_warp;  // multiple of kChunk
    int CS, RB;         // column splits, row blocks
    int Npad, Mpad;     // RB*kTileRows, CS*kWarps*cols_per_warp
    float* rp_min;      // [B][CS][Npad]
    int* rp_chunk;      // [B][CS][Npad]   global chunk id (column / kChunk)
    uint2* colpart;     // [B][RB][Mpad]   {float bits of min over the row block, lane ballot}
    unsigned* counter;  // zeroed here for the finalize kernel's last-block reduction
};

template <bool kFma>
__global__ void __launch_bounds__(kThreads, 4) chamfer_sweep_kernel(SweepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int BN = kWarps * p.cols_per_warp;
    float4* s_xy = reinterpret_cast<float4*>(smem_raw);               // [BN/2] {x0,x1,y0,y1}
    float2* s_z = reinterpret_cast<float2*>(s_xy + BN / 2);           // [BN/2] {z0,z1}
    uint4* s_col = reinterpret_cast<uint4*>(s_z + BN / 2);            // [BN/2] {min0,ballot0,min1,ballot1}
    float* s_rmin = reinterpret_cast<float*>(s_col + BN / 2);         // [kWarps][kTileRows]
    int* s_rchunk = reinterpret_cast<int*>(s_rmin + kWarps * kTileRows);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cs = blockIdx.x, rb = blockIdx.y, b = blockIdx.z;
    const int col0 = cs * BN;
    const int row0 = rb * kTileRows;

    if (tid == 0 && cs == 0 && rb == 0 && b == 0) *p.counter = 0u;

    // ---- stage the column tile: AoS global [M][3] → packed column pairs in shared memory --------
    {
        const float* gB = p.Bp + (size_t)b * p.M * 3;
        for (int pp = tid; pp < BN / 2; pp += kThreads) {
            int j = col0 + 2 * pp;
            float x0 = kPadB, y0 = kPadB, z0 = kPadB, x1 = kPadB, y1 = kPadB, z1 = kPadB;
            if (j < p.M) { x0 = __ldg(gB + 3 * j); y0 = __ldg(gB + 3 * j + 1); z0 = __ldg(gB + 3 * j + 2); }
            if (j + 1 < p.M) { x1 = __ldg(gB + 3 * j + 3); y1 = __ldg(gB + 3 * j + 4); z1 = __ldg(gB + 3 * j + 5); }
            s_xy[pp] = make_float4(x0, x1, y0, y1);
            s_z[pp] = make_float2(z0, z1);
        }
    }
    // ---- this lane's 8 rows → registers (every warp of the CTA holds the same 256 rows) ----------
    float ax[kRowsPerLane], ay[kRowsPerLane], az[kRowsPerLane];
    {
        const float* gA = p.A + (size_t)b * p.N * 3;
#pragma unroll
        for (int r = 0; r < kRowsPerLane; ++r) {
            int i = row0 + lane * kRowsPerLane + r;
            if (i < p.N) {
                ax[r] = __ldg(gA + 3 * i); ay[r] = __ldg(gA + 3 * i + 1); az[r] = __ldg(gA + 3 * i + 2);
            } else {
                ax[r] = kPadA; ay[r] = kPadA; az[r] = kPadA;
            }
        }
    }
    float amin[kRowsPerLane], aprev[kRowsPerLane];
    int abest[kRowsPerLane];
#pragma unroll
    for (int r = 0; r < kRowsPerLane; ++r) { amin[r] = INFINITY; aprev[r] = INFINITY; abest[r] = 0; }
    __syncthreads();

    // ---- sweep this warp's columns -----------------------------------------------------------------
    const int wp0 = warp * (p.cols_per_warp / 2);  // first column pair of this warp inside the tile
    const int nchunks = p.cols_per_warp / kChunk;
    const int gchunk0 = (col0 + warp * p.cols_per_warp) / kChunk;
    for (int ch = 0; ch < nchunks; ++ch) {
#pragma unroll 2
        for (int q = 0; q < kChunk / 2; ++q) {
            const int pp = wp0 + ch * (kChunk / 2) + q;
            const float4 xy = s_xy[pp];
            const float2 zz = s_z[pp];
            const u64 bx = pack2(xy.x, xy.y), by = pack2(xy.z, xy.w), bz = pack2(zz.x, zz.y);
            float c0 = INFINITY, c1 = INFINITY;
#pragma unroll
            for (int r = 0; r < kRowsPerLane; ++r) {
                const u64 dx = sub2(pack2(ax[r], ax[r]), bx);
                const u64 dy = sub2(pack2(ay[r], ay[r]), by);
                const u64 dz = sub2(pack2(az[r], az[r]), bz);
                float d0, d1;
                if (kFma) {
                    u64 s = mul2(dx, dx);
                    s = fma2(dy, dy, s);
                    s = fma2(dz, dz, s);
                    unpack2(s, d0, d1);
                } else {
                    float x0, x1, y0, y1, z0, z1;
                    unpack2(mul2(dx, dx), x0, x1);
                    unpack2(mul2(dy, dy), y0, y1);
                    unpack2(mul2(dz, dz), z0, z1);
                    d0 = __fadd_rn(__fadd_rn(x0, y0), z0);
                    d1 = __fadd_rn(__fadd_rn(x1, y1), z1);
                }
                amin[r] = fminf(amin[r], fminf(d0, d1));
                c0 = fminf(c0, d0);
                c1 = fminf(c1, d1);
            }
            // d >= 0, so the IEEE bit patterns order like unsigned integers: one REDUX per column.
            const unsigned u0 = __float_as_uint(c0), u1 = __float_as_uint(c1);
            const unsigned m0 = __reduce_min_sync(0xffffffffu, u0);
            const unsigned m1 = __reduce_min_sync(0xffffffffu, u1);
            const unsigned bal0 = __ballot_sync(0xffffffffu, u0 == m0);
            const unsigned bal1 = __ballot_sync(0xffffffffu, u1 == m1);
            if (lane == 0) s_col[pp] = make_uint4(m0, bal0, m1, bal1);
        }
        // chunk locator: strict '<' keeps the EARLIEST chunk that reached the running minimum
#pragma unroll
        for (int r = 0; r < kRowsPerLane; ++r) {
            if (amin[r] < aprev[r]) abest[r] = gchunk0 + ch;
            aprev[r] = amin[r];
        }
    }

    // ---- merge the 4 warps' row minima (ascending column order, strict '<'), store partials -------
#pragma unroll
    for (int r = 0; r < kRowsPerLane; ++r) {
        s_rmin[warp * kTileRows + lane * kRowsPerLane + r] = amin[r];
        s_rchunk[warp * kTileRows + lane * kRowsPerLane + r] = abest[r];
    }
    __syncthreads();
    {
        const size_t base = ((size_t)b * p.CS + cs) * p.Npad + row0;
        for (int rr = tid; rr < kTileRows; rr += kThreads) {
            float best = s_rmin[rr];
            int bc = s_rchunk[rr];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) {
                float v = s_rmin[w * kTileRows + rr];
                if (v < best) { best = v; bc = s_rchunk[w * kTileRows + rr]; }
            }
            p.rp_min[base + rr] = best;
            p.rp_chunk[base + rr] = bc;
        }
        uint2* gcol = p.colpart + ((size_t)b * p.RB + rb) * p.Mpad + col0;
        uint4* gcol4 = reinterpret_cast<uint4*>(gcol);
        for (int pp = tid; pp < BN / 2; pp += kThreads) gcol4[pp] = s_col[pp];
    }
}

struct FinalizeParams {
    const float* A;
    const float* Bp;
    int B, N, M, CS, RB, Npad, Mpad;
    const float* rp_min;
    const int* rp_chunk;
    const uint2* colpart;
    int32_t* nnA;  // may be null
    int32_t* nnB;  // may be null
    double* partial;  // [nbA + nbB]
    unsigned* counter;
    int nbA, nbB;
    float w1, w2;
    double denomA, denomB;  // N*B_total, M*B_total
    float* loss;            // [1]
    float* terms;           // [2] or null
};

constexpr int kFinThreads = 256;

template <bool kFma>
__global__ void __launch_bounds__(kFinThreads) chamfer_finalize_kernel(FinalizeParams p) {
    __shared__ double s_red[kFinThreads / 32];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    double mine = 0.0;
    if ((int)blockIdx.x < p.nbA) {
        // ---- rows: A → B --------------------------------------------------------------------------
        long t = (long)blockIdx.x * kFinThreads + tid;
        if (t < (long)p.B * p.N) {
            int b = (int)(t / p.N), i = (int)(t % p.N);
            float best = INFINITY;
            int chunk = 0;
            for (int cs = 0; cs < p.CS; ++cs) {
                size_t o = ((size_t)b * p.CS + cs) * p.Npad + i;
                float v = p.rp_min[o];
                if (v < best) { best = v; chunk = p.rp_chunk[o]; }
            }
            const float* a = p.A + ((size_t)b * p.N + i) * 3;
            const float ax = a[0], ay = a[1], az = a[2];
            const float* gB = p.Bp + (size_t)b * p.M * 3;
            int j0 = chunk * kChunk, j1 = min(j0 + kChunk, p.M);
            float dmin = INFINITY;
            int jmin = j0;
            for (int j = j0; j < j1; ++j) {
                float d = sqdist3<kFma>(ax, ay, az, gB[3 * j], gB[3 * j + 1], gB[3 * j + 2]);
                if (d < dmin) { dmin = d; jmin = j; }
            }
            if (p.nnA) p.nnA[t] = jmin;
            mine = (double)dmin;
        }
    } else {
        // ---- columns: B → A -----------------------------------------------------------------------
        long t = (long)(blockIdx.x - p.nbA) * kFinThreads + tid;
        if (t < (long)p.B * p.M) {
            int b = (int)(t / p.M), j = (int)(t % p.M);
            float best = INFINITY;
            int brb = 0;
            unsigned bal = 1u;
            for (int rb = 0; rb < p.RB; ++rb) {
                uint2 e = p.colpart[((size_t)b * p.RB + rb) * p.Mpad + j];
                float v = __uint_as_float(e.x);
                if (v < best) { best = v; brb = rb; bal = e.y; }
            }
            const float* q = p.Bp + ((size_t)b * p.M + j) * 3;
            const float bx = q[0], by = q[1], bz = q[2];
            const float* gA = p.A + (size_t)b * p.N * 3;
            int i0 = brb * kTileRows + (__ffs(bal) - 1) * kRowsPerLane, i1 = min(i0 + kRowsPerLane, p.N);
            float dmin = INFINITY;
            int imin = i0;
            for (int i = i0; i < i1; ++i) {
                float d = sqdist3<kFma>(gA[3 * i], gA[3 * i + 1], gA[3 * i + 2], bx, by, bz);
                if (d < dmin) { dmin = d; imin = i; }
            }
            if (p.nnB) p.nnB[t] = imin;
            mine = (double)dmin;
        }
    }
    // ---- block partial sum (fixed tree → run-to-run deterministic) -----------------------------------
    mine = warp_sum(mine);
    if ((tid & 31) == 0) s_red[tid >> 5] = mine;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kFinThreads / 32; ++w) s += s_red[w];
        p.partial[blockIdx.x] = s;
        __threadfence();
        unsigned done = atomicAdd(p.counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    // ---- last block: reduce the per-block partials in a fixed order, emit the loss ----------------------
    __threadfence();
    double sa = 0.0, sb = 0.0;
    for (int k = tid; k < p.nbA; k += kFinThreads) sa += __ldcg(p.partial + k);
    for (int k = tid; k < p.nbB; k += kFinThreads) sb += __ldcg(p.partial + p.nbA + k);
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    __shared__ double s_a[kFinThreads / 32], s_b[kFinThreads / 32];
    if ((tid & 31) == 0) { s_a[tid >> 5] = sa; s_b[tid >> 5] = sb; }
    __syncthreads();
    if (tid == 0) {
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < kFinThreads / 32; ++w) { ta += s_a[w]; tb += s_b[w]; }
        // dist_A_to_B = mean((A .- B[:,nn]).^2) * 3  ==  Σ_rows d_min / (N*B)      pcloud.jl:47-48
        float dAB = (float)(ta / p.denomA), dBA = (float)(tb / p.denomB);
        if (p.terms) { p.terms[0] = dAB; p.terms[1] = dBA; }
        p.loss[0] = __fadd_rn(__fmul_rn(p.w1, dAB), __fmul_rn(p.w2, dBA));  // pcloud.jl:50
    }
}

struct Plan {
    int cols_per_warp, BN, CS, RB, Npad, Mpad, nbA, nbB;
    size_t off_rp_min, off_rp_chunk, off_colpart, off_partial, off_counter, total;
};

Plan make_plan(int B, int N, int M) {
    Plan pl;
    int cpw = (int)align_up((size_t)(M + kWarps - 1) / kWarps, kChunk);
    if (cpw > kMaxColsPerWarp) cpw = kMaxColsPerWarp;
    pl.RB = (N + kTileRows - 1) / kTileRows;
    // small problems: narrower column tiles until the grid fills the machine once (4 CTAs x 148 SMs) — wide tiles are
    // ≈ 5 % faster per pair, but not if most SMs have nothing to do (B=2 N=M=1024: 27.7 → 20.8 µs at 64 columns per warp; 32 is too narrow: 26.8 µs)
    while (cpw > 2 * kChunk && cpw % (2 * kChunk) == 0 && (long)((M + kWarps * cpw - 1) / (kWarps * cpw)) * pl.RB * B < 592) cpw /= 2;
    pl.cols_per_warp = cpw;
    pl.BN = kWarps * cpw;
    pl.CS = (M + pl.BN - 1) / pl.BN;
    pl.Npad = pl.RB * kTileRows;
    pl.Mpad = pl.CS * pl.BN;
    pl.nbA = (int)(((long)B * N + kFinThreads - 1) / kFinThreads);
    pl.nbB = (int)(((long)B * M + kFinThreads - 1) / kFinThreads);
    size_t o = 0;
    pl.off_rp_min = o;   o = align_up(o + sizeof(float) * (size_t)B * pl.CS * pl.Npad, 256);
    pl.off_rp_chunk = o; o = align_up(o + sizeof(int) * (size_t)B * pl.CS * pl.Npad, 256);
    pl.off_colpart = o;  o = align_up(o + sizeof(uint2) * (size_t)B * pl.RB * pl.Mpad, 256);
    pl.off_partial = o;  o = align_up(o + sizeof(double) * (size_t)(pl.nbA + pl.nbB), 256);
    pl.off_counter = o;  o = align_up(o + sizeof(unsigned), 256);
    pl.total = o;
    return pl;
}

size_t sweep_smem_bytes(int BN) {
    return (size_t)(BN / 2) * (sizeof(float4) + sizeof(float2) + sizeof(uint4)) +
           (size_t)kWarps * kTileRows * (sizeof(float) + sizeof(int));
}


// =====================================================================================================
// Filtered sweep (default path).  Same result, bit for bit, as the exact sweep above — obtained with
// about half the FP32 work per pair:
//
//   1. FILTER.  Every pair gets an approximate squared distance in the expanded form on centred points
//        f_ij = ((nb_j - 2 a'x b'x) - 2 a'y b'y) - 2 a'z b'z + na_i        (3 FFMA2 + 1 FADD2 per TWO pairs)
//      a' = a - c, b' = b - c (c: a centre of the batch element), na = |a'|², nb = |b'|².  The exact form
//      needs 3 FADD2 + 3 FMUL2 + 4 FADD per two pairs.  |f_ij - d_ij| <= 15.03u (na_i + nb_j) + 5.001u d_ij, d_ij the
//      reference-arithmetic distance, u = 2^-24 (derivation in DESIGN.md §3.1: 11.02u from the filter's own
//      roundings incl. the norms, 4.01u from centring, 5.001u relative from the reference's roundings).
//   2. LOCATE.  Rows keep, per column split, (b1, c1, b2): the filter minimum, the 32-column chunk where
//      it was first reached, and the minimum over all OTHER chunks.  Columns keep, per 256-row block,
//      the REDUX minimum m and the ballot of lanes (8 rows each) whose minimum is within the window of m.
//   3. CERTIFY + RESCAN (finalize).  If every filter value outside the located chunk / lanes exceeds the
//      located minimum b1 by more than kWinAbs (na+nb) + kWinRel b1, the exact argmin (with the lowest-index tie rule) provably lies
//      inside: those <=32 / <=8 candidates are re-evaluated in the reference arithmetic.  Otherwise
//      (~0.2 % of rows on uniform clouds; every row of tie-heavy inputs) every tile whose filter minimum is
//      within the window is re-evaluated exactly.  Every reported index and distance therefore comes from the exact arithmetic.
// =====================================================================================================
// Certificate constants: |f - d| <= 15.03u (na+nb) + 5.001u d with u = 2^-24, hence "every value
// outside the located set exceeds the located minimum b1 by more than  kWinAbs (na+nb) + kWinRel b1" proves the
// exact argmin is inside it (30.1u = 1.794e-6 <= kWinAbs, 10.01u = 5.97e-7 <= kWinRel).
constexpr float kWinAbs = 2.0e-6f;
constexpr float kWinRel = 6.2e-7f;
// in-sweep lane ballot: the same window plus 15.1u (na+nb) of slack because REDUX on raw bit patterns may return
// any of several slightly negative values (all clamp to 0); applied as thr = m * kBallotRel + kBallotAbs (na+nb)
constexpr float kBallotAbs = 3.0e-6f;   // >= 45.2u = 2.69e-6
constexpr float kBallotRel = 1.0000007f;  // >= 1 + 10.01u
constexpr float kPadF = 3.0e38f;     // padded rows / columns: filter value ~3e38 (or +inf), never a minimum
constexpr float kNormLimit = 1.0e29f;  // above this the filter arithmetic could overflow: certify nothing

struct FiltParams {
    const float* A;
    const float* Bp;
    int N, M;
    int cols_per_warp, CS, RB, Npad, Mpad;
    float4* rowpart;   // [B][CS][Npad] {b1, b2, c1 (int bits), -}
    uint2* colpart;    // [B][RB][Mpad] {m (float bits, may be slightly negative), lane ballot}
    // prepared operands (chamfer_prepare_kernel; null in upload mode, where every tile stages its own operands):
    const float4* Ap;   // [B][Npad]    {a'x, a'y, a'z, |a'|²}, pads {0, 0, 0, kPadF}
    const float4* Bxy;  // [B][Mpad/2]  column pairs {-2b'x0, -2b'x1, -2b'y0, -2b'y1}: a tile is one contiguous bulk copy
    const float4* Bzn;  // [B][Mpad/2]  {-2b'z0, -2b'z1, |b'0|², |b'1|²}, pads {0, 0, kPadF, kPadF}
    float* maxna;      // [B][RB]  max |a'|² over the valid rows of the block
    float* maxnb;      // [B][CS]  max |b'|² over the valid columns of the tile
    float* centre;     // [B][4]   the centre used for this batch element (finalize recomputes |a'|², |b'|² with it)
    unsigned* counter;  // header of the zeroed region (kHdr*: finalize blocks done, CTAs started, upload timeout); one memset per call zeroes it with:
    int* rowdone;       // [B][RB]  tiles of the row block that have published their partials (target CS)
    int* coldone;       // [B][CS]  tiles of the column split that have published their partials (target RB)
    // host-array pipeline (chamfer_pipe.cu): the first up.U CTAs of the grid (by start ticket) do not sweep — they pull the
    // clouds out of page-locked HOST memory over PCIe (up.hA / up.hB, device-accessible) into A / Bp, batch element by
    // batch element, and count each element's arrival in up.arrived[b]; a sweeping CTA waits for its element.
    struct Upload {
        const float* hA;     // [B][N][3] host; null when the inputs are resident (U = 0)
        const float* hB;     // [B][M][3] host
        float* dA;           // == A, writable
        float* dB;           // == Bp, writable
        unsigned* arrived;   // [B]  uploader CTAs that have delivered their share of the element (target U)
        unsigned* timeout;   // set if a CTA gave up waiting (the loss is then reported as NaN, never a wrong number)
        int U;
    } up;
    int B;  // batch elements of this call (upload mode: how many elements the uploaders deliver)
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- upload mode: pull both clouds out of page-locked host memory, batch element by batch element -----------------
// U CTAs x 128 threads stream the element with 16-byte loads, four per thread in flight (≈ 256 KB outstanding over
// PCIe), into the staging copy the sweep reads; each CTA then counts the element as delivered (release).  A chunked
// cudaMemcpyAsync schedule does the same job with ~3.5 µs of HOST time per call (profiles/r01g: 12 chunk copies + 6
// flag writes = 131 µs against 67 µs for two plain copies) — here the host launches one grid and the copy paces itself.
__device__ __forceinline__ float4 ld_host_f4(const float* p) {
    float4 v;  // volatile: never served from a stale cache line of a previous step's bytes at the same host address
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_host_f1(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
// floats [s, e) of src -> dst (same index space, both bases 16-byte aligned): scalar head/tail, 16-byte body
__device__ __forceinline__ void upload_span(const float* __restrict__ src, float* __restrict__ dst, size_t s, size_t e, int u, int U, int tid) {
    size_t s4 = (s + 3) & ~(size_t)3, e4 = e & ~(size_t)3;
    if (s4 > e4) s4 = e4 = e;  // fewer than one aligned quad: everything is "head"
    if (u == 0) {
        if (s + tid < s4) dst[s + tid] = ld_host_f1(src + s + tid);    // < 4 floats each
        if (e4 + tid < e && e4 >= s4) dst[e4 + tid] = ld_host_f1(src + e4 + tid);
    }
    const size_t n4 = (e4 - s4) >> 2, stride = (size_t)U * kThreads;
    for (size_t i = (size_t)u * kThreads + tid; i < n4; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i + k * stride < n4) v[k] = ld_host_f4(src + s4 + 4 * (i + k * stride));
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i + k * stride < n4) __stcg(reinterpret_cast<float4*>(dst + s4) + i + k * stride, v[k]);
    }
}
__device__ void upload_clouds(const FiltParams& p, const int u, const int tid) {
    const size_t ea = (size_t)p.N * 3, eb = (size_t)p.M * 3;
    for (int b = 0; b < p.B; ++b) {
        upload_span(p.up.hA, p.up.dA, b * ea, (b + 1) * ea, u, p.up.U, tid);
        upload_span(p.up.hB, p.up.dB, b * eb, (b + 1) * eb, u, p.up.U, tid);
        __threadfence();   // this thread's stores are visible device-wide ...
        __syncthreads();   // ... for every thread of the CTA ...
        if (tid == 0) atomicAdd(p.up.arrived + b, 1u);  // ... before the element counts as delivered by this CTA
    }
}

// ---- TMA bulk copy (cp.async.bulk, completion on an mbarrier) -----------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// ---- prepare: centre every cloud, pre-scale, take the norms — ONCE per call instead of once per tile ----------------
// Without it each of the B*CS*RB tiles re-derives its 1024 columns and 256 rows from the raw points (54 scalar loads per
// thread, a shuffle reduction for the centre, the transform, two barriers): ≈ 25 % of a tile's life at cfg2, repeated RB = 16
// times per column tile.  Here one thread per point writes the operands in exactly the layout the sweep keeps in shared
// memory / registers, so a tile's prologue is two TMA bulk copies (8 KB each) and eight 16-byte loads per lane.
// grid (blocks of 256 points, B, 2): z = 0 rows of A, z = 1 columns of B.  Same arithmetic, bit for bit, as the in-tile
// staging of upload mode and as the finalize's recomputation of |q'|².
#ifndef F3D_SWEEP_CTAS_PER_SM
#define F3D_SWEEP_CTAS_PER_SM 0
#endif
constexpr int kSweepCtasPerSm = F3D_SWEEP_CTAS_PER_SM;  // > 0: persistent sweep grid of this many CTAs per SM (experiment); 0: one tile per CTA
constexpr int kPrepThreads = 256;
#ifndef F3D_PREP_MIN_RB
#define F3D_PREP_MIN_RB 32
#endif
constexpr int kPrepMinRowBlocks = F3D_PREP_MIN_RB;  // use the prepare grid from this many 256-row blocks per cloud on
__global__ void __launch_bounds__(kPrepThreads) chamfer_prepare_kernel(FiltParams p, float4* __restrict__ Ap, float4* __restrict__ Bxy,
                                                                       float4* __restrict__ Bzn) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the sweep's launch overlaps this grid; it waits before reading
    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.y;
    const bool isA = blockIdx.z == 0;
    const int BN = kWarps * p.cols_per_warp;
    if (isA ? (int)blockIdx.x * kPrepThreads >= p.Npad : (int)blockIdx.x * kPrepThreads >= p.Mpad) return;
    const float* gA = p.A + (size_t)b * p.N * 3;
    const float* gB = p.Bp + (size_t)b * p.M * 3;
    // the centre of the batch element: the mean of 32 + 32 strided sample points (any point works: the certificate
    // uses the norms actually obtained); every warp computes it with the same operations => same bits everywhere
    const int ia = (int)(((long)lane * p.N) >> 5), ib = (int)(((long)lane * p.M) >> 5);
    float cx = __ldg(gA + 3 * ia) + __ldg(gB + 3 * ib);
    float cy = __ldg(gA + 3 * ia + 1) + __ldg(gB + 3 * ib + 1);
    float cz = __ldg(gA + 3 * ia + 2) + __ldg(gB + 3 * ib + 2);
    const int i = blockIdx.x * kPrepThreads + tid;  // point index (row of A / column of B)
    const int n = isA ? p.N : p.M;
    const float* src = (isA ? gA : gB) + 3 * (size_t)min(i, n - 1);
    const float rx = __ldg(src), ry = __ldg(src + 1), rz = __ldg(src + 2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cx += __shfl_xor_sync(0xffffffffu, cx, o);
        cy += __shfl_xor_sync(0xffffffffu, cy, o);
        cz += __shfl_xor_sync(0xffffffffu, cz, o);
    }
    cx *= (1.0f / 64.0f); cy *= (1.0f / 64.0f); cz *= (1.0f / 64.0f);
    if (isA && blockIdx.x == 0 && tid == 0) { p.centre[4 * b] = cx; p.centre[4 * b + 1] = cy; p.centre[4 * b + 2] = cz; }
    float x = 0.f, y = 0.f, z = 0.f, nrm = kPadF, mx = 0.f;
    if (i < n) {
        x = rx - cx; y = ry - cy; z = rz - cz;
        nrm = fmaf(z, z, fmaf(y, y, x * x));
        mx = nrm;
    }
    mx = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(mx)));  // norms are >= 0
    if (isA) {
        if (i < p.Npad) Ap[(size_t)b * p.Npad + i] = make_float4(x, y, z, nrm);
        // 32 consecutive rows lie in one 256-row block
        if (lane == 0 && i < p.N) atomicMax(reinterpret_cast<unsigned*>(p.maxna) + (size_t)b * p.RB + i / kTileRows, __float_as_uint(mx));
    } else {
        // columns travel in pairs: the even lane writes {-2x0, -2x1, -2y0, -2y1}, the odd lane {-2z0, -2z1, n0, n1}
        const float ox = __shfl_xor_sync(0xffffffffu, x, 1), oy = __shfl_xor_sync(0xffffffffu, y, 1);
        const float oz = __shfl_xor_sync(0xffffffffu, z, 1), on = __shfl_xor_sync(0xffffffffu, nrm, 1);
        if (i < p.Mpad) {
            const size_t pp = ((size_t)b * p.Mpad + i) >> 1;
            if ((lane & 1) == 0) Bxy[pp] = make_float4(-2.0f * x, -2.0f * ox, -2.0f * y, -2.0f * oy);
            else Bzn[pp] = make_float4(-2.0f * oz, -2.0f * z, on, nrm);
        }
        // 32 consecutive columns lie in one column tile (BN is a multiple of 128)
        if (lane == 0 && i < p.M) atomicMax(reinterpret_cast<unsigned*>(p.maxnb) + (size_t)b * p.CS + i / BN, __float_as_uint(mx));
    }
}

#ifndef F3D_FILT_MINB
#define F3D_FILT_MINB 4
#endif
#ifndef F3D_FILT_UNROLL
#define F3D_FILT_UNROLL 2
#endif
// kLoop = false: one tile per CTA on a (CS, RB, B) grid — the default for resident inputs; kLoop = true: 1-D grid, tile loop
// c, c + G, ... with roles by start ticket (upload mode; persistent-grid experiments)
template <bool kLoop>
__global__ void __launch_bounds__(kThreads, F3D_FILT_MINB) chamfer_filter_sweep_kernel(FiltParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int BN = kWarps * p.cols_per_warp;
    float4* s_xy = reinterpret_cast<float4*>(smem_raw);          // [BN/2] {-2x0,-2x1,-2y0,-2y1}
    float4* s_zn = s_xy + BN / 2;                                // [BN/2] {-2z0,-2z1,nb0,nb1}
    float4* s_row = s_zn + BN / 2;                               // [kWarps][kTileRows] {b1,b2,c1,-}; first used as s_rm
    __shared__ unsigned s_maxnb;
    __shared__ int s_ticket;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Programmatic dependent launch: let the finalize grid become resident as soon as every CTA of this grid has
    // started; its blocks wait on rowdone/coldone, so they soak up the SM slots the last partial wave leaves idle.
    // A CTA sweeps tiles c, c + G, c + 2G, ... (G = grid size).  By default G = number of tiles — one tile per CTA and
    // the hardware block scheduler keeps every SM slot busy; a persistent grid (G = 3 or 4 CTAs per SM, the finalize
    // resident beside it from the start) measured 13–40 % slower (profiles/r01g §4): its CTAs march through prologue,
    // loop and epilogue in step, and nothing fills the FMA pipe meanwhile.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int first = blockIdx.x, stride = gridDim.x;
    if (kLoop && p.up.U) {
        // Upload mode (U extra CTAs).  Roles go by START order, not by blockIdx: the first U CTAs to start are the
        // uploaders, so every CTA that waits for data below started after the CTAs that deliver it — forward progress
        // needs no assumption about the dispatch order.
        if (tid == 0) s_ticket = (int)atomicAdd(p.counter + kHdrStarted, 1u);
        __syncthreads();
        const int ticket = s_ticket;
        if (ticket < p.up.U) {
            upload_clouds(p, ticket, tid);
            return;
        }
        first = ticket - p.up.U;
        stride = gridDim.x - p.up.U;
    }
    const int ntiles = p.CS * p.RB * p.B;
    __shared__ __align__(8) unsigned long long s_bar;
    const unsigned bar = smem_u32(&s_bar);
    if (p.Ap) {
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        asm volatile("griddepcontrol.wait;" ::: "memory");  // the prepare grid has completed and its writes are visible
    }
    unsigned bar_phase = 0;
    int arrived_b = -1;  // upload mode: batch elements up to this one have landed
  for (int tile = first; kLoop ? tile < ntiles : tile == first; tile += kLoop ? stride : 1) {
    const int cs = kLoop ? tile % p.CS : (int)blockIdx.x, rb = kLoop ? (tile / p.CS) % p.RB : (int)blockIdx.y,
              b = kLoop ? tile / (p.CS * p.RB) : (int)blockIdx.z;
    if (tid == 0) s_maxnb = 0u;
    if (kLoop && p.up.U && b > arrived_b) {
        // wait until this batch element has landed; the bounded wait (2 s) turns a lost upload into a reported error
        // instead of a hung device
        if (tid == 0) {
            const int* f = reinterpret_cast<const int*>(p.up.arrived) + b;
            unsigned long long t0 = 0;
            for (unsigned spins = 0; ld_acquire(f) < p.up.U; ++spins) {
                __nanosleep(100);
                if ((spins & 1023u) == 1023u) {
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > 2000000000ull) { atomicExch(p.up.timeout, 1u); break; }
                }
            }
        }
        __syncthreads();
        arrived_b = b;
    }
    const int col0 = cs * BN, row0 = rb * kTileRows;
    const float* gA = p.A + (size_t)b * p.N * 3;
    const float* gB = p.Bp + (size_t)b * p.M * 3;

    float ax[kRowsPerLane], ay[kRowsPerLane], az[kRowsPerLane], na[kRowsPerLane];
    float maxna, maxnb;
    if (p.Ap) {
        // ---- prologue, prepared operands: the column tile arrives by TMA (two bulk copies onto one mbarrier), this
        // lane's 8 rows by eight 16-byte loads — one memory latency, no arithmetic, no barrier besides the mbarrier ----
        if (tid == 0) {
            const unsigned bytes = (unsigned)(BN / 2) * (unsigned)sizeof(float4);
            mbar_expect_tx(bar, 2 * bytes);
            tma_bulk_g2s(smem_u32(s_xy), p.Bxy + (((size_t)b * p.Mpad + col0) >> 1), bytes, bar);
            tma_bulk_g2s(smem_u32(s_zn), p.Bzn + (((size_t)b * p.Mpad + col0) >> 1), bytes, bar);
        }
        const float4* gAp = p.Ap + (size_t)b * p.Npad + row0 + lane * kRowsPerLane;
#pragma unroll
        for (int r = 0; r < kRowsPerLane; ++r) {
            const float4 v = __ldg(gAp + r);
            ax[r] = v.x; ay[r] = v.y; az[r] = v.z; na[r] = v.w;
        }
        maxna = __ldg(p.maxna + (size_t)b * p.RB + rb);
        maxnb = __ldg(p.maxnb + (size_t)b * p.CS + cs);
        while (!mbar_try_wait(bar, bar_phase)) {}
        bar_phase ^= 1u;
    } else {
        // ---- prologue: issue EVERY global load first (centre samples, this thread's column pairs, this lane's rows),
        // so the CTA pays one memory latency instead of one per dependent step -------------------------------------
        constexpr int kMaxPairsPerThread = kMaxColsPerWarp * kWarps / 2 / kThreads;  // 4
        const int ia = (int)(((long)lane * p.N) >> 5), ib = (int)(((long)lane * p.M) >> 5);
        const float sx = __ldg(gA + 3 * ia) + __ldg(gB + 3 * ib);
        const float sy = __ldg(gA + 3 * ia + 1) + __ldg(gB + 3 * ib + 1);
        const float sz = __ldg(gA + 3 * ia + 2) + __ldg(gB + 3 * ib + 2);
        float rawc[kMaxPairsPerThread][6];
#pragma unroll
        for (int k = 0; k < kMaxPairsPerThread; ++k) {
            const int pp = tid + k * kThreads;
            const int j = col0 + 2 * pp;
#pragma unroll
            for (int e = 0; e < 6; ++e) rawc[k][e] = (pp < BN / 2 && j + e / 3 < p.M) ? __ldg(gB + 3 * (size_t)j + e) : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < kRowsPerLane; ++r) {
            const int i = row0 + lane * kRowsPerLane + r;
            const bool ok = i < p.N;
            ax[r] = ok ? __ldg(gA + 3 * (size_t)i) : 0.0f;
            ay[r] = ok ? __ldg(gA + 3 * (size_t)i + 1) : 0.0f;
            az[r] = ok ? __ldg(gA + 3 * (size_t)i + 2) : 0.0f;
        }
        // the centre of the batch element: every warp computes it redundantly (same operations => same bits)
        float cx = sx, cy = sy, cz = sz;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cx += __shfl_xor_sync(0xffffffffu, cx, o);
            cy += __shfl_xor_sync(0xffffffffu, cy, o);
            cz += __shfl_xor_sync(0xffffffffu, cz, o);
        }
        cx *= (1.0f / 64.0f); cy *= (1.0f / 64.0f); cz *= (1.0f / 64.0f);
        // every tile publishes the (identical) centre and its norm maxima: a finalize block may only rely on the tiles
        // of ITS row block / column split having finished
        if (tid == 0) { p.centre[4 * b] = cx; p.centre[4 * b + 1] = cy; p.centre[4 * b + 2] = cz; }
        __syncthreads();  // s_maxnb = 0 visible

        // ---- stage the column tile: centred, pre-scaled by -2, with |b'|² -----------------------------------
        {
            float mynb = 0.0f;
#pragma unroll
            for (int k = 0; k < kMaxPairsPerThread; ++k) {
                const int pp = tid + k * kThreads;
                if (pp < BN / 2) {
                    const int j = col0 + 2 * pp;
                    float x0 = 0.f, y0 = 0.f, z0 = 0.f, n0 = kPadF, x1 = 0.f, y1 = 0.f, z1 = 0.f, n1 = kPadF;
                    if (j < p.M) {
                        x0 = rawc[k][0] - cx; y0 = rawc[k][1] - cy; z0 = rawc[k][2] - cz;
                        n0 = fmaf(z0, z0, fmaf(y0, y0, x0 * x0));
                        mynb = fmaxf(mynb, n0);
                    }
                    if (j + 1 < p.M) {
                        x1 = rawc[k][3] - cx; y1 = rawc[k][4] - cy; z1 = rawc[k][5] - cz;
                        n1 = fmaf(z1, z1, fmaf(y1, y1, x1 * x1));
                        mynb = fmaxf(mynb, n1);
                    }
                    s_xy[pp] = make_float4(-2.0f * x0, -2.0f * x1, -2.0f * y0, -2.0f * y1);
                    s_zn[pp] = make_float4(-2.0f * z0, -2.0f * z1, n0, n1);
                }
            }
            mynb = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(mynb)));  // norms are >= 0
            if (lane == 0) atomicMax(&s_maxnb, __float_as_uint(mynb));
        }
        // ---- this lane's 8 rows: centred coordinates and |a'|² ------------------------------------------------
        float myna = 0.0f;
#pragma unroll
        for (int r = 0; r < kRowsPerLane; ++r) {
            const int i = row0 + lane * kRowsPerLane + r;
            if (i < p.N) {
                ax[r] -= cx; ay[r] -= cy; az[r] -= cz;
                na[r] = fmaf(az[r], az[r], fmaf(ay[r], ay[r], ax[r] * ax[r]));
                myna = fmaxf(myna, na[r]);
            } else {
                ax[r] = 0.f; ay[r] = 0.f; az[r] = 0.f; na[r] = kPadF;
            }
        }
        maxna = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(myna)));
        __syncthreads();
        maxnb = __uint_as_float(s_maxnb);
        if (tid == 0) {
            p.maxnb[(size_t)b * p.CS + cs] = maxnb;
            p.maxna[(size_t)b * p.RB + rb] = maxna;
        }
    }
    const float wt = kBallotAbs * (maxna + maxnb);

    const int wp0 = warp * (p.cols_per_warp / 2);
    const int nchunks = p.cols_per_warp / kChunk;
    const int gchunk0 = (col0 + warp * p.cols_per_warp) / kChunk;
    // column partials {m0, ballot0, m1, ballot1} go straight to global memory (lane 0, 16 B per column pair): a
    // shared-memory store here would order against the next pair's LDS and stop the compiler from overlapping
    // one pair's min/REDUX tail with the next pair's FFMA2s.
    uint4* __restrict__ gcol4 = reinterpret_cast<uint4*>(p.colpart + ((size_t)b * p.RB + rb) * p.Mpad + col0);
    // per-chunk row minima are parked in shared memory ([warp][chunk][r][lane], conflict-free) and turned into
    // (b1, c1, b2) after the sweep: keeping that state out of the loop's registers lets ptxas emit the FFMA2s
    // stage by stage with operand reuse (each FFMA2 then reads <= 2 registers per bank; the row-by-row order it
    // picks under register pressure costs 3 cycles per FFMA2 instead of 2 — profiles/r01b notes).
    // (the LAST chunk's minima never leave the registers — they are still live when the loop ends — so nchunks - 1 chunks
    // are parked: 28 KB instead of 32 KB at 256 columns per warp, which is what lets a fifth CTA fit the SM's 228 KB)
    float* s_rm = reinterpret_cast<float*>(s_row) + (size_t)warp * (nchunks - 1) * kTileRows;
    float rm[kRowsPerLane];
    for (int ch = 0; ch < nchunks; ++ch) {
#pragma unroll
        for (int r = 0; r < kRowsPerLane; ++r) rm[r] = INFINITY;
#pragma unroll 2
        for (int q = 0; q < kChunk / 2; ++q) {
            const int pp = wp0 + ch * (kChunk / 2) + q;
            const float4 xy = s_xy[pp];
            const float4 zn = s_zn[pp];
            const u64 X2 = pack2(xy.x, xy.y), Y2 = pack2(xy.z, xy.w), Z2 = pack2(zn.x, zn.y), NB = pack2(zn.z, zn.w);
            u64 t[kRowsPerLane];
#pragma unroll
            for (int r = 0; r < kRowsPerLane; ++r) t[r] = fma2(X2, pack2(ax[r], ax[r]), NB);
#pragma unroll
            for (int r = 0; r < kRowsPerLane; ++r) t[r] = fma2(Y2, pack2(ay[r], ay[r]), t[r]);
#pragma unroll
            for (int r = 0; r < kRowsPerLane; ++r) t[r] = fma2(Z2, pack2(az[r], az[r]), t[r]);
            float c0 = INFINITY, c1v = INFINITY;
#pragma unroll
            for (int r = 0; r < kRowsPerLane; ++r) {
                float f0, f1;
                unpack2(add2(t[r], pack2(na[r], na[r])), f0, f1);
                rm[r] = fminf(rm[r], fminf(f0, f1));
                c0 = fminf(c0, f0);
                c1v = fminf(c1v, f1);
            }
            const int m0 = __reduce_min_sync(0xffffffffu, __float_as_int(c0));
            const int m1 = __reduce_min_sync(0xffffffffu, __float_as_int(c1v));
            const unsigned bal0 = __ballot_sync(0xffffffffu, c0 <= fmaf(__int_as_float(m0), kBallotRel, wt));
            const unsigned bal1 = __ballot_sync(0xffffffffu, c1v <= fmaf(__int_as_float(m1), kBallotRel, wt));
            if (lane == 0) gcol4[pp] = make_uint4((unsigned)m0, bal0, (unsigned)m1, bal1);
        }
        if (ch + 1 < nchunks) {
#pragma unroll
            for (int r = 0; r < kRowsPerLane; ++r) s_rm[(ch * kRowsPerLane + r) * 32 + lane] = rm[r];
        }
    }

    // ---- (b1, c1, b2) per row over this warp's chunks: minimum, the EARLIEST chunk reaching it, and the minimum
    // over the other chunks (each lane reads back only what it stored) ----------------------------------------
    float b1[kRowsPerLane], b2[kRowsPerLane];
    int c1[kRowsPerLane];
#pragma unroll
    for (int r = 0; r < kRowsPerLane; ++r) { b1[r] = INFINITY; b2[r] = INFINITY; c1[r] = 0; }
    for (int ch = 0; ch < nchunks; ++ch) {
#pragma unroll
        for (int r = 0; r < kRowsPerLane; ++r) {
            const float cm = ch + 1 < nchunks ? s_rm[(ch * kRowsPerLane + r) * 32 + lane] : rm[r];
            b2[r] = fminf(b2[r], fmaxf(b1[r], cm));
            if (cm < b1[r]) c1[r] = gchunk0 + ch;
            b1[r] = fminf(b1[r], cm);
        }
    }
    __syncthreads();  // s_row below aliases the s_rm regions of all warps
    // ---- merge the 4 warps' row triples (ascending column order), store the partials ----------------------
#pragma unroll
    for (int r = 0; r < kRowsPerLane; ++r)
        s_row[warp * kTileRows + lane * kRowsPerLane + r] = make_float4(b1[r], b2[r], __int_as_float(c1[r]), 0.f);
    __syncthreads();
    {
        const size_t base = ((size_t)b * p.CS + cs) * p.Npad + row0;
        for (int rr = tid; rr < kTileRows; rr += kThreads) {
            float4 best = s_row[rr];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) {
                const float4 o = s_row[w * kTileRows + rr];
                if (o.x < best.x) { best.y = fminf(best.x, o.y); best.x = o.x; best.z = o.z; }
                else best.y = fminf(best.y, o.x);
            }
            p.rowpart[base + rr] = best;
        }
    }
    // publish: all of this tile's partials are in global memory before the counters move (release)
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        atomicAdd(p.rowdone + (size_t)b * p.RB + rb, 1);
        atomicAdd(p.coldone + (size_t)b * p.CS + cs, 1);
    }
    // (the barrier above also separates this tile's last use of the shared tile / s_row from the next tile's staging)
  }
}

// ---- finalize for the filtered sweep: certify, rescan exactly, reduce the loss ------------------------------
struct FiltFinalizeParams {
    const float* A;
    const float* Bp;
    int B, N, M, CS, RB, Npad, Mpad, BN;
    const float4* rowpart;
    const uint2* colpart;
    const float* maxna;
    const float* maxnb;
    const float* centre;
    const int* rowdone;
    const int* coldone;
    int32_t* nnA;
    int32_t* nnB;
    double* partial;
    unsigned* counter;
    int nbA, nbB;
    float w1, w2;
    double denomA, denomB;
    float* loss;
    float* terms;
    const unsigned* upload_timeout;  // null unless the sweep ran in upload mode
    ChamferPeerSum peer;             // nranks > 1: sum the loss over the ranks through peer memory (NVLink), see below
};


// exact (reference-arithmetic) argmin of |q - P[j]|² over j in [j0, j1), cooperatively by the whole block: thread t takes
// candidates j0 + 4t .. j0 + 4t + 3 (+ 4·kFinThreads per further trip) and merges them into ITS running (dmin, jmin); a
// thread's candidates ascend, so '<' keeps the lowest index.  Four points are 48 bytes: three 16-byte loads when the
// cloud is 16-byte aligned there (N, M multiples of 4), twelve scalar loads otherwise — all issued before any arithmetic.
template <bool kQIsRow>
__device__ __forceinline__ void block_exact_scan(const float* __restrict__ P, int j0, int j1, float qx, float qy, float qz,
                                                 int tid, float& dmin, int& jmin) {
    for (int jb = j0 + 4 * tid; jb < j1; jb += 4 * kFinThreads) {
        const float* pp = P + 3 * (size_t)jb;
        float c[12];
        if (jb + 4 <= j1 && (reinterpret_cast<uintptr_t>(pp) & 15u) == 0) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(pp));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(pp) + 1);
            const float4 v2 = __ldg(reinterpret_cast<const float4*>(pp) + 2);
            c[0] = v0.x; c[1] = v0.y; c[2] = v0.z; c[3] = v0.w; c[4] = v1.x; c[5] = v1.y;
            c[6] = v1.z; c[7] = v1.w; c[8] = v2.x; c[9] = v2.y; c[10] = v2.z; c[11] = v2.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int jj = min(jb + k, j1 - 1);  // clamped duplicates are harmless (same value, same index)
                c[3 * k] = __ldg(P + 3 * (size_t)jj); c[3 * k + 1] = __ldg(P + 3 * (size_t)jj + 1); c[3 * k + 2] = __ldg(P + 3 * (size_t)jj + 2);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // operand order of the reference: (a - b) with a from the first cloud
            const float d = kQIsRow ? sqdist3<false>(qx, qy, qz, c[3 * k], c[3 * k + 1], c[3 * k + 2])
                                    : sqdist3<false>(c[3 * k], c[3 * k + 1], c[3 * k + 2], qx, qy, qz);
            if (d < dmin) { dmin = d; jmin = min(jb + k, j1 - 1); }
        }
    }
}

// Block = 8 warps; a warp owns 32 consecutive items (rows of A, or columns = points of B).
//   phase 1  lane <-> item: merge the sweep's partials, decide "certified" vs "ambiguous"
//   phase 2  lane <-> item: re-evaluate the located candidates exactly (16-byte loads, batched)
//   phase 3  the (rare) ambiguous items, one at a time, whole block: exact scan of every tile within the window
#ifndef F3D_FIN_MINB
#define F3D_FIN_MINB 5
#endif
__global__ void __launch_bounds__(kFinThreads, F3D_FIN_MINB) chamfer_filter_finalize_kernel(FiltFinalizeParams p) {
    __shared__ double s_red[kFinThreads / 32];
    __shared__ bool s_last;
    constexpr int kMaxSel = 32;  // phase 3: tiles within an ambiguous item's window that are scanned as one candidate range
    __shared__ int s_sel[kMaxSel], s_nsel;
    __shared__ unsigned s_amb[kFinThreads / 32], s_wd[kFinThreads / 32];  // phase 3: ambiguous-item queue, block argmin
    __shared__ int s_wj[kFinThreads / 32], s_ib[kFinThreads], s_iq[kFinThreads];
    __shared__ float s_ilim[kFinThreads];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Blocks are numbered in the order the persistent sweep completes their inputs: batch element by batch element,
    // its row blocks (fbA = ceil(N/256) of them), then its column blocks (fbB = ceil(M/256)).
    const int fbA = (p.N + kFinThreads - 1) / kFinThreads, fbB = (p.M + kFinThreads - 1) / kFinThreads;
    const int fe = (int)blockIdx.x / (fbA + fbB), fj = (int)blockIdx.x - fe * (fbA + fbB);
    const bool rows = fj < fbA;
    const int Q = rows ? p.N : p.M;        // items per batch element (queries)
    const long t = (long)fe * Q + (long)(rows ? fj : fj - fbA) * kFinThreads + tid;  // flat item index b*Q + q
    const int R = rows ? p.M : p.N;        // points searched per query
    const float* gQ = rows ? p.A : p.Bp;   // queries
    const float* gP = rows ? p.Bp : p.A;   // searched cloud
    const bool valid = (rows ? fj : fj - fbA) * kFinThreads + tid < Q;
    double mine = 0.0;

    // ---- phase 1: merge the sweep's partials ------------------------------------------------------------------------
    // Everything here and in phase 2 is a chain of L2 round trips (≈ 0.4 µs each on B200), so the loads of a step are
    // issued together: one dependent-issue load per loop trip made this kernel 45 µs; batches make it a handful of trips.
    int b = 0, q = 0, loc = 0;     // loc: chunk id (rows) or row block (columns)
    unsigned bal = 0u;             // columns: lane ballot of the located block
    float best = 0.0f, win = 0.0f; // clamped filter minimum and the certificate window
    float qx = 0.f, qy = 0.f, qz = 0.f;  // the query point (raw coordinates)
    bool amb = false;
    if (valid) { b = (int)(t / Q); q = (int)(t - (long)b * Q); }
    {
        // wait until every sweep tile of this item's row block / column split has published (this grid is launched
        // programmatically and may be resident while the sweep's last wave is still running).  The lanes of a warp
        // almost always share one flag: one lane per distinct flag polls it.
        const int* flag = !valid ? nullptr : (rows ? p.rowdone + (size_t)b * p.RB + q / kTileRows : p.coldone + (size_t)b * p.CS + q / p.BN);
        const int target = rows ? p.CS : p.RB;
        const unsigned peers = __match_any_sync(0xffffffffu, reinterpret_cast<unsigned long long>(flag));
        if (flag && lane == __ffs(peers) - 1)
            while (ld_acquire(flag) < target) __nanosleep(100);
        __syncwarp();
    }
    if (valid) {
        const float* c = p.centre + 4 * b;
        const float* pt = gQ + ((size_t)b * Q + q) * 3;
        const float cx = __ldcg(c), cy = __ldcg(c + 1), cz = __ldcg(c + 2);
        qx = __ldg(pt); qy = __ldg(pt + 1); qz = __ldg(pt + 2);
        float second = INFINITY, other = 0.0f;
        best = INFINITY;
        if (rows) {
            for (int cs0 = 0; cs0 < p.CS; cs0 += 4) {
                float4 e[4];
                float mo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int cs = min(cs0 + k, p.CS - 1);
                    e[k] = __ldcg(p.rowpart + ((size_t)b * p.CS + cs) * p.Npad + q);
                    mo[k] = __ldcg(p.maxnb + (size_t)b * p.CS + cs);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (cs0 + k < p.CS) {
                        const float e1 = fmaxf(e[k].x, 0.0f), e2 = fmaxf(e[k].y, 0.0f);
                        other = fmaxf(other, mo[k]);
                        if (e1 < best) { second = fminf(best, e2); best = e1; loc = __float_as_int(e[k].z); }
                        else second = fminf(second, e1);
                    }
                }
            }
        } else {
            for (int rb0 = 0; rb0 < p.RB; rb0 += 8) {
                uint2 e[8];
                float mo[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int rb = min(rb0 + k, p.RB - 1);
                    e[k] = __ldcg(p.colpart + ((size_t)b * p.RB + rb) * p.Mpad + q);
                    mo[k] = __ldcg(p.maxna + (size_t)b * p.RB + rb);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (rb0 + k < p.RB) {
                        const float v = fmaxf(__uint_as_float(e[k].x), 0.0f);
                        other = fmaxf(other, mo[k]);
                        if (v < best) { second = best; best = v; loc = rb0 + k; bal = e[k].y; }
                        else second = fminf(second, v);
                    }
                }
            }
        }
        const float x = qx - cx, y = qy - cy, z = qz - cz;
        const float nq = fmaf(z, z, fmaf(y, y, x * x));  // the sweep's |a'|² / |b'|², bit for bit
        win = fmaf(kWinRel, best, kWinAbs * (nq + other));
        // written so that NaN / inf / out-of-range norms can only make the item ambiguous, never certified
        amb = !(nq <= kNormLimit && other <= kNormLimit && second > best + win);
    }
    // ---- phase 2: certified items — each lane re-evaluates its own item's located candidates exactly -------------------
    // rows: the 32 columns of the located chunk (384 contiguous bytes); columns: the 8 rows of every balloted lane (96
    // contiguous bytes each, almost always one).  16-byte loads when the cloud is 16-byte aligned there (N, M % 4 == 0).
    if (valid && !amb) {
        const float* P = gP + (size_t)b * R * 3;
        float d = INFINITY;
        int j = 0x7fffffff;
        if (rows) {
            const int j0 = loc * kChunk, j1 = min(j0 + kChunk, R);
            const float* pp = P + 3 * (size_t)j0;
            if ((kChunk % 8) == 0 && j0 + kChunk <= R && (reinterpret_cast<uintptr_t>(pp) & 15u) == 0) {
#pragma unroll
                for (int h = 0; h < kChunk / 8; ++h) {  // 8 candidates = six 16-byte loads per trip
                    float4 v[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(pp) + h * 6 + i);
                    const float c[24] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w, v[2].x, v[2].y, v[2].z, v[2].w,
                                         v[3].x, v[3].y, v[3].z, v[3].w, v[4].x, v[4].y, v[4].z, v[4].w, v[5].x, v[5].y, v[5].z, v[5].w};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float dd = sqdist3<false>(qx, qy, qz, c[3 * k], c[3 * k + 1], c[3 * k + 2]);
                        if (dd < d) { d = dd; j = j0 + h * 8 + k; }  // ascending candidates: '<' keeps the lowest index
                    }
                }
            } else {
#pragma unroll 4
                for (int jj = j0; jj < j1; ++jj) {
                    const float dd = sqdist3<false>(qx, qy, qz, __ldg(P + 3 * (size_t)jj), __ldg(P + 3 * (size_t)jj + 1), __ldg(P + 3 * (size_t)jj + 2));
                    if (dd < d) { d = dd; j = jj; }
                }
            }
        } else {
            for (unsigned bits = bal; bits; bits &= bits - 1) {  // ascending lanes => ascending row indices
                const int i0 = loc * kTileRows + (__ffs(bits) - 1) * kRowsPerLane, i1 = min(i0 + kRowsPerLane, R);
                const float* pp = P + 3 * (size_t)i0;
                if (kRowsPerLane == 8 && i0 + 8 <= R && (reinterpret_cast<uintptr_t>(pp) & 15u) == 0) {
                    float4 v[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(pp) + i);
                    const float c[24] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w, v[2].x, v[2].y, v[2].z, v[2].w,
                                         v[3].x, v[3].y, v[3].z, v[3].w, v[4].x, v[4].y, v[4].z, v[4].w, v[5].x, v[5].y, v[5].z, v[5].w};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        // operand order of the reference: (a - b) with a from the first cloud
                        const float dd = sqdist3<false>(c[3 * k], c[3 * k + 1], c[3 * k + 2], qx, qy, qz);
                        if (dd < d) { d = dd; j = i0 + k; }
                    }
                } else {
                    for (int ii = i0; ii < i1; ++ii) {
                        const float dd = sqdist3<false>(__ldg(P + 3 * (size_t)ii), __ldg(P + 3 * (size_t)ii + 1), __ldg(P + 3 * (size_t)ii + 2), qx, qy, qz);
                        if (dd < d) { d = dd; j = ii; }
                    }
                }
            }
        }
        int32_t* nn = rows ? p.nnA : p.nnB;
        if (nn) nn[t] = j;
        mine += (double)d;
    }
    // ---- phase 3: ambiguous items, one at a time, by the WHOLE BLOCK -------------------------------------------
    // An ambiguous item needs an exact scan of every tile within its window (up to BN = 1024 candidates each).  Done by
    // the owning warp alone that is a chain of dependent L2 round trips, and the few warps that own two or three such
    // items set the kernel's tail; spread over the block every thread takes 4 candidates of a tile, all loads of a
    // tile are in flight at once, and an item costs about one memory latency.  The order of the queue is fixed (warp,
    // lane), each owner adds its own item's distance: the block's partial sum stays run-to-run deterministic.
    {
        const unsigned ambmask = __ballot_sync(0xffffffffu, valid && amb);
        if (lane == 0) s_amb[warp] = ambmask;
        if (valid && amb) { s_ib[tid] = b; s_iq[tid] = q; s_ilim[tid] = best + win; }
        __syncthreads();
        for (int w = 0; w < kFinThreads / 32; ++w) {
            for (unsigned rem = s_amb[w]; rem; rem &= rem - 1) {
                const int src = w * 32 + __ffs(rem) - 1;
                const int bb = s_ib[src], qq = s_iq[src];
                const float lim = s_ilim[src];  // NaN/inf → scan everything (comparison below)
                const float* qp = gQ + ((size_t)bb * Q + qq) * 3;
                const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
                const float* P = gP + (size_t)bb * R * 3;
                float d = INFINITY;
                int j = 0x7fffffff;
                // (a) which tiles lie within the window: one thread per tile, so ONE round trip instead of one per tile
                const int ntile = rows ? p.CS : p.RB, tw = rows ? p.BN : kTileRows;
                if (tid == 0) s_nsel = 0;
                __syncthreads();
                for (int t0 = 0; t0 < ntile; t0 += kFinThreads) {
                    const int t = t0 + tid;
                    if (t < ntile) {
                        const float e1 = rows ? __ldcg(&p.rowpart[((size_t)bb * p.CS + t) * p.Npad + qq].x)
                                              : __uint_as_float(__ldcg(&p.colpart[((size_t)bb * p.RB + t) * p.Mpad + qq].x));
                        if (!(fmaxf(e1, 0.0f) > lim)) {
                            const int pos = atomicAdd(&s_nsel, 1);
                            if (pos < kMaxSel) s_sel[pos] = t;
                        }
                    }
                }
                __syncthreads();
                const int nsel = s_nsel;
                if (nsel <= kMaxSel) {
                    // (b) the selected tiles as ONE candidate range: thread t takes candidates 4t .. 4t+3 (+ 1024 per trip), so
                    // the loads of all tiles are in flight together.  The order of s_sel is arbitrary: full (d, j) comparison.
                    const int total = nsel * tw;
                    for (int idx = 4 * tid; idx < total; idx += 4 * kFinThreads) {
                        const int sidx = idx / tw, jb = s_sel[sidx] * tw + (idx - sidx * tw);  // tw is a multiple of 4
                        if (jb >= R) continue;
                        const float* pp = P + 3 * (size_t)jb;
                        float c[12];
                        if (jb + 4 <= R && (reinterpret_cast<uintptr_t>(pp) & 15u) == 0) {
                            const float4 v0 = __ldg(reinterpret_cast<const float4*>(pp));
                            const float4 v1 = __ldg(reinterpret_cast<const float4*>(pp) + 1);
                            const float4 v2 = __ldg(reinterpret_cast<const float4*>(pp) + 2);
                            c[0] = v0.x; c[1] = v0.y; c[2] = v0.z; c[3] = v0.w; c[4] = v1.x; c[5] = v1.y;
                            c[6] = v1.z; c[7] = v1.w; c[8] = v2.x; c[9] = v2.y; c[10] = v2.z; c[11] = v2.w;
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int jj = min(jb + k, R - 1);  // clamped duplicates are harmless (same value, same index)
                                c[3 * k] = __ldg(P + 3 * (size_t)jj); c[3 * k + 1] = __ldg(P + 3 * (size_t)jj + 1); c[3 * k + 2] = __ldg(P + 3 * (size_t)jj + 2);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // operand order of the reference: (a - b) with a from the first cloud
                            const float dd = rows ? sqdist3<false>(qx, qy, qz, c[3 * k], c[3 * k + 1], c[3 * k + 2])
                                                  : sqdist3<false>(c[3 * k], c[3 * k + 1], c[3 * k + 2], qx, qy, qz);
                            const int jj = min(jb + k, R - 1);
                            if (dd < d || (dd == d && jj < j)) { d = dd; j = jj; }
                        }
                    }
                } else if (rows) {  // tie-heavy input: more tiles within the window than the list holds — scan them in order
                    for (int cs = 0; cs < p.CS; ++cs) {
                        const float e1 = fmaxf(__ldcg(&p.rowpart[((size_t)bb * p.CS + cs) * p.Npad + qq].x), 0.0f);
                        if (!(e1 > lim)) block_exact_scan<true>(P, cs * p.BN, min((cs + 1) * p.BN, R), qx, qy, qz, tid, d, j);
                    }
                } else {
                    for (int rb = 0; rb < p.RB; ++rb) {
                        const float v = fmaxf(__uint_as_float(__ldcg(&p.colpart[((size_t)bb * p.RB + rb) * p.Mpad + qq].x)), 0.0f);
                        if (!(v > lim)) block_exact_scan<false>(P, rb * kTileRows, min((rb + 1) * kTileRows, R), qx, qy, qz, tid, d, j);
                    }
                }
                // (d, j) minimum over the block, lowest j on ties (d >= 0 or +inf: bit order == value order)
                const unsigned mb = __reduce_min_sync(0xffffffffu, __float_as_uint(d));
                const int jm = __reduce_min_sync(0xffffffffu, (__float_as_uint(d) == mb) ? j : 0x7fffffff);
                if (lane == 0) { s_wd[warp] = mb; s_wj[warp] = jm; }
                __syncthreads();
                if (tid == src) {
                    unsigned bd = s_wd[0];
                    int bj = s_wj[0];
#pragma unroll
                    for (int k = 1; k < kFinThreads / 32; ++k) {
                        const unsigned od = s_wd[k];
                        const int oj = s_wj[k];
                        if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
                    }
                    int32_t* nn = rows ? p.nnA : p.nnB;
                    if (nn) nn[(size_t)bb * Q + qq] = bj;
                    mine += (double)__uint_as_float(bd);
                }
                __syncthreads();  // s_wd / s_wj are reused by the next item
            }
        }
    }
    // ---- block partial sum → last block reduces in a fixed order (deterministic) ------------------------------
    mine = warp_sum(mine);
    if (lane == 0) s_red[warp] = mine;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kFinThreads / 32; ++w) s += s_red[w];
        p.partial[blockIdx.x] = s;
        __threadfence();
        s_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double sa = 0.0, sb = 0.0;
    for (int k = tid; k < p.nbA + p.nbB; k += kFinThreads) {
        const double v = __ldcg(p.partial + k);
        if (k % (fbA + fbB) < fbA) sa += v; else sb += v;
    }
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    __shared__ double s_a[kFinThreads / 32], s_b[kFinThreads / 32];
    if (lane == 0) { s_a[warp] = sa; s_b[warp] = sb; }
    __syncthreads();
    if (tid == 0) {
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < kFinThreads / 32; ++w) { ta += s_a[w]; tb += s_b[w]; }
        const float dAB = (float)(ta / p.denomA), dBA = (float)(tb / p.denomB);  // pcloud.jl:47-48
        if (p.terms) { p.terms[0] = dAB; p.terms[1] = dBA; }
        float l = __fadd_rn(__fmul_rn(p.w1, dAB), __fmul_rn(p.w2, dBA));          // pcloud.jl:50
        if (p.upload_timeout && __ldcg(p.upload_timeout) != 0u) l = __int_as_float(0x7fc00000);  // an upload never landed
        if (p.peer.nranks <= 1) p.loss[0] = l;
        s_a[0] = (double)l;
    }
    if (p.peer.nranks <= 1) return;

    // ---- the one exchange of the sharded path, fused into this kernel (peer mailboxes over NVLink: f3d_common.cuh) ----
    __shared__ float s_v[kMaxPeerRanks];
    __syncthreads();
    if (tid < p.peer.nranks) s_v[tid] = peer_exchange(p.peer, tid, (float)s_a[0]);
    __syncthreads();
    if (tid == 0) {
        float sum = 0.0f;
        for (int r = 0; r < p.peer.nranks; ++r) sum = __fadd_rn(sum, s_v[r]);
        p.loss[0] = sum;
    }
}

struct FiltPlan {
    int cols_per_warp, BN, CS, RB, Npad, Mpad, nbA, nbB;
    size_t off_rowpart, off_colpart, off_Ap, off_Bxy, off_Bzn, off_maxna, off_maxnb, off_centre, off_partial, off_counter, counter_bytes,
        zero_from, zero_bytes, total;
};

FiltPlan make_filt_plan(int B, int N, int M) {
    FiltPlan pl;
    int cpw = (int)align_up((size_t)(M + kWarps - 1) / kWarps, kChunk);
    if (cpw > kMaxColsPerWarp) cpw = kMaxColsPerWarp;
    pl.cols_per_warp = cpw;
    pl.BN = kWarps * cpw;
    pl.CS = (M + pl.BN - 1) / pl.BN;
    pl.RB = (N + kTileRows - 1) / kTileRows;
    pl.Npad = pl.RB * kTileRows;
    pl.Mpad = pl.CS * pl.BN;
    pl.nbA = B * ((N + kFinThreads - 1) / kFinThreads);  // finalize blocks: 256 rows of ONE batch element ...
    pl.nbB = B * ((M + kFinThreads - 1) / kFinThreads);  // ... or 256 of its columns
    size_t o = 0;
    pl.off_rowpart = o; o = align_up(o + sizeof(float4) * (size_t)B * pl.CS * pl.Npad, 256);
    pl.off_colpart = o; o = align_up(o + sizeof(uint2) * (size_t)B * pl.RB * pl.Mpad, 256);
    pl.off_Ap = o;      o = align_up(o + sizeof(float4) * (size_t)B * pl.Npad, 256);
    pl.off_Bxy = o;     o = align_up(o + sizeof(float4) * (size_t)B * (pl.Mpad / 2), 256);
    pl.off_Bzn = o;     o = align_up(o + sizeof(float4) * (size_t)B * (pl.Mpad / 2), 256);
    pl.zero_from = o;   // everything from here on is zeroed by ONE memset per call: norm maxima (atomicMax), counters, flags
    pl.off_maxna = o;   o = align_up(o + sizeof(float) * (size_t)B * pl.RB, 256);
    pl.off_maxnb = o;   o = align_up(o + sizeof(float) * (size_t)B * pl.CS, 256);
    pl.off_centre = o;  o = align_up(o + sizeof(float) * 4 * (size_t)B, 256);
    pl.off_partial = o; o = align_up(o + sizeof(double) * (size_t)(pl.nbA + pl.nbB), 256);
    // header (kHdr*: finalize blocks done, CTAs started (upload mode), upload timeout flag) |
    // rowdone [B][RB] | coldone [B][CS] | arrived [B] (upload mode)  — one memset zeroes all of it per call
    pl.counter_bytes = sizeof(int) * kHdrInts + sizeof(int) * ((size_t)B * pl.RB + (size_t)B * pl.CS + (size_t)B);
    pl.off_counter = o; o = align_up(o + pl.counter_bytes, 256);
    pl.zero_bytes = o - pl.zero_from;
    pl.total = o;
    return pl;
}

size_t filt_smem_bytes(int BN) {
    const size_t rm_bytes = (size_t)(BN / kChunk - kWarps) * kTileRows * sizeof(float);  // [kWarps][chunks per warp - 1][8][32]
    return (size_t)(BN / 2) * (2 * sizeof(float4)) + std::max(rm_bytes, (size_t)kWarps * kTileRows * sizeof(float4));
}

}  // namespace
}  // namespace f3d


extern "C" size_t f3d_chamfer_workspace_bytes(int32_t B, int32_t N, int32_t M) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    return std::max(std::max(f3d::make_plan(B, N, M).total, f3d::make_filt_plan(B, N, M).total), f3d::chamfer_tc_workspace_bytes(B, N, M));
}

extern "C" int32_t f3d_chamfer_fwd(const float* A, const float* Bp, int32_t B, int32_t N, int32_t M,
                                   float w1, float w2, int32_t B_total, float* loss_dev,
                                   float* terms_dev, int32_t* nnA_dev, int32_t* nnB_dev, void* ws,
                                   size_t ws_bytes, int32_t flags, f3d_stream_t stream_) {
    return f3d::chamfer_fwd_launch(A, Bp, B, N, M, w1, w2, B_total, loss_dev, terms_dev, nnA_dev, nnB_dev, ws, ws_bytes, flags,
                                   static_cast<cudaStream_t>(stream_), nullptr, nullptr);
}

int32_t f3d::chamfer_fwd_launch(const float* A, const float* Bp, int32_t B, int32_t N, int32_t M, float w1, float w2,
                                int32_t B_total, float* loss_dev, float* terms_dev, int32_t* nnA_dev, int32_t* nnB_dev,
                                void* ws, size_t ws_bytes, int32_t flags, cudaStream_t stream, const ChamferUpload* upload,
                                const ChamferPeerSum* peer) {
    using namespace f3d;
    if (!A || !Bp || !loss_dev) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: null A/B/loss pointer");
    if (B <= 0 || N <= 0 || M <= 0) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: B, N, M must be positive (got %d, %d, %d)", B, N, M);
    if (B > 65535) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: B must be <= 65535 per call");
    if (B_total == 0) B_total = B;
    if (B_total < B) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: B_total (%d) < B (%d)", B_total, B);
    const size_t need = f3d_chamfer_workspace_bytes(B, N, M);
    if (!ws || ws_bytes < need) return fail(F3D_ERR_WORKSPACE, "f3d_chamfer_fwd: workspace %zu < required %zu bytes", ws_bytes, need);
    if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0) return fail(F3D_ERR_MISALIGNED, "f3d_chamfer_fwd: workspace must be 256-byte aligned");
    unsigned char* w = static_cast<unsigned char*>(ws);
    const bool fma = (flags & F3D_FLAG_FMA) != 0;
    if ((upload || peer) && (fma || (flags & (F3D_FLAG_EXACT_SWEEP | F3D_FLAG_SWEEP_ONLY))))
        return fail(F3D_ERR_INVALID, "chamfer_fwd_launch: the in-grid upload / peer sum exist only for the default (filtered) sweep");
    if (peer && (peer->nranks < 1 || peer->nranks > kMaxPeerRanks || peer->rank < 0 || peer->rank >= peer->nranks))
        return fail(F3D_ERR_INVALID, "chamfer_fwd_launch: bad peer rank %d of %d", peer->rank, peer->nranks);

    // ---- default for problems that fill the machine: the filter sweep on the tensor cores (chamfer_tc.cu) --------------
    // (F3D_FLAG_TENSOR forces it for any shape: tests drive small and ragged shapes through it that way)
    if (!fma && !(flags & (F3D_FLAG_EXACT_SWEEP | F3D_FLAG_CUDA_CORES)) && (!upload || chamfer_tc_upload_possible(N, M)) &&
        ((flags & F3D_FLAG_TENSOR) ? chamfer_tc_possible(B, N, M) : chamfer_tc_supported(B, N, M)))
        return chamfer_tc_launch(A, Bp, B, N, M, w1, w2, B_total, loss_dev, terms_dev, nnA_dev, nnB_dev, ws, ws_bytes, flags, stream, upload, peer);

    if (!fma && !(flags & F3D_FLAG_EXACT_SWEEP)) {
        // ---- CUDA-core filtered sweep + certified exact finalize (bit-identical results, ~half the FP32 work) ----
        FiltPlan fl = make_filt_plan(B, N, M);
        if (fl.RB > 65535) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: N too large");
        FiltParams sp;
        sp.A = A; sp.Bp = Bp; sp.N = N; sp.M = M;
        sp.cols_per_warp = fl.cols_per_warp; sp.CS = fl.CS; sp.RB = fl.RB; sp.Npad = fl.Npad; sp.Mpad = fl.Mpad;
        sp.rowpart = reinterpret_cast<float4*>(w + fl.off_rowpart);
        sp.colpart = reinterpret_cast<uint2*>(w + fl.off_colpart);
        sp.maxna = reinterpret_cast<float*>(w + fl.off_maxna);
        sp.maxnb = reinterpret_cast<float*>(w + fl.off_maxnb);
        sp.centre = reinterpret_cast<float*>(w + fl.off_centre);
        sp.counter = reinterpret_cast<unsigned*>(w + fl.off_counter);
        sp.rowdone = reinterpret_cast<int*>(w + fl.off_counter) + kHdrInts;
        sp.coldone = sp.rowdone + (size_t)B * fl.RB;
        F3D_CUDA(cudaMemsetAsync(w + fl.zero_from, 0, fl.zero_bytes, stream));
        sp.Ap = nullptr; sp.Bxy = nullptr; sp.Bzn = nullptr;
        sp.up.hA = nullptr; sp.up.hB = nullptr; sp.up.dA = nullptr; sp.up.dB = nullptr; sp.up.U = 0;
        sp.up.arrived = reinterpret_cast<unsigned*>(sp.coldone + (size_t)B * fl.CS);
        sp.up.timeout = sp.counter + kHdrTimeout;
        if (upload) {
            if (upload->uploaders < 1 || upload->uploaders > 1024) return fail(F3D_ERR_INVALID, "chamfer_fwd_launch: bad uploader count %d", upload->uploaders);
            sp.up.hA = upload->A_host_dev; sp.up.hB = upload->B_host_dev;
            sp.up.dA = const_cast<float*>(A); sp.up.dB = const_cast<float*>(Bp);
            sp.up.U = upload->uploaders;
        }
        const size_t smem = filt_smem_bytes(fl.BN);
        int dev_id = 0;
        {
            // opt in to the largest tile's shared memory once per device (idempotent; racing threads at worst set it twice)
            static unsigned char attr_done[256];
            int dev = 0;
            F3D_CUDA(cudaGetDevice(&dev));
            dev_id = dev;
            if (dev < 0 || dev >= 256 || !attr_done[dev]) {
                F3D_CUDA(cudaFuncSetAttribute(chamfer_filter_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)filt_smem_bytes(kWarps * kMaxColsPerWarp)));
                F3D_CUDA(cudaFuncSetAttribute(chamfer_filter_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)filt_smem_bytes(kWarps * kMaxColsPerWarp)));
                if (dev >= 0 && dev < 256) attr_done[dev] = 1;
            }
        }
        sp.B = B;
        // grid: one CTA per tile (kSweepCtasPerSm > 0: a persistent grid instead, each CTA sweeping tiles c, c + G, ...)
        static int sm_count[256];
        if (kSweepCtasPerSm > 0 && dev_id >= 0 && dev_id < 256 && sm_count[dev_id] == 0)
            F3D_CUDA(cudaDeviceGetAttribute(&sm_count[dev_id], cudaDevAttrMultiProcessorCount, dev_id));
        const int sms = (dev_id >= 0 && dev_id < 256 && sm_count[dev_id] > 0) ? sm_count[dev_id] : 148;
        const long long ntiles = (long long)fl.CS * fl.RB * B;
        if (ntiles > 0x7fffffffLL - 2048) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: too many tiles (%lld)", ntiles);
        unsigned G = (unsigned)(kSweepCtasPerSm > 0 ? std::min<long long>(ntiles, (long long)kSweepCtasPerSm * sms) : ntiles);
        const bool loop = G != (unsigned)ntiles;  // a persistent grid (experiments); else one tile per CTA, 3-D grid
        if (fl.RB > 65535 && !loop && !sp.up.U) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: N too large");
        if (sp.up.U) {
            // upload mode: the operands do not exist yet — every tile stages its own once its batch element has landed
            chamfer_filter_sweep_kernel<true><<<dim3(G + sp.up.U), kThreads, smem, stream>>>(sp);
        } else if (fl.RB < kPrepMinRowBlocks) {
            // few row blocks: re-deriving a column tile RB times costs less than one more grid in front of the sweep
            // (cfg2, RB = 16: 158.7 µs against 161.1 µs with the prepare grid — profiles/r01g)
            if (loop) chamfer_filter_sweep_kernel<true><<<dim3(G), kThreads, smem, stream>>>(sp);
            else chamfer_filter_sweep_kernel<false><<<dim3(fl.CS, fl.RB, B), kThreads, smem, stream>>>(sp);
        } else {
            // many row blocks: prepare the operands once, then the sweep — launched programmatically, it waits
            // (griddepcontrol.wait) only where it first reads them (B=16 N=M=10000: 430 → 419 µs)
            float4* Ap = reinterpret_cast<float4*>(w + fl.off_Ap);
            float4* Bxy = reinterpret_cast<float4*>(w + fl.off_Bxy);
            float4* Bzn = reinterpret_cast<float4*>(w + fl.off_Bzn);
            chamfer_prepare_kernel<<<dim3((std::max(fl.Npad, fl.Mpad) + kPrepThreads - 1) / kPrepThreads, B, 2), kPrepThreads, 0, stream>>>(sp, Ap, Bxy, Bzn);
            F3D_CHECK_LAUNCH("chamfer_prepare_kernel");
            sp.Ap = Ap; sp.Bxy = Bxy; sp.Bzn = Bzn;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = loop ? dim3(G) : dim3(fl.CS, fl.RB, B); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            if (loop) F3D_CUDA(cudaLaunchKernelEx(&cfg, chamfer_filter_sweep_kernel<true>, sp));
            else F3D_CUDA(cudaLaunchKernelEx(&cfg, chamfer_filter_sweep_kernel<false>, sp));
        }
        F3D_CHECK_LAUNCH("chamfer_filter_sweep_kernel");
        if (flags & F3D_FLAG_SWEEP_ONLY) return F3D_OK;
        FiltFinalizeParams fp;
        fp.A = A; fp.Bp = Bp; fp.B = B; fp.N = N; fp.M = M;
        fp.CS = fl.CS; fp.RB = fl.RB; fp.Npad = fl.Npad; fp.Mpad = fl.Mpad; fp.BN = fl.BN;
        fp.rowpart = sp.rowpart; fp.colpart = sp.colpart; fp.maxna = sp.maxna; fp.maxnb = sp.maxnb; fp.centre = sp.centre; fp.rowdone = sp.rowdone; fp.coldone = sp.coldone;
        fp.nnA = nnA_dev; fp.nnB = nnB_dev;
        fp.partial = reinterpret_cast<double*>(w + fl.off_partial);
        fp.counter = sp.counter;
        fp.nbA = fl.nbA; fp.nbB = fl.nbB;
        fp.w1 = w1; fp.w2 = w2;
        fp.denomA = (double)N * (double)B_total;
        fp.denomB = (double)M * (double)B_total;
        fp.loss = loss_dev; fp.terms = terms_dev;
        fp.upload_timeout = upload ? sp.up.timeout : nullptr;
        if (peer) fp.peer = *peer;
        else { fp.peer.mailboxes = nullptr; fp.peer.nranks = 0; fp.peer.rank = 0; fp.peer.seq = 0; fp.peer.timeout_ns = 0; fp.peer.fault = nullptr; }
        {
            // programmatic dependent launch: blocks may start while the sweep's last wave is still running
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(fl.nbA + fl.nbB); cfg.blockDim = dim3(kFinThreads); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            F3D_CUDA(cudaLaunchKernelEx(&cfg, chamfer_filter_finalize_kernel, fp));
        }
        F3D_CHECK_LAUNCH("chamfer_filter_finalize_kernel");
        return F3D_OK;
    }

    Plan pl = make_plan(B, N, M);
    if (pl.RB > 65535) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: N too large");

    SweepParams sp;
    sp.A = A; sp.Bp = Bp; sp.N = N; sp.M = M;
    sp.cols_per_warp = pl.cols_per_warp; sp.CS = pl.CS; sp.RB = pl.RB; sp.Npad = pl.Npad; sp.Mpad = pl.Mpad;
    sp.rp_min = reinterpret_cast<float*>(w + pl.off_rp_min);
    sp.rp_chunk = reinterpret_cast<int*>(w + pl.off_rp_chunk);
    sp.colpart = reinterpret_cast<uint2*>(w + pl.off_colpart);
    sp.counter = reinterpret_cast<unsigned*>(w + pl.off_counter);
    size_t smem = sweep_smem_bytes(pl.BN);
    dim3 grid(pl.CS, pl.RB, B);
    if (fma) {
        F3D_CUDA(cudaFuncSetAttribute(chamfer_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chamfer_sweep_kernel<true><<<grid, kThreads, smem, stream>>>(sp);
    } else {
        F3D_CUDA(cudaFuncSetAttribute(chamfer_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chamfer_sweep_kernel<false><<<grid, kThreads, smem, stream>>>(sp);
    }
    F3D_CHECK_LAUNCH("chamfer_sweep_kernel");
    if (flags & F3D_FLAG_SWEEP_ONLY) return F3D_OK;

    FinalizeParams fp;
    fp.A = A; fp.Bp = Bp; fp.B = B; fp.N = N; fp.M = M;
    fp.CS = pl.CS; fp.RB = pl.RB; fp.Npad = pl.Npad; fp.Mpad = pl.Mpad;
    fp.rp_min = sp.rp_min; fp.rp_chunk = sp.rp_chunk; fp.colpart = sp.colpart;
    fp.nnA = nnA_dev; fp.nnB = nnB_dev;
    fp.partial = reinterpret_cast<double*>(w + pl.off_partial);
    fp.counter = sp.counter;
    fp.nbA = pl.nbA; fp.nbB = pl.nbB;
    fp.w1 = w1; fp.w2 = w2;
    fp.denomA = (double)N * (double)B_total;
    fp.denomB = (double)M * (double)B_total;
    fp.loss = loss_dev; fp.terms = terms_dev;
    if (fma) chamfer_finalize_kernel<true><<<pl.nbA + pl.nbB, kFinThreads, 0, stream>>>(fp);
    else chamfer_finalize_kernel<false><<<pl.nbA + pl.nbB, kFinThreads, 0, stream>>>(fp);
    F3D_CHECK_LAUNCH("chamfer_finalize_kernel");
    return F3D_OK;
}
