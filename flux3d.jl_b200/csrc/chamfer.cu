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
#include "f3d_common.cuh"

namespace f3d {
namespace {

constexpr int kRowsPerLane = 8;
constexpr int kTileRows = 32 * kRowsPerLane;  // 256
constexpr int kWarps = 4;
constexpr int kThreads = 32 * kWarps;
constexpr int kChunk = 32;         // columns per argmin-locator chunk
constexpr int kMaxColsPerWarp = 256;
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
    pl.cols_per_warp = cpw;
    pl.BN = kWarps * cpw;
    pl.CS = (M + pl.BN - 1) / pl.BN;
    pl.RB = (N + kTileRows - 1) / kTileRows;
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

}  // namespace
}  // namespace f3d

extern "C" size_t f3d_chamfer_workspace_bytes(int32_t B, int32_t N, int32_t M) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    return f3d::make_plan(B, N, M).total;
}

extern "C" int32_t f3d_chamfer_fwd(const float* A, const float* Bp, int32_t B, int32_t N, int32_t M,
                                   float w1, float w2, int32_t B_total, float* loss_dev,
                                   float* terms_dev, int32_t* nnA_dev, int32_t* nnB_dev, void* ws,
                                   size_t ws_bytes, int32_t flags, f3d_stream_t stream_) {
    using namespace f3d;
    if (!A || !Bp || !loss_dev) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: null A/B/loss pointer");
    if (B <= 0 || N <= 0 || M <= 0) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: B, N, M must be positive (got %d, %d, %d)", B, N, M);
    if (B > 65535) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: B must be <= 65535 per call");
    if (B_total == 0) B_total = B;
    if (B_total < B) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: B_total (%d) < B (%d)", B_total, B);
    Plan pl = make_plan(B, N, M);
    if (!ws || ws_bytes < pl.total) return fail(F3D_ERR_WORKSPACE, "f3d_chamfer_fwd: workspace %zu < required %zu bytes", ws_bytes, pl.total);
    if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0) return fail(F3D_ERR_MISALIGNED, "f3d_chamfer_fwd: workspace must be 256-byte aligned");
    if (pl.RB > 65535) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd: N too large");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    unsigned char* w = static_cast<unsigned char*>(ws);
    const bool fma = (flags & F3D_FLAG_FMA) != 0;

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
