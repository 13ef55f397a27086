// chamfer_bwd.cu — pullback of the Chamfer loss for sm_100a.
//
// Replaces the Zygote pullback of Flux3D.jl src/metrics/pcloud.jl:47-50 (gather B[:,nn] -> broadcast
// subtract/square -> mean; the indices are constants, :45 is @ignore):
//   gA_i = cA (A_i - B_nnA(i))  -  Σ_{j : nnB(j) = i} cB (B_j - A_i),   cA = 2 w1 g / (N B_total)
//   gB_j = cB (B_j - A_nnB(j))  -  Σ_{i : nnA(i) = j} cA (A_i - B_j),   cB = 2 w2 g / (M B_total)
// HBM-bound (reads 2 clouds + 2 index arrays, writes 2 gradients: 32 B per point).  Two launches:
// the direct terms overwrite the outputs (no memset needed), then the scatter terms are added with
// RED.ADD.F32.  The scatter order is not fixed, so the last bit of a gradient entry that receives
// several contributions may vary run to run (the reference pins this to atol 1e-2 / rtol 1e-3,
// test/metrics.jl:112-114).
#include "f3d_common.cuh"

namespace f3d {
namespace {

constexpr int kBT = 256;

struct BwdParams {
    const float* A;
    const float* Bp;
    int N, M;
    const int32_t* nnA;
    const int32_t* nnB;
    const float* gout;
    float cA, cB;  // without the upstream gradient
    float* gA;
    float* gB;
    long totA, totB;  // B*N, B*M
};

template <bool kScatter>
__global__ void __launch_bounds__(kBT) chamfer_bwd_kernel(BwdParams p) {
    const long t = (long)blockIdx.x * kBT + threadIdx.x;
    const float g = __ldg(p.gout);
    if (t < p.totA) {
        const long b = t / p.N;
        const long o = (long)b * p.M + __ldg(p.nnA + t);
        const float c = p.cA * g;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float v = c * (__ldg(p.A + 3 * t + d) - __ldg(p.Bp + 3 * o + d));
            if (kScatter) atomicAdd(p.gB + 3 * o + d, -v);
            else p.gA[3 * t + d] = v;
        }
    } else if (t - p.totA < p.totB) {
        const long u = t - p.totA;
        const long b = u / p.M;
        const long o = (long)b * p.N + __ldg(p.nnB + u);
        const float c = p.cB * g;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float v = c * (__ldg(p.Bp + 3 * u + d) - __ldg(p.A + 3 * o + d));
            if (kScatter) atomicAdd(p.gA + 3 * o + d, -v);
            else p.gB[3 * u + d] = v;
        }
    }
}

}  // namespace
}  // namespace f3d

extern "C" int32_t f3d_chamfer_bwd(const float* A, const float* Bp, int32_t B, int32_t N, int32_t M, float w1, float w2,
                                   int32_t B_total, const int32_t* nnA_dev, const int32_t* nnB_dev,
                                   const float* gout_dev, float* gA, float* gB, f3d_stream_t stream_) {
    using namespace f3d;
    if (!A || !Bp || !nnA_dev || !nnB_dev || !gout_dev || !gA || !gB) return fail(F3D_ERR_INVALID, "f3d_chamfer_bwd: null pointer");
    if (B <= 0 || N <= 0 || M <= 0) return fail(F3D_ERR_INVALID, "f3d_chamfer_bwd: B, N, M must be positive (got %d, %d, %d)", B, N, M);
    if (B_total == 0) B_total = B;
    if (B_total < B) return fail(F3D_ERR_INVALID, "f3d_chamfer_bwd: B_total (%d) < B (%d)", B_total, B);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    BwdParams p;
    p.A = A; p.Bp = Bp; p.N = N; p.M = M; p.nnA = nnA_dev; p.nnB = nnB_dev; p.gout = gout_dev;
    p.cA = (float)(2.0 * (double)w1 / ((double)N * (double)B_total));
    p.cB = (float)(2.0 * (double)w2 / ((double)M * (double)B_total));
    p.gA = gA; p.gB = gB;
    p.totA = (long)B * N; p.totB = (long)B * M;
    const long tot = p.totA + p.totB;
    const unsigned grid = (unsigned)((tot + kBT - 1) / kBT);
    chamfer_bwd_kernel<false><<<grid, kBT, 0, stream>>>(p);
    F3D_CHECK_LAUNCH("chamfer_bwd_kernel<direct>");
    chamfer_bwd_kernel<true><<<grid, kBT, 0, stream>>>(p);
    F3D_CHECK_LAUNCH("chamfer_bwd_kernel<scatter>");
    return F3D_OK;
}
