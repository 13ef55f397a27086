// microbench.cu — measures the sm_100a instruction throughputs the chamfer design rests on
// (packed FP32 FADD2/FMUL2/FFMA2, scalar FADD/FMUL/FFMA, FMNMX/FMNMX3, CREDUX, VOTE, and the
// mixed per-pair recipe).  Output: lane-ops per clock per SM, from in-kernel clock64 and from events.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned long long u64;
#define DEV __device__ __forceinline__
DEV u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r; }
DEV void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }

constexpr int U = 8;        // independent chains per thread
constexpr int ITERS = 4096;

template <int OP>
__global__ void __launch_bounds__(256) k_op(float* out, long long* cyc, float seed) {
    float v[U], w[U];
    u64 p[U];
#pragma unroll
    for (int i = 0; i < U; ++i) { v[i] = seed + i + threadIdx.x; w[i] = seed * 0.5f + i; p[i] = pk(v[i], w[i]); }
    u64 q = pk(seed, seed + 1.0f);
    unsigned acc = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(seed));
            if (OP == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(seed));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(seed), "f"(w[i]));
            if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q));
            if (OP == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q));
            if (OP == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(q));
            if (OP == 6) asm volatile("min.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(w[i]));
            if (OP == 7) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(w[i]), "f"(seed));
            if (OP == 8) { unsigned r; asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(__float_as_uint(v[i]))); acc += r; }
            if (OP == 9) { unsigned r; asm volatile("{ .reg .pred pp; setp.eq.f32 pp, %1, %2; vote.sync.ballot.b32 %0, pp, 0xffffffff; }" : "=r"(r) : "f"(v[i]), "f"(w[i])); acc += r; }
            if (OP == 10) asm volatile("add.rn.f32x2 %0, %1, %0;" : "+l"(p[i]) : "l"(pk(seed, seed)));  // broadcast-operand form
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < U; ++i) { float a, b; upk(p[i], a, b); s += v[i] + a + b; }
    if (s == 123.456f || acc == 0x12345u) out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// The per-pair recipe of the exact chamfer sweep: per (8 rows x 2 columns): 24 FADD2, 24 FMUL2, 32 FADD,
// 16 FMNMX3 (+ optional 2 CREDUX + 2 VOTE).  MODE 0: exact; 1: fma (24 FADD2, 8 FMUL2, 16 FFMA2); +2: with redux/vote
template <int MODE>
__global__ void __launch_bounds__(128, 4) k_recipe(const float* __restrict__ in, float* out, long long* cyc, int iters) {
    float ax[8], ay[8], az[8], amin[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { ax[r] = in[threadIdx.x + r]; ay[r] = in[threadIdx.x + 8 + r]; az[r] = in[threadIdx.x + 16 + r]; amin[r] = 1e30f; }
    const u64 one2 = pk(in[1000], in[1000]);  // 1.0f at run time
    __shared__ float4 sxy[64];
    __shared__ float2 sz[64];
    if (threadIdx.x < 64) { sxy[threadIdx.x] = make_float4(in[threadIdx.x], in[threadIdx.x+1], in[threadIdx.x+2], in[threadIdx.x+3]); sz[threadIdx.x] = make_float2(in[threadIdx.x+4], in[threadIdx.x+5]); }
    __syncthreads();
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int q = 0; q < 64; ++q) {
            float4 xy = sxy[q]; float2 zz = sz[q];
            u64 bx = pk(xy.x, xy.y), by = pk(xy.z, xy.w), bz = pk(zz.x, zz.y);
            float c0 = 1e30f, c1 = 1e30f;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                u64 dx, dy, dz; float d0, d1;
                asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(pk(ax[r], ax[r])), "l"(bx));
                asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(pk(ay[r], ay[r])), "l"(by));
                asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(pk(az[r], az[r])), "l"(bz));
                if (MODE & 4) {   // exact, adds issued as FFMA2 x 1.0 (ptxas cannot contract a runtime multiplier)
                    u64 sx, sy, s2, t;
                    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(sx) : "l"(dx));
                    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(sy) : "l"(dy));
                    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(s2) : "l"(dz));
                    asm("fma.rn.f32x2 %0, %1, %3, %2;" : "=l"(t) : "l"(sx), "l"(sy), "l"(one2));
                    asm("fma.rn.f32x2 %0, %1, %3, %2;" : "=l"(t) : "l"(t), "l"(s2), "l"(one2));
                    upk(t, d0, d1);
                } else if (MODE & 1) {
                    u64 s;
                    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(s) : "l"(dx));
                    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(s) : "l"(dy));
                    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(s) : "l"(dz));
                    upk(s, d0, d1);
                } else {
                    u64 sx, sy, s2; float x0, x1, y0, y1, z0, z1;
                    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(sx) : "l"(dx));
                    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(sy) : "l"(dy));
                    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(s2) : "l"(dz));
                    upk(sx, x0, x1); upk(sy, y0, y1); upk(s2, z0, z1);
                    d0 = __fadd_rn(__fadd_rn(x0, y0), z0); d1 = __fadd_rn(__fadd_rn(x1, y1), z1);
                }
                amin[r] = fminf(amin[r], fminf(d0, d1));
                c0 = fminf(c0, d0); c1 = fminf(c1, d1);
            }
            if (MODE & 2) {
                unsigned u0 = __float_as_uint(c0), u1 = __float_as_uint(c1);
                unsigned m0 = __reduce_min_sync(0xffffffffu, u0), m1 = __reduce_min_sync(0xffffffffu, u1);
                acc += __ballot_sync(0xffffffffu, u0 == m0) + __ballot_sync(0xffffffffu, u1 == m1) + m0 + m1;
            } else {
                acc += __float_as_uint(c0) ^ __float_as_uint(c1);
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += amin[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("device %s SMs %d clock %d kHz\n", prop.name, sms, prop.clockRate);
    float* out; long long* cyc; float* in;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8 * 4096); cudaMalloc(&in, 4096);
    std::vector<float> h(1024); for (int i = 0; i < 1024; ++i) h[i] = (float)rand() / RAND_MAX; h[1000] = 1.0f; cudaMemcpy(in, h.data(), 4096, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "FMNMX", "FMNMX3", "REDUX.MIN", "SETP+VOTE", "FADD2(bcast)"};
    const int lanes_per_op[] = {1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 2};
    for (int bps = 1; bps <= 8; bps *= 2) {  // blocks (256 thr) per SM
        printf("--- %d warps/SM\n", bps * 8);
        for (int op = 0; op <= 10; ++op) {
            int grid = sms * bps; float ms = 0; std::vector<long long> hc(grid);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                switch (op) {
                    case 0: k_op<0><<<grid, 256>>>(out, cyc, 1.0001f); break; case 1: k_op<1><<<grid, 256>>>(out, cyc, 1.0001f); break;
                    case 2: k_op<2><<<grid, 256>>>(out, cyc, 1.0001f); break; case 3: k_op<3><<<grid, 256>>>(out, cyc, 1.0001f); break;
                    case 4: k_op<4><<<grid, 256>>>(out, cyc, 1.0001f); break; case 5: k_op<5><<<grid, 256>>>(out, cyc, 1.0001f); break;
                    case 6: k_op<6><<<grid, 256>>>(out, cyc, 1.0001f); break; case 7: k_op<7><<<grid, 256>>>(out, cyc, 1.0001f); break;
                    case 8: k_op<8><<<grid, 256>>>(out, cyc, 1.0001f); break; case 9: k_op<9><<<grid, 256>>>(out, cyc, 1.0001f); break;
                    case 10: k_op<10><<<grid, 256>>>(out, cyc, 1.0001f); break;
                }
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            }
            cudaMemcpy(hc.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (auto c : hc) avg += c; avg /= grid;
            double warp_instr_per_sm = (double)bps * 8 * ITERS * U;
            printf("%-14s cyc/warp-instr/SM %.3f   lane-ops/clk/SM %.1f   (%.3f ms, eff clk %.0f MHz)\n", names[op],
                   avg / warp_instr_per_sm, warp_instr_per_sm * 32 * lanes_per_op[op] / avg, ms, avg / (ms * 1e3));
        }
    }
    for (int mode = 0; mode < 8; ++mode) {
        if (mode == 5 || mode == 7) continue;
        for (int bps = 1; bps <= 4; ++bps) {
            int grid = sms * bps, iters = 64; float ms = 0; std::vector<long long> hc(grid);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k_recipe<0><<<grid, 128>>>(in, out, cyc, iters); if (mode == 1) k_recipe<1><<<grid, 128>>>(in, out, cyc, iters);
                if (mode == 2) k_recipe<2><<<grid, 128>>>(in, out, cyc, iters); if (mode == 3) k_recipe<3><<<grid, 128>>>(in, out, cyc, iters);
                if (mode == 4) k_recipe<4><<<grid, 128>>>(in, out, cyc, iters); if (mode == 6) k_recipe<6><<<grid, 128>>>(in, out, cyc, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            }
            cudaMemcpy(hc.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (auto c : hc) avg += c; avg /= grid;
            double pairs_per_sm = (double)bps * 128 * iters * 64 * 16;
            printf("recipe mode %d (%s%s) %2d warps/SM: cycles/pair/lane %.3f  pairs/clk/SM %.2f  -> %.3e pairs/s chip @event-time (%.3f ms)\n", mode,
                   (mode & 4) ? "exact-ffma2x1" : (mode & 1) ? "fma" : "exact", (mode & 2) ? "+redux" : "", bps * 4, avg * 128 / pairs_per_sm, pairs_per_sm / avg,
                   pairs_per_sm * sms / (ms * 1e-3), ms);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
