// Development aid: what does a chunked H2D upload cost?  Per-op overheads of cudaMemcpyAsync / flag copies /
// cuStreamWriteValue32 on one or two copy streams (3.1 MB = cfg2's two clouds, pinned host memory).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
typedef CUresult (*wv32_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
int main() {
    const int B = 32; const size_t per = 4096 * 3 * sizeof(float);  // bytes per batch element per cloud
    float *hA, *hB, *dA, *dB; unsigned *flags, *hone;
    CK(cudaHostAlloc(&hA, B * per, 0)); CK(cudaHostAlloc(&hB, B * per, 0)); CK(cudaHostAlloc(&hone, 64, 0)); hone[0] = 1;
    CK(cudaMalloc(&dA, B * per)); CK(cudaMalloc(&dB, B * per)); CK(cudaMalloc(&flags, 256));
    cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, j; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
    wv32_t wv32 = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuStreamWriteValue32", (void**)&wv32, cudaEnableDefault, &qr));
    printf("cuStreamWriteValue32 %s\n", (wv32 && qr == cudaDriverEntryPointSuccess) ? "available" : "MISSING");
    const int sizes[6] = {1, 2, 4, 8, 8, 9};
    for (int mode = 0; mode < 7; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 20; ++rep) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, s1));
            if (mode == 0) { CK(cudaMemcpyAsync(dA, hA, B * per, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(dB, hB, B * per, cudaMemcpyHostToDevice, s1)); }
            else if (mode == 6) {  // unchunked, two streams
                CK(cudaEventRecord(j, s1)); CK(cudaStreamWaitEvent(s2, j, 0));
                CK(cudaMemcpyAsync(dA, hA, B * per, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(dB, hB, B * per, cudaMemcpyHostToDevice, s2));
                CK(cudaEventRecord(j, s2)); CK(cudaStreamWaitEvent(s1, j, 0));
            } else {
                const bool two = mode >= 4;
                if (two) { CK(cudaEventRecord(j, s1)); CK(cudaStreamWaitEvent(s2, j, 0)); }
                int b0 = 0;
                for (int c = 0; c < 6; ++c) {
                    CK(cudaMemcpyAsync((char*)dA + b0 * per, (char*)hA + b0 * per, sizes[c] * per, cudaMemcpyHostToDevice, s1));
                    CK(cudaMemcpyAsync((char*)dB + b0 * per, (char*)hB + b0 * per, sizes[c] * per, cudaMemcpyHostToDevice, two ? s2 : s1));
                    if (mode == 2) CK(cudaMemcpyAsync(flags + c, hone, 4, cudaMemcpyHostToDevice, s1));
                    if (mode == 3 || mode == 5) {
                        if (wv32(s1, (CUdeviceptr)(flags + c), 1u, 0) != CUDA_SUCCESS) { printf("wv32 failed\n"); return 1; }
                        if (two && wv32(s2, (CUdeviceptr)(flags + 16 + c), 1u, 0) != CUDA_SUCCESS) { printf("wv32 failed\n"); return 1; }
                    }
                    b0 += sizes[c];
                }
                if (two) { CK(cudaEventRecord(j, s2)); CK(cudaStreamWaitEvent(s1, j, 0)); }
            }
            CK(cudaEventRecord(e1, s1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        const char* names[7] = {"2 copies, one stream", "12 chunk copies, one stream", "12 chunk copies + 6 flag memcpys", "12 chunk copies + 6 writeValue32",
                                "12 chunk copies, two streams", "12 chunk copies + writeValue32, two streams", "2 copies, two streams"};
        printf("%-48s best %.1f us  (%.1f GB/s)\n", names[mode], best * 1e3, 2.0 * B * per / best / 1e6);
    }
    return 0;
}
