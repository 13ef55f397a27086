// capi.cu — version / error plumbing of the C ABI (include/flux3d_b200.h).
#include <cstring>

#include "f3d_common.cuh"

namespace f3d {
namespace {
thread_local char g_err[512] = "";
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int32_t fail(int32_t code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int32_t cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return F3D_ERR_CUDA;
}
}  // namespace f3d

extern "C" int32_t f3d_version(void) { return 100; /* 0.1.0 */ }

extern "C" int32_t f3d_last_error(char* buf, size_t n) {
    size_t len = strlen(f3d::g_err);
    if (buf && n > 0) {
        size_t c = len < n - 1 ? len : n - 1;
        memcpy(buf, f3d::g_err, c);
        buf[c] = '\0';
    }
    return (int32_t)len;
}
