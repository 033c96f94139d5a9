// C-ABI housekeeping: error text, device query, workspace sizing.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace fb200 {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return (dev < 0 || dev >= MAX_DEVICES) ? 0 : dev;
}

int sm_count() {
    static std::mutex m;
    static int n[MAX_DEVICES] = {};
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(m);
    if (n[dev] == 0) {
        if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0) {
            cudaGetLastError();
            n[dev] = 148;   // B200
        }
    }
    return n[dev];
}

size_t dense_partial_elems(int64_t M, int64_t N);

}  // namespace fb200

extern "C" int fb200_abi_version(void) { return FB200_ABI_VERSION; }

extern "C" const char* fb200_last_error(void) { return fb200::g_error; }

extern "C" size_t fb200_workspace_bytes(int64_t M, int64_t N) {
    if (M < 1) M = 1;
    if (N < 1) N = 1;
    return fb200::DENSE_OFF + fb200::dense_partial_elems(M, N) * sizeof(double);
}

// The per-trial snapshot off the compute stream: record `ev_main` on the stream the kernels run on, make the side stream
// wait for it, copy `bytes` from the device ring slot to pinned host memory there and record `ev_done` behind the copy.
// The compute stream never sees a memcpy node (its next kernel starts right behind the deciding kernel); the host waits
// on `ev_done`.  Events and streams are the caller's (cudaEvent_t / cudaStream_t passed as void*).
extern "C" int fb200_snapshot_copy(void* dst_host, const void* src_dev, size_t bytes, void* main_stream, void* side_stream,
                                   void* ev_main, void* ev_done) {
    cudaStream_t ms = static_cast<cudaStream_t>(main_stream), ss = static_cast<cudaStream_t>(side_stream);
    cudaEvent_t e0 = static_cast<cudaEvent_t>(ev_main), e1 = static_cast<cudaEvent_t>(ev_done);
    cudaError_t e = cudaEventRecord(e0, ms);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ss, e0, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ss);
    if (e == cudaSuccess) e = cudaEventRecord(e1, ss);
    if (e != cudaSuccess) { fb200::set_error("snapshot_copy: %s", cudaGetErrorString(e)); cudaGetLastError(); return 1; }
    return 0;
}
