// Library plumbing: error slot, launch counter, init, memory, copies.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "adb_internal.h"

namespace adb {
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int g_precision = 0;
}  // namespace adb

extern "C" {

const char* adb_last_error(void) { return adb::g_err; }
const char* adb_version(void) { return "adb200 0.1 (sm_100a)"; }
int64_t adb_launch_count(void) { return adb::g_launches.load(); }

int adb_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return adb::set_error("adb_init: no CUDA device visible (%s) — this backend has no CPU path",
                              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    ADB_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp p;
    ADB_CUDA_OK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) return adb::set_error("adb_init: device %s is sm_%d%d; this library is built for sm_100a only", p.name, p.major, p.minor);
    // keep freed workspaces in the stream-ordered pool: the default threshold (0) hands memory back to the driver at
    // every synchronisation, which turns each cudaMallocAsync after a sync into a millisecond-scale allocation
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    return 0;
}

int adb_set_precision(int precision) {
    if (precision != 0 && precision != 1) return adb::set_error("adb_set_precision: unknown precision %d", precision);
    adb::g_precision = precision;
    return 0;
}
int adb_get_precision(void) { return adb::g_precision; }

int adb_sync(void* stream) {
    ADB_CUDA_OK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}

int adb_malloc(void** out, size_t bytes, void* stream) {
    ADB_CUDA_OK(cudaMallocAsync(out, bytes ? bytes : 16, reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}
int adb_free(void* ptr, void* stream) {
    if (!ptr) return 0;
    ADB_CUDA_OK(cudaFreeAsync(ptr, reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}
int adb_memcpy_h2d(void* dst, const void* host_src, size_t bytes, void* stream) {
    if (!bytes) return 0;
    ADB_CUDA_OK(cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}
int adb_memcpy_d2h(void* host_dst, const void* src, size_t bytes, void* stream) {
    if (!bytes) return 0;
    ADB_CUDA_OK(cudaMemcpyAsync(host_dst, src, bytes, cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}
int adb_memcpy2d_h2d(void* dst, size_t dst_pitch, const void* host_src, size_t src_pitch, size_t width_bytes, size_t rows, void* stream) {
    if (!width_bytes || !rows) return 0;
    ADB_CUDA_OK(cudaMemcpy2DAsync(dst, dst_pitch, host_src, src_pitch, width_bytes, rows, cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}
int adb_memcpy2d_d2h(void* host_dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows, void* stream) {
    if (!width_bytes || !rows) return 0;
    ADB_CUDA_OK(cudaMemcpy2DAsync(host_dst, dst_pitch, src, src_pitch, width_bytes, rows, cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}
int adb_host_alloc(void** out, size_t bytes) {
    ADB_CUDA_OK(cudaMallocHost(out, bytes ? bytes : 16));
    return 0;
}
int adb_host_free(void* ptr) {
    if (!ptr) return 0;
    ADB_CUDA_OK(cudaFreeHost(ptr));
    return 0;
}

int adb_memset_async(void* ptr, int byte_value, size_t bytes, void* stream) {
    if (!bytes) return 0;
    ADB_CUDA_OK(cudaMemsetAsync(ptr, byte_value, bytes, reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}
int adb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
    if (!bytes) return 0;
    ADB_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}

// ---- CUDA-graph replay of launch tapes (core/tape.py).  Relaxed mode: the planner may call cudaMalloc / cudaFuncSetAttribute
// while a recording is open (workspaces are warmed up by an eager pass first, so in practice it does not).
int adb_graph_begin(void* capture_stream) {
    if (!capture_stream) return adb::set_error("adb_graph_begin: the legacy default stream cannot be captured");
    ADB_CUDA_OK(cudaStreamBeginCapture(reinterpret_cast<cudaStream_t>(capture_stream), cudaStreamCaptureModeRelaxed));
    return 0;
}
int adb_graph_end(void* capture_stream, void** exec) {
    *exec = nullptr;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(reinterpret_cast<cudaStream_t>(capture_stream), &graph);
    if (e != cudaSuccess || !graph) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        return adb::set_error("adb_graph_end: capture failed (%s)", cudaGetErrorString(e));
    }
    cudaGraphExec_t ex = nullptr;
    e = cudaGraphInstantiate(&ex, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return adb::set_error("adb_graph_end: cudaGraphInstantiate failed (%s)", cudaGetErrorString(e));
    }
    *exec = ex;
    return 0;
}
int adb_graph_launch(void* exec, void* stream) {
    ADB_CUDA_OK(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(exec), reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}
int adb_graph_destroy(void* exec) {
    if (exec) ADB_CUDA_OK(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(exec)));
    return 0;
}

}  // extern "C"
