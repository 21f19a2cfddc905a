// Device / memory / stream / event / graph plumbing of the C ABI (include/neuromah_b200.h).
#include "nm_common.cuh"
#include <cuda_profiler_api.h>
#include <mutex>
#include <unordered_map>

namespace nm {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
}

extern "C" {

const char* nm_last_error(void) { return nm::g_err; }
int nm_version(void) { return 1; }

int nm_device_count(int* count) {
    NM_REQUIRE(count, "count is NULL");
    NM_CUDA(cudaGetDeviceCount(count));
    return NM_OK;
}

int nm_init(int device) {
    NM_CUDA(cudaSetDevice(device));
    NM_CUDA(cudaFree(0));
    return NM_OK;
}

int nm_device_props(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
    cudaDeviceProp p;
    NM_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return NM_OK;
}

int nm_malloc(void** dptr, size_t bytes) {
    NM_REQUIRE(dptr, "dptr is NULL");
    *dptr = nullptr;
    if (bytes == 0) return NM_OK;
    NM_CUDA(cudaMalloc(dptr, bytes));
    return NM_OK;
}
int nm_free(void* dptr) { if (dptr) NM_CUDA(cudaFree(dptr)); return NM_OK; }
int nm_host_alloc(void** hptr, size_t bytes) {
    NM_REQUIRE(hptr, "hptr is NULL");
    NM_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1));
    return NM_OK;
}
int nm_host_free(void* hptr) { if (hptr) NM_CUDA(cudaFreeHost(hptr)); return NM_OK; }

int nm_memset(void* dptr, int byte, size_t bytes, void* stream) {
    if (bytes) NM_CUDA(cudaMemsetAsync(dptr, byte, bytes, nm::S(stream)));
    return NM_OK;
}
int nm_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes) NM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, nm::S(stream)));
    return NM_OK;
}
int nm_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes) NM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, nm::S(stream)));
    return NM_OK;
}
int nm_d2d(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes) NM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, nm::S(stream)));
    return NM_OK;
}

int nm_stream_create(void** stream) {
    NM_REQUIRE(stream, "stream is NULL");
    cudaStream_t s;
    NM_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return NM_OK;
}
int nm_stream_destroy(void* stream) { if (stream) NM_CUDA(cudaStreamDestroy(nm::S(stream))); return NM_OK; }
int nm_stream_sync(void* stream) { NM_CUDA(cudaStreamSynchronize(nm::S(stream))); return NM_OK; }
int nm_device_sync(void) { NM_CUDA(cudaDeviceSynchronize()); return NM_OK; }

int nm_event_create(void** event) {
    NM_REQUIRE(event, "event is NULL");
    cudaEvent_t e;
    NM_CUDA(cudaEventCreate(&e));
    *event = e;
    return NM_OK;
}
int nm_event_destroy(void* event) { if (event) NM_CUDA(cudaEventDestroy((cudaEvent_t)event)); return NM_OK; }
int nm_event_record(void* event, void* stream) { NM_CUDA(cudaEventRecord((cudaEvent_t)event, nm::S(stream))); return NM_OK; }
int nm_event_sync(void* event) { NM_CUDA(cudaEventSynchronize((cudaEvent_t)event)); return NM_OK; }
int nm_event_elapsed_ms(void* start, void* stop, float* ms) {
    NM_REQUIRE(ms, "ms is NULL");
    NM_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return NM_OK;
}
int nm_stream_wait_event(void* stream, void* event) {
    NM_CUDA(cudaStreamWaitEvent(nm::S(stream), (cudaEvent_t)event, 0));
    return NM_OK;
}

// kernels captured per graph, so that nm_launch_count() also counts launches replayed from a graph
static std::mutex g_graph_mu;
static std::unordered_map<void*, unsigned long long> g_graph_kernels;
static thread_local unsigned long long g_capture_start = 0;

// Relaxed mode: a cudaFree / cudaMalloc issued by this thread while the capture is open (a Python finalizer run by the
// cyclic GC, a lazily created workspace) neither fails nor invalidates the capture -- in thread-local mode it did both
// (cudaErrorStreamCaptureInvalidated at the next captured launch).  The host side still defers frees (device.py).
int nm_graph_begin(void* stream) {
    NM_CUDA(cudaStreamBeginCapture(nm::S(stream), cudaStreamCaptureModeRelaxed));
    g_capture_start = nm::g_launches.load();
    return NM_OK;
}
// Close a capture whose body failed: end it, drop whatever was captured, clear the sticky capture error.
int nm_graph_abort(void* stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(nm::S(stream), &st);
    if (st != cudaStreamCaptureStatusNone) {
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(nm::S(stream), &g);
        if (g) cudaGraphDestroy(g);
    }
    cudaGetLastError();
    const unsigned long long captured = nm::g_launches.load() - g_capture_start;
    nm::g_launches.fetch_sub(captured);
    return NM_OK;
}
int nm_graph_end(void* stream, void** graph_exec) {
    NM_REQUIRE(graph_exec, "graph_exec is NULL");
    cudaGraph_t g = nullptr;
    NM_CUDA(cudaStreamEndCapture(nm::S(stream), &g));
    cudaGraphExec_t ge = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return nm::cuda_fail(e, "cudaGraphInstantiate");
    *graph_exec = ge;
    const unsigned long long captured = nm::g_launches.load() - g_capture_start;
    nm::g_launches.fetch_sub(captured);           // captured, not executed: they count when the graph is launched
    std::lock_guard<std::mutex> lk(g_graph_mu);
    g_graph_kernels[ge] = captured;
    return NM_OK;
}
int nm_graph_launch(void* graph_exec, void* stream) {
    NM_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, nm::S(stream)));
    std::lock_guard<std::mutex> lk(g_graph_mu);
    auto it = g_graph_kernels.find(graph_exec);
    if (it != g_graph_kernels.end()) nm::g_launches.fetch_add(it->second);
    return NM_OK;
}
int nm_graph_destroy(void* graph_exec) {
    if (graph_exec) {
        NM_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
        std::lock_guard<std::mutex> lk(g_graph_mu);
        g_graph_kernels.erase(graph_exec);
    }
    return NM_OK;
}

int nm_profiler_start(void) { NM_CUDA(cudaProfilerStart()); return NM_OK; }
int nm_profiler_stop(void) { NM_CUDA(cudaProfilerStop()); return NM_OK; }

int nm_launch_count(unsigned long long* count) {
    NM_REQUIRE(count, "count is NULL");
    *count = nm::g_launches.load();
    return NM_OK;
}

}  // extern "C"
