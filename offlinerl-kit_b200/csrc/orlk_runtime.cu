// Runtime plumbing of the C ABI: error text, CUDA-graph capture/replay, copies, events.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "orlk_common.cuh"

namespace orlk {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("ORLK_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}
static unsigned long long* g_trace = nullptr;
unsigned long long* trace_buffer() { return g_trace; }
void set_trace_buffer(unsigned long long* p) { g_trace = p; }

void prepare_kernel(const void* fn) {
    static const void* seen[256];
    static int n_seen = 0;
    static int carve = -2;
    if (carve == -2) {
        const char* e = getenv("ORLK_CARVEOUT");
        carve = e ? atoi(e) : -1;
    }
    if (carve < 0) return;
    for (int i = 0; i < n_seen; ++i)
        if (seen[i] == fn) return;
    if (n_seen < 256) seen[n_seen++] = fn;
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
}
}  // namespace orlk
using namespace orlk;

extern "C" {

int orlk_abi_version(void) { return ORLK_ABI_VERSION; }
const char* orlk_last_error(void) { return g_err; }
int orlk_sizeof_gemm_desc(void) { return (int)sizeof(OrlkGemmDesc); }
int orlk_sizeof_adam_desc(void) { return (int)sizeof(OrlkAdamDesc); }
int orlk_sizeof_adam_group(void) { return (int)sizeof(OrlkAdamGroup); }
int orlk_sizeof_concat_seg(void) { return (int)sizeof(OrlkConcatSeg); }
int orlk_sizeof_sample_use(void) { return (int)sizeof(OrlkSampleUse); }

int orlk_device_info(int device, int* out4) {
    ORLK_REQUIRE(out4 != nullptr, "out4 is NULL");
    cudaDeviceProp p;
    int rc = check(cudaGetDeviceProperties(&p, device), "cudaGetDeviceProperties");
    if (rc) return rc;
    out4[0] = p.multiProcessorCount;
    out4[1] = p.major;
    out4[2] = p.minor;
    out4[3] = (int)p.sharedMemPerBlockOptin;
    return 0;
}

int orlk_graph_begin(void* stream) {
    return check(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
}

int orlk_capture_status(void* stream) {     /* 0 = not capturing, 1 = capturing, 2 = capture invalidated, < 0 = query failed */
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing((cudaStream_t)stream, &st) != cudaSuccess) { cudaGetLastError(); return -1; }
    return st == cudaStreamCaptureStatusActive ? 1 : (st == cudaStreamCaptureStatusInvalidated ? 2 : 0);
}

int orlk_graph_end(void* stream, void** graph_exec_out) {
    ORLK_REQUIRE(graph_exec_out != nullptr, "graph_exec_out is NULL");
    cudaGraph_t g = nullptr;
    int rc = check(cudaStreamEndCapture((cudaStream_t)stream, &g), "cudaStreamEndCapture");
    if (rc) return rc;
    if (getenv("ORLK_GRAPH_DEBUG")) {
        size_t n_nodes = 0, n_edges = 0;
        cudaGraphGetNodes(g, nullptr, &n_nodes);
        cudaGraphGetEdges_v2(g, nullptr, nullptr, nullptr, &n_edges);
        cudaGraphNode_t* from = (cudaGraphNode_t*)malloc(sizeof(cudaGraphNode_t) * (n_edges + 1));
        cudaGraphNode_t* to = (cudaGraphNode_t*)malloc(sizeof(cudaGraphNode_t) * (n_edges + 1));
        cudaGraphEdgeData* ed = (cudaGraphEdgeData*)malloc(sizeof(cudaGraphEdgeData) * (n_edges + 1));
        size_t prog = 0;
        if (cudaGraphGetEdges_v2(g, from, to, ed, &n_edges) == cudaSuccess)
            for (size_t i = 0; i < n_edges; ++i) prog += ed[i].type == cudaGraphDependencyTypeProgrammatic;
        fprintf(stderr, "[orlk] graph: %zu nodes, %zu edges, %zu programmatic\n", n_nodes, n_edges, prog);
        free(from); free(to); free(ed);
    }
    cudaGraphExec_t ex = nullptr;
    // per-node priorities (the optimiser launches, orlk::launch_high_priority) are honoured only with this flag
    rc = check(cudaGraphInstantiate(&ex, g, cudaGraphInstantiateFlagUseNodePriority), "cudaGraphInstantiate");
    cudaGraphDestroy(g);
    if (rc) return rc;
    *graph_exec_out = (void*)ex;
    return 0;
}

int orlk_graph_launch(void* graph_exec, void* stream) {
    return check(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream), "cudaGraphLaunch");
}

int orlk_graph_launch_sync(void* graph_exec, void* stream) {
    int rc = check(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream), "cudaGraphLaunch");
    if (rc) return rc;
    return check(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize");
}

int orlk_graph_launch_wait_event(void* graph_exec, void* stream, void* ev) {
    int rc = check(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream), "cudaGraphLaunch");
    if (rc) return rc;
    return check(cudaEventSynchronize((cudaEvent_t)ev), "cudaEventSynchronize");
}

int orlk_event_record_external(void* ev, void* stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    int rc = check(cudaStreamIsCapturing((cudaStream_t)stream, &st), "cudaStreamIsCapturing");
    if (rc) return rc;
    if (st == cudaStreamCaptureStatusActive)        // an event-record NODE: every launch of the graph records the event
        return check(cudaEventRecordWithFlags((cudaEvent_t)ev, (cudaStream_t)stream, cudaEventRecordExternal), "cudaEventRecordWithFlags");
    return check(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream), "cudaEventRecord");
}

int orlk_graph_destroy(void* graph_exec) {
    return check(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec), "cudaGraphExecDestroy");
}

int orlk_stream_create(void** stream_out) {
    ORLK_REQUIRE(stream_out != nullptr, "stream_out is NULL");
    cudaStream_t s;
    int rc = check(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
    if (rc) return rc;
    *stream_out = (void*)s;
    return 0;
}
int orlk_stream_destroy(void* stream) { return check(cudaStreamDestroy((cudaStream_t)stream), "cudaStreamDestroy"); }
int orlk_stream_wait_event(void* stream, void* ev) {
    return check(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0), "cudaStreamWaitEvent");
}
int orlk_event_create_notiming(void** ev_out) {
    ORLK_REQUIRE(ev_out != nullptr, "ev_out is NULL");
    cudaEvent_t e;
    int rc = check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    if (rc) return rc;
    *ev_out = (void*)e;
    return 0;
}

int orlk_stream_sync(void* stream) { return check(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize"); }

int orlk_memcpy_h2d_async(void* dst, const void* src_host, size_t bytes, void* stream) {
    return check(cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), "memcpy h2d");
}
int orlk_memcpy_d2h_async(void* dst_host, const void* src, size_t bytes, void* stream) {
    return check(cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "memcpy d2h");
}
int orlk_memcpy_d2d_async(void* dst, const void* src, size_t bytes, void* stream) {
    return check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream), "memcpy d2d");
}
int orlk_memset_async(void* dst, int value, size_t bytes, void* stream) {
    return check(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream), "memset");
}

int orlk_event_create(void** ev_out) {
    ORLK_REQUIRE(ev_out != nullptr, "ev_out is NULL");
    cudaEvent_t e;
    int rc = check(cudaEventCreate(&e), "cudaEventCreate");
    if (rc) return rc;
    *ev_out = (void*)e;
    return 0;
}
int orlk_event_record(void* ev, void* stream) {
    return check(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream), "cudaEventRecord");
}
int orlk_event_sync(void* ev) { return check(cudaEventSynchronize((cudaEvent_t)ev), "cudaEventSynchronize"); }
int orlk_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms_out) {
    int rc = check(cudaEventSynchronize((cudaEvent_t)ev_stop), "cudaEventSynchronize");
    if (rc) return rc;
    return check(cudaEventElapsedTime(ms_out, (cudaEvent_t)ev_start, (cudaEvent_t)ev_stop), "cudaEventElapsedTime");
}
int orlk_event_destroy(void* ev) { return check(cudaEventDestroy((cudaEvent_t)ev), "cudaEventDestroy"); }

}  // extern "C"
