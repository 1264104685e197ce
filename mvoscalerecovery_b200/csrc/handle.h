// Internal to libmvosr.so: the handle behind the opaque mvosr_handle of include/mvosr.h and the error plumbing shared by the
// translation units (api.cu, five_point_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../include/mvosr.h"

extern thread_local char g_cuda_err[256];

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "%s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return MVOSR_E_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

static const int NCOUNTERS = 16;

struct mvosr_handle {
    mvosr_config cfg;
    int device;
    int num_sms;
    int smem_optin;
    int cap_max;
    int *work_counter;           // device: NCOUNTERS dynamic-scheduler counters (one per in-flight launch)
    int counter_slot;            // round-robin
    cudaStream_t s_copy, s_comp[2]; cudaEvent_t ev_copy[8]; int streams_ready;    // host-buffer pipeline
    int64_t launches;
    // host-API staging (grown on demand)
    void *d_stage; size_t stage_bytes;
    // large-frame staging (frames beyond the shared-memory capacity): two sets of per-CTA slabs used alternately, so that two
    // launches in flight (the two compute streams of the host pipeline) never share one; a launch waits for the previous user
    // of its set through that set's event
    void *d_ws[2]; size_t ws_bytes[2]; cudaEvent_t ev_ws[2]; int ev_ws_ready[2]; int ws_slot;
    // single-frame host entry (mvosr_scale_frame_host_f64): device staging, its pinned mirror, a private stream
    void *d_frame, *h_frame; size_t frame_bytes; cudaStream_t s_frame; int s_frame_ready;
    long long *phase_cycles;     // optional profiling sink (device), set by mvosr_set_phase_timing
};
