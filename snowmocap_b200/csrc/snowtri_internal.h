// Internals shared by the translation units of libsnowtri.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "snowtri.h"
#include "snowtri_math.cuh"

namespace snowtri {
struct Params {
    double kst, ast, dthr, cond_tol, score_tol;
    float kst_f;  // smallest float >= kst: (float s < kst_f) <=> ((double)s < kst)
    int num_tol, center;
};
}  // namespace snowtri

char* snowtri_global_error();  // message buffer used when there is no handle (create failures)

#define SNOWTRI_PIPE_EVENTS 4

struct snowtri_handle {
    int device, C, sm_count, max_smem;
    size_t total_mem;         // device memory (sizes the scratch of the several-persons path)
    double* d_cam;  // (C,12) M = R*inv(K), t
    double* cam_host;
    double* kinv_host;  // (C,9) inverse intrinsics and (C,9) camera->world rotations, kept for the DLT mode
    double* r_host;
    int smem_per_sm;
    snowtri::Params prm;
    int precision;
    int tune_G, tune_ctas, tune_threads, no_fly, last_fly, no_p1;
    long long launches;
    int last_grid, last_block, last_smem, last_G;
    // device staging owned by the handle (snowtri_run_host only)
    void* stage[6];
    // two-stream pipeline of snowtri_run_host
    cudaStream_t pipe_in, pipe_k, pipe_out;   // host->device copies, kernels, device->host copies of snowtri_run_host
    cudaEvent_t pipe_ev[SNOWTRI_PIPE_EVENTS], pipe_evh[SNOWTRI_PIPE_EVENTS], pipe_start;
    int tune_chunk;  // frames per pipeline chunk (0 = automatic)
    void* gen_scratch;        // candidate scratch of the streaming general path
    size_t gen_scratch_bytes;
    void* p1_args;            // host copy of the single-person kernel's argument block
    void* mf_args;            // host copy of the multi-person fuse kernel's argument block
    size_t mf_args_bytes;
    int gen1_only;            // tests / comparisons: first-generation general kernels only
    int last_gen2;            // bit 0: second-generation match kernel ran, bit 1: second-generation fuse
    size_t p1_args_bytes;
    int jit_mode;             // 0 off, 1 auto (long batches), 2 always
    void* jit_cache;          // rig-specialised kernels (snowtri_jit.cu)
    // the last rig-specialised launch (snowtri_p1.cu): same argument block and batch shape -> same kernel, no source
    // string is rebuilt and hashed on the way to the launch
    unsigned long long p1_fast_key;
    void* p1_fast_fn;
    int p1_fast_grid, p1_fast_gw;
    size_t p1_fast_smem;
    char jit_status[512];
    void* nccl_comm;          // communicator owned by the handle (snowtri_comm.cu), or NULL
    int nccl_nranks, nccl_rank;
    int allow_f32_multi;  // tests only: float32 general kernel with several persons per camera
    size_t stage_cap[6];
    char err[512];
};

static inline int fail(snowtri_t* h, int code, const char* fmt, ...) {
    char* dst = h ? h->err : snowtri_global_error();
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(h, call)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(h, SNOWTRI_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                   \
    } while (0)


// streaming general path (snowtri_general.cu)
bool snowtri_general_eligible(const snowtri_t* h);
int snowtri_general_run(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts, int F, int P,
                        int J, int keypoint_num, int Pout, float* d_out, float* d_pscores, int* d_nout, void* stream);

// runtime specialisation (snowtri_jit.cu)
#ifdef __cplusplus
#include <string>
bool snowtri_jit_cached(const snowtri_t* h, const std::string& source);
void* snowtri_jit_get(snowtri_t* h, const std::string& source, const char* name, size_t smem);
int snowtri_jit_launch(void* fn, int grid, int block, size_t smem, void* stream, void* args);
void snowtri_jit_free(snowtri_t* h);
#endif

// single-person path (snowtri_p1.cu)
bool snowtri_p1_eligible(const snowtri_t* h, int P, int Pout, int keypoint_num);
int snowtri_p1_run(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts, int F, int J,
                   int keypoint_num, int Pout, float* d_out, float* d_pscores, int* d_nout, void* stream);
