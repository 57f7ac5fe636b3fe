// ctx.h -- context object behind the opaque b200_ctx handle.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "b200aqc.h"

namespace b200 {
extern thread_local std::string g_last_error;
int set_error(const std::string& msg);

constexpr size_t PARTIAL_DOUBLES = (size_t)4096 * 64;  // per-block partials of the reductions
constexpr size_t OUT_DOUBLES = 4096;                   // small result staging (device + pinned host)

struct MpsState;  // mps.cu
}  // namespace b200

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            cudaGetLastError();                                                                     \
            return b200::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + \
                                   __FILE__ + ":" + std::to_string(__LINE__) + ")");                \
        }                                                                                           \
    } while (0)

struct b200_ctx {
    int device = 0;
    int num_sms = 0;
    bool coop_ok = false;  // device supports cooperative (grid-synchronised) launches
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_plan = nullptr;
    bool timing = false, timing_pending = false;
    bool plan_in_flight = false;

    // statevector slots
    int nq = 0;
    std::vector<void*> slots;
    std::vector<char> owned;

    // scratch
    double* d_partial = nullptr;
    double* d_out = nullptr;
    double* h_out = nullptr;  // pinned
    void* d_plan = nullptr;
    size_t d_plan_cap = 0;
    void* h_plan = nullptr;  // pinned
    size_t h_plan_cap = 0;

    // launch geometry of the sweep kernel
    int sweep_occ_smem = 1, sweep_occ_nosmem = 1, fused_occ = 1;
    int grid_mult = 1;
    int sweep_mode = 0;   // 0 = direct-load kernel (default), 1 = pipelined bulk-copy kernel (B200AQC_SWEEP=pipe)

    uint64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    // bench marks (b200_ctx_mark / b200_ctx_elapsed_ms)
    cudaEvent_t ev_mark[2] = {nullptr, nullptr};

    // per-kernel-class profile (b200_ctx_profile*): event pairs around every launch
    bool profiling = false;
    struct ProfRec { int cls; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    std::vector<float> sweep_log;   // per-launch ms of the sweep kernel while profiling (b200_ctx_profile_sweeps)
    double prof_ms[B200_PROF_CLASSES] = {0};
    uint64_t prof_n[B200_PROF_CLASSES] = {0};

    b200::MpsState* mps = nullptr;
};

// releases everything the MPS path allocated on this context (mps.cu)
void b200_mps_release(b200_ctx* ctx);
