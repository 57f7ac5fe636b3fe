// api.cu -- C-ABI entry points of libb200aqc.so (statevector path).  See include/b200aqc.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b200aqc.h"
#include "ctx.h"
#include "sv_kernels.cuh"
#include "sv_plan.h"

using namespace b200;

namespace b200 {
thread_local std::string g_last_error;
int set_error(const std::string& msg) {
    g_last_error = msg;
    return -1;
}
}  // namespace b200

namespace {

int ensure_scratch(b200_ctx* ctx) {
    if (ctx->d_partial) return 0;
    CUDA_TRY(cudaMalloc(&ctx->d_partial, PARTIAL_DOUBLES * sizeof(double)));
    CUDA_TRY(cudaMalloc(&ctx->d_out, OUT_DOUBLES * sizeof(double)));
    CUDA_TRY(cudaMallocHost(&ctx->h_out, OUT_DOUBLES * sizeof(double)));
    return 0;
}

int check_slot(b200_ctx* ctx, int slot) {
    if (!ctx) return set_error("null context");
    if (ctx->nq <= 0) return set_error("statevector not allocated (call b200_sv_alloc)");
    if (slot < 0 || slot >= (int)ctx->slots.size() || !ctx->slots[slot])
        return set_error("invalid statevector slot " + std::to_string(slot));
    return 0;
}

// grows a device buffer to hold `bytes`
int reserve_dev(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const size_t want = std::max(bytes * 2, (size_t)1 << 16);
    CUDA_TRY(cudaMalloc(p, want));
    *cap = want;
    return 0;
}

int reserve_host(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return 0;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    const size_t want = std::max(bytes * 2, (size_t)1 << 16);
    CUDA_TRY(cudaMallocHost(p, want));
    *cap = want;
    return 0;
}

struct Timer {
    b200_ctx* ctx;
    explicit Timer(b200_ctx* c) : ctx(c) {
        if (ctx->timing) cudaEventRecord(ctx->ev0, ctx->stream);
    }
    void stop() {
        if (ctx->timing) {
            cudaEventRecord(ctx->ev1, ctx->stream);
            ctx->timing_pending = true;
        }
    }
};

// Brackets one kernel launch: counts it and, in profile mode, records an event pair around it.
struct KScope {
    b200_ctx* c;
    int cls;
    cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t get_ev(b200_ctx* c) {
        if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    KScope(b200_ctx* ctx, int k) : c(ctx), cls(k) {
        c->counters[0] += 1;
        if (c->profiling) {
            a = get_ev(c); b = get_ev(c);
            cudaEventRecord(a, c->stream);
        }
    }
    ~KScope() {
        if (a) {
            cudaEventRecord(b, c->stream);
            c->prof_recs.push_back({cls, a, b});
        }
    }
};

// Folds finished event pairs into the per-class accumulators (synchronises the stream).
int prof_drain(b200_ctx* ctx) {
    if (ctx->prof_recs.empty()) return 0;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (auto& r : ctx->prof_recs) {
        float f = 0;
        CUDA_TRY(cudaEventElapsedTime(&f, r.a, r.b));
        ctx->prof_ms[r.cls] += f;
        if (r.cls == B200_PROF_SWEEP && ctx->sweep_log.size() < 65536) ctx->sweep_log.push_back(f);
        ctx->prof_n[r.cls] += 1;
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof_recs.clear();
    return 0;
}

// Upload the small-path program (ops, mat2 table) in one pinned staging copy.  (The tiled path needs no
// upload: each sweep's program travels as a kernel parameter.)
int upload_plan(b200_ctx* ctx, const Plan& plan, const DevOp** d_ops, const double** d_mat2) {
    const size_t b_ops = plan.ops.size() * sizeof(DevOp);
    const size_t b_mat2 = plan.mat2.size() * sizeof(double);
    auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t o_mat2 = up16(b_ops), total = o_mat2 + up16(b_mat2) + 16;
    // The staging buffer may still be in flight from the previous call on this stream.
    if (ctx->plan_in_flight) { CUDA_TRY(cudaEventSynchronize(ctx->ev_plan)); ctx->plan_in_flight = false; }
    if (reserve_host(&ctx->h_plan, &ctx->h_plan_cap, total)) return -1;
    if (ctx->d_plan_cap < total) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // old buffer may be in use by a running kernel
        if (reserve_dev(&ctx->d_plan, &ctx->d_plan_cap, total)) return -1;
    }
    char* h = (char*)ctx->h_plan;
    if (b_ops) std::memcpy(h, plan.ops.data(), b_ops);
    if (b_mat2) std::memcpy(h + o_mat2, plan.mat2.data(), b_mat2);
    CUDA_TRY(cudaMemcpyAsync(ctx->d_plan, h, total, cudaMemcpyHostToDevice, ctx->stream));
    ctx->counters[4] += total;
    CUDA_TRY(cudaEventRecord(ctx->ev_plan, ctx->stream));
    ctx->plan_in_flight = true;
    *d_ops = (const DevOp*)ctx->d_plan;
    *d_mat2 = (const double*)((char*)ctx->d_plan + o_mat2);
    return 0;
}

const EmbedSrc kNoEmbed = {};

int run_plan(b200_ctx* ctx, int dst_slot, int src_slot, const Plan& plan, const EmbedSrc* es = nullptr) {
    const int n = ctx->nq;
    const uint64_t dim = 1ull << n;
    double2* dst = (double2*)ctx->slots[dst_slot];
    const double2* src = src_slot >= 0 ? (const double2*)ctx->slots[src_slot] : nullptr;
    Timer tm(ctx);

    if (plan.small) {
        const DevOp* d_ops; const double* d_mat2;
        if (upload_plan(ctx, plan, &d_ops, &d_mat2)) return -1;
        const size_t smem = dim * sizeof(double2);
        {
            KScope ks(ctx, B200_PROF_SMALL);
            sv_small_kernel<<<1, 256, smem, ctx->stream>>>(src, dst, n, d_ops, (int)plan.ops.size(), d_mat2,
                                                            src == nullptr ? 1 : 0);
        }
        CUDA_TRY(cudaGetLastError());
        ctx->counters[1] += 1;
        ctx->counters[2] += plan.n_gates_in; ctx->counters[3] += 32 * dim;
        tm.stop();
        return 0;
    }

    if (src == nullptr && (plan.sweeps.empty() || ctx->sweep_mode == 1)) {
        // (the direct sweep kernel takes |0..0> as an implicit source: no fill pass, no read pass)
        {
            KScope ks(ctx, B200_PROF_FILL);
            sv_init_zero_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(dst, dim);
        }
        CUDA_TRY(cudaGetLastError());
        src = dst;
    }
    if (plan.sweeps.empty()) {
        if (src != dst) {
            KScope ks(ctx, B200_PROF_FILL);
            CUDA_TRY(cudaMemcpyAsync(dst, src, dim * sizeof(double2), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        tm.stop();
        return 0;
    }
    const uint32_t ntiles = (uint32_t)(dim >> TILE_BITS);
    for (const SweepProg& sw : plan.sweeps) {
        const int nr = sw.nrounds;
        if (ctx->sweep_mode == 1) {
            const uint32_t grid = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)ctx->num_sms);
            KScope ks(ctx, B200_PROF_SWEEP);
            sv_sweep_pipe_kernel<REG_BITS><<<grid, PIPE_THREADS, PIPE_STAGES * TILE_BYTES, ctx->stream>>>(src, dst, sw, ntiles);
        } else {
            const size_t smem = nr > 1 ? ((size_t)1 << TILE_BITS) * sizeof(double2) : 0;
            const int per_sm = nr > 1 ? ctx->sweep_occ_smem : ctx->sweep_occ_nosmem;
            const uint32_t grid = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)ctx->num_sms * per_sm * ctx->grid_mult);
            // a sweep from the implicit |0..0> only writes (16 * 2^n bytes): accounted with the fills, so
            // that the SWEEP class holds read+write passes only (roofline accounting, bench.py)
            KScope ks(ctx, src == nullptr ? B200_PROF_FILL : B200_PROF_SWEEP);
            if (es != nullptr && src == nullptr) {
                EmbedSrc e1 = *es;
                embed_prepare(e1, sw);
                sv_sweep_kernel<REG_BITS, true><<<grid, SWEEP_THREADS, smem, ctx->stream>>>(src, dst, sw, ntiles, e1);
            }
            else
                sv_sweep_kernel<REG_BITS><<<grid, SWEEP_THREADS, smem, ctx->stream>>>(src, dst, sw, ntiles, kNoEmbed);
        }
        CUDA_TRY(cudaGetLastError());
        ctx->counters[1] += 1; ctx->counters[3] += (src == nullptr ? 16 : 32) * dim;
        src = dst;
        // host -> device bytes of the program that rode along as the kernel parameter
        ctx->counters[4] += offsetof(SweepProg, ops) + (size_t)sw.nops * sizeof(POp) + (size_t)sw.nmat2 * 32 * sizeof(double);
    }
    ctx->counters[2] += plan.n_gates_in;
    tm.stop();
    return 0;
}

int make_plan(int nq, const b200_gate* gates, int n_gates, const double* mats, int n_mats, bool inverse,
              Plan& plan, bool fold_perm = true) {
    if (n_gates < 0 || (n_gates > 0 && !gates)) return set_error("null gate array");
    std::vector<COp> ops;
    const std::string err = canonicalize(nq, gates, n_gates, mats, n_mats, inverse, ops);
    if (!err.empty()) return set_error(err);
    fuse_single_qubit_runs(ops);
    fuse_diagonals(ops);
    build_plan(nq, ops, plan, fold_perm);
    plan.n_gates_in = n_gates;
    return 0;
}

int sv_run_impl(b200_ctx* ctx, int dst_slot, int src_slot, const b200_gate* gates, int n_gates,
                const double* mats, int n_mats, bool inverse) {
    if (check_slot(ctx, dst_slot)) return -1;
    if (src_slot >= 0 && check_slot(ctx, src_slot)) return -1;
    CUDA_TRY(cudaSetDevice(ctx->device));
    Plan plan;
    // the pipelined (bulk-copy) variant lands / picks up tiles in a linear layout and keeps executing X-type ops
    if (make_plan(ctx->nq, gates, n_gates, mats, n_mats, inverse, plan, ctx->sweep_mode == 0)) return -1;
    ctx->counters[6] += 1;
    return run_plan(ctx, dst_slot, src_slot, plan);
}

// copy `count` doubles of d_out to the caller (sync)
int fetch_out(b200_ctx* ctx, double* out, int count) {
    CUDA_TRY(cudaMemcpyAsync(ctx->h_out, ctx->d_out, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    std::memcpy(out, ctx->h_out, count * sizeof(double));
    ctx->counters[5] += count * sizeof(double);
    return 0;
}

int red_grid(b200_ctx* ctx, uint64_t work_items) {
    const uint64_t want = (work_items + RED_THREADS - 1) / RED_THREADS;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)ctx->num_sms * 8));
}

}  // namespace

extern "C" {

int b200_abi_version(void) { return B200AQC_ABI_VERSION; }

const char* b200_last_error(void) { return g_last_error.c_str(); }

int b200_device_count(int* count) {
    if (!count) return set_error("null pointer");
    CUDA_TRY(cudaGetDeviceCount(count));
    return 0;
}

int b200_ctx_create(int device, b200_ctx** out) {
    if (!out) return set_error("null pointer");
    *out = nullptr;
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count)
        return set_error("device " + std::to_string(device) + " not present (" + std::to_string(count) + " visible)");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error(std::string("libb200aqc is built for sm_100a only; device is ") + prop.name + " (sm_" +
                         std::to_string(prop.major) + std::to_string(prop.minor) + ")");
    b200_ctx* ctx = new b200_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->coop_ok = prop.cooperativeLaunch != 0 && !std::getenv("B200AQC_NO_COOP");
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&ctx->ev0));
    CUDA_TRY(cudaEventCreate(&ctx->ev1));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_plan, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreate(&ctx->ev_mark[0]));
    CUDA_TRY(cudaEventCreate(&ctx->ev_mark[1]));
    const size_t tile_bytes = ((size_t)1 << TILE_BITS) * sizeof(double2);
    CUDA_TRY(cudaFuncSetAttribute(sv_sweep_kernel<REG_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tile_bytes));
    CUDA_TRY(cudaFuncSetAttribute(sv_sweep_pipe_kernel<REG_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(PIPE_STAGES * TILE_BYTES)));
    if (const char* e = std::getenv("B200AQC_SWEEP")) ctx->sweep_mode = std::strcmp(e, "pipe") == 0 ? 1 : 0;
    CUDA_TRY(cudaFuncSetAttribute(sv_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(((size_t)1 << SMALL_MAX_QUBITS) * sizeof(double2))));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->sweep_occ_smem, sv_sweep_kernel<REG_BITS>,
                                                           SWEEP_THREADS, tile_bytes));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->sweep_occ_nosmem, sv_sweep_kernel<REG_BITS>,
                                                           SWEEP_THREADS, 0));
    if (ctx->sweep_occ_smem < 1 || ctx->sweep_occ_nosmem < 1) { delete ctx; return set_error("sweep kernel does not fit on an SM"); }
    CUDA_TRY(cudaFuncSetAttribute(sv_sweep_inner2_kernel<REG_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)FUSED_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(sv_sweep_inner2_kernel<REG_BITS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)FUSED_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(sv_sweep_kernel<REG_BITS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tile_bytes));
    CUDA_TRY(cudaFuncSetAttribute(sv_sweep_project_kernel<REG_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tile_bytes));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->fused_occ, sv_sweep_inner2_kernel<REG_BITS>,
                                                           SWEEP_THREADS, FUSED_SMEM_BYTES));
    if (ctx->fused_occ < 1) { delete ctx; return set_error("fused sweep kernel does not fit on an SM"); }
    if (const char* e = std::getenv("B200AQC_GRID_MULT")) ctx->grid_mult = std::max(1, std::atoi(e));
    if (ensure_scratch(ctx)) { delete ctx; return -1; }
    *out = ctx;
    return 0;
}

int b200_ctx_destroy(b200_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (size_t i = 0; i < ctx->slots.size(); ++i)
        if (ctx->slots[i] && ctx->owned[i]) cudaFree(ctx->slots[i]);
    if (ctx->d_partial) cudaFree(ctx->d_partial);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->d_plan) cudaFree(ctx->d_plan);
    if (ctx->h_plan) cudaFreeHost(ctx->h_plan);
    b200_mps_release(ctx);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->ev_plan);
    cudaEventDestroy(ctx->ev_mark[0]);
    cudaEventDestroy(ctx->ev_mark[1]);
    for (auto& r : ctx->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ctx->prof_pool) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

int b200_ctx_sync(b200_ctx* ctx) {
    if (!ctx) return set_error("null context");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200_ctx_counters(b200_ctx* ctx, uint64_t out[8]) {
    if (!ctx || !out) return set_error("null pointer");
    for (int k = 0; k < 8; ++k) out[k] = ctx->counters[k];
    return 0;
}

int b200_ctx_mark(b200_ctx* ctx, int which) {
    if (!ctx) return set_error("null context");
    if (which < 0 || which > 1) return set_error("mark index must be 0 or 1");
    CUDA_TRY(cudaEventRecord(ctx->ev_mark[which], ctx->stream));
    return 0;
}

int b200_ctx_elapsed_ms(b200_ctx* ctx, double* ms) {
    if (!ctx || !ms) return set_error("null pointer");
    CUDA_TRY(cudaEventSynchronize(ctx->ev_mark[1]));
    float f = 0;
    CUDA_TRY(cudaEventElapsedTime(&f, ctx->ev_mark[0], ctx->ev_mark[1]));
    *ms = f;
    return 0;
}

int b200_ctx_profile(b200_ctx* ctx, int enable) {
    if (!ctx) return set_error("null context");
    if (prof_drain(ctx)) return -1;
    ctx->profiling = enable != 0;
    if (enable)
        for (int k = 0; k < B200_PROF_CLASSES; ++k) { ctx->prof_ms[k] = 0; ctx->prof_n[k] = 0; }
        ctx->sweep_log.clear();
    return 0;
}

int b200_ctx_profile_read(b200_ctx* ctx, double ms[B200_PROF_CLASSES], uint64_t launches[B200_PROF_CLASSES]) {
    if (!ctx || !ms || !launches) return set_error("null pointer");
    if (prof_drain(ctx)) return -1;
    for (int k = 0; k < B200_PROF_CLASSES; ++k) { ms[k] = ctx->prof_ms[k]; launches[k] = ctx->prof_n[k]; }
    return 0;
}

int b200_ctx_profile_sweeps(b200_ctx* ctx, double* ms, int max_entries, int* n) {
    if (!ctx || !n || (max_entries > 0 && !ms)) return set_error("null pointer");
    if (prof_drain(ctx)) return -1;
    *n = (int)ctx->sweep_log.size();
    for (int k = 0; k < *n && k < max_entries; ++k) ms[k] = ctx->sweep_log[k];
    return 0;
}

int b200_ctx_set_timing(b200_ctx* ctx, int enable) {
    if (!ctx) return set_error("null context");
    ctx->timing = enable != 0;
    ctx->timing_pending = false;
    return 0;
}

int b200_ctx_last_ms(b200_ctx* ctx, double* ms) {
    if (!ctx || !ms) return set_error("null pointer");
    if (!ctx->timing || !ctx->timing_pending) return set_error("no timed call pending (b200_ctx_set_timing)");
    CUDA_TRY(cudaEventSynchronize(ctx->ev1));
    float f = 0;
    CUDA_TRY(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
    *ms = f;
    return 0;
}

int b200_sv_alloc(b200_ctx* ctx, int num_qubits, int n_slots) {
    if (!ctx) return set_error("null context");
    if (num_qubits < 1 || num_qubits > 40) return set_error("num_qubits out of range [1,40]");
    if (n_slots < 1 || n_slots > 64) return set_error("n_slots out of range [1,64]");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < ctx->slots.size(); ++i)
        if (ctx->slots[i] && ctx->owned[i]) cudaFree(ctx->slots[i]);
    ctx->slots.assign(n_slots, nullptr);
    ctx->owned.assign(n_slots, 0);
    ctx->nq = 0;
    const size_t bytes = ((size_t)1 << num_qubits) * sizeof(double2);
    for (int i = 0; i < n_slots; ++i) {
        cudaError_t e = cudaMalloc(&ctx->slots[i], bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int j = 0; j < i; ++j) { cudaFree(ctx->slots[j]); ctx->slots[j] = nullptr; }
            return set_error("cudaMalloc of " + std::to_string(bytes) + " bytes for slot " + std::to_string(i) +
                             " failed: " + cudaGetErrorString(e));
        }
        ctx->owned[i] = 1;
    }
    ctx->nq = num_qubits;
    return 0;
}

int b200_sv_reserve_slots(b200_ctx* ctx, int num_qubits, int n_slots) {
    if (!ctx) return set_error("null context");
    if (num_qubits < 1 || num_qubits > 40) return set_error("num_qubits out of range [1,40]");
    if (n_slots < 1 || n_slots > 64) return set_error("n_slots out of range [1,64]");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < ctx->slots.size(); ++i)
        if (ctx->slots[i] && ctx->owned[i]) cudaFree(ctx->slots[i]);
    ctx->slots.assign(n_slots, nullptr);
    ctx->owned.assign(n_slots, 0);
    ctx->nq = num_qubits;
    return 0;
}

int b200_sv_attach(b200_ctx* ctx, int slot, void* device_ptr) {
    if (!ctx) return set_error("null context");
    if (ctx->nq <= 0) return set_error("statevector not allocated (call b200_sv_alloc)");
    if (slot < 0 || slot >= (int)ctx->slots.size()) return set_error("invalid slot");
    if (!device_ptr || ((uintptr_t)device_ptr & 15)) return set_error("device pointer must be 16-byte aligned");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->slots[slot] && ctx->owned[slot]) cudaFree(ctx->slots[slot]);
    ctx->slots[slot] = device_ptr;
    ctx->owned[slot] = 0;
    return 0;
}

int b200_sv_device_ptr(b200_ctx* ctx, int slot, void** out) {
    if (check_slot(ctx, slot)) return -1;
    *out = ctx->slots[slot];
    return 0;
}

int b200_sv_ipc_export(b200_ctx* ctx, int slot, unsigned char handle[64]) {
    if (check_slot(ctx, slot)) return -1;
    if (!handle) return set_error("null pointer");
    if (!ctx->owned[slot]) return set_error("ipc_export: slot memory is caller-owned (attach): export it from its owner");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, ctx->slots[slot]));
    std::memcpy(handle, &h, 64);
    return 0;
}

int b200_sv_ipc_open(b200_ctx* ctx, const unsigned char handle[64], void** peer_ptr) {
    if (!ctx || !handle || !peer_ptr) return set_error("null pointer");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int b200_sv_ipc_close(b200_ctx* ctx, void* peer_ptr) {
    if (!ctx) return set_error("null context");
    if (!peer_ptr) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}

int b200_sv_peer_swap(b200_ctx* ctx, int slot, void* const* peer_ptrs, int world, int rank) {
    if (check_slot(ctx, slot)) return -1;
    if (!peer_ptrs) return set_error("null pointer");
    if (world < 2 || world > PEER_MAX_WORLD || (world & (world - 1))) return set_error("peer_swap: world must be a power of two in [2,16]");
    if (rank < 0 || rank >= world) return set_error("peer_swap: rank out of range");
    const uint64_t dim = 1ull << ctx->nq;
    if (dim / world < 2) return set_error("peer_swap: slice too small for this world size");
    CUDA_TRY(cudaSetDevice(ctx->device));
    PeerTable pt;
    for (int p = 0; p < PEER_MAX_WORLD; ++p) pt.p[p] = nullptr;
    for (int p = 0; p < world; ++p) {
        if (p != rank && !peer_ptrs[p]) return set_error("peer_swap: missing peer pointer");
        pt.p[p] = (double2*)peer_ptrs[p];
    }
    const uint64_t chunk = dim / world;
    Timer tm(ctx);
    {
        KScope ks(ctx, B200_PROF_FILL);
        sv_peer_swap_kernel<<<ctx->num_sms * 4, 512, 0, ctx->stream>>>((double2*)ctx->slots[slot], pt, world, rank, chunk);
    }
    CUDA_TRY(cudaGetLastError());
    ctx->counters[6] += 1;
    tm.stop();
    return 0;
}

int b200_sv_peer_swap_strided(b200_ctx* ctx, int slot, void* const* peer_ptrs, int world, int rank, const int32_t* positions) {
    if (check_slot(ctx, slot)) return -1;
    if (!peer_ptrs || !positions) return set_error("null pointer");
    if (world < 2 || world > PEER_MAX_WORLD || (world & (world - 1))) return set_error("peer_swap: world must be a power of two in [2,16]");
    if (rank < 0 || rank >= world) return set_error("peer_swap: rank out of range");
    int g = 0;
    while ((1 << g) < world) ++g;
    if (ctx->nq - g < 1) return set_error("peer_swap: slice too small for this world size");
    PeerStride ps;
    ps.g = g;
    for (int j = 0; j < g; ++j) {
        if (positions[j] < 0 || positions[j] >= ctx->nq) return set_error("peer_swap_strided: position out of range");
        for (int k = 0; k < j; ++k)
            if (positions[k] == positions[j]) return set_error("peer_swap_strided: positions must differ");
        ps.pos_sorted[j] = positions[j];
    }
    std::sort(ps.pos_sorted, ps.pos_sorted + g);
    for (int r = 0; r < PEER_MAX_WORLD; ++r) {
        ps.pattern[r] = 0;
        for (int j = 0; j < g; ++j) ps.pattern[r] |= (uint64_t)((r >> j) & 1) << positions[j];
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    PeerTable pt;
    for (int p = 0; p < PEER_MAX_WORLD; ++p) pt.p[p] = nullptr;
    for (int p = 0; p < world; ++p) {
        if (p != rank && !peer_ptrs[p]) return set_error("peer_swap: missing peer pointer");
        pt.p[p] = (double2*)peer_ptrs[p];
    }
    const uint64_t sub = (1ull << ctx->nq) >> g;
    Timer tm(ctx);
    {
        KScope ks(ctx, B200_PROF_FILL);
        sv_peer_swap_strided_kernel<<<ctx->num_sms * 4, 512, 0, ctx->stream>>>((double2*)ctx->slots[slot], pt, world, rank, sub, ps);
    }
    CUDA_TRY(cudaGetLastError());
    ctx->counters[6] += 1;
    tm.stop();
    return 0;
}

int b200_sv_num_qubits(b200_ctx* ctx, int* out) {
    if (!ctx || !out) return set_error("null pointer");
    *out = ctx->nq;
    return 0;
}

int b200_sv_init_zero(b200_ctx* ctx, int slot) {
    if (check_slot(ctx, slot)) return -1;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t dim = 1ull << ctx->nq;
    const int grid = (int)std::min<uint64_t>((dim + 255) / 256, (uint64_t)ctx->num_sms * 8);
    {
        KScope ks(ctx, B200_PROF_FILL);
        sv_init_zero_kernel<<<grid, 256, 0, ctx->stream>>>((double2*)ctx->slots[slot], dim);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int b200_sv_copy(b200_ctx* ctx, int dst_slot, int src_slot) {
    if (check_slot(ctx, dst_slot) || check_slot(ctx, src_slot)) return -1;
    if (dst_slot == src_slot) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    KScope ks(ctx, B200_PROF_FILL);
    CUDA_TRY(cudaMemcpyAsync(ctx->slots[dst_slot], ctx->slots[src_slot], ((size_t)1 << ctx->nq) * sizeof(double2),
                             cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

int b200_sv_run(b200_ctx* ctx, int dst_slot, int src_slot, const b200_gate* gates, int n_gates,
                const double* mats, int n_mats) {
    return sv_run_impl(ctx, dst_slot, src_slot, gates, n_gates, mats, n_mats, false);
}

int b200_sv_run_inverse(b200_ctx* ctx, int dst_slot, int src_slot, const b200_gate* gates, int n_gates,
                        const double* mats, int n_mats) {
    return sv_run_impl(ctx, dst_slot, src_slot, gates, n_gates, mats, n_mats, true);
}

int b200_sv_amp(b200_ctx* ctx, int slot, uint64_t index, double out[2]) {
    if (check_slot(ctx, slot)) return -1;
    if (index >> ctx->nq) return set_error("amplitude index out of range");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpyAsync(ctx->h_out, (const double2*)ctx->slots[slot] + index, sizeof(double2),
                             cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    out[0] = ctx->h_out[0];
    out[1] = ctx->h_out[1];
    ctx->counters[5] += sizeof(double2);
    ctx->counters[6] += 1;
    return 0;
}

int b200_sv_expz(b200_ctx* ctx, int slot, double* out) {
    if (check_slot(ctx, slot)) return -1;
    if (!out) return set_error("null pointer");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int n = ctx->nq;
    const double2* psi = (const double2*)ctx->slots[slot];
    Timer tm(ctx);
    if (n <= SMALL_MAX_QUBITS) {
        {
            KScope ks(ctx, B200_PROF_EXPZ);
            sv_expz_small_kernel<<<1, 64, 0, ctx->stream>>>(psi, n, ctx->d_out);
        }
        CUDA_TRY(cudaGetLastError());
    } else {
        const int tb = std::min(n - 3, 17), kb = n - tb;
        const int nblocks = 1 << (tb - 8);
        {
            KScope ks(ctx, B200_PROF_EXPZ);
            sv_expz_kernel<<<nblocks, RED_THREADS, 0, ctx->stream>>>(psi, tb, kb, ctx->d_partial);
        }
        CUDA_TRY(cudaGetLastError());
        {
            KScope ks(ctx, B200_PROF_REDUCE);
            sv_expz_final_kernel<<<n + 1, RED_THREADS, 0, ctx->stream>>>(ctx->d_partial, nblocks, n, tb, ctx->d_out);
        }
        CUDA_TRY(cudaGetLastError());
    }
    ctx->counters[3] += 16ull << n;
    ctx->counters[6] += 1;
    tm.stop();
    return fetch_out(ctx, out, n + 1);
}

int b200_sv_pair_rdm(b200_ctx* ctx, int slot, const int32_t* pairs, int n_pairs, double* out) {
    return b200_sv_pair_rdm_part(ctx, slot, pairs, n_pairs, 0, 1, out);
}

// The same pass list as b200_sv_pair_rdm (it depends only on the requested pairs), of which only the passes
// k = part (mod n_parts) are launched: a pair owned by a pass of another part comes back as zeros, so that the SUM over
// the parts (one all-reduce of 32 * n_pairs doubles between ranks holding replicas of the state) is, bit for bit, the
// result of the undivided call.
int b200_sv_pair_rdm_part(b200_ctx* ctx, int slot, const int32_t* pairs, int n_pairs, int part, int n_parts, double* out) {
    if (check_slot(ctx, slot)) return -1;
    if (n_pairs < 0 || (n_pairs > 0 && (!pairs || !out))) return set_error("null pointer");
    if (n_parts < 1 || part < 0 || part >= n_parts) return set_error("pair_rdm_part: part out of range");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int n = ctx->nq;
    const double2* psi = (const double2*)ctx->slots[slot];
    for (int p = 0; p < n_pairs; ++p) {
        const int a = pairs[2 * p], b = pairs[2 * p + 1];
        if (a < 0 || b < 0 || a >= n || b >= n || a == b)
            return set_error("pair " + std::to_string(p) + ": qubits out of range");
    }
    // Cover the requested pairs with qubit quadruples (six pair-RDMs per read pass), triples (three) or
    // single pairs, greedily by newly covered pairs per pass.  Results of many passes are staged in d_out
    // and fetched together (one synchronisation per ~40 passes instead of one per pass).
    std::vector<char> want(n * n, 0), have(n * n, 0);
    std::vector<double> acc((size_t)n * n * 16, 0.0);  // indexed [lo*n+hi][16]
    for (int p = 0; p < n_pairs; ++p) {
        const int lo = std::min(pairs[2 * p], pairs[2 * p + 1]), hi = std::max(pairs[2 * p], pairs[2 * p + 1]);
        want[lo * n + hi] = 1;
    }
    auto need = [&](int u, int v) { const int l = std::min(u, v), h = std::max(u, v); return want[l * n + h] && !have[l * n + h] ? 1 : 0; };
    struct Pending { int off, npairs; int pl[6][2]; };
    std::vector<Pending> pending;
    int out_off = 0;
    int pass_index = 0;
    auto flush = [&]() -> int {
        if (pending.empty()) return 0;
        std::vector<double> r(out_off);
        if (fetch_out(ctx, r.data(), out_off)) return -1;
        for (const Pending& pd : pending)
            for (int t = 0; t < pd.npairs; ++t)
                if (pd.pl[t][0] >= 0)
                    std::memcpy(&acc[(size_t)(pd.pl[t][0] * n + pd.pl[t][1]) * 16], r.data() + pd.off + 16 * t, 16 * sizeof(double));
        pending.clear();
        out_off = 0;
        return 0;
    };
    Timer tm(ctx);
    for (int lo = 0; lo < n; ++lo)
        for (int hi = lo + 1; hi < n; ++hi) {
            if (!want[lo * n + hi] || have[lo * n + hi]) continue;
            int c3 = -1, g3 = 1, c4a = -1, c4b = -1, g4 = 1;
            for (int c = 0; c < n && n >= 3; ++c) {
                if (c == lo || c == hi) continue;
                const int gc = 1 + need(lo, c) + need(hi, c);
                if (gc > g3) { g3 = gc; c3 = c; }
                for (int d = c + 1; d < n && n >= 4; ++d) {
                    if (d == lo || d == hi) continue;
                    const int gd = gc + need(lo, d) + need(hi, d) + need(c, d);
                    if (gd > g4) { g4 = gd; c4a = c; c4b = d; }
                }
            }
            int q[4] = {lo, hi, -1, -1}, nq = 2;
            if (g4 > g3 && g4 >= 3) { q[2] = c4a; q[3] = c4b; nq = 4; }      // a quad pass costs ~1.2 triple passes
            else if (g3 >= 2) { q[2] = c3; nq = 3; }
            std::sort(q, q + nq);
            // a pair belongs to the FIRST pass that covers it (a later quadruple may contain it again)
            Pending pd;
            pd.npairs = 0;
            for (int i = 0; i < nq; ++i)
                for (int j = i + 1; j < nq; ++j) {
                    pd.pl[pd.npairs][0] = have[q[i] * n + q[j]] ? -1 : q[i];
                    pd.pl[pd.npairs][1] = q[j];
                    have[q[i] * n + q[j]] = 1;
                    ++pd.npairs;
                }
            if ((pass_index++ % n_parts) != part) continue;       // another rank's pass
            const int width = nq == 4 ? RDM4_WIDTH : (nq == 3 ? RDM3_WIDTH : 16);
            if (out_off + width > (int)OUT_DOUBLES && flush()) return -1;
            const int grid = red_grid(ctx, 1ull << (n - nq));
            {
                KScope ks(ctx, B200_PROF_RDM);
                if (nq == 4) sv_rdm4_kernel<<<grid, RED_THREADS, 0, ctx->stream>>>(psi, n, q[0], q[1], q[2], q[3], ctx->d_partial);
                else if (nq == 3) sv_rdm3_kernel<<<grid, RED_THREADS, 0, ctx->stream>>>(psi, n, q[0], q[1], q[2], ctx->d_partial);
                else sv_rdm2_kernel<<<grid, RED_THREADS, 0, ctx->stream>>>(psi, n, q[0], q[1], ctx->d_partial);
            }
            CUDA_TRY(cudaGetLastError());
            {
                KScope ks(ctx, B200_PROF_REDUCE);
                reduce_partials_kernel<<<width, 32, 0, ctx->stream>>>(ctx->d_partial, grid, width, ctx->d_out + out_off);
            }
            CUDA_TRY(cudaGetLastError());
            ctx->counters[3] += 16ull << n;
            pd.off = out_off;
            pending.push_back(pd);
            out_off += width;
        }
    if (flush()) return -1;
    tm.stop();
    // expand the packed Hermitian accumulators into row-major 4x4 complex
    static const int off_r[6] = {1, 2, 3, 2, 3, 3}, off_c[6] = {0, 0, 0, 1, 1, 2};
    for (int p = 0; p < n_pairs; ++p) {
        const int lo = std::min(pairs[2 * p], pairs[2 * p + 1]), hi = std::max(pairs[2 * p], pairs[2 * p + 1]);
        const double* a = &acc[(size_t)(lo * n + hi) * 16];
        double* o = out + (size_t)p * 32;
        for (int k = 0; k < 32; ++k) o[k] = 0.0;
        for (int d = 0; d < 4; ++d) o[2 * (5 * d)] = a[d];
        for (int k = 0; k < 6; ++k) {
            const int r = off_r[k], c = off_c[k];
            o[2 * (4 * r + c)] = a[4 + 2 * k];
            o[2 * (4 * r + c) + 1] = a[5 + 2 * k];
            o[2 * (4 * c + r)] = a[4 + 2 * k];
            o[2 * (4 * c + r) + 1] = -a[5 + 2 * k];
        }
    }
    return 0;
}

int b200_sv_inner(b200_ctx* ctx, int l_slot, int r_slot, int q, double out[8]) {
    if (check_slot(ctx, l_slot) || check_slot(ctx, r_slot)) return -1;
    if (q >= ctx->nq) return set_error("qubit out of range");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int n = ctx->nq;
    Timer tm(ctx);
    const int grid = red_grid(ctx, q < 0 ? (1ull << n) : (1ull << (n - 1)));
    {
        KScope ks(ctx, B200_PROF_INNER);
        sv_inner_kernel<<<grid, RED_THREADS, 0, ctx->stream>>>((const double2*)ctx->slots[l_slot],
                                                               (const double2*)ctx->slots[r_slot], n, q, ctx->d_partial);
    }
    CUDA_TRY(cudaGetLastError());
    {
        KScope ks(ctx, B200_PROF_REDUCE);
        reduce_partials_kernel<<<INNER_WIDTH, 32, 0, ctx->stream>>>(ctx->d_partial, grid, INNER_WIDTH, ctx->d_out);
    }
    CUDA_TRY(cudaGetLastError());
    ctx->counters[3] += 32ull << n;
    ctx->counters[6] += 1;
    tm.stop();
    return fetch_out(ctx, out, 8);
}

int b200_sv_inner2(b200_ctx* ctx, int l_slot, int r_slot, int qa, int qb, double out[32]) {
    if (check_slot(ctx, l_slot) || check_slot(ctx, r_slot)) return -1;
    const int n = ctx->nq;
    if (qa < 0 || qb < 0 || qa >= n || qb >= n || qa == qb) return set_error("inner2: qubits out of range");
    if (!out) return set_error("null pointer");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int lo = std::min(qa, qb), hi = std::max(qa, qb);
    Timer tm(ctx);
    const int grid = red_grid(ctx, 1ull << (n - 2));
    {
        KScope ks(ctx, B200_PROF_INNER);
        sv_inner2_kernel<<<grid, RED_THREADS, 0, ctx->stream>>>((const double2*)ctx->slots[l_slot],
                                                                (const double2*)ctx->slots[r_slot], n, lo, hi, ctx->d_partial);
    }
    CUDA_TRY(cudaGetLastError());
    {
        KScope ks(ctx, B200_PROF_REDUCE);
        reduce_partials_kernel<<<INNER2_WIDTH, 32, 0, ctx->stream>>>(ctx->d_partial, grid, INNER2_WIDTH, ctx->d_out);
    }
    CUDA_TRY(cudaGetLastError());
    ctx->counters[3] += 32ull << n;
    ctx->counters[6] += 1;
    tm.stop();
    double t[32];
    if (fetch_out(ctx, t, 32)) return -1;
    if (qa < qb) {
        std::memcpy(out, t, sizeof t);
    } else {  // caller's index = bit(qa) + 2 bit(qb) with qa the higher qubit: swap the two index bits
        auto sw2 = [](int i) { return ((i & 1) << 1) | (i >> 1); };
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                out[2 * (4 * sw2(i) + sw2(j))] = t[2 * (4 * i + j)];
                out[2 * (4 * sw2(i) + sw2(j)) + 1] = t[2 * (4 * i + j) + 1];
            }
    }
    return 0;
}

// shared body of b200_sv_run_inner2 / b200_sv_run_embedded_inner2 (es != nullptr: the source is the embedded state)
static int run_inner2_impl(b200_ctx* ctx, int dst_slot, int src_slot, const EmbedSrc* es, const b200_gate* gates, int n_gates,
                           const double* mats, int n_mats, int inverse, int other_slot, int qa, int qb, double out[32],
                           int* stored = nullptr) {
    const int n = ctx->nq;
    if (qa < 0 || qb < 0 || qa >= n || qb >= n || qa == qb) return set_error("run_inner2: qubits out of range");
    if (!out) return set_error("null pointer");
    if (other_slot == dst_slot) return set_error("run_inner2: `other` must not be the destination");
    if (n <= SMALL_MAX_QUBITS) return set_error("run_inner2: register too small for the tiled path (use b200_sv_run + b200_sv_inner2)");
    if (n_gates < 0 || (n_gates > 0 && !gates)) return set_error("null gate array");
    CUDA_TRY(cudaSetDevice(ctx->device));
    Plan plan;
    {
        std::vector<COp> ops;
        const std::string err = canonicalize(n, gates, n_gates, mats, n_mats, inverse != 0, ops);
        if (!err.empty()) return set_error(err);
        fuse_single_qubit_runs(ops);
        fuse_diagonals(ops);
        build_plan(n, ops, plan, true, qa, qb, es != nullptr ? es->outside : 0);
        plan.n_gates_in = n_gates;
    }
    EpiProg ep;
    if (plan.sweeps.empty() || !make_epilogue(plan.sweeps.back(), qa, qb, ep)) return set_error("run_inner2: planner did not place the pair in the last tile");
    ctx->counters[6] += 2;
    const uint64_t dim = 1ull << n;
    const uint32_t ntiles = (uint32_t)(dim >> TILE_BITS);
    const double2* src = (es != nullptr || src_slot < 0) ? nullptr : (const double2*)ctx->slots[src_slot];
    const bool from_zero = es == nullptr && src_slot < 0;
    double2* dst = (double2*)ctx->slots[dst_slot];
    // *stored == 0 on entry: the caller only wants T -- the swept state is not written when ONE sweep does the whole
    // program (32 bytes per amplitude instead of 48); a longer program needs dst for its intermediate state anyway
    const bool keep = stored != nullptr && *stored == 0 && plan.sweeps.size() == 1;
    if (stored != nullptr) *stored = keep ? 0 : 1;
    const size_t tile_bytes = ((size_t)1 << TILE_BITS) * sizeof(double2);
    Timer tm(ctx);
    uint32_t grid = 1;
    for (size_t k = 0; k < plan.sweeps.size(); ++k) {
        const SweepProg& sw = plan.sweeps[k];
        const bool embed = es != nullptr && k == 0;
        EmbedSrc e1 = embed ? *es : kNoEmbed;
        if (embed) embed_prepare(e1, sw);
        if (k + 1 < plan.sweeps.size()) {
            const size_t smem = sw.nrounds > 1 ? tile_bytes : 0;
            const int per_sm = sw.nrounds > 1 ? ctx->sweep_occ_smem : ctx->sweep_occ_nosmem;
            const uint32_t g = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)ctx->num_sms * per_sm * ctx->grid_mult);
            const bool no_read = embed || (from_zero && k == 0);
            KScope ks(ctx, no_read ? B200_PROF_FILL : B200_PROF_SWEEP);
            if (embed) sv_sweep_kernel<REG_BITS, true><<<g, SWEEP_THREADS, smem, ctx->stream>>>(src, dst, sw, ntiles, e1);
            else sv_sweep_kernel<REG_BITS><<<g, SWEEP_THREADS, smem, ctx->stream>>>(src, dst, sw, ntiles, kNoEmbed);
            ctx->counters[3] += (no_read ? 16 : 32) * dim;
        } else {
            grid = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)ctx->num_sms * ctx->fused_occ);
            // from an embedded source the pass reads `other` and writes dst only (32 B per amplitude): its own class, so
            // that the FUSED class holds 48-byte passes only (roofline accounting, bench.py)
            const bool no_read = embed || (from_zero && k == 0);
            KScope ks(ctx, no_read ? (keep ? B200_PROF_FUSED_EMBED_READ : B200_PROF_FUSED_EMBED)
                                   : (keep ? B200_PROF_FUSED_READ : B200_PROF_FUSED));
            if (embed)
                sv_sweep_inner2_kernel<REG_BITS, true><<<grid, SWEEP_THREADS, FUSED_SMEM_BYTES, ctx->stream>>>(
                    src, keep ? nullptr : dst, (const double2*)ctx->slots[other_slot], sw, ep, ntiles, ctx->d_partial, e1);
            else
                sv_sweep_inner2_kernel<REG_BITS><<<grid, SWEEP_THREADS, FUSED_SMEM_BYTES, ctx->stream>>>(
                    src, keep ? nullptr : dst, (const double2*)ctx->slots[other_slot], sw, ep, ntiles, ctx->d_partial, kNoEmbed);
            ctx->counters[3] += ((no_read && keep) ? 16 : ((no_read || keep) ? 32 : 48)) * dim;
        }
        CUDA_TRY(cudaGetLastError());
        ctx->counters[1] += 1;
        ctx->counters[4] += offsetof(SweepProg, ops) + (size_t)sw.nops * sizeof(POp) + (size_t)sw.nmat2 * 32 * sizeof(double);
        src = dst;
    }
    {
        KScope ks(ctx, B200_PROF_REDUCE);
        reduce_partials_kernel<<<INNER2_WIDTH, 32, 0, ctx->stream>>>(ctx->d_partial, (int)grid, INNER2_WIDTH, ctx->d_out);
    }
    CUDA_TRY(cudaGetLastError());
    ctx->counters[2] += n_gates;
    tm.stop();
    double t[32];
    if (fetch_out(ctx, t, 32)) return -1;
    if (qa < qb) {
        std::memcpy(out, t, sizeof t);
    } else {
        auto sw2 = [](int i) { return ((i & 1) << 1) | (i >> 1); };
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                out[2 * (4 * sw2(i) + sw2(j))] = t[2 * (4 * i + j)];
                out[2 * (4 * sw2(i) + sw2(j)) + 1] = t[2 * (4 * i + j) + 1];
            }
    }
    return 0;
}

int b200_sv_run_inner2(b200_ctx* ctx, int dst_slot, int src_slot, const b200_gate* gates, int n_gates, const double* mats,
                       int n_mats, int inverse, int other_slot, int qa, int qb, double out[32], int* stored) {
    if (check_slot(ctx, dst_slot) || check_slot(ctx, other_slot)) return -1;
    if (src_slot >= 0 && check_slot(ctx, src_slot)) return -1;      // src_slot = -1: the source is |0..0> (as in b200_sv_run)
    if (src_slot < 0 && stored) *stored = 1;                        // (nothing to keep: the state exists only once swept)
    return run_inner2_impl(ctx, dst_slot, src_slot, nullptr, gates, n_gates, mats, n_mats, inverse, other_slot, qa, qb, out, stored);
}

// fills `es` from the caller's (compact_state, K, qmap)
static int make_embed(b200_ctx* ctx, const void* compact_state, int K, const int32_t* qmap, EmbedSrc& es) {
    if (!qmap || !compact_state) return set_error("null pointer");
    const int n = ctx->nq;
    if (K < 1 || K > n || K > 40) return set_error("embedded source: K out of range");
    std::memset(&es, 0, sizeof es);
    uint64_t inside = 0;
    for (int b = 0; b < K; ++b) {
        if (qmap[b] < 0 || qmap[b] >= n || (inside >> qmap[b] & 1)) return set_error("embedded source: qmap must hold distinct qubits of the register");
        inside |= 1ull << qmap[b];
        es.q[b] = qmap[b];
    }
    es.phi = (const double2*)compact_state;
    es.K = K;
    es.nq = n;
    es.outside = ~inside & ((1ull << n) - 1ull);
    return 0;
}

int b200_sv_run_embedded(b200_ctx* ctx, int dst_slot, const void* compact_state, int K, const int32_t* qmap,
                         const b200_gate* gates, int n_gates, const double* mats, int n_mats, int inverse) {
    if (check_slot(ctx, dst_slot)) return -1;
    EmbedSrc es;
    if (make_embed(ctx, compact_state, K, qmap, es)) return -1;
    CUDA_TRY(cudaSetDevice(ctx->device));
    Plan plan;
    if (make_plan(ctx->nq, gates, n_gates, mats, n_mats, inverse != 0, plan, true)) return -1;
    ctx->counters[6] += 1;
    if (plan.small || plan.sweeps.empty() || ctx->sweep_mode == 1) {     // no sweep to ride on: embed first, then the program
        if (b200_sv_scatter(ctx, dst_slot, qmap, K, compact_state)) return -1;
        return run_plan(ctx, dst_slot, dst_slot, plan);
    }
    return run_plan(ctx, dst_slot, -1, plan, &es);
}

int b200_sv_run_project(b200_ctx* ctx, int scratch_slot, int src_slot, const b200_gate* gates, int n_gates, const double* mats,
                        int n_mats, int inverse, void* compact_dst, int K, const int32_t* qmap, int* scratch_used) {
    if (check_slot(ctx, scratch_slot) || check_slot(ctx, src_slot)) return -1;
    if (scratch_slot == src_slot) return set_error("run_project: the scratch slot must differ from the source");
    EmbedSrc es;
    if (make_embed(ctx, compact_dst, K, qmap, es)) return -1;
    const int n = ctx->nq;
    if (n <= SMALL_MAX_QUBITS) return set_error("run_project: register too small for the tiled path (use b200_sv_run + b200_sv_gather)");
    CUDA_TRY(cudaSetDevice(ctx->device));
    Plan plan;
    {
        if (n_gates < 0 || (n_gates > 0 && !gates)) return set_error("null gate array");
        std::vector<COp> ops;
        const std::string err = canonicalize(n, gates, n_gates, mats, n_mats, inverse != 0, ops);
        if (!err.empty()) return set_error(err);
        fuse_single_qubit_runs(ops);
        fuse_diagonals(ops);
        build_plan(n, ops, plan, true, -2, -2);      // (-2: at least one sweep, nothing forced into the tile)
        plan.n_gates_in = n_gates;
    }
    if (plan.sweeps.empty()) return set_error("run_project: empty plan");
    ProjectDst pd;
    std::memset(&pd, 0, sizeof pd);
    pd.phi = (double2*)compact_dst; pd.outside = es.outside; pd.K = K;
    for (int b = 0; b < K; ++b) pd.q[b] = es.q[b];
    ctx->counters[6] += 1;
    const uint64_t dim = 1ull << n;
    const uint32_t ntiles = (uint32_t)(dim >> TILE_BITS);
    const double2* src = (const double2*)ctx->slots[src_slot];
    double2* scratch = (double2*)ctx->slots[scratch_slot];
    const size_t tile_bytes = ((size_t)1 << TILE_BITS) * sizeof(double2);
    if (scratch_used) *scratch_used = plan.sweeps.size() > 1 ? 1 : 0;
    Timer tm(ctx);
    for (size_t k = 0; k < plan.sweeps.size(); ++k) {
        const SweepProg& sw = plan.sweeps[k];
        const size_t smem = sw.nrounds > 1 ? tile_bytes : 0;
        const int per_sm = sw.nrounds > 1 ? ctx->sweep_occ_smem : ctx->sweep_occ_nosmem;
        const uint32_t g = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)ctx->num_sms * per_sm * ctx->grid_mult);
        if (k + 1 < plan.sweeps.size()) {
            KScope ks(ctx, B200_PROF_SWEEP);
            sv_sweep_kernel<REG_BITS><<<g, SWEEP_THREADS, smem, ctx->stream>>>(src, scratch, sw, ntiles, kNoEmbed);
            ctx->counters[3] += 32 * dim;
            src = scratch;
        } else {
            project_prepare(pd, sw);
            KScope ks(ctx, B200_PROF_PROJECT);
            sv_sweep_project_kernel<REG_BITS><<<g, SWEEP_THREADS, smem, ctx->stream>>>(src, sw, ntiles, pd);
            ctx->counters[3] += 16 * dim + (16ull << K);
        }
        CUDA_TRY(cudaGetLastError());
        ctx->counters[1] += 1;
        ctx->counters[4] += offsetof(SweepProg, ops) + (size_t)sw.nops * sizeof(POp) + (size_t)sw.nmat2 * 32 * sizeof(double);
    }
    ctx->counters[2] += n_gates;
    tm.stop();
    return 0;
}

int b200_sv_run_embedded_inner2(b200_ctx* ctx, int dst_slot, const void* compact_state, int K, const int32_t* qmap,
                                const b200_gate* gates, int n_gates, const double* mats, int n_mats, int inverse,
                                int other_slot, int qa, int qb, double out[32], int* stored) {
    if (check_slot(ctx, dst_slot) || check_slot(ctx, other_slot)) return -1;
    EmbedSrc es;
    if (make_embed(ctx, compact_state, K, qmap, es)) return -1;
    return run_inner2_impl(ctx, dst_slot, -1, &es, gates, n_gates, mats, n_mats, inverse, other_slot, qa, qb, out, stored);
}

int b200_sv_inner2_gather(b200_ctx* ctx, int r_slot, const void* compact_state, int K, const int32_t* qmap, int qa,
                          int qb, double out[32]) {
    if (check_slot(ctx, r_slot)) return -1;
    const int n = ctx->nq;
    if (!compact_state || !qmap || !out) return set_error("null pointer");
    if (K < 2 || K > n || K > 40) return set_error("inner2_gather: compact size out of range");
    QMap qm;
    uint64_t seen = 0;
    int ca = -1, cb = -1;
    for (int b = 0; b < K; ++b) {
        if (qmap[b] < 0 || qmap[b] >= n || (seen >> qmap[b] & 1)) return set_error("inner2_gather: bad qubit map");
        seen |= 1ull << qmap[b];
        qm.q[b] = qmap[b];
        if (qmap[b] == qa) ca = b;
        if (qmap[b] == qb) cb = b;
    }
    if (ca < 0 || cb < 0 || ca == cb) return set_error("inner2_gather: open qubits must be in the map");
    CUDA_TRY(cudaSetDevice(ctx->device));
    Timer tm(ctx);
    const int lo = std::min(ca, cb), hi = std::max(ca, cb);
    const int grid = red_grid(ctx, 1ull << (K - 2));
    {
        KScope ks(ctx, B200_PROF_INNER);
        sv_inner2_gather_kernel<<<grid, RED_THREADS, 0, ctx->stream>>>((const double2*)compact_state, K,
                                                                       (const double2*)ctx->slots[r_slot], qm, lo, hi, ctx->d_partial);
    }
    CUDA_TRY(cudaGetLastError());
    {
        KScope ks(ctx, B200_PROF_REDUCE);
        reduce_partials_kernel<<<INNER2_WIDTH, 32, 0, ctx->stream>>>(ctx->d_partial, grid, INNER2_WIDTH, ctx->d_out);
    }
    CUDA_TRY(cudaGetLastError());
    ctx->counters[3] += 32ull << K;
    ctx->counters[6] += 1;
    tm.stop();
    double t[32];
    if (fetch_out(ctx, t, 32)) return -1;
    if (ca < cb) {
        std::memcpy(out, t, sizeof t);
    } else {
        auto sw2 = [](int i) { return ((i & 1) << 1) | (i >> 1); };
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                out[2 * (4 * sw2(i) + sw2(j))] = t[2 * (4 * i + j)];
                out[2 * (4 * sw2(i) + sw2(j)) + 1] = t[2 * (4 * i + j) + 1];
            }
    }
    return 0;
}

static int gather_impl(b200_ctx* ctx, int slot, const int32_t* qmap, int K, void* dst, int rank_bits, int rank) {
    if (check_slot(ctx, slot)) return -1;
    if (!qmap || !dst) return set_error("null pointer");
    const int n = ctx->nq;
    if (K < 1 || K > n + rank_bits || K > 40) return set_error("gather: K out of range");
    QMap qm;
    uint64_t seen = 0;
    uint32_t covered = 0;
    for (int b = 0; b < K; ++b) {
        if (qmap[b] < 0 || qmap[b] >= n + rank_bits || (seen >> qmap[b] & 1))
            return set_error("gather: qmap must hold distinct qubits of the register");
        seen |= 1ull << qmap[b];
        if (qmap[b] >= n) covered |= 1u << (qmap[b] - n);
        qm.q[b] = qmap[b];
    }
    const int contributes = ((uint32_t)rank & ~covered) == 0 ? 1 : 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    Timer tm(ctx);
    {
        KScope ks(ctx, B200_PROF_INNER);
        sv_gather_kernel<<<red_grid(ctx, 1ull << K), RED_THREADS, 0, ctx->stream>>>(
            (const double2*)ctx->slots[slot], qm, K, (double2*)dst, n, (uint32_t)rank & covered, contributes);
    }
    CUDA_TRY(cudaGetLastError());
    ctx->counters[3] += 32ull << K;
    ctx->counters[6] += 1;
    tm.stop();
    return 0;
}

int b200_sv_gather(b200_ctx* ctx, int slot, const int32_t* qmap, int K, void* dst) {
    return gather_impl(ctx, slot, qmap, K, dst, 0, 0);
}

int b200_sv_gather_ranked(b200_ctx* ctx, int slot, const int32_t* qmap, int K, int rank_bits, int rank, void* dst) {
    if (rank_bits < 0 || rank_bits > 8 || rank < 0 || rank >= (1 << rank_bits)) return set_error("gather: rank out of range");
    return gather_impl(ctx, slot, qmap, K, dst, rank_bits, rank);
}

int b200_sv_scatter(b200_ctx* ctx, int slot, const int32_t* qmap, int K, const void* src) {
    if (check_slot(ctx, slot)) return -1;
    if (!qmap || !src) return set_error("null pointer");
    const int n = ctx->nq;
    if (K < 1 || K > n || K > 40) return set_error("scatter: K out of range");
    QMap qm;
    uint64_t inside = 0;
    for (int b = 0; b < K; ++b) {
        if (qmap[b] < 0 || qmap[b] >= n || (inside >> qmap[b] & 1)) return set_error("scatter: qmap must hold distinct qubits of the register");
        inside |= 1ull << qmap[b];
        qm.q[b] = qmap[b];
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t dim = 1ull << n;
    Timer tm(ctx);
    {   // a plain zero fill at write bandwidth, then 2^K scattered stores (one pass with the test inside ran at half of it)
        KScope ks(ctx, B200_PROF_FILL);
        sv_zero_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>((double2*)ctx->slots[slot], dim);
    }
    CUDA_TRY(cudaGetLastError());
    {
        KScope ks(ctx, B200_PROF_FILL);
        sv_scatter_kernel<<<red_grid(ctx, 1ull << K), RED_THREADS, 0, ctx->stream>>>((double2*)ctx->slots[slot], qm, K, (const double2*)src);
    }
    CUDA_TRY(cudaGetLastError());
    ctx->counters[3] += 16 * dim;
    ctx->counters[6] += 1;
    tm.stop();
    return 0;
}

int b200_sv_download(b200_ctx* ctx, int slot, uint64_t offset, uint64_t count, double* host) {
    if (check_slot(ctx, slot)) return -1;
    if (!host && count) return set_error("null host buffer");
    if (offset > (1ull << ctx->nq) || count > (1ull << ctx->nq) - offset) return set_error("download range out of bounds");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpyAsync(host, (const double2*)ctx->slots[slot] + offset, count * sizeof(double2),
                             cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->counters[5] += count * sizeof(double2);
    return 0;
}

int b200_sv_upload(b200_ctx* ctx, int slot, uint64_t offset, uint64_t count, const double* host) {
    if (check_slot(ctx, slot)) return -1;
    if (!host && count) return set_error("null host buffer");
    if (offset > (1ull << ctx->nq) || count > (1ull << ctx->nq) - offset) return set_error("upload range out of bounds");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpyAsync((double2*)ctx->slots[slot] + offset, host, count * sizeof(double2),
                             cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->counters[4] += count * sizeof(double2);
    return 0;
}

int b200_sv_plan_stats(int num_qubits, const b200_gate* gates, int n_gates, const double* mats, int n_mats,
                       int32_t out[4]) {
    if (!out) return set_error("null pointer");
    if (num_qubits < 1 || num_qubits > 40) return set_error("num_qubits out of range [1,40]");
    Plan plan;
    if (make_plan(num_qubits, gates, n_gates, mats, n_mats, false, plan)) return -1;
    out[0] = plan.small ? 1 : (int32_t)plan.sweeps.size();
    out[1] = plan.small ? 1 : (int32_t)plan.n_rounds();
    out[2] = (int32_t)plan.n_ops();
    out[3] = plan.small ? 1 : 0;
    return 0;
}

int b200_sv_plan_detail(int num_qubits, const b200_gate* gates, int n_gates, const double* mats, int n_mats,
                        int32_t* out, int max_sweeps, int32_t* n_sweeps) {
    if (!out || !n_sweeps) return set_error("null pointer");
    if (num_qubits < 1 || num_qubits > 40) return set_error("num_qubits out of range [1,40]");
    Plan plan;
    if (make_plan(num_qubits, gates, n_gates, mats, n_mats, false, plan)) return -1;
    *n_sweeps = (int32_t)plan.sweeps.size();
    for (int k = 0; k < (int)plan.sweeps.size() && k < max_sweeps; ++k) {
        const SweepProg& sw = plan.sweeps[k];
        const int ops = sw.nops;
        int dense = 0;
        for (int o = 0; o < sw.nops; ++o)
            if (sw.ops[o].kind == P_MAT1 || sw.ops[o].kind == P_MAT2 || sw.ops[o].kind == P_MAT1LANE) ++dense;
        out[4 * k] = sw.nrounds;
        out[4 * k + 1] = ops;
        out[4 * k + 2] = dense;
        out[4 * k + 3] = sw.c;
    }
    return 0;
}

}  // extern "C"
