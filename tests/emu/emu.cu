// emu.cu -- TEST INFRASTRUCTURE: runs the product's planner and the __host__ __device__ bodies of
// the sweep / small kernels on the CPU, one loop iteration per CUDA thread, so that the planner
// (commutation-aware scheduling, tile/round mapping) and the kernels' index math can be checked
// against the oracle without a GPU (`pytest -m "not gpu"`).  Never loaded by the product.
#include <cstring>
#include <string>
#include <vector>

#include "b200aqc.h"
#include "sv_kernels.cuh"
#include "sv_plan.h"

using namespace b200;

static std::string g_err;

extern "C" const char* emu_last_error() { return g_err.c_str(); }

// state: 2^nq complex128 (host), updated in place.  src_is_zero: start from |0..0>.
extern "C" int emu_sv_run(int nq, double* state_ri, int src_is_zero, const b200_gate* gates, int n_gates,
                          const double* mats, int n_mats, int inverse, int32_t stats[4]) {
    std::vector<COp> ops;
    g_err = canonicalize(nq, gates, n_gates, mats, n_mats, inverse != 0, ops);
    if (!g_err.empty()) return -1;
    fuse_single_qubit_runs(ops);
    Plan plan;
    build_plan(nq, ops, plan);
    double2* psi = reinterpret_cast<double2*>(state_ri);
    const uint64_t dim = 1ull << nq;
    if (src_is_zero) {
        for (uint64_t i = 0; i < dim; ++i) psi[i] = make_double2(i == 0 ? 1.0 : 0.0, 0.0);
    }
    if (stats) {
        stats[0] = plan.small ? 1 : (int32_t)plan.sweeps.size();
        stats[1] = plan.small ? 1 : (int32_t)plan.rounds.size();
        stats[2] = (int32_t)plan.ops.size();
        stats[3] = plan.small ? 1 : 0;
    }
    const double* mat2 = plan.mat2.empty() ? nullptr : plan.mat2.data();
    if (plan.small) {
        const uint32_t nthreads = 256;
        for (const DevOp& op : plan.ops)
            for (uint32_t tid = 0; tid < nthreads; ++tid) small_apply_op(psi, (uint32_t)dim, &op, mat2, tid, nthreads);
        return 0;
    }
    std::vector<double2> smem((size_t)1 << TILE_BITS);
    const uint32_t ntiles = (uint32_t)(dim >> TILE_BITS);
    for (const DevSweep& sw : plan.sweeps) {
        const int nr = sw.round_end - sw.round_begin;
        for (uint32_t tile = 0; tile < ntiles; ++tile) {
            const uint64_t base = sweep_tile_base(sw, tile);
            for (int r = 0; r < nr; ++r)
                for (uint32_t tid = 0; tid < (uint32_t)SWEEP_THREADS; ++tid)
                    sweep_round<REG_BITS>(psi, psi, smem.data(), sw, plan.rounds.data() + sw.round_begin + r,
                                          plan.ops.data(), mat2, base, tid, r == 0, r == nr - 1);
        }
    }
    return 0;
}
