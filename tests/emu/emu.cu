// emu.cu -- TEST INFRASTRUCTURE: runs the product's planner and the __host__ __device__ bodies of
// the sweep / small kernels on the CPU, one loop iteration per CUDA thread, so that the planner
// (commutation-aware scheduling, tile/round mapping) and the kernels' index math can be checked
// against the oracle without a GPU (`pytest -m "not gpu"`).  Never loaded by the product.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b200aqc.h"
#include "sv_kernels.cuh"
#include "sv_plan.h"

using namespace b200;

static std::string g_err;

static int g_variant = 1;   // 0 = direct-load kernel (sv_sweep_kernel), 1 = pipelined kernel (sv_sweep_pipe_kernel)

extern "C" const char* emu_last_error() { return g_err.c_str(); }
extern "C" void emu_set_variant(int v) { g_variant = v; }

// state: 2^nq complex128 (host), updated in place.  src_is_zero: start from |0..0>.
// fused (qa >= 0): the last sweep keeps its tiles in "shared memory" and runs the epilogue of
// sv_sweep_inner2_kernel against `other`; T (32 doubles, bit0 = lower qubit) comes back in t_out.
static int emu_run_impl(int nq, double* state_ri, int src_is_zero, const b200_gate* gates, int n_gates,
                        const double* mats, int n_mats, int inverse, int32_t stats[4], int qa, int qb,
                        const double* other_ri, int write_back, double* t_out, const EmbedSrc* es = nullptr,
                        ProjectDst* pd = nullptr) {
    const bool fused = qa >= 0;
    std::vector<COp> ops;
    g_err = canonicalize(nq, gates, n_gates, mats, n_mats, inverse != 0, ops);
    if (!g_err.empty()) return -1;
    fuse_single_qubit_runs(ops);
    fuse_diagonals(ops);
    Plan plan;
    build_plan(nq, ops, plan, /*fold_perm=*/g_variant == 0 || fused, fused ? qa : (pd ? -2 : -1), fused ? qb : (pd ? -2 : -1),
               (fused && es != nullptr) ? es->outside : 0);
    if (pd != nullptr && (plan.small || plan.sweeps.empty() || g_variant != 0)) { g_err = "projected store: tiled direct kernel only"; return -1; }
    if (fused && (plan.small || g_variant != 0)) { g_err = "fused path: tiled direct kernel only"; return -1; }
    if (es != nullptr && (plan.small || plan.sweeps.empty() || g_variant != 0)) { g_err = "embedded source: needs a tiled sweep to ride on"; return -1; }
    double2* psi = reinterpret_cast<double2*>(state_ri);
    const uint64_t dim = 1ull << nq;
    if (src_is_zero) {
        for (uint64_t i = 0; i < dim; ++i) psi[i] = make_double2(i == 0 ? 1.0 : 0.0, 0.0);
    }
    if (stats) {
        stats[0] = plan.small ? 1 : (int32_t)plan.sweeps.size();
        stats[1] = plan.small ? 1 : (int32_t)plan.n_rounds();
        stats[2] = (int32_t)plan.n_ops();
        stats[3] = plan.small ? 1 : 0;
    }
    const double* mat2 = plan.mat2.empty() ? nullptr : plan.mat2.data();
    if (plan.small) {
        const uint32_t nthreads = 256;
        for (const DevOp& op : plan.ops)
            for (uint32_t tid = 0; tid < nthreads; ++tid) small_apply_op(psi, (uint32_t)dim, &op, mat2, tid, nthreads);
        return 0;
    }
    // Tiled path: the CTA's 256 threads are stepped in lock-step op by op, so the warp shuffles of the
    // lane ops can be served from a snapshot of the registers (SnapExchange).
    constexpr int NA = 1 << REG_BITS;
    struct Regs { double2 a[NA]; };
    std::vector<double2> smem((size_t)1 << TILE_BITS);
    std::vector<Regs> regs(SWEEP_THREADS), snap(SWEEP_THREADS);
    std::vector<double2> pend(SWEEP_THREADS);
    std::vector<uint64_t> gidx(SWEEP_THREADS);
    std::vector<uint32_t> tls(SWEEP_THREADS), tlin(SWEEP_THREADS);
    struct SnapExchange {
        const Regs* snap; uint32_t tid;
        __host__ __device__ double2 operator()(const double2, const int j, const int lane_mask) const { return snap[tid ^ (uint32_t)lane_mask].a[j]; }
    };
    const uint32_t ntiles = (uint32_t)(dim >> TILE_BITS);
    if (std::getenv("EMU_DUMP")) {
        for (const SweepProg& sp : plan.sweeps) {
            std::printf("sweep: rounds=%d ops=%d c=%d tileq=", sp.nrounds, sp.nops, sp.c);
            for (int i = 0; i < TILE_BITS; ++i) std::printf("%d ", sp.tileq[i]);
            std::printf("\n");
            for (int r = 0; r < sp.nrounds; ++r) {
                const PRound& rd = sp.rounds[r];
                std::printf("  round %d regpos=%d,%d,%d,%d pend=%d\n", r, rd.regpos[0], rd.regpos[1], rd.regpos[2], rd.regpos[3], rd.has_pend);
                for (int o = rd.op_begin; o < rd.op_end; ++o) {
                    const POp& op = sp.ops[o];
                    std::printf("    kind=%d r0=%d r1=%d cq=%d dq0=%d dq1=%d flush=%d\n", op.kind, op.r0, op.r1, op.cq, op.dq0, op.dq1, op.flush & 1);
                }
            }
        }
    }
    bool first_sweep = true;
    EpiProg ep;
    std::vector<double2> tacc((size_t)SWEEP_THREADS * 4, make_double2(0.0, 0.0));
    const double2* other = reinterpret_cast<const double2*>(other_ri);
    if (fused && !make_epilogue(plan.sweeps.back(), qa, qb, ep)) { g_err = "pair not in the last tile"; return -1; }
    for (size_t si = 0; si < plan.sweeps.size(); ++si) {
        const SweepProg& sp = plan.sweeps[si];
        const bool tail = fused && si + 1 == plan.sweeps.size();
        EmbedSrc e1;
        if (es != nullptr && si == 0) { e1 = *es; embed_prepare(e1, sp); }
        const bool proj = pd != nullptr && si + 1 == plan.sweeps.size();
        if (proj) project_prepare(*pd, sp);
        const int nr = sp.nrounds;
        // the direct kernel takes |0..0> as an IMPLICIT source (src == nullptr) in its first sweep
        const double2* hbm_src = (first_sweep && src_is_zero && g_variant == 0) ? nullptr : psi;
        first_sweep = false;
        for (uint32_t tile = 0; tile < ntiles; ++tile) {
            const uint64_t base = sweep_tile_base(sp, tile);
            if (tail && es != nullptr && si == 0 && !(base & e1.outside_nontile) && e1.n_skip > 0) {
                // the kernel's compact enumeration (T only) must reach this tile: its number without the skipped bits
                uint32_t t = 0; int k = 0;
                for (int b = 0; b < e1.nq; ++b)
                    if (!((e1.skipmask >> b) & 1ull)) { t |= (uint32_t)((base >> b) & 1ull) << k; ++k; }
                if (embed_tile_base(e1, t) != base || t >= (ntiles >> e1.n_skip)) { g_err = "compact tile enumeration is off"; return -1; }
            }
            if (tail && es != nullptr && si == 0 && (base & e1.outside_nontile)) {     // the kernel's zero-tile shortcut
                if (write_back)
                    for (uint32_t tid = 0; tid < (uint32_t)SWEEP_THREADS; ++tid)
                        epi_zero_tile<REG_BITS>(psi, ep, epi_index(sp, ep, tid), base);
                continue;
            }
            const uint32_t rows = 1u << (TILE_BITS - sp.c), row_len = 1u << sp.c;
            if (g_variant == 1)   // the producer warp's bulk copies: tile rows -> linear buffer
                for (uint32_t row = 0; row < rows; ++row)
                    std::memcpy(smem.data() + ((size_t)row << sp.c), psi + sweep_row_index(sp, base, row), row_len * sizeof(double2));
            for (int r = 0; r < nr; ++r) {
                const PRound& rd = sp.rounds[r];
                for (uint32_t tid = 0; tid < (uint32_t)SWEEP_THREADS; ++tid) {
                    uint32_t tl;
                    round_index<REG_BITS>(sp, rd, base, tid, tl, gidx[tid]);
                    tls[tid] = swz(tl);
                    tlin[tid] = tl;
                    if (r == 0 && g_variant == 1) round_load_lin<REG_BITS>(regs[tid].a, smem.data(), rd, tl);
                    else if (r == 0) round_load_hbm<REG_BITS>(regs[tid].a, hbm_src, sp, rd, gidx[tid], (es != nullptr && si == 0) ? &e1 : nullptr);
                    else round_load_smem<REG_BITS>(regs[tid].a, smem.data(), rd, tls[tid], gidx[tid]);
                    pend[tid] = make_double2(1.0, 0.0);
                }
                if (tail && r == nr - 1)
                    for (int f = 0; f < rd.n_trail; ++f)
                        if (rd.trail[f].smask == 0) { g_err = "lane fold on the store side of a fused tail"; return -1; }
                for (int o = rd.op_begin; o < rd.op_end; ++o) {
                    POp op = sp.ops[o];
                    const bool lane_op = op.kind == P_XLANE || op.kind == P_MAT1LANE;
                    if (lane_op) {
                        if (r != 0 && r != nr - 1) { g_err = "lane op scheduled in a shared-memory round"; return -1; }

                        if (op.flush & 1) {   // every lane applies its pending phase before the exchange
                            for (uint32_t tid = 0; tid < (uint32_t)SWEEP_THREADS; ++tid) {
                                apply_pend<REG_BITS>(regs[tid].a, pend[tid]);
                                pend[tid] = make_double2(1.0, 0.0);
                            }
                            op.flush &= ~1;
                        }
                        snap = regs;
                    }
                    for (uint32_t tid = 0; tid < (uint32_t)SWEEP_THREADS; ++tid) {
                        const SnapExchange ex{snap.data(), tid};
                        apply_op<REG_BITS>(regs[tid].a, sp, op, gidx[tid], tid & 31u, pend[tid], ex);
                    }
                }
                for (uint32_t tid = 0; tid < (uint32_t)SWEEP_THREADS; ++tid) {
                    if (rd.has_pend) apply_pend<REG_BITS>(regs[tid].a, pend[tid]);
                    if (r == nr - 1 && g_variant == 1) round_store_lin<REG_BITS>(regs[tid].a, smem.data(), rd, tlin[tid]);
                    else if (r == nr - 1 && !tail) round_store_hbm<REG_BITS>(regs[tid].a, psi, sp, rd, gidx[tid], proj ? pd : nullptr);
                    else round_store_smem<REG_BITS>(regs[tid].a, smem.data(), rd, tls[tid], gidx[tid]);
                }
            }
            if (g_variant == 1)
                for (uint32_t row = 0; row < rows; ++row)
                    std::memcpy(psi + sweep_row_index(sp, base, row), smem.data() + ((size_t)row << sp.c), row_len * sizeof(double2));
            if (tail)
                for (uint32_t tid = 0; tid < (uint32_t)SWEEP_THREADS; ++tid) {
                    const EpiIdx ix = epi_index(sp, ep, tid);
                    double2 t[4];
                    for (int i = 0; i < 4; ++i) t[i] = tacc[(size_t)i * SWEEP_THREADS + tid];
                    epi_tile<REG_BITS>(smem.data(), other, write_back ? psi : nullptr, ep, ix, base, t);
                    for (int i = 0; i < 4; ++i) tacc[(size_t)i * SWEEP_THREADS + tid] = t[i];
                }
        }
    }
    if (fused) {   // the kernel's CTA reduction
        for (int x = 0; x < 32; ++x) {
            const int k = x >> 1, part = x & 1, i = k >> 2, j = k & 3;
            double sum = 0.0;
            for (uint32_t u = 0; u < (uint32_t)SWEEP_THREADS; ++u) {
                const int ju = (int)(((u >> ep.ja) & 1u) | (((u >> ep.jb) & 1u) << 1));
                const double2 v = tacc[(size_t)i * SWEEP_THREADS + u];
                if (ju == j) sum += part ? v.y : v.x;
            }
            t_out[x] = sum;
        }
    }
    return 0;
}

extern "C" int emu_sv_run(int nq, double* state_ri, int src_is_zero, const b200_gate* gates, int n_gates,
                          const double* mats, int n_mats, int inverse, int32_t stats[4]) {
    return emu_run_impl(nq, state_ri, src_is_zero, gates, n_gates, mats, n_mats, inverse, stats, -1, -1, nullptr, 0, nullptr);
}

// state <- gates applied to state (write_back != 0), t_out = T[i][j] = sum conj(state'[i,rest]) other[j,rest], index bit 0 = the
// LOWER of (qa, qb) (b200_sv_run_inner2 reorders for the caller; the test does the same)
extern "C" int emu_sv_run_inner2(int nq, double* state_ri, const double* other_ri, const b200_gate* gates, int n_gates,
                                 const double* mats, int n_mats, int inverse, int qa, int qb, int write_back, double* t_out,
                                 int32_t stats[4]) {
    const int saved = g_variant;
    g_variant = 0;
    const int rc = emu_run_impl(nq, state_ri, 0, gates, n_gates, mats, n_mats, inverse, stats, qa, qb, other_ri, write_back, t_out);
    g_variant = saved;
    return rc;
}


// state <- gates applied to the embedded state (phi on the qubits qmap[0..K), |0> elsewhere); qa >= 0: also the fused
// transfer pass against `other` (see emu_sv_run_inner2).  rc 1: no tiled sweep to ride on (the product scatters first).
extern "C" int emu_sv_run_embedded(int nq, double* state_ri, const double* phi_ri, int K, const int32_t* qmap,
                                   const b200_gate* gates, int n_gates, const double* mats, int n_mats, int inverse, int qa,
                                   int qb, const double* other_ri, double* t_out, int32_t stats[4]) {
    EmbedSrc es;
    std::memset(&es, 0, sizeof es);
    uint64_t inside = 0;
    for (int b = 0; b < K; ++b) { es.q[b] = qmap[b]; inside |= 1ull << qmap[b]; }
    es.phi = reinterpret_cast<const double2*>(phi_ri);
    es.K = K;
    es.nq = nq;
    es.outside = ~inside & ((1ull << nq) - 1ull);
    const int saved = g_variant;
    g_variant = 0;
    const int rc = emu_run_impl(nq, state_ri, 0, gates, n_gates, mats, n_mats, inverse, stats, qa, qb, other_ri, 1, t_out, &es);
    g_variant = saved;
    return rc;
}


// phi (2^K amplitudes) <- projection of (gates applied to state) onto |0> of every qubit outside qmap, through the
// projected-store sweep; `state` holds intermediate sweeps of longer programs (the product's scratch slot) and is NOT the
// swept state afterwards.  stats[0] = number of sweeps.
extern "C" int emu_sv_run_project(int nq, double* state_ri, double* phi_ri, int K, const int32_t* qmap, const b200_gate* gates,
                                  int n_gates, const double* mats, int n_mats, int inverse, int32_t stats[4]) {
    ProjectDst pd;
    std::memset(&pd, 0, sizeof pd);
    uint64_t inside = 0;
    for (int b = 0; b < K; ++b) { pd.q[b] = qmap[b]; inside |= 1ull << qmap[b]; }
    pd.phi = reinterpret_cast<double2*>(phi_ri);
    pd.K = K;
    pd.outside = ~inside & ((1ull << nq) - 1ull);
    const int saved = g_variant;
    g_variant = 0;
    const int rc = emu_run_impl(nq, state_ri, 0, gates, n_gates, mats, n_mats, inverse, stats, -1, -1, nullptr, 0, nullptr, nullptr, &pd);
    g_variant = saved;
    return rc;
}
