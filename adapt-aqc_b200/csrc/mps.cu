// mps.cu -- C-ABI entry points of the matrix-product-state path (b200_mps_*).  See include/b200aqc.h.
//
// A b200_mps is a Vidal-form MPS resident in HBM: Gamma_i as [2][chi_{i-1}][chi_i] complex128,
// lambda_i as chi_i doubles -- the same content as the reference's QiskitMPS wire format
// (adaptaqc/utils/constants.py:17).  Gate application follows the algorithm of the simulator the
// reference calls (qiskit-aer matrix_product_state: contract two sites with the surrounding
// lambdas, apply the gate, SVD, truncate with Aer's rule, divide the outer lambdas back out; swaps
// for non-neighbours), with every tensor operation on the device:
//   contraction / transfer matrices : zgemm_dmma_kernel (FP64 tensor cores)
//   SVD                              : one-sided Jacobi (jacobi_cta_kernel / jacobi_block_kernel; coop + per-round fallbacks)
//   truncation                       : singular values (<= 4 KB) are read back, Aer's reduce_zeros
//                                      rule picks the kept count (the launch geometry of the
//                                      following kernels depends on it), the rest stays on device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "b200aqc.h"
#include "ctx.h"
#include "mps_kernels.cuh"
#include "sv_plan.h"

using namespace b200;
using cplx = std::complex<double>;

namespace b200 {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// per-context scratch of the MPS path
struct MpsState {
    DevBuf packA, packB, theta, X, W, sigma, perm, kept, flag, table, env[6], bits, outz;
    std::vector<b200_mps*> live;
};

}  // namespace b200

struct b200_mps {
    b200_ctx* ctx = nullptr;
    int n = 0;
    double thr = 1e-16;
    int max_chi = 0;                 // 0 = unlimited
    std::vector<int> chi;            // chi[i] = bond to the right of site i (chi[n-1] = 1)
    std::vector<DevBuf> gam;         // [2][chi_l][chi_r] double2
    std::vector<DevBuf> lam;         // chi[i] doubles, i < n-1
    uint64_t svd_count = 0, svd_sweeps = 0, svd_flops = 0;
    int chiL(int i) const { return i == 0 ? 1 : chi[i - 1]; }
};

namespace {

int reserve(DevBuf& b, size_t bytes, cudaStream_t stream) {
    if (b.cap >= bytes) return 0;
    if (b.p) {
        CUDA_TRY(cudaStreamSynchronize(stream));
        cudaFree(b.p);
    }
    b.p = nullptr; b.cap = 0;
    const size_t want = std::max(bytes + bytes / 2, (size_t)4096);
    CUDA_TRY(cudaMalloc(&b.p, want));
    b.cap = want;
    return 0;
}

void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
}

MpsState* state(b200_ctx* ctx) {
    if (!ctx->mps) ctx->mps = new MpsState();
    return ctx->mps;
}

struct MScope {  // counts + (in profile mode) times one MPS kernel launch, like KScope in api.cu
    b200_ctx* c;
    cudaEvent_t a = nullptr, b = nullptr;
    int cls;
    explicit MScope(b200_ctx* ctx, int cls_ = B200_PROF_MPS) : c(ctx), cls(cls_) {
        c->counters[0] += 1;
        if (c->profiling) {
            auto get = [&]() { if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
                               cudaEvent_t e = nullptr; cudaEventCreate(&e); return e; };
            a = get(); b = get();
            cudaEventRecord(a, c->stream);
        }
    }
    ~MScope() {
        if (a) { cudaEventRecord(b, c->stream); c->prof_recs.push_back({cls, a, b}); }
    }
};

int grid_for(size_t work, int threads = 256) { return (int)std::max<size_t>(1, std::min<size_t>((work + threads - 1) / threads, 1184)); }

int gemm(b200_ctx* ctx, const GemmArgs& g) {
    if (g.M <= 0 || g.N <= 0) return 0;
    // 64 x 64 tiles with a register-prefetch pipeline once the problem fills the machine with them; the
    // 32 x 32 kernel keeps more CTAs in flight on small ones
    const long long tiles64 = (long long)((g.N + G2_TILE - 1) / G2_TILE) * ((g.M + G2_TILE - 1) / G2_TILE);
    if (tiles64 >= ctx->num_sms / 4 && g.K >= 2 * G2_K && !std::getenv("B200AQC_GEMM32")) {
        dim3 grid((g.N + G2_TILE - 1) / G2_TILE, (g.M + G2_TILE - 1) / G2_TILE);
        MScope ms(ctx, B200_PROF_GEMM);
        zgemm_dmma64_kernel<<<grid, 256, 0, ctx->stream>>>(g);
    } else {
        dim3 grid((g.N + GM_TILE - 1) / GM_TILE, (g.M + GM_TILE - 1) / GM_TILE);
        MScope ms(ctx, B200_PROF_GEMM);
        zgemm_dmma_kernel<<<grid, 128, 0, ctx->stream>>>(g);
    }
    CUDA_TRY(cudaGetLastError());
    ctx->counters[7] += (uint64_t)8 * g.M * g.N * g.K;   // real flops of the complex GEMM
    return 0;
}

// C (M x N, row-major, ldc = N) = opA * B with plain row-major operands
GemmArgs gemm_nn(const double2* A, const double2* B, double2* C, int M, int N, int K) {
    GemmArgs g;
    std::memset(&g, 0, sizeof g);
    g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K;
    g.sam = K; g.sak = 1; g.sbk = N; g.sbn = 1; g.scm = N; g.scn = 1;
    g.rs_mod = g.cs_mod = 1;
    return g;
}

int check_mps(b200_mps* m) {
    if (!m || !m->ctx) return set_error("null MPS handle");
    return 0;
}

int set_site(b200_mps* m, int i, int chi_l, int chi_r) {
    if (reserve(m->gam[i], (size_t)2 * chi_l * chi_r * sizeof(double2), m->ctx->stream)) return -1;
    return 0;
}

// 2x2 / 4x4 matrix of one gate record via the statevector planner's canonical form
int gate_matrix(int nq, const b200_gate& g, const double* mats, int n_mats, cplx out[16], int& nqubits, int& qa, int& qb) {
    if (g.op == B200_OP_SWAP) {
        if (g.q0 < 0 || g.q1 < 0 || g.q0 >= nq || g.q1 >= nq || g.q0 == g.q1) return set_error("swap: qubits out of range");
        for (int k = 0; k < 16; ++k) out[k] = 0;
        out[0] = out[6] = out[9] = out[15] = 1;
        nqubits = 2; qa = g.q0; qb = g.q1;
        return 0;
    }
    std::vector<COp> ops;
    const std::string err = canonicalize(nq, &g, 1, mats, n_mats, false, ops);
    if (!err.empty()) return set_error(err);
    if (ops.empty()) { nqubits = 0; return 0; }
    const COp& o = ops[0];
    for (int k = 0; k < 16; ++k) out[k] = 0;
    if (o.kind == K_MAT1 && o.c < 0) { nqubits = 1; qa = o.t0; for (int k = 0; k < 4; ++k) out[k] = o.m[k]; }
    else if (o.kind == K_X && o.c < 0) { nqubits = 1; qa = o.t0; out[1] = out[2] = 1; }
    else if (o.kind == K_DIAG && o.d1 < 0) { nqubits = 1; qa = o.d0; out[0] = o.m[0]; out[3] = o.m[1]; }
    else if (o.kind == K_X) {            // cx: index = bit(control) + 2 bit(target)
        nqubits = 2; qa = o.c; qb = o.t0;
        out[0 * 4 + 0] = 1; out[2 * 4 + 2] = 1; out[1 * 4 + 3] = 1; out[3 * 4 + 1] = 1;
    } else if (o.kind == K_DIAG) { nqubits = 2; qa = o.d0; qb = o.d1; for (int k = 0; k < 4; ++k) out[5 * k] = o.m[k]; }
    else if (o.kind == K_MAT2) { nqubits = 2; qa = o.t0; qb = o.t1; for (int k = 0; k < 16; ++k) out[k] = o.m[k]; }
    else return set_error("unsupported controlled gate on the MPS path");
    return 0;
}

int apply_1q(b200_mps* m, int q, const cplx g[4]) {
    b200_ctx* ctx = m->ctx;
    const int sz = m->chiL(q) * m->chi[q];
    {
        MScope ms(ctx);
        mps_apply1q_kernel<<<grid_for(sz), 256, 0, ctx->stream>>>(
            (double2*)m->gam[q].p, sz, make_double2(g[0].real(), g[0].imag()), make_double2(g[1].real(), g[1].imag()),
            make_double2(g[2].real(), g[2].imag()), make_double2(g[3].real(), g[3].imag()));
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Aer's reduce_zeros (qiskit-aer 0.16, src/simulators/matrix_product_state/svd.cpp -- not vendored in the reference
// tree, restated from the published source): number of singular values kept (S descending) and their renormalised
// values.  Two details of that function are easy to get wrong, so both are a named rule (b200_mps_set_chop_rule):
//   B200_CHOP_AER (default): num_of_SV counts values with std::norm(S[i]) > CHOP_THRESHOLD -- std::norm of a real is its
//     SQUARE, i.e. sigma^2 > 1e-16 (sigma > 1e-8); and the tail-drop loop only lowers the kept count when it BREAKS: if
//     every value down to S[1] fits under the threshold the count is left unchanged (nothing is dropped).
//   B200_CHOP_SIGMA: the round-1 reading -- sigma > 1e-16, and a loop that runs out keeps exactly one value.
constexpr double CHOP_THRESHOLD = 1e-16;
int g_chop_rule = 0;   // B200_CHOP_AER

int reduce_zeros(const std::vector<double>& S, int max_chi, double thr, std::vector<double>& kept) {
    const bool aer = g_chop_rule == 0;
    int sv_num = 0;
    for (double s : S) if ((aer ? s * s : s) > CHOP_THRESHOLD) ++sv_num;
    int new_num = sv_num;
    if (max_chi > 0 && max_chi < sv_num) new_num = max_chi;
    double sum_sq = 0.0;
    int i = new_num - 1;
    bool broke = false;
    for (; i > 0; --i) {
        if (sum_sq + S[i] * S[i] < thr) sum_sq += S[i] * S[i];
        else { broke = true; break; }
    }
    if (broke || !aer) new_num = i + 1;
    new_num = std::max(1, new_num);
    kept.assign(S.begin(), S.begin() + new_num);
    if (new_num < sv_num) {
        double nrm = 0;
        for (double s : kept) nrm += s * s;
        nrm = std::sqrt(nrm);
        for (double& s : kept) s /= nrm;
    }
    return new_num;
}

// 2-qubit gate on neighbouring sites (i, i+1); u4 indexed bit(site i) + 2 bit(site i+1)
int apply_adjacent(b200_mps* m, int i, const cplx u4[16]) {
    b200_ctx* ctx = m->ctx;
    MpsState* st = state(ctx);
    cudaStream_t s = ctx->stream;
    const int chi_l = m->chiL(i), chi_m = m->chi[i], chi_r = m->chi[i + 1];
    const int mm = 2 * chi_l, nn = 2 * chi_r;
    const double* ll = i > 0 ? (const double*)m->lam[i - 1].p : nullptr;
    const double* lm = (const double*)m->lam[i].p;
    const double* lr = (i + 1 < m->n - 1) ? (const double*)m->lam[i + 1].p : nullptr;

    // A' = ll Gamma_i lm  (2 chi_l x chi_m),  B'[beta, (b', gamma)] = Gamma_{i+1}[b'][beta][gamma] lr[gamma]
    if (reserve(st->packA, (size_t)mm * chi_m * sizeof(double2), s)) return -1;
    if (reserve(st->theta, (size_t)mm * nn * sizeof(double2), s)) return -1;
    {
        MScope ms(ctx);
        mps_pack_scaled_kernel<<<grid_for((size_t)mm * chi_m), 256, 0, s>>>((const double2*)m->gam[i].p, chi_l, chi_m, ll, lm,
                                                                             (double2*)st->packA.p);
    }
    CUDA_TRY(cudaGetLastError());
    // theta_raw[(b,al), (b',gm)] = sum_beta A'[(b,al),beta] * Gamma_{i+1}[b'][beta][gm] * lr[gm]: one GEMM per b'
    for (int bp = 0; bp < 2; ++bp) {
        GemmArgs g = gemm_nn((const double2*)st->packA.p, (const double2*)m->gam[i + 1].p + (size_t)bp * chi_m * chi_r,
                             (double2*)st->theta.p + (size_t)bp * chi_r, mm, chi_r, chi_m);
        g.scm = nn;
        g.colscale = lr; g.cs_mod = chi_r;
        if (gemm(ctx, g)) return -1;
    }
    // gate + column-major layout in the tall orientation
    const int tall = mm >= nn ? 1 : 0;
    const int p = tall ? mm : nn, q = tall ? nn : mm;
    if (reserve(st->X, (size_t)p * q * sizeof(double2), s)) return -1;
    if (reserve(st->W, (size_t)q * q * sizeof(double2), s)) return -1;
    if (reserve(st->sigma, (size_t)q * sizeof(double), s)) return -1;
    if (reserve(st->perm, (size_t)q * sizeof(int), s)) return -1;
    if (reserve(st->kept, (size_t)q * sizeof(double), s)) return -1;
    if (reserve(st->flag, 64, s)) return -1;
    Gate4 U;
    for (int k = 0; k < 16; ++k) U.u[k] = make_double2(u4[k].real(), u4[k].imag());
    {
        MScope ms(ctx);
        mps_theta_gate_kernel<<<grid_for((size_t)chi_l * chi_r), 256, 0, s>>>((const double2*)st->theta.p, chi_l, chi_r, U, tall,
                                                                               (double2*)st->X.p);
    }
    CUDA_TRY(cudaGetLastError());
    {
        MScope ms(ctx);
        mps_set_identity_kernel<<<grid_for((size_t)q * q), 256, 0, s>>>((double2*)st->W.p, q);
    }
    CUDA_TRY(cudaGetLastError());

    // ---- one-sided Jacobi SVD ----
    const int max_sweeps = 40;
    int sweeps = 0;
    if (q == 1) {
        sweeps = 0;
    } else if (q <= JACOBI_CTA_MAX_Q) {
        const int pairs = ((q + 1) / 2);
        const int threads = std::min(1024, std::max(32, pairs * 32));
        {
            MScope ms(ctx, B200_PROF_SVD);
            jacobi_cta_kernel<<<1, threads, 0, s>>>((double2*)st->X.p, (double2*)st->W.p, p, q, max_sweeps, (int*)st->flag.p);
        }
        CUDA_TRY(cudaGetLastError());
    } else {
        const int N = (q + 1) & ~1;
        double* fro2 = (double*)((char*)st->flag.p + 32);
        {
            MScope ms(ctx);
            jacobi_fro_kernel<<<1, 256, 0, s>>>((const double2*)st->X.p, (size_t)p * q, fro2);
        }
        CUDA_TRY(cudaGetLastError());
        bool coop_done = false;
        // blocked cooperative SVD: 2*WB columns of X and W per CTA in shared memory
        constexpr int WB = 4;
        const size_t jb_smem = (size_t)2 * WB * (p + q) * sizeof(double2);
        if (ctx->coop_ok && jb_smem <= 200 * 1024 && !std::getenv("B200AQC_JACOBI_SCALAR")) {
            const int NB = (((q + WB - 1) / WB) + 1) & ~1;
            // (A variant that orthogonalises the whole block pair per round through its 8x8 Gram matrix -- one warp
            // diagonalises it, one pass applies the accumulated unitary -- was built and measured in round 2: same sweep
            // count, 73.9 ms instead of 40.0 ms per 512x512 SVD: profiles/r2j_svd_gram_negative.txt.  Removed.)
            const void* jkernel = (const void*)jacobi_block_kernel<WB>;
            CUDA_TRY(cudaFuncSetAttribute(jkernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            int per_sm = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jkernel, JB_GROUP * WB, jb_smem));
            if ((long long)per_sm * ctx->num_sms >= NB / 2) {
                CUDA_TRY(cudaMemsetAsync(st->flag.p, 0, 4 * sizeof(int), s));
                double2* Xp = (double2*)st->X.p; double2* Wp = (double2*)st->W.p;
                int pp = p, qq = q, NN = NB, ms_ = max_sweeps;
                const double* fr = fro2; int* ctrl = (int*)st->flag.p;
                void* args[] = {&Xp, &Wp, &pp, &qq, &NN, &ms_, &fr, &ctrl};
                {
                    MScope ms(ctx, B200_PROF_SVD);
                    CUDA_TRY(cudaLaunchCooperativeKernel(jkernel, dim3(NB / 2), dim3(JB_GROUP * WB), args, jb_smem, s));
                }
                int done[3] = {0, 0, 0};
                CUDA_TRY(cudaMemcpyAsync(done, st->flag.p, sizeof done, cudaMemcpyDeviceToHost, s));
                CUDA_TRY(cudaStreamSynchronize(s));
                ctx->counters[5] += sizeof done;
                sweeps = done[2];
                coop_done = true;
            }
        }
        if (!coop_done && ctx->coop_ok) {
            // one cooperative launch for the whole SVD (needs all N/2 CTAs co-resident)
            int per_sm = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_coop_kernel, 128, 0));
            if ((long long)per_sm * ctx->num_sms >= N / 2) {
                CUDA_TRY(cudaMemsetAsync(st->flag.p, 0, 3 * sizeof(int), s));
                double2* Xp = (double2*)st->X.p; double2* Wp = (double2*)st->W.p;
                int pp = p, qq = q, NN = N, ms_ = max_sweeps;
                const double* fr = fro2; int* ctrl = (int*)st->flag.p;
                void* args[] = {&Xp, &Wp, &pp, &qq, &NN, &ms_, &fr, &ctrl};
                {
                    MScope ms(ctx, B200_PROF_SVD);
                    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)jacobi_coop_kernel, dim3(N / 2), dim3(128), args, 0, s));
                }
                int done[3] = {0, 0, 0};
                CUDA_TRY(cudaMemcpyAsync(done, st->flag.p, sizeof done, cudaMemcpyDeviceToHost, s));
                CUDA_TRY(cudaStreamSynchronize(s));
                ctx->counters[5] += sizeof done;
                sweeps = done[2];
                coop_done = true;
            }
        }
        for (; !coop_done && sweeps < max_sweeps; ++sweeps) {
            CUDA_TRY(cudaMemsetAsync(st->flag.p, 0, sizeof(int), s));
            for (int r = 0; r < N - 1; ++r) {
                MScope ms(ctx, B200_PROF_SVD);
                jacobi_round_kernel<<<N / 2, 128, 0, s>>>((double2*)st->X.p, (double2*)st->W.p, p, q, N, r, fro2, (int*)st->flag.p);
            }
            CUDA_TRY(cudaGetLastError());
            int rotated = 0;
            CUDA_TRY(cudaMemcpyAsync(&rotated, st->flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            ctx->counters[5] += sizeof(int);
            if (!rotated) { ++sweeps; break; }
        }
        if (sweeps >= max_sweeps) return set_error("Jacobi SVD did not converge in " + std::to_string(max_sweeps) + " sweeps");
    }
    {
        MScope ms(ctx);
        jacobi_sigma_kernel<<<(q + 7) / 8, 256, 0, s>>>((const double2*)st->X.p, p, q, (double*)st->sigma.p);
    }
    CUDA_TRY(cudaGetLastError());
    std::vector<double> sig(q);
    int cta_sweeps = 0;
    CUDA_TRY(cudaMemcpyAsync(sig.data(), st->sigma.p, q * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (q > 1 && q <= JACOBI_CTA_MAX_Q)
        CUDA_TRY(cudaMemcpyAsync(&cta_sweeps, st->flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    ctx->counters[5] += q * sizeof(double);
    if (q > 1 && q <= JACOBI_CTA_MAX_Q) {
        sweeps = cta_sweeps;
        if (sweeps >= max_sweeps) return set_error("Jacobi SVD (single-CTA) did not converge");
    }
    m->svd_count += 1;
    m->svd_sweeps += sweeps;
    // algorithmic flop model of one-sided Jacobi: per column pair and sweep, three length-p complex inner
    // products (Gram entries: 2 + 2 + 8 flop per row) and the rotation of two columns of X (p rows) and of W (q rows),
    // 28 flop per row -- (40 p + 28 q) flop per pair, q (q - 1) / 2 pairs per sweep
    m->svd_flops += (uint64_t)sweeps * ((uint64_t)q * (q - 1) / 2) * (uint64_t)(40 * p + 28 * q);

    // ---- truncation (Aer reduce_zeros) ----
    std::vector<int> order(q);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sig[a] > sig[b]; });
    std::vector<double> S(q), kept;
    for (int j = 0; j < q; ++j) S[j] = sig[order[j]];
    const int k = reduce_zeros(S, m->max_chi, m->thr, kept);
    CUDA_TRY(cudaMemcpyAsync(st->perm.p, order.data(), k * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(st->kept.p, kept.data(), k * sizeof(double), cudaMemcpyHostToDevice, s));
    ctx->counters[4] += k * (sizeof(int) + sizeof(double));

    // ---- new site tensors ----
    if (set_site(m, i, chi_l, k) || set_site(m, i + 1, k, chi_r)) return -1;
    if (reserve(m->lam[i], (size_t)k * sizeof(double), s)) return -1;
    {
        MScope ms(ctx);
        mps_write_sites_kernel<<<grid_for((size_t)(mm + nn) * k + k), 256, 0, s>>>(
            (const double2*)st->X.p, (const double2*)st->W.p, (const double*)st->sigma.p, (const int*)st->perm.p,
            (const double*)st->kept.p, chi_l, chi_r, k, tall, ll, lr, (double2*)m->gam[i].p, (double2*)m->gam[i + 1].p,
            (double*)m->lam[i].p);
    }
    CUDA_TRY(cudaGetLastError());
    // order/kept are host vectors read by the async copies above
    CUDA_TRY(cudaStreamSynchronize(s));
    m->chi[i] = k;
    return 0;
}

int apply_2q(b200_mps* m, int q0, int q1, const cplx u_in[16]) {
    cplx u[16];
    if (q0 > q1) {   // re-index so that the first index bit belongs to the lower site
        static const int sw[4] = {0, 2, 1, 3};
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) u[4 * sw[r] + sw[c]] = u_in[4 * r + c];
        std::swap(q0, q1);
    } else {
        for (int k = 0; k < 16; ++k) u[k] = u_in[k];
    }
    cplx swp[16];
    for (int k = 0; k < 16; ++k) swp[k] = 0;
    swp[0] = swp[6] = swp[9] = swp[15] = 1;
    for (int j = q1 - 1; j > q0; --j)
        if (apply_adjacent(m, j, swp)) return -1;
    if (apply_adjacent(m, q0, u)) return -1;
    for (int j = q0 + 1; j < q1; ++j)
        if (apply_adjacent(m, j, swp)) return -1;
    return 0;
}

// uploads the pointer/bond table of an MPS into `buf` and returns the device view
int make_table(b200_mps* m, DevBuf& buf, SiteTable& out) {
    b200_ctx* ctx = m->ctx;
    const int n = m->n;
    const size_t bytes = (size_t)n * (2 * sizeof(void*) + sizeof(int)) + 64;
    if (reserve(buf, bytes, ctx->stream)) return -1;
    std::vector<char> host(bytes, 0);
    const void** gp = (const void**)host.data();
    const void** lp = gp + n;
    int* cp = (int*)(lp + n);
    for (int i = 0; i < n; ++i) {
        gp[i] = m->gam[i].p;
        lp[i] = i < n - 1 ? m->lam[i].p : nullptr;
        cp[i] = m->chi[i];
    }
    CUDA_TRY(cudaMemcpyAsync(buf.p, host.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->counters[4] += bytes;
    out.gam = (const double2* const*)buf.p;
    out.lam = (const double* const*)((const void**)buf.p + n);
    out.chi = (const int*)((const void**)buf.p + 2 * n);
    out.n = n;
    return 0;
}

int max_bond(const b200_mps* m) {
    int c = 1;
    for (int x : m->chi) c = std::max(c, x);
    return c;
}

// dst (+)= diag(la) A_sa^H (env . B_sb . diag(lb)): one term of a transfer-matrix step at site i of
// <a|b>, with physical index sa on the bra and sb on the ket.  envs are row-major (chi_a x chi_b),
// bra index first; A_s = Gamma_s lambda (the "preprocessed" tensors of aqc_research).
int transfer_one(b200_ctx* ctx, const b200_mps* a, const b200_mps* b, int i, const double2* env, double2* tmp,
                 double2* dst, int sa, int sb, bool accumulate) {
    const int al = a->chiL(i), ar = a->chi[i], bl = b->chiL(i), br = b->chi[i];
    const double* la = i < a->n - 1 ? (const double*)a->lam[i].p : nullptr;
    const double* lb = i < b->n - 1 ? (const double*)b->lam[i].p : nullptr;
    GemmArgs g1 = gemm_nn(env, (const double2*)b->gam[i].p + (size_t)sb * bl * br, tmp, al, br, bl);
    g1.colscale = lb; g1.cs_mod = br;
    if (gemm(ctx, g1)) return -1;
    GemmArgs g2;
    std::memset(&g2, 0, sizeof g2);
    g2.A = (const double2*)a->gam[i].p + (size_t)sa * al * ar; g2.B = tmp; g2.C = dst;
    g2.M = ar; g2.N = br; g2.K = al;
    g2.sam = 1; g2.sak = ar; g2.conj_a = 1;
    g2.sbk = br; g2.sbn = 1; g2.scm = br; g2.scn = 1;
    g2.rowscale = la; g2.rs_mod = ar; g2.cs_mod = 1;
    g2.accumulate = accumulate ? 1 : 0;
    return gemm(ctx, g2);
}

int transfer_step(b200_ctx* ctx, const b200_mps* a, const b200_mps* b, int i, const double2* env, double2* tmp, double2* dst) {
    if (transfer_one(ctx, a, b, i, env, tmp, dst, 0, 0, false)) return -1;
    return transfer_one(ctx, a, b, i, env, tmp, dst, 1, 1, true);
}

// One step of the right sweep of <a|b> at site i:
// Gout[x,y] = sum_s sum_{x',y'} conj(A_s[x,x']) Gin[x',y'] B_s[y,y'],  A = Gamma(a) lambda(a), B likewise.
int right_step(b200_ctx* ctx, const b200_mps* a, const b200_mps* b, int i, const double2* Gin, double2* tmp, double2* Gout) {
    MpsState* st = state(ctx);
    cudaStream_t s = ctx->stream;
    const int al = a->chiL(i), ar = a->chi[i], bl = b->chiL(i), br = b->chi[i];
    const double* la = i < a->n - 1 ? (const double*)a->lam[i].p : nullptr;
    const double* lb = i < b->n - 1 ? (const double*)b->lam[i].p : nullptr;
    if (reserve(st->packA, (size_t)2 * al * ar * sizeof(double2), s)) return -1;
    if (reserve(st->packB, (size_t)2 * bl * br * sizeof(double2), s)) return -1;
    {
        MScope ms(ctx);
        mps_pack_scaled_kernel<<<grid_for((size_t)2 * al * ar), 256, 0, s>>>((const double2*)a->gam[i].p, al, ar, nullptr, la,
                                                                              (double2*)st->packA.p);
    }
    {
        MScope ms(ctx);
        mps_pack_scaled_kernel<<<grid_for((size_t)2 * bl * br), 256, 0, s>>>((const double2*)b->gam[i].p, bl, br, nullptr, lb,
                                                                              (double2*)st->packB.p);
    }
    CUDA_TRY(cudaGetLastError());
    for (int sp = 0; sp < 2; ++sp) {
        const double2* A = (const double2*)st->packA.p + (size_t)sp * al * ar;   // A_s (al x ar)
        const double2* B = (const double2*)st->packB.p + (size_t)sp * bl * br;   // B_s (bl x br)
        GemmArgs g1;   // T1 (ar x bl)[x', y] = sum_{y'} Gin[x', y'] B_s[y, y']
        std::memset(&g1, 0, sizeof g1);
        g1.A = Gin; g1.B = B; g1.C = tmp;
        g1.M = ar; g1.N = bl; g1.K = br;
        g1.sam = br; g1.sak = 1; g1.sbk = 1; g1.sbn = br;
        g1.scm = bl; g1.scn = 1; g1.rs_mod = g1.cs_mod = 1;
        if (gemm(ctx, g1)) return -1;
        GemmArgs g2;   // Gout (al x bl)[x, y] (+)= sum_{x'} conj(A_s[x, x']) T1[x', y]
        std::memset(&g2, 0, sizeof g2);
        g2.A = A; g2.B = tmp; g2.C = Gout;
        g2.M = al; g2.N = bl; g2.K = ar;
        g2.sam = ar; g2.sak = 1; g2.conj_a = 1;
        g2.sbk = bl; g2.sbn = 1; g2.scm = bl; g2.scn = 1; g2.rs_mod = g2.cs_mod = 1;
        g2.accumulate = sp;
        if (gemm(ctx, g2)) return -1;
    }
    return 0;
}

// Right environments of <a|b>: F_i (chi^a_{i-1} x chi^b_{i-1}), i = 0..n, packed back to back in `buf`.
int right_envs(b200_mps* a, b200_mps* b, DevBuf& buf, DevBuf& tmp, std::vector<size_t>& foff) {
    b200_ctx* ctx = a->ctx;
    cudaStream_t s = ctx->stream;
    const int n = a->n;
    foff.assign(n + 2, 0);
    for (int i = 0; i <= n; ++i) {
        const size_t ca = i == 0 ? 1 : a->chi[i - 1], cb = i == 0 ? 1 : b->chi[i - 1];
        foff[i + 1] = foff[i] + ca * cb;
    }
    if (reserve(buf, foff[n + 1] * sizeof(double2), s)) return -1;
    const size_t mc2 = (size_t)max_bond(a) * max_bond(b);
    if (reserve(tmp, mc2 * sizeof(double2), s)) return -1;
    double2* F = (double2*)buf.p;
    const double2 one = make_double2(1.0, 0.0);
    CUDA_TRY(cudaMemcpyAsync(F + foff[n], &one, sizeof(double2), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int i = n - 1; i >= 0; --i)
        if (right_step(ctx, a, b, i, F + foff[i + 1], (double2*)tmp.p, F + foff[i])) return -1;
    return 0;
}
int right_envs(b200_mps* m, DevBuf& buf, DevBuf& tmp, std::vector<size_t>& foff) { return right_envs(m, m, buf, tmp, foff); }

const double2 kOne = {1.0, 0.0};

int set_env_one(b200_ctx* ctx, DevBuf& b) {
    if (reserve(b, sizeof(double2), ctx->stream)) return -1;
    CUDA_TRY(cudaMemcpyAsync(b.p, &kOne, sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

}  // namespace

void b200_mps_release(b200_ctx* ctx) {
    if (!ctx->mps) return;
    MpsState* st = ctx->mps;
    for (b200_mps* m : std::vector<b200_mps*>(st->live)) b200_mps_destroy(m);
    DevBuf* all[] = {&st->packA, &st->packB, &st->theta, &st->X, &st->W, &st->sigma, &st->perm, &st->kept, &st->flag,
                     &st->table, &st->env[0], &st->env[1], &st->env[2], &st->env[3], &st->env[4], &st->env[5], &st->bits,
                     &st->outz};
    for (DevBuf* b : all) release(*b);
    delete st;
    ctx->mps = nullptr;
}

extern "C" {

int b200_mps_create(b200_ctx* ctx, int num_qubits, double truncation_threshold, int max_bond_dimension, b200_mps** out) {
    if (!ctx || !out) return set_error("null pointer");
    if (num_qubits < 1 || num_qubits > 4096) return set_error("num_qubits out of range [1,4096]");
    if (truncation_threshold < 0) return set_error("negative truncation threshold");
    CUDA_TRY(cudaSetDevice(ctx->device));
    b200_mps* m = new b200_mps();
    m->ctx = ctx; m->n = num_qubits; m->thr = truncation_threshold; m->max_chi = std::max(0, max_bond_dimension);
    m->chi.assign(num_qubits, 1);
    m->gam.resize(num_qubits);
    m->lam.resize(std::max(0, num_qubits - 1));
    state(ctx)->live.push_back(m);
    *out = m;
    return b200_mps_init_zero(m);
}

int b200_mps_destroy(b200_mps* m) {
    if (!m) return 0;
    if (m->ctx) {
        cudaSetDevice(m->ctx->device);
        cudaStreamSynchronize(m->ctx->stream);
        if (m->ctx->mps) {
            auto& live = m->ctx->mps->live;
            live.erase(std::remove(live.begin(), live.end(), m), live.end());
        }
    }
    for (auto& b : m->gam) release(b);
    for (auto& b : m->lam) release(b);
    delete m;
    return 0;
}

int b200_mps_set_truncation(b200_mps* m, double truncation_threshold, int max_bond_dimension) {
    if (check_mps(m)) return -1;
    if (truncation_threshold < 0) return set_error("negative truncation threshold");
    m->thr = truncation_threshold;
    m->max_chi = std::max(0, max_bond_dimension);
    return 0;
}

int b200_mps_init_zero(b200_mps* m) {
    if (check_mps(m)) return -1;
    b200_ctx* ctx = m->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const double2 site[2] = {{1.0, 0.0}, {0.0, 0.0}};
    const double one = 1.0;
    for (int i = 0; i < m->n; ++i) {
        m->chi[i] = 1;
        if (set_site(m, i, 1, 1)) return -1;
        CUDA_TRY(cudaMemcpyAsync(m->gam[i].p, site, sizeof site, cudaMemcpyHostToDevice, ctx->stream));
        if (i < m->n - 1) {
            if (reserve(m->lam[i], sizeof(double), ctx->stream)) return -1;
            CUDA_TRY(cudaMemcpyAsync(m->lam[i].p, &one, sizeof one, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200_mps_set(b200_mps* m, const int32_t* bond_dims, const double* gammas, const double* lambdas) {
    if (check_mps(m)) return -1;
    if (!bond_dims && m->n > 1) return set_error("null bond_dims");
    if (!gammas || (!lambdas && m->n > 1)) return set_error("null gammas / lambdas");
    b200_ctx* ctx = m->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    for (int i = 0; i + 1 < m->n; ++i)
        if (bond_dims[i] < 1 || bond_dims[i] > 65536) return set_error("bond dimension out of range at bond " + std::to_string(i));
    size_t goff = 0, loff = 0;
    for (int i = 0; i < m->n; ++i) {
        m->chi[i] = i < m->n - 1 ? bond_dims[i] : 1;
        const int cl = m->chiL(i), cr = m->chi[i];
        if (set_site(m, i, cl, cr)) return -1;
        const size_t cnt = (size_t)2 * cl * cr;
        CUDA_TRY(cudaMemcpyAsync(m->gam[i].p, gammas + 2 * goff, cnt * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
        goff += cnt;
        ctx->counters[4] += cnt * sizeof(double2);
        if (i < m->n - 1) {
            if (reserve(m->lam[i], (size_t)cr * sizeof(double), ctx->stream)) return -1;
            CUDA_TRY(cudaMemcpyAsync(m->lam[i].p, lambdas + loff, cr * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            loff += cr;
            ctx->counters[4] += cr * sizeof(double);
        }
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200_mps_num_qubits(b200_mps* m, int* out) {
    if (check_mps(m) || !out) return set_error("null pointer");
    *out = m->n;
    return 0;
}

int b200_mps_bond_dims(b200_mps* m, int32_t* out) {
    if (check_mps(m)) return -1;
    if (!out && m->n > 1) return set_error("null pointer");
    for (int i = 0; i + 1 < m->n; ++i) out[i] = m->chi[i];
    return 0;
}

int b200_mps_get(b200_mps* m, double* gammas, double* lambdas) {
    if (check_mps(m)) return -1;
    if (!gammas || (!lambdas && m->n > 1)) return set_error("null pointer");
    b200_ctx* ctx = m->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    size_t goff = 0, loff = 0;
    for (int i = 0; i < m->n; ++i) {
        const size_t cnt = (size_t)2 * m->chiL(i) * m->chi[i];
        CUDA_TRY(cudaMemcpyAsync(gammas + 2 * goff, m->gam[i].p, cnt * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
        goff += cnt;
        ctx->counters[5] += cnt * sizeof(double2);
        if (i < m->n - 1) {
            CUDA_TRY(cudaMemcpyAsync(lambdas + loff, m->lam[i].p, m->chi[i] * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            loff += m->chi[i];
            ctx->counters[5] += m->chi[i] * sizeof(double);
        }
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200_mps_copy(b200_mps* dst, b200_mps* src) {
    if (check_mps(dst) || check_mps(src)) return -1;
    if (dst->ctx != src->ctx || dst->n != src->n) return set_error("mps_copy: handles differ in context or size");
    if (dst == src) return 0;
    b200_ctx* ctx = dst->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    for (int i = 0; i < src->n; ++i) {
        dst->chi[i] = src->chi[i];
        const size_t bytes = (size_t)2 * src->chiL(i) * src->chi[i] * sizeof(double2);
        if (reserve(dst->gam[i], bytes, ctx->stream)) return -1;
        CUDA_TRY(cudaMemcpyAsync(dst->gam[i].p, src->gam[i].p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        if (i < src->n - 1) {
            if (reserve(dst->lam[i], src->chi[i] * sizeof(double), ctx->stream)) return -1;
            CUDA_TRY(cudaMemcpyAsync(dst->lam[i].p, src->lam[i].p, src->chi[i] * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    return 0;
}

static int mps_apply_impl(b200_mps* m, const b200_gate* gates, int n_gates, const double* mats, int n_mats, bool inverse) {
    if (check_mps(m)) return -1;
    if (n_gates < 0 || (n_gates > 0 && !gates)) return set_error("null gate array");
    b200_ctx* ctx = m->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx->counters[6] += 1;
    for (int kk = 0; kk < n_gates; ++kk) {
        const int k = inverse ? n_gates - 1 - kk : kk;
        cplx u[16];
        int nq = 0, qa = -1, qb = -1;
        if (gate_matrix(m->n, gates[k], mats, n_mats, u, nq, qa, qb)) return -1;
        if (nq == 0) continue;
        if (inverse) {   // conjugate transpose
            const int d = nq == 1 ? 2 : 4;
            cplx t[16];
            for (int r = 0; r < d; ++r)
                for (int c = 0; c < d; ++c) t[d * r + c] = std::conj(u[d * c + r]);
            for (int e = 0; e < d * d; ++e) u[e] = t[e];
        }
        if (nq == 1) { if (apply_1q(m, qa, u)) return -1; }
        else if (apply_2q(m, qa, qb, u)) return -1;
        ctx->counters[2] += 1;
    }
    return 0;
}

int b200_mps_apply(b200_mps* m, const b200_gate* gates, int n_gates, const double* mats, int n_mats) {
    return mps_apply_impl(m, gates, n_gates, mats, n_mats, false);
}

int b200_mps_apply_inverse(b200_mps* m, const b200_gate* gates, int n_gates, const double* mats, int n_mats) {
    return mps_apply_impl(m, gates, n_gates, mats, n_mats, true);
}

// out = T[i][j] = <a| (|i><j| on `qubits`) |b>, i = bra, j = ket physical indices (bit(qubits[0]) +
// 2 bit(qubits[1])); n_open = 0 (plain <a|b>, 1 complex), 1 (2x2) or 2 (4x4), row-major.
int b200_mps_transfer(b200_mps* a, b200_mps* b, const int32_t* qubits, int n_open, double* out) {
    if (check_mps(a) || check_mps(b)) return -1;
    if (a->ctx != b->ctx || a->n != b->n) return set_error("mps_transfer: handles differ in context or size");
    if (!out || n_open < 0 || n_open > 2 || (n_open > 0 && !qubits)) return set_error("mps_transfer: bad arguments");
    if (n_open == 0) return b200_mps_dot(a, b, out);
    b200_ctx* ctx = a->ctx;
    MpsState* st = state(ctx);
    cudaStream_t s = ctx->stream;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int n = a->n;
    for (int k = 0; k < n_open; ++k)
        if (qubits[k] < 0 || qubits[k] >= n) return set_error("mps_transfer: qubit out of range");
    if (n_open == 2 && qubits[0] == qubits[1]) return set_error("mps_transfer: qubits must differ");
    const int lo = n_open == 1 ? qubits[0] : std::min(qubits[0], qubits[1]);
    const int hi = n_open == 1 ? qubits[0] : std::max(qubits[0], qubits[1]);
    const size_t mc2 = (size_t)max_bond(a) * max_bond(b);
    for (int k : {0, 1, 2, 3})
        if (reserve(st->env[k], std::max(mc2, (size_t)1) * sizeof(double2), s)) return -1;
    if (reserve(st->env[4], 8 * mc2 * sizeof(double2), s)) return -1;
    if (reserve(st->env[5], 2 * mc2 * sizeof(double2), s)) return -1;
    if (reserve(st->outz, 16 * sizeof(double2), s)) return -1;
    double2* tmp = (double2*)st->env[2].p;
    // right environment down to site hi+1 (ping-pong inside env[5])
    double2* G0 = (double2*)st->env[5].p; double2* G1 = G0 + mc2;
    const double2 one = make_double2(1.0, 0.0);
    CUDA_TRY(cudaMemcpyAsync(G0, &one, sizeof(double2), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int i = n - 1; i > hi; --i) {
        if (right_step(ctx, a, b, i, G0, tmp, G1)) return -1;
        std::swap(G0, G1);
    }
    // left environment up to site lo-1
    if (set_env_one(ctx, st->env[0])) return -1;
    int cur = 0;
    for (int i = 0; i < lo; ++i) {
        if (transfer_step(ctx, a, b, i, (const double2*)st->env[cur].p, tmp, (double2*)st->env[1 - cur].p)) return -1;
        cur = 1 - cur;
    }
    const double2* E = (const double2*)st->env[cur].p;
    double2* O = (double2*)st->env[4].p; double2* O2 = O + 4 * mc2;
    for (int t = 0; t < 2; ++t)
        for (int sb = 0; sb < 2; ++sb)
            if (transfer_one(ctx, a, b, lo, E, tmp, O + (size_t)(sb + 2 * t) * mc2, sb, t, false)) return -1;
    double2 host[16];
    if (n_open == 1) {
        const int nel = a->chi[lo] * b->chi[lo];
        for (int t = 0; t < 2; ++t)
            for (int sb = 0; sb < 2; ++sb) {
                MScope ms(ctx);
                mps_dot_elem_kernel<<<1, 256, 0, s>>>(O + (size_t)(sb + 2 * t) * mc2, G0, nel, (double2*)st->outz.p + 2 * sb + t);
            }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(host, st->outz.p, 4 * sizeof(double2), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        for (int k = 0; k < 4; ++k) { out[2 * k] = host[k].x; out[2 * k + 1] = host[k].y; }
        ctx->counters[5] += 4 * sizeof(double2);
        ctx->counters[6] += 1;
        return 0;
    }
    for (int site = lo + 1; site < hi; ++site) {
        for (int k = 0; k < 4; ++k)
            if (transfer_step(ctx, a, b, site, O + (size_t)k * mc2, tmp, O2 + (size_t)k * mc2)) return -1;
        std::swap(O, O2);
    }
    double2* Y = (double2*)st->env[3].p;
    const int nel = a->chi[hi] * b->chi[hi];
    for (int k = 0; k < 4; ++k)
        for (int tp = 0; tp < 2; ++tp)
            for (int sp = 0; sp < 2; ++sp) {
                if (transfer_one(ctx, a, b, hi, O + (size_t)k * mc2, tmp, Y, sp, tp, false)) return -1;
                const int sb = k & 1, t = k >> 1;
                const int bra = sb + 2 * sp, ket = t + 2 * tp;
                {
                    MScope ms(ctx);
                    mps_dot_elem_kernel<<<1, 256, 0, s>>>(Y, G0, nel, (double2*)st->outz.p + 4 * bra + ket);
                }
                CUDA_TRY(cudaGetLastError());
            }
    CUDA_TRY(cudaMemcpyAsync(host, st->outz.p, 16 * sizeof(double2), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    ctx->counters[5] += 16 * sizeof(double2);
    ctx->counters[6] += 1;
    const bool swapped = qubits[0] > qubits[1];
    auto sw2 = [](int i) { return ((i & 1) << 1) | (i >> 1); };
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const int ii = swapped ? sw2(i) : i, jj = swapped ? sw2(j) : j;
            out[2 * (4 * ii + jj)] = host[4 * i + j].x;
            out[2 * (4 * ii + jj) + 1] = host[4 * i + j].y;
        }
    return 0;
}

int b200_mps_set_chop_rule(int rule) {
    if (rule != 0 && rule != 1) return set_error("chop rule must be B200_CHOP_AER (0) or B200_CHOP_SIGMA (1)");
    g_chop_rule = rule;
    return 0;
}

int b200_mps_reduce_zeros(const double* s_desc, int n, int max_bond_dimension, double truncation_threshold, int* n_kept,
                          double* kept_out) {
    if (!s_desc || !n_kept || !kept_out || n <= 0) return set_error("b200_mps_reduce_zeros: null or empty input");
    std::vector<double> S(s_desc, s_desc + n), kept;
    *n_kept = reduce_zeros(S, max_bond_dimension, truncation_threshold, kept);
    for (int i = 0; i < *n_kept; ++i) kept_out[i] = kept[i];
    return 0;
}

int b200_mps_stats(b200_mps* m, uint64_t out[4]) {
    if (check_mps(m) || !out) return set_error("null pointer");
    out[0] = m->svd_count; out[1] = m->svd_sweeps; out[2] = (uint64_t)max_bond(m); out[3] = m->svd_flops;
    return 0;
}

int b200_mps_amps(b200_mps* m, const uint64_t* bitstrings, int count, double* out) {
    if (check_mps(m)) return -1;
    if (count < 0 || (count > 0 && (!bitstrings || !out))) return set_error("null pointer");
    if (m->n > 64) return set_error("bitstring amplitudes are limited to 64 qubits");
    if (count == 0) return 0;
    b200_ctx* ctx = m->ctx;
    MpsState* st = state(ctx);
    CUDA_TRY(cudaSetDevice(ctx->device));
    SiteTable tab;
    if (make_table(m, st->table, tab)) return -1;
    if (reserve(st->bits, (size_t)count * sizeof(uint64_t), ctx->stream)) return -1;
    if (reserve(st->outz, (size_t)count * sizeof(double2), ctx->stream)) return -1;
    CUDA_TRY(cudaMemcpyAsync(st->bits.p, bitstrings, count * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    ctx->counters[4] += count * sizeof(uint64_t);
    const int maxchi = max_bond(m);
    const size_t smem = (size_t)(2 + AMPS_SLICES) * maxchi * sizeof(double2);
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(mps_amps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        MScope ms(ctx);
        mps_amps_kernel<<<count, AMPS_THREADS, smem, ctx->stream>>>(tab, (const uint64_t*)st->bits.p, maxchi, (double2*)st->outz.p);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, st->outz.p, count * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->counters[5] += count * sizeof(double2);
    ctx->counters[6] += 1;
    return 0;
}

int b200_mps_dot(b200_mps* a, b200_mps* b, double out[2]) {
    if (check_mps(a) || check_mps(b)) return -1;
    if (a->ctx != b->ctx || a->n != b->n) return set_error("mps_dot: handles differ in context or size");
    if (!out) return set_error("null pointer");
    b200_ctx* ctx = a->ctx;
    MpsState* st = state(ctx);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t cap = (size_t)max_bond(a) * max_bond(b) * sizeof(double2);
    for (int k = 0; k < 3; ++k)
        if (reserve(st->env[k], std::max(cap, sizeof(double2)), ctx->stream)) return -1;
    if (set_env_one(ctx, st->env[0])) return -1;
    int cur = 0;
    for (int i = 0; i < a->n; ++i) {
        if (transfer_step(ctx, a, b, i, (const double2*)st->env[cur].p, (double2*)st->env[2].p, (double2*)st->env[1 - cur].p))
            return -1;
        cur = 1 - cur;
    }
    double2 r;
    CUDA_TRY(cudaMemcpyAsync(&r, st->env[cur].p, sizeof r, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->counters[5] += sizeof r;
    ctx->counters[6] += 1;
    out[0] = r.x; out[1] = r.y;
    return 0;
}

// <Z_q> for every qubit + <psi|psi> from one right sweep (stores F_i) and one left sweep.
// Full contractions (no canonical-form shortcut), like aqc_research.mps_expectation
// (adaptaqc/backends/aer_mps_backend.py:80-86), but all n values from two sweeps instead of n.
int b200_mps_expz(b200_mps* m, double* out) {
    if (check_mps(m) || !out) return set_error("null pointer");
    b200_ctx* ctx = m->ctx;
    MpsState* st = state(ctx);
    cudaStream_t s = ctx->stream;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int n = m->n;
    const size_t mc = max_bond(m);
    for (int k : {0, 1, 2, 4, 5})
        if (reserve(st->env[k], mc * mc * sizeof(double2), s)) return -1;
    if (reserve(st->outz, (size_t)(n + 1) * sizeof(double2), s)) return -1;
    std::vector<size_t> foff;
    if (right_envs(m, st->env[3], st->env[2], foff)) return -1;
    const double2* F = (const double2*)st->env[3].p;
    // left sweep: X_s = A_s^H E_i A_s ; z_i = <X_0 - X_1, F_{i+1}> ; E_{i+1} = X_0 + X_1
    if (set_env_one(ctx, st->env[0])) return -1;
    int cur = 0;
    for (int i = 0; i < n; ++i) {
        const int cr = m->chi[i];
        double2* X0 = (double2*)st->env[4].p; double2* X1 = (double2*)st->env[5].p;
        if (transfer_one(ctx, m, m, i, (const double2*)st->env[cur].p, (double2*)st->env[2].p, X0, 0, 0, false)) return -1;
        if (transfer_one(ctx, m, m, i, (const double2*)st->env[cur].p, (double2*)st->env[2].p, X1, 1, 1, false)) return -1;
        double2* En = (double2*)st->env[1 - cur].p;
        {
            MScope ms(ctx);
            mps_sum_diff_kernel<<<grid_for((size_t)cr * cr), 256, 0, s>>>(X0, X1, cr * cr, En, X0);   // X0 <- X0 - X1
        }
        {
            MScope ms(ctx);
            mps_dot_elem_kernel<<<1, 256, 0, s>>>(X0, F + foff[i + 1], cr * cr, (double2*)st->outz.p + i);
        }
        CUDA_TRY(cudaGetLastError());
        cur = 1 - cur;
    }
    CUDA_TRY(cudaMemcpyAsync((double2*)st->outz.p + n, st->env[cur].p, sizeof(double2), cudaMemcpyDeviceToDevice, s));
    std::vector<double2> host(n + 1);
    CUDA_TRY(cudaMemcpyAsync(host.data(), st->outz.p, (n + 1) * sizeof(double2), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    ctx->counters[5] += (n + 1) * sizeof(double2);
    ctx->counters[6] += 1;
    for (int i = 0; i <= n; ++i) out[i] = host[i].x;
    return 0;
}

// Shared body of b200_mps_pair_rdm / b200_mps_pair_transfer: for every requested pair, the 16 values
//   V[(s,s'),(t,t')] = <a| (|s s'><t t'| on (lo, hi)) |b>      (s, t: physical indices at lo; s', t' at hi; s = bra)
// written at out16[p] + (transposed ? 4*ket + bra : 4*bra + ket), bra = s + 2 s', ket = t + 2 t'.
// Left/right environments of <a|b> are built once; pairs sharing their lower qubit share the propagation of the
// four open environments.
static int pair_open_values(b200_mps* ma, b200_mps* mb, const int32_t* pairs, int n_pairs, bool transposed, double* out) {
    b200_ctx* ctx = ma->ctx;
    MpsState* st = state(ctx);
    cudaStream_t s = ctx->stream;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int n = ma->n;
    for (int p = 0; p < n_pairs; ++p) {
        const int a = pairs[2 * p], b = pairs[2 * p + 1];
        if (a < 0 || b < 0 || a >= n || b >= n || a == b) return set_error("pair " + std::to_string(p) + ": qubits out of range");
    }
    const size_t mc2 = (size_t)max_bond(ma) * max_bond(mb);
    // env[0], env[1]: left env ping-pong; env[2]: GEMM temp; env[3]: all F; env[4]: 4 open envs (+4 ping-pong);
    // env[5]: closing temps
    for (int k : {0, 1, 2})
        if (reserve(st->env[k], mc2 * sizeof(double2), s)) return -1;
    if (reserve(st->env[4], 8 * mc2 * sizeof(double2), s)) return -1;
    if (reserve(st->env[5], 2 * mc2 * sizeof(double2), s)) return -1;
    if (reserve(st->outz, (size_t)16 * n_pairs * sizeof(double2), s)) return -1;
    std::vector<size_t> foff;
    if (right_envs(ma, mb, st->env[3], st->env[2], foff)) return -1;
    const double2* F = (const double2*)st->env[3].p;

    // group the requested pairs by their lower qubit
    std::vector<std::vector<std::pair<int, int>>> by_lo(n);   // (hi, request index)
    int max_lo = -1;
    for (int p = 0; p < n_pairs; ++p) {
        const int lo = std::min(pairs[2 * p], pairs[2 * p + 1]), hi = std::max(pairs[2 * p], pairs[2 * p + 1]);
        by_lo[lo].push_back({hi, p});
        max_lo = std::max(max_lo, lo);
    }
    if (set_env_one(ctx, st->env[0])) return -1;
    int cur = 0;
    double2* tmp = (double2*)st->env[2].p;
    for (int lo = 0; lo <= max_lo; ++lo) {
        const double2* E = (const double2*)st->env[cur].p;
        if (!by_lo[lo].empty()) {
            int far = lo;
            for (auto& hp : by_lo[lo]) far = std::max(far, hp.first);
            // open site lo: O[s + 2t] (s = bra, t = ket physical index)
            double2* O = (double2*)st->env[4].p;
            double2* O2 = O + 4 * mc2;
            for (int t = 0; t < 2; ++t)
                for (int sb = 0; sb < 2; ++sb)
                    if (transfer_one(ctx, ma, mb, lo, E, tmp, O + (size_t)(sb + 2 * t) * mc2, sb, t, false)) return -1;
            for (int site = lo + 1; site <= far; ++site) {
                for (auto& hp : by_lo[lo]) {
                    if (hp.first != site) continue;
                    const int nel = ma->chi[site] * mb->chi[site];
                    // close at `site`: value[(s,t),(s',t')] = < A_{s'}^H O_{st} B_{t'}, F_{site+1} >
                    for (int st_ = 0; st_ < 4; ++st_)
                        for (int tp = 0; tp < 2; ++tp)
                            for (int sp = 0; sp < 2; ++sp) {
                                double2* Y = (double2*)st->env[5].p;
                                if (transfer_one(ctx, ma, mb, site, O + (size_t)st_ * mc2, tmp, Y, sp, tp, false)) return -1;
                                const int sb = st_ & 1, t = st_ >> 1;
                                const int ket = t + 2 * tp, bra = sb + 2 * sp;
                                {
                                    MScope ms(ctx);
                                    mps_dot_elem_kernel<<<1, 256, 0, s>>>(Y, F + foff[site + 1], nel,
                                        (double2*)st->outz.p + (size_t)16 * hp.second + (transposed ? 4 * ket + bra : 4 * bra + ket));
                                }
                                CUDA_TRY(cudaGetLastError());
                            }
                }
                if (site < far) {   // propagate the four open environments through `site`
                    for (int st_ = 0; st_ < 4; ++st_)
                        if (transfer_step(ctx, ma, mb, site, O + (size_t)st_ * mc2, tmp, O2 + (size_t)st_ * mc2)) return -1;
                    std::swap(O, O2);
                }
            }
        }
        if (lo < max_lo) {
            if (transfer_step(ctx, ma, mb, lo, E, tmp, (double2*)st->env[1 - cur].p)) return -1;
            cur = 1 - cur;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(out, st->outz.p, (size_t)16 * n_pairs * sizeof(double2), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    ctx->counters[5] += (size_t)16 * n_pairs * sizeof(double2);
    ctx->counters[6] += 1;
    return 0;
}

// out[p] = 4x4 reduced density matrix (row-major, 32 doubles) of pairs[2p], pairs[2p+1]; the lower
// qubit is the least-significant index; rho[ket][bra].  Replaces aqc_research partial_trace
// (adaptaqc/utils/entanglement_measures.py:76-79), once per candidate pair per layer
// (adaptaqc/compilers/adapt/adapt_compiler.py:960-975).
int b200_mps_pair_rdm(b200_mps* m, const int32_t* pairs, int n_pairs, double* out) {
    if (check_mps(m)) return -1;
    if (n_pairs < 0 || (n_pairs > 0 && (!pairs || !out))) return set_error("null pointer");
    if (n_pairs == 0) return 0;
    return pair_open_values(m, m, pairs, n_pairs, true, out);
}

// out[p] = T_p[i][j] = <a| (|i><j| on pairs[2p], pairs[2p+1]) |b>, i = bra, j = ket, index = bit(pairs[2p]) +
// 2 bit(pairs[2p+1]) (the order the pair is given in), row-major 4x4 complex.  All pairs from ONE left and ONE right
// environment sweep of <a|b>.  With <a| = <s| (starting state) and |b> = |psi>, sum_ij O[i][j] T_p[i][j] = <s|O_p|psi>
// for any two-qubit operator O: every (pair, generator) overlap of general_grad_of_pairs
// (adaptaqc/utils/gradients.py:23-124) is 4x4 host algebra on these matrices.
int b200_mps_pair_transfer(b200_mps* a, b200_mps* b, const int32_t* pairs, int n_pairs, double* out) {
    if (check_mps(a) || check_mps(b)) return -1;
    if (a->ctx != b->ctx || a->n != b->n) return set_error("mps_pair_transfer: handles differ in context or size");
    if (n_pairs < 0 || (n_pairs > 0 && (!pairs || !out))) return set_error("null pointer");
    if (n_pairs == 0) return 0;
    if (pair_open_values(a, b, pairs, n_pairs, false, out)) return -1;
    auto sw2 = [](int i) { return ((i & 1) << 1) | (i >> 1); };
    for (int p = 0; p < n_pairs; ++p) {
        if (pairs[2 * p] < pairs[2 * p + 1]) continue;      // given as (hi, lo): re-index bit(first) + 2 bit(second)
        double t[32];
        double* o = out + (size_t)32 * p;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                t[2 * (4 * sw2(i) + sw2(j))] = o[2 * (4 * i + j)];
                t[2 * (4 * sw2(i) + sw2(j)) + 1] = o[2 * (4 * i + j) + 1];
            }
        std::memcpy(o, t, sizeof t);
    }
    return 0;
}

}  // extern "C"
