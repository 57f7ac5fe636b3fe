// mps_kernels.cuh -- sm_100a kernels of the matrix-product-state path (complex128).
//
//   zgemm_dmma_kernel     K8/K10: complex GEMM on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64;
//                         tcgen05 has no f64 kind).  Generic element strides + optional conjugation
//                         of A, so the same kernel serves the two-site contraction  (lambda Gamma
//                         lambda)(Gamma lambda)  and every transfer-matrix step  A^H . E . B.
//   mps_pack_*            lambda scaling + packing of site tensors into GEMM operands.
//   mps_theta_gate_kernel applies the 4x4 gate to the contracted two-site tensor and lays the
//                         (2 chi_l x 2 chi_r) matrix out column-major (tall orientation) for the SVD.
//   jacobi_*              K9: one-sided (Hestenes) Jacobi SVD, complex128, round-robin ordering.
//                         jacobi_cta_kernel does a whole SVD in one CTA (small bonds: one launch per
//                         gate); jacobi_round_kernel is one tournament round with one CTA per
//                         column pair (large bonds).
//   mps_write_sites_kernel truncated U / V^H -> Gamma_i, Gamma_{i+1}, lambda_i (divides the outer
//                         lambdas back out, as Aer does).
//   mps_amps_kernel       batched <bitstring|psi> (one CTA per bitstring, whole chain in one launch).
//   mps_apply1q_kernel, mps_dot_pairs_kernel, small helpers.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace b200 {

__device__ __forceinline__ double2 z_mul(const double2 a, const double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 z_fma(const double2 a, const double2 b, double2 c) {
    c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
    return c;
}
__device__ __forceinline__ double2 z_conj(const double2 a) { return make_double2(a.x, -a.y); }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// ---------------------------------------------------------------------------------------------
// K8/K10: complex GEMM on FP64 tensor cores
// ---------------------------------------------------------------------------------------------
// C(m,n) = rowscale[m % rs_mod] * colscale[n % cs_mod] * sum_k opA(m,k) * B(k,n)   (+ C if accumulate)
//   opA(m,k) = A[m*sam + k*sak]  (conjugated if conj_a);   B(k,n) = B[k*sbk + n*sbn];
//   C(m,n) at C[m*scm + n*scn].   rowscale / colscale may be null.
struct GemmArgs {
    const double2* A; const double2* B; double2* C;
    int M, N, K;
    long long sam, sak, sbk, sbn, scm, scn;
    int conj_a, accumulate;
    const double* rowscale; int rs_mod;
    const double* colscale; int cs_mod;
};

constexpr int GM_TILE = 32;   // CTA tile 32 x 32, 4 warps, each warp a 16 x 16 sub-tile
constexpr int GK_TILE = 16;
constexpr int G_PAD = 1;

__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128)
zgemm_dmma_kernel(const GemmArgs g) {
    __shared__ double As_re[GM_TILE][GK_TILE + G_PAD], As_im[GM_TILE][GK_TILE + G_PAD];
    __shared__ double Bs_re[GK_TILE][GM_TILE + G_PAD], Bs_im[GK_TILE][GM_TILE + G_PAD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * GM_TILE, n0 = blockIdx.x * GM_TILE;
    const int wm = (warp >> 1) * 16, wn = (warp & 1) * 16;   // warp sub-tile origin inside the CTA tile
    double cre[2][2][2], cim[2][2][2];                         // [m frag][n frag][2 columns]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) { cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0; }

    // The next K tile travels global -> registers while the DMMAs of the current one are issued (these GEMMs
    // are skinny: few CTAs, one K loop of 16-64 tiles each -- exposed load latency was most of the time).
    double2 pa[4], pb[4];
    auto fetch = [&](const int k0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int e = tid + 128 * t;
            const int r = e / GK_TILE, c = e % GK_TILE;
            pa[t] = make_double2(0.0, 0.0);
            if (m0 + r < g.M && k0 + c < g.K) pa[t] = g.A[(long long)(m0 + r) * g.sam + (long long)(k0 + c) * g.sak];
            const int rb = e / GM_TILE, cb = e % GM_TILE;
            pb[t] = make_double2(0.0, 0.0);
            if (k0 + rb < g.K && n0 + cb < g.N) pb[t] = g.B[(long long)(k0 + rb) * g.sbk + (long long)(n0 + cb) * g.sbn];
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int e = tid + 128 * t;
            As_re[e / GK_TILE][e % GK_TILE] = pa[t].x;
            As_im[e / GK_TILE][e % GK_TILE] = g.conj_a ? -pa[t].y : pa[t].y;
            Bs_re[e / GM_TILE][e % GM_TILE] = pb[t].x;
            Bs_im[e / GM_TILE][e % GM_TILE] = pb[t].y;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < g.K; k0 += GK_TILE) {
        stage();
        __syncthreads();
        if (k0 + GK_TILE < g.K) fetch(k0 + GK_TILE);
#pragma unroll
        for (int kk = 0; kk < GK_TILE; kk += 4) {
            double are[2], aim[2], bre[2], bim[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                are[i] = As_re[wm + 8 * i + (lane >> 2)][kk + (lane & 3)];
                aim[i] = As_im[wm + 8 * i + (lane >> 2)][kk + (lane & 3)];
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                bre[j] = Bs_re[kk + (lane & 3)][wn + 8 * j + (lane >> 2)];
                bim[j] = Bs_im[kk + (lane & 3)][wn + 8 * j + (lane >> 2)];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    dmma884(cre[i][j][0], cre[i][j][1], are[i], bre[j]);
                    dmma884(cre[i][j][0], cre[i][j][1], -aim[i], bim[j]);
                    dmma884(cim[i][j][0], cim[i][j][1], are[i], bim[j]);
                    dmma884(cim[i][j][0], cim[i][j][1], aim[i], bre[j]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int m = m0 + wm + 8 * i + (lane >> 2), n = n0 + wn + 8 * j + 2 * (lane & 3) + c;
                if (m < g.M && n < g.N) {
                    double s = 1.0;
                    if (g.rowscale) s *= g.rowscale[m % g.rs_mod];
                    if (g.colscale) s *= g.colscale[n % g.cs_mod];
                    double2 v = make_double2(cre[i][j][c] * s, cim[i][j][c] * s);
                    double2* dst = g.C + (long long)m * g.scm + (long long)n * g.scn;
                    if (g.accumulate) { const double2 o = *dst; v.x += o.x; v.y += o.y; }
                    *dst = v;
                }
            }
}

// Large-problem variant: 64 x 64 CTA tile, 8 warps (4 x 2, each a 16 x 32 sub-tile = 2 x 4 DMMA fragments,
// 32 accumulator doubles per thread), K tile 16.  The next K tile is fetched from global memory into
// registers BEFORE the DMMAs of the current one are issued and written to shared memory after them, so
// the global-load latency hides behind the tensor work.  Row strides 20 / 68 doubles make every half-warp
// fragment load hit 16 distinct 8-byte banks.  Per k4-step and warp: 12 LDS.64 feed 32 DMMAs.
constexpr int G2_TILE = 64;
constexpr int G2_K = 16;
constexpr int G2_LDA = G2_K + 4;      // 20
constexpr int G2_LDB = G2_TILE + 4;   // 68

__global__ void __launch_bounds__(256, 2)
zgemm_dmma64_kernel(const GemmArgs g) {
    __shared__ double As_re[G2_TILE][G2_LDA], As_im[G2_TILE][G2_LDA];
    __shared__ double Bs_re[G2_K][G2_LDB], Bs_im[G2_K][G2_LDB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * G2_TILE, n0 = blockIdx.x * G2_TILE;
    const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;
    double cre[2][4][2], cim[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0; }

    // element (r, c) of the A tile handled by this thread in pass t: the unit-stride dimension runs fastest
    const bool a_k_fast = g.sak <= g.sam;     // A(m,k): k contiguous (row-major) or m contiguous
    const bool b_n_fast = g.sbn <= g.sbk;
    double2 pa[4], pb[4];
    auto fetch = [&](const int k0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int e = tid + 256 * t;
            const int r = a_k_fast ? e / G2_K : e % G2_TILE, c = a_k_fast ? e % G2_K : e / G2_TILE;
            pa[t] = make_double2(0.0, 0.0);
            if (m0 + r < g.M && k0 + c < g.K) pa[t] = g.A[(long long)(m0 + r) * g.sam + (long long)(k0 + c) * g.sak];
            const int rb = b_n_fast ? e / G2_TILE : e % G2_K, cb = b_n_fast ? e % G2_TILE : e / G2_K;
            pb[t] = make_double2(0.0, 0.0);
            if (k0 + rb < g.K && n0 + cb < g.N) pb[t] = g.B[(long long)(k0 + rb) * g.sbk + (long long)(n0 + cb) * g.sbn];
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int e = tid + 256 * t;
            const int r = a_k_fast ? e / G2_K : e % G2_TILE, c = a_k_fast ? e % G2_K : e / G2_TILE;
            As_re[r][c] = pa[t].x;
            As_im[r][c] = g.conj_a ? -pa[t].y : pa[t].y;
            const int rb = b_n_fast ? e / G2_TILE : e % G2_K, cb = b_n_fast ? e % G2_TILE : e / G2_K;
            Bs_re[rb][cb] = pb[t].x;
            Bs_im[rb][cb] = pb[t].y;
        }
    };
    fetch(0);
    stage();
    __syncthreads();
    for (int k0 = 0; k0 < g.K; k0 += G2_K) {
        const bool more = k0 + G2_K < g.K;
        if (more) fetch(k0 + G2_K);                 // in flight during the DMMAs below
#pragma unroll
        for (int kk = 0; kk < G2_K; kk += 4) {
            double are[2], aim[2], bre[4], bim[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                are[i] = As_re[wm + 8 * i + (lane >> 2)][kk + (lane & 3)];
                aim[i] = As_im[wm + 8 * i + (lane >> 2)][kk + (lane & 3)];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bre[j] = Bs_re[kk + (lane & 3)][wn + 8 * j + (lane >> 2)];
                bim[j] = Bs_im[kk + (lane & 3)][wn + 8 * j + (lane >> 2)];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma884(cre[i][j][0], cre[i][j][1], are[i], bre[j]);
                    dmma884(cre[i][j][0], cre[i][j][1], -aim[i], bim[j]);
                    dmma884(cim[i][j][0], cim[i][j][1], are[i], bim[j]);
                    dmma884(cim[i][j][0], cim[i][j][1], aim[i], bre[j]);
                }
        }
        __syncthreads();
        if (more) { stage(); __syncthreads(); }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int m = m0 + wm + 8 * i + (lane >> 2), n = n0 + wn + 8 * j + 2 * (lane & 3) + c;
                if (m < g.M && n < g.N) {
                    double sc = 1.0;
                    if (g.rowscale) sc *= g.rowscale[m % g.rs_mod];
                    if (g.colscale) sc *= g.colscale[n % g.cs_mod];
                    double2 v = make_double2(cre[i][j][c] * sc, cim[i][j][c] * sc);
                    double2* dst = g.C + (long long)m * g.scm + (long long)n * g.scn;
                    if (g.accumulate) { const double2 o = *dst; v.x += o.x; v.y += o.y; }
                    *dst = v;
                }
            }
}

// ---------------------------------------------------------------------------------------------
// site-tensor helpers.  Gamma_i is stored [2][chi_l][chi_r] (physical index outermost).
// ---------------------------------------------------------------------------------------------
// Gamma^{s'} = sum_s G[s'][s] Gamma^{s}   (G row-major 2x2 complex in gm[8])
__global__ void mps_apply1q_kernel(double2* __restrict__ gam, const int sz, const double2 g00, const double2 g01,
                                   const double2 g10, const double2 g11) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < sz; e += gridDim.x * blockDim.x) {
        const double2 a = gam[e], b = gam[sz + e];
        gam[e] = z_fma(g01, b, z_mul(g00, a));
        gam[sz + e] = z_fma(g11, b, z_mul(g10, a));
    }
}

// out[(s,a), b] = ll[a] * Gamma[s][a][b] * lm[b]     (ll / lm may be null = ones)
__global__ void mps_pack_scaled_kernel(const double2* __restrict__ gam, const int chi_l, const int chi_r,
                                       const double* __restrict__ ll, const double* __restrict__ lr,
                                       double2* __restrict__ out) {
    const int total = 2 * chi_l * chi_r;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int b = e % chi_r, a = (e / chi_r) % chi_l;
        const double s = (ll ? ll[a] : 1.0) * (lr ? lr[b] : 1.0);
        const double2 v = gam[e];
        out[e] = make_double2(v.x * s, v.y * s);
    }
}

// C: raw two-site tensor, row-major (2 chi_l) x (2 chi_r), rows (b, alpha), cols (b', gamma).
// M[(p,alpha),(q,gamma)] = sum_{b,b'} U[p + 2q][b + 2b'] C[(b,alpha),(b',gamma)].
// X = M column-major (tall == 1, leading dimension m = 2 chi_l) or X = M^H column-major
// (tall == 0, leading dimension nn = 2 chi_r).
struct Gate4 { double2 u[16]; };
__global__ void mps_theta_gate_kernel(const double2* __restrict__ C, const int chi_l, const int chi_r,
                                      const Gate4 U, const int tall, double2* __restrict__ X) {
    const int total = chi_l * chi_r;
    const int m = 2 * chi_l, nn = 2 * chi_r;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int gm = e % chi_r, al = e / chi_r;
        double2 c[4];  // index b + 2 b'
#pragma unroll
        for (int bp = 0; bp < 2; ++bp)
#pragma unroll
            for (int b = 0; b < 2; ++b)
                c[b + 2 * bp] = C[(size_t)(b * chi_l + al) * nn + bp * chi_r + gm];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                double2 v = z_mul(U.u[4 * (p + 2 * q)], c[0]);
                v = z_fma(U.u[4 * (p + 2 * q) + 1], c[1], v);
                v = z_fma(U.u[4 * (p + 2 * q) + 2], c[2], v);
                v = z_fma(U.u[4 * (p + 2 * q) + 3], c[3], v);
                const int r = p * chi_l + al, cc = q * chi_r + gm;
                if (tall) X[(size_t)cc * m + r] = v;
                else X[(size_t)r * nn + cc] = z_conj(v);
            }
    }
}

__global__ void mps_set_identity_kernel(double2* __restrict__ W, const int q) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < q * q; e += gridDim.x * blockDim.x)
        W[e] = make_double2((e / q) == (e % q) ? 1.0 : 0.0, 0.0);
}

// ---------------------------------------------------------------------------------------------
// K9: one-sided Jacobi SVD.  X: p x q column-major (p >= q is the caller's job), W: q x q
// column-major (starts as identity).  On exit the columns of X are orthogonal: X_in = Q S W^H.
// ---------------------------------------------------------------------------------------------
// Columns count as orthogonal when |<a|b>| <= tol |a||b| with tol = 2 sqrt(p) eps: the rounding noise
// of a length-p complex dot product (LAPACK's zgesvj uses the same sqrt(m) eps scale).  A tighter
// test keeps rotating noise and doubles the number of sweeps without changing any singular triplet.
__host__ __device__ __forceinline__ double jacobi_tol(const int p) { return 4.5e-16 * sqrt((double)p); }

// round-robin tournament on N (even) players: pair k of round r
__device__ __forceinline__ void rr_pair(const int N, const int r, const int k, int& a, int& b) {
    if (k == 0) { a = N - 1; b = r; }
    else { a = (r + k) % (N - 1); b = (r - k + (N - 1)) % (N - 1); }
    if (a > b) { const int t = a; a = b; b = t; }
}

// One Hestenes rotation of columns (ca, cb), executed by `nth` cooperating threads (thread t of
// nth); red = reduction functor summing a double over those threads.  Returns 1 if rotated.
template <class Reduce>
__device__ __forceinline__ int jacobi_rotate(double2* __restrict__ X, double2* __restrict__ W, const int p,
                                             const int q, const int ca, const int cb, const int t, const int nth,
                                             const double tiny2, Reduce red) {
    double2* xa = X + (size_t)ca * p; double2* xb = X + (size_t)cb * p;
    double al = 0, be = 0, gr = 0, gi = 0;
    for (int i = t; i < p; i += nth) {
        const double2 u = xa[i], v = xb[i];
        al = fma(u.x, u.x, fma(u.y, u.y, al));
        be = fma(v.x, v.x, fma(v.y, v.y, be));
        gr = fma(u.x, v.x, fma(u.y, v.y, gr));      // conj(u) * v
        gi = fma(u.x, v.y, fma(-u.y, v.x, gi));
    }
    al = red(al); be = red(be); gr = red(gr); gi = red(gi);
    const double g = sqrt(gr * gr + gi * gi);
    // columns below 1e-17 |X|_F are numerical zeros (Aer chops singular values <= 1e-16): rotating
    // them against each other never converges and cannot change any kept singular triplet
    if (g == 0.0 || g <= jacobi_tol(p) * sqrt(al * be) || al <= tiny2 || be <= tiny2) return 0;
    const double zeta = (be - al) / (2.0 * g);
    const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
    const double2 ph = make_double2(gr / g, gi / g);   // e^{i phi}
    const double2 sm = make_double2(-s * ph.x, s * ph.y);   // -s e^{-i phi}
    const double2 sp = make_double2(s * ph.x, s * ph.y);    //  s e^{+i phi}
    for (int i = t; i < p; i += nth) {
        const double2 u = xa[i], v = xb[i];
        double2 nu = z_mul(sm, v); nu.x = fma(c, u.x, nu.x); nu.y = fma(c, u.y, nu.y);
        double2 nv = z_mul(sp, u); nv.x = fma(c, v.x, nv.x); nv.y = fma(c, v.y, nv.y);
        xa[i] = nu; xb[i] = nv;
    }
    double2* wa = W + (size_t)ca * q; double2* wb = W + (size_t)cb * q;
    for (int i = t; i < q; i += nth) {
        const double2 u = wa[i], v = wb[i];
        double2 nu = z_mul(sm, v); nu.x = fma(c, u.x, nu.x); nu.y = fma(c, u.y, nu.y);
        double2 nv = z_mul(sp, u); nv.x = fma(c, v.x, nv.x); nv.y = fma(c, v.y, nv.y);
        wa[i] = nu; wb[i] = nv;
    }
    return 1;
}

struct WarpReduce {
    __device__ __forceinline__ double operator()(double v) const { return warp_sum_d(v); }
};

// Whole SVD in one CTA (q <= JACOBI_CTA_MAX_Q): each warp owns pairs k = warp, warp + nwarps, ...
// of every round; rounds are separated by __syncthreads; sweeps repeat until no rotation fired.
constexpr int JACOBI_CTA_MAX_Q = 96;
__global__ void __launch_bounds__(1024)
jacobi_cta_kernel(double2* __restrict__ X, double2* __restrict__ W, const int p, const int q, const int max_sweeps,
                  int* __restrict__ sweeps_done) {
    __shared__ int rotated;
    __shared__ double fro_w[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int N = (q + 1) & ~1;
    double fro = 0;
    for (size_t e = threadIdx.x; e < (size_t)p * q; e += blockDim.x) { const double2 v = X[e]; fro = fma(v.x, v.x, fma(v.y, v.y, fro)); }
    fro = warp_sum_d(fro);
    if (lane == 0) fro_w[warp] = fro;
    __syncthreads();
    fro = 0;
    for (int w = 0; w < nwarps; ++w) fro += fro_w[w];
    const double tiny2 = 1e-34 * fro;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (threadIdx.x == 0) rotated = 0;
        __syncthreads();
        for (int r = 0; r < N - 1; ++r) {
            for (int k = warp; k < N / 2; k += nwarps) {
                int a, b;
                rr_pair(N, r, k, a, b);
                if (b < q) {
                    const int did = jacobi_rotate(X, W, p, q, a, b, lane, 32, tiny2, WarpReduce());
                    if (did && lane == 0) rotated = 1;
                }
            }
            __syncthreads();
        }
        const int any = rotated;
        __syncthreads();
        if (!any) { ++sweep; break; }
    }
    if (threadIdx.x == 0) *sweeps_done = sweep;
}

// One tournament round, one CTA (128 threads) per column pair.
__global__ void __launch_bounds__(128)
jacobi_round_kernel(double2* __restrict__ X, double2* __restrict__ W, const int p, const int q, const int N,
                    const int r, const double* __restrict__ fro2, int* __restrict__ rotated) {
    __shared__ double red_buf[4][4];
    int a, b;
    rr_pair(N, r, blockIdx.x, a, b);
    if (b >= q) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int slot = 0;
    auto red = [&](double v) -> double {
        v = warp_sum_d(v);
        if (lane == 0) red_buf[slot][warp] = v;
        __syncthreads();
        const double s = (red_buf[slot][0] + red_buf[slot][1]) + (red_buf[slot][2] + red_buf[slot][3]);
        ++slot;
        return s;
    };
    const int did = jacobi_rotate(X, W, p, q, a, b, threadIdx.x, 128, 1e-34 * (*fro2), red);
    if (did && threadIdx.x == 0) atomicOr(rotated, 1);
}

// Whole SVD in ONE cooperative launch for large bonds: CTA k owns pair k of every tournament round,
// rounds are separated by grid-wide barriers (the ~2 us barrier replaces a ~9 us kernel launch per
// round: a 512-column SVD is ~5000-9000 sequential rounds).  ctrl[0..1] = per-sweep rotation flags
// (ping-pong), ctrl[2] = sweeps done; fro2 = |X|_F^2.
__global__ void __launch_bounds__(128)
jacobi_coop_kernel(double2* __restrict__ X, double2* __restrict__ W, const int p, const int q, const int N,
                   const int max_sweeps, const double* __restrict__ fro2, int* __restrict__ ctrl) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double red_buf[4][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double tiny2 = 1e-34 * (*fro2);
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        int* flag = ctrl + (sweep & 1);
        int mine = 0;
        for (int r = 0; r < N - 1; ++r) {
            int a, b;
            rr_pair(N, r, blockIdx.x, a, b);
            if (b < q) {
                int slot = 0;
                auto red = [&](double v) -> double {
                    v = warp_sum_d(v);
                    if (lane == 0) red_buf[slot][warp] = v;
                    __syncthreads();
                    const double s = (red_buf[slot][0] + red_buf[slot][1]) + (red_buf[slot][2] + red_buf[slot][3]);
                    ++slot;
                    return s;
                };
                mine |= jacobi_rotate(X, W, p, q, a, b, threadIdx.x, 128, tiny2, red);
            }
            grid.sync();
        }
        if (mine && threadIdx.x == 0) atomicOr(flag, 1);
        if (blockIdx.x == 0 && threadIdx.x == 0) ctrl[(sweep + 1) & 1] = 0;   // next sweep's flag
        grid.sync();
        const int any = *((volatile int*)flag);
        if (!any) { ++sweep; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl[2] = sweep;
}

// Blocked variant of the cooperative SVD: the columns are grouped in blocks of WB; in every outer round
// CTA k owns one PAIR of blocks (round-robin tournament over the blocks), keeps its 2*WB columns of X
// and of W in shared memory and rotates every cross pair (i in block A, j in block B) there -- WB local
// rounds of WB disjoint pairs, one warp per pair -- before the grid-wide barrier.  The first outer
// round of a sweep also rotates the pairs inside each block, so a sweep still visits every column pair
// exactly once.  Compared with one pair per CTA per barrier this divides the number of grid barriers
// (and of trips through L2) per sweep by WB; the sweep count is the same (checked against the scalar
// ordering on the C4 two-site matrices).
__device__ __forceinline__ void jacobi_grid_barrier(unsigned* bar, const unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (*((volatile unsigned*)bar) < target) { }
        __threadfence();
    }
    __syncthreads();
}

// ctrl[0..1] = per-sweep rotation flags (ping-pong), ctrl[2] = sweeps done, ctrl[3] = barrier counter (all
// zero at launch).  NB = number of blocks (even, NB * WB >= q).  Dynamic shared memory: 2*WB*(p+q) double2.
// 128 threads (4 warps) cooperate on one column pair: WB pairs per local round -> 128*WB threads per CTA
// (one warp per pair left every latency exposed: 20 us per outer round instead of ~6).
// Copies the 2*WB columns of block pair (A, B) of a column-major matrix G (`rows` rows, `q` valid columns)
// between global memory and rows [row0, row0 + rows) of the shared-memory columns (stride ld); LOAD: global
// -> shared (L2 loads), else shared -> global.  Element e = c * rows + i is walked with stride NT.
template <int WB, int NT, bool LOAD>
__device__ __forceinline__ void jb_copy(double2* __restrict__ smem, double2* __restrict__ G, const int rows, const int q,
                                        const int row0, const int ld, const int A, const int B) {
    const int total = 2 * WB * rows;
    int c = 0, i = threadIdx.x;
    while (i >= rows) { i -= rows; ++c; }
    const int dc = NT / rows, di = NT - dc * rows;       // advance of (c, i) per step (loop-invariant division)
    for (int e = threadIdx.x; e < total; e += 4 * NT) {
        int cc[4], ii[4];
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            cc[u] = c; ii[u] = i;
            c += dc; i += di;
            if (i >= rows) { i -= rows; ++c; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int col = cc[u] < WB ? A * WB + cc[u] : B * WB + (cc[u] - WB);
            const bool ok = cc[u] < 2 * WB && col < q;
            if (LOAD) { if (ok) v[u] = __ldcg(G + (size_t)col * rows + ii[u]); }
            else { if (ok) v[u] = smem[(size_t)cc[u] * ld + row0 + ii[u]]; }
            cc[u] = ok ? cc[u] : -1;
            if (!LOAD && ok) __stcg(G + (size_t)col * rows + ii[u], v[u]);
        }
        if (LOAD) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (cc[u] >= 0) smem[(size_t)cc[u] * ld + row0 + ii[u]] = v[u];
        }
    }
}

constexpr int JB_GROUP = 128;
template <int WB>
__global__ void __launch_bounds__(JB_GROUP * WB)
jacobi_block_kernel(double2* __restrict__ X, double2* __restrict__ W, const int p, const int q, const int NB,
                    const int max_sweeps, const double* __restrict__ fro2, int* __restrict__ ctrl) {
    extern __shared__ __align__(16) double2 jb_smem[];
    __shared__ double red[WB][JB_GROUP / 32][4];
    constexpr int NT = JB_GROUP * WB;
    const int lane = threadIdx.x & 31;
    const int grp = threadIdx.x / JB_GROUP, gt = threadIdx.x % JB_GROUP, gw = gt >> 5;
    const int ld = p + q;
    const double tiny2 = 1e-34 * (*fro2);
    const double tol = jacobi_tol(p);
    unsigned* bar = (unsigned*)(ctrl + 3);
    unsigned epoch = 0;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        int* flag = ctrl + (sweep & 1);
        int mine = 0;
        for (int r = 0; r < NB - 1; ++r) {
            int A, B;
            rr_pair(NB, r, blockIdx.x, A, B);
            // ---- block pair -> shared memory (column c of the pair at jb_smem + c*ld: X rows, then W rows) ----
            // (c, i) advance incrementally: no integer division on the copy path; 4 loads in flight per thread
            jb_copy<WB, NT, true>(jb_smem, X, p, q, 0, ld, A, B);
            jb_copy<WB, NT, true>(jb_smem, W, q, q, p, ld, A, B);
            __syncthreads();
            const int nloc = r == 0 ? 2 * WB - 1 : WB;
            for (int t = 0; t < nloc; ++t) {
                int i, j;
                if (r == 0) rr_pair(2 * WB, t, grp, i, j);
                else { i = grp; j = WB + (grp + t) % WB; }
                const int ci = i < WB ? A * WB + i : B * WB + (i - WB);
                const int cj = j < WB ? A * WB + j : B * WB + (j - WB);
                const bool valid = ci < q && cj < q;
                double2* ca = jb_smem + (size_t)i * ld;
                double2* cb = jb_smem + (size_t)j * ld;
                double al = 0, be = 0, gr = 0, gi = 0;
                if (valid) {
#pragma unroll 4
                    for (int k = gt; k < p; k += JB_GROUP) {
                        const double2 u = ca[k], v = cb[k];
                        al = fma(u.x, u.x, fma(u.y, u.y, al));
                        be = fma(v.x, v.x, fma(v.y, v.y, be));
                        gr = fma(u.x, v.x, fma(u.y, v.y, gr));
                        gi = fma(u.x, v.y, fma(-u.y, v.x, gi));
                    }
                }
                al = warp_sum_d(al); be = warp_sum_d(be); gr = warp_sum_d(gr); gi = warp_sum_d(gi);
                if (lane == 0) { red[grp][gw][0] = al; red[grp][gw][1] = be; red[grp][gw][2] = gr; red[grp][gw][3] = gi; }
                __syncthreads();
                al = (red[grp][0][0] + red[grp][1][0]) + (red[grp][2][0] + red[grp][3][0]);
                be = (red[grp][0][1] + red[grp][1][1]) + (red[grp][2][1] + red[grp][3][1]);
                gr = (red[grp][0][2] + red[grp][1][2]) + (red[grp][2][2] + red[grp][3][2]);
                gi = (red[grp][0][3] + red[grp][1][3]) + (red[grp][2][3] + red[grp][3][3]);
                const double g = sqrt(gr * gr + gi * gi);
                if (valid && !(g == 0.0 || g <= tol * sqrt(al * be) || al <= tiny2 || be <= tiny2)) {
                    const double zeta = (be - al) / (2.0 * g);
                    const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + tt * tt), sn = c * tt;
                    const double2 ph = make_double2(gr / g, gi / g);
                    const double2 sm = make_double2(-sn * ph.x, sn * ph.y);
                    const double2 sp = make_double2(sn * ph.x, sn * ph.y);
#pragma unroll 4
                    for (int k = gt; k < ld; k += JB_GROUP) {
                        const double2 u = ca[k], v = cb[k];
                        double2 nu = z_mul(sm, v); nu.x = fma(c, u.x, nu.x); nu.y = fma(c, u.y, nu.y);
                        double2 nv = z_mul(sp, u); nv.x = fma(c, v.x, nv.x); nv.y = fma(c, v.y, nv.y);
                        ca[k] = nu; cb[k] = nv;
                    }
                    mine = 1;
                }
                __syncthreads();
            }
            // ---- back to global ----
            jb_copy<WB, NT, false>(jb_smem, X, p, q, 0, ld, A, B);
            jb_copy<WB, NT, false>(jb_smem, W, q, q, p, ld, A, B);
            jacobi_grid_barrier(bar, ++epoch * gridDim.x);
        }
        if (__syncthreads_or(mine) && threadIdx.x == 0) atomicOr(flag, 1);
        if (blockIdx.x == 0 && threadIdx.x == 0) ctrl[(sweep + 1) & 1] = 0;   // next sweep's flag
        jacobi_grid_barrier(bar, ++epoch * gridDim.x);
        const int any = *((volatile int*)flag);
        if (!any) { ++sweep; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl[2] = sweep;
}

// fro2[0] = |X|_F^2 (one CTA)
__global__ void __launch_bounds__(256)
jacobi_fro_kernel(const double2* __restrict__ X, const size_t nelem, double* __restrict__ fro2) {
    __shared__ double sh[8];
    double s = 0;
    for (size_t e = threadIdx.x; e < nelem; e += blockDim.x) { const double2 v = X[e]; s = fma(v.x, v.x, fma(v.y, v.y, s)); }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < 8; ++w) t += sh[w]; *fro2 = t; }
}

// sigma[j] = || X[:, j] ||   (one warp per column)
__global__ void jacobi_sigma_kernel(const double2* __restrict__ X, const int p, const int q, double* __restrict__ sigma) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= q) return;
    double s = 0;
    for (int i = lane; i < p; i += 32) { const double2 v = X[(size_t)j * p + i]; s = fma(v.x, v.x, fma(v.y, v.y, s)); }
    s = warp_sum_d(s);
    if (lane == 0) sigma[j] = sqrt(s);
}

// Truncated factors -> site tensors.  perm[j] = column of X/W holding the j-th largest singular
// value, kept[j] its (renormalised) value, sig[perm[j]] the raw norm of that column.
//   tall:  U[(b,al), j] = X[perm_j][(b,al)] / sig      Vh[j, (b',gm)] = conj(W[perm_j][(b',gm)])
//   wide:  U[(b,al), j] = W[perm_j][(b,al)]            Vh[j, (b',gm)] = conj(X[perm_j][(b',gm)]) / sig
// Gamma_i[b][al][j] = U / ll[al];  Gamma_{i+1}[b'][j][gm] = Vh / lr[gm];  lambda_i[j] = kept[j].
__global__ void mps_write_sites_kernel(const double2* __restrict__ X, const double2* __restrict__ W,
                                       const double* __restrict__ sig, const int* __restrict__ perm,
                                       const double* __restrict__ kept, const int chi_l, const int chi_r,
                                       const int k, const int tall, const double* __restrict__ ll,
                                       const double* __restrict__ lr, double2* __restrict__ g0,
                                       double2* __restrict__ g1, double* __restrict__ lam) {
    const int m = 2 * chi_l, nn = 2 * chi_r;
    const int n0 = m * k, n1 = k * nn;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n0 + n1 + k; e += gridDim.x * blockDim.x) {
        if (e < n0) {
            const int j = e % k, r = e / k;            // r = (b, al)
            const int col = perm[j];
            double2 v;
            if (tall) { v = X[(size_t)col * m + r]; const double s = 1.0 / sig[col]; v.x *= s; v.y *= s; }
            else v = W[(size_t)col * m + r];
            const double d = ll ? 1.0 / ll[r % chi_l] : 1.0;
            g0[e] = make_double2(v.x * d, v.y * d);     // [b][al][j] == r * k + j
        } else if (e < n0 + n1) {
            const int f = e - n0;
            const int gm = f % chi_r, j = (f / chi_r) % k, bp = f / (chi_r * k);
            const int col = perm[j], c = bp * chi_r + gm;
            double2 v;
            if (tall) v = z_conj(W[(size_t)col * nn + c]);
            else { v = z_conj(X[(size_t)col * nn + c]); const double s = 1.0 / sig[col]; v.x *= s; v.y *= s; }
            const double d = lr ? 1.0 / lr[gm] : 1.0;
            g1[f] = make_double2(v.x * d, v.y * d);     // [b'][j][gm]
        } else {
            lam[e - n0 - n1] = kept[e - n0 - n1];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// read-outs
// ---------------------------------------------------------------------------------------------
struct SiteTable {            // device arrays describing an MPS (built per call)
    const double2* const* gam; // n pointers
    const double* const* lam;  // n pointers (lam[n-1] = null)
    const int* chi;            // chi[i] = right bond of site i; left bond = chi[i-1] (1 for i = 0)
    int n;
};

// out[t] = <bits_t | psi>: one CTA per bitstring, the whole chain of vector-matrix products in one
// launch; the running vector lives in shared memory.  1024 threads: 256 output columns x 4 slices of the
// contracted bond, 8 independent loads in flight per thread (one thread per column with a serial loop over
// the bond was latency-bound: 2.1 ms per amplitude at chi = 256).  Shared memory: 6 * maxchi double2.
constexpr int AMPS_THREADS = 1024;
constexpr int AMPS_SLICES = 4;
__global__ void __launch_bounds__(AMPS_THREADS)
mps_amps_kernel(const SiteTable st, const uint64_t* __restrict__ bits, const int maxchi, double2* __restrict__ out) {
    extern __shared__ double2 vsm[];
    double2* cur = vsm;
    double2* nxt = vsm + maxchi;
    double2* part = vsm + 2 * maxchi;                 // [AMPS_SLICES][maxchi]
    const uint64_t b = bits[blockIdx.x];
    const int col0 = threadIdx.x % (AMPS_THREADS / AMPS_SLICES), slice = threadIdx.x / (AMPS_THREADS / AMPS_SLICES);
    if (threadIdx.x == 0) cur[0] = make_double2(1.0, 0.0);
    __syncthreads();
    for (int i = 0; i < st.n; ++i) {
        const int chi_l = i ? st.chi[i - 1] : 1, chi_r = st.chi[i];
        const double2* G = st.gam[i] + (size_t)((b >> i) & 1ull) * chi_l * chi_r;
        const double* lam = st.lam[i];
        for (int be = col0; be < chi_r; be += AMPS_THREADS / AMPS_SLICES) {
            double2 acc = make_double2(0.0, 0.0);
#pragma unroll 8
            for (int al = slice; al < chi_l; al += AMPS_SLICES) acc = z_fma(cur[al], G[(size_t)al * chi_r + be], acc);
            part[slice * maxchi + be] = acc;
        }
        __syncthreads();
        for (int be = threadIdx.x; be < chi_r; be += AMPS_THREADS) {
            double2 acc = part[be];
#pragma unroll
            for (int s = 1; s < AMPS_SLICES; ++s) { acc.x += part[s * maxchi + be].x; acc.y += part[s * maxchi + be].y; }
            if (lam) { acc.x *= lam[be]; acc.y *= lam[be]; }
            nxt[be] = acc;
        }
        __syncthreads();
        double2* t = cur; cur = nxt; nxt = t;
    }
    if (threadIdx.x == 0) out[blockIdx.x] = cur[0];
}

// sum_e X[e] * Y[e] (no conjugation) -> partial per block; used for <D_i, F_{i+1}> contractions
__global__ void __launch_bounds__(256)
mps_dot_elem_kernel(const double2* __restrict__ X, const double2* __restrict__ Y, const int nelem,
                    double2* __restrict__ out) {
    __shared__ double sre[8], sim[8];
    double re = 0, im = 0;
    for (int e = threadIdx.x; e < nelem; e += blockDim.x) {
        const double2 p = z_mul(X[e], Y[e]);
        re += p.x; im += p.y;
    }
    re = warp_sum_d(re); im = warp_sum_d(im);
    if ((threadIdx.x & 31) == 0) { sre[threadIdx.x >> 5] = re; sim[threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sre[w]; b += sim[w]; }
        *out = make_double2(a, b);
    }
}

// Z = X - Y and S = X + Y elementwise
__global__ void mps_sum_diff_kernel(const double2* X, const double2* Y, const int nelem, double2* S, double2* D) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += gridDim.x * blockDim.x) {
        const double2 a = X[e], b = Y[e];
        S[e] = make_double2(a.x + b.x, a.y + b.y);
        D[e] = make_double2(a.x - b.x, a.y - b.y);
    }
}

}  // namespace b200
