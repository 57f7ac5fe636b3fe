// sv_kernels.cuh -- sm_100a statevector kernels (complex128, little-endian qubit order).
//
//   sv_sweep_kernel<R>  K1/K2: fused gate sweep.  One CTA = one tile of 2^12 amplitudes
//                       (5 lane qubits + 7 arbitrary qubits); 2^R amplitudes per thread live in
//                       registers for a whole round of gates; rounds exchange through swizzled
//                       shared memory; first round loads from HBM with 512 B-per-warp coalesced
//                       128-bit accesses and the last round stores the same way.
//   sv_small_kernel     n <= 11: whole state in one CTA's shared memory (launch-latency path).
//   sv_expz_kernel      K4: all <Z_q> in one read pass (deterministic two-stage reduction).
//   sv_rdm3_kernel      K5: three pair-RDMs per read pass, 8 amplitudes per thread in registers.
//   sv_inner_kernel     K6: 2x2 transfer matrix <L|(|i><j|_q)|R> in one read pass over L and R.
//
// HBM-bound: algorithmic bytes per sweep = 32 * 2^n (read + write every amplitude once).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "sv_plan.h"

// The per-thread bodies are __host__ __device__ so that tests/emu can run the very same code
// on the CPU (one loop iteration per CUDA thread) to check the planner and the index math
// without a GPU.  The product only ever launches the __global__ kernels.
#define B200_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define B200_LDG(p) __ldg(p)
#else
#define B200_LDG(p) (*(p))
#endif

namespace b200 {

constexpr int SWEEP_THREADS = 1 << (TILE_BITS - REG_BITS);  // 256
constexpr int RED_THREADS = 256;
constexpr int EXPZ_WIDTH = 32;   // doubles per block partial
constexpr int RDM3_WIDTH = 48;
constexpr int INNER_WIDTH = 8;
constexpr int INNER2_WIDTH = 32;

B200_HD uint64_t ins0_64(uint64_t x, int pos) {
    const uint64_t lo = x & ((1ull << pos) - 1ull);
    return ((x >> pos) << (pos + 1)) | lo;
}
B200_HD uint32_t ins0_32(uint32_t x, int pos) {
    const uint32_t lo = x & ((1u << pos) - 1u);
    return ((x >> pos) << (pos + 1)) | lo;
}
// 16-byte-slot swizzle: the slot-in-row bits (low 3) are XOR-folded with every higher 3-bit group of
// the 12-bit tile index, so a quarter-warp whose lanes vary ANY three consecutive tile bits (the
// three lowest non-register positions of a round) lands on the 8 distinct 16 B slots of a 128 B row.
B200_HD uint32_t swz(uint32_t i) { return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7u); }

B200_HD double2 cmul(const double2 a, const double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// a*b + c
B200_HD double2 cfma(const double2 a, const double2 b, double2 c) {
    c.x = fma(a.x, b.x, c.x);
    c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y);
    c.y = fma(a.y, b.x, c.y);
    return c;
}
// conj(a)*b + c
B200_HD double2 cjfma(const double2 a, const double2 b, double2 c) {
    c.x = fma(a.x, b.x, c.x);
    c.x = fma(a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y);
    c.y = fma(-a.y, b.x, c.y);
    return c;
}

// ---------------------------------------------------------------------------------------------
// register-resident op bodies
// ---------------------------------------------------------------------------------------------
template <int R, int TB>
B200_HD void op_mat1(double2 (&a)[1 << R], const double* __restrict__ m, const int cmask) {
    const double2 m00 = make_double2(m[0], m[1]), m01 = make_double2(m[2], m[3]);
    const double2 m10 = make_double2(m[4], m[5]), m11 = make_double2(m[6], m[7]);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        if (j & (1 << TB)) continue;
        if (cmask == 0 || (j & cmask)) {
            const double2 x = a[j], y = a[j | (1 << TB)];
            a[j] = cfma(m01, y, cmul(m00, x));
            a[j | (1 << TB)] = cfma(m11, y, cmul(m10, x));
        }
    }
}

template <int R, int TB>
B200_HD void op_x(double2 (&a)[1 << R], const int cmask) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        if (j & (1 << TB)) continue;
        if (cmask == 0 || (j & cmask)) {
            const double2 x = a[j];
            a[j] = a[j | (1 << TB)];
            a[j | (1 << TB)] = x;
        }
    }
}

template <int R, int TB0, int TB1>
B200_HD void op_mat2(double2 (&a)[1 << R], const double* __restrict__ mm, const int cmask) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        if (j & ((1 << TB0) | (1 << TB1))) continue;
        if (cmask == 0 || (j & cmask)) {
            const double2 v0 = a[j], v1 = a[j | (1 << TB0)], v2 = a[j | (1 << TB1)],
                          v3 = a[j | (1 << TB0) | (1 << TB1)];
            double2 o[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double2* row = reinterpret_cast<const double2*>(mm) + 4 * r;
                double2 acc = cmul(B200_LDG(row + 0), v0);
                acc = cfma(B200_LDG(row + 1), v1, acc);
                acc = cfma(B200_LDG(row + 2), v2, acc);
                acc = cfma(B200_LDG(row + 3), v3, acc);
                o[r] = acc;
            }
            a[j] = o[0];
            a[j | (1 << TB0)] = o[1];
            a[j | (1 << TB1)] = o[2];
            a[j | (1 << TB0) | (1 << TB1)] = o[3];
        }
    }
}

template <int R>
B200_HD void op_diag(double2 (&a)[1 << R], const DevOp* __restrict__ op, const uint64_t g) {
    const int dq0 = op->dq0, dq1 = op->dq1, dm0 = op->dmask0, dm1 = op->dmask1;
    const int u0 = dq0 >= 0 ? (int)((g >> dq0) & 1ull) : 0;
    const int u1 = dq1 >= 0 ? (int)((g >> dq1) & 1ull) : 0;
    const double2* ph = reinterpret_cast<const double2*>(op->m);
    if (dm0 == 0 && dm1 == 0) {
        const double2 p = ph[u0 + 2 * u1];
        if (p.x == 1.0 && p.y == 0.0) return;
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) a[j] = cmul(a[j], p);
    } else if (dm0 != 0 && dm1 != 0) {
        const double2 p0 = ph[0], p1 = ph[1], p2 = ph[2], p3 = ph[3];
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) {
            const double2 lo = (j & dm0) ? p1 : p0, hi = (j & dm0) ? p3 : p2;
            a[j] = cmul(a[j], (j & dm1) ? hi : lo);
        }
    } else {
        const int dm = dm0 | dm1;
        const double2 p0 = dm0 ? ph[2 * u1] : ph[u0];
        const double2 p1 = dm0 ? ph[1 + 2 * u1] : ph[u0 + 2];
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) a[j] = cmul(a[j], (j & dm) ? p1 : p0);
    }
}

// `pend`: product of the phases of diagonal ops whose qubits are all thread-uniform.  Such a phase
// multiplies all 2^R amplitudes of the thread alike, so it commutes with every other op of the round
// and is applied ONCE at the end of the round (1 complex multiply per op instead of 2^R).
template <int R>
B200_HD void apply_op(double2 (&a)[1 << R], const DevOp* __restrict__ op, const uint64_t g,
                                         const double* __restrict__ mat2tab, double2& pend) {
    const int kind = op->kind;
    if (kind == K_DIAG) {
        if (op->dmask0 == 0 && op->dmask1 == 0) {
            const int u0 = op->dq0 >= 0 ? (int)((g >> op->dq0) & 1ull) : 0;
            const int u1 = op->dq1 >= 0 ? (int)((g >> op->dq1) & 1ull) : 0;
            pend = cmul(pend, reinterpret_cast<const double2*>(op->m)[u0 + 2 * u1]);
            return;
        }
        op_diag<R>(a, op, g);
        return;
    }
    const int cq = op->cq;
    if (cq >= 0 && !((g >> cq) & 1ull)) return;
    const int cmask = op->cmask;
    const int t0 = op->treg0;
    if (kind == K_X) {
        switch (t0) {
        case 0: op_x<R, 0>(a, cmask); break;
        case 1: op_x<R, 1>(a, cmask); break;
        case 2: op_x<R, 2>(a, cmask); break;
        default: op_x<R, 3>(a, cmask); break;
        }
    } else if (kind == K_MAT1) {
        switch (t0) {
        case 0: op_mat1<R, 0>(a, op->m, cmask); break;
        case 1: op_mat1<R, 1>(a, op->m, cmask); break;
        case 2: op_mat1<R, 2>(a, op->m, cmask); break;
        default: op_mat1<R, 3>(a, op->m, cmask); break;
        }
    } else {  // K_MAT2, treg0 < treg1
        const double* mm = mat2tab + op->mat2;
        switch (t0 * 4 + op->treg1) {
        case 1: op_mat2<R, 0, 1>(a, mm, cmask); break;
        case 2: op_mat2<R, 0, 2>(a, mm, cmask); break;
        case 3: op_mat2<R, 0, 3>(a, mm, cmask); break;
        case 6: op_mat2<R, 1, 2>(a, mm, cmask); break;
        case 7: op_mat2<R, 1, 3>(a, mm, cmask); break;
        default: op_mat2<R, 2, 3>(a, mm, cmask); break;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1/K2: fused sweep
// ---------------------------------------------------------------------------------------------
// Base index of tile `tile`: zero bits inserted at the tile's high qubits.
B200_HD uint64_t sweep_tile_base(const DevSweep& sw, const uint32_t tile) {
    uint64_t y = tile;
    for (int i = sw.c; i < TILE_BITS; ++i) y = ins0_64(y, sw.tileq[i] - sw.c);
    return y << sw.c;
}

// One round of one thread: load 2^R amplitudes (HBM in the first round, shared memory after),
// apply the round's ops in registers, store (HBM in the last round, shared memory before).
// The caller separates rounds with a block barrier.
template <int R>
B200_HD void sweep_round(const double2* __restrict__ src, double2* __restrict__ dst, double2* tile_smem,
                         const DevSweep& sw, const DevRound* __restrict__ rd, const DevOp* __restrict__ ops,
                         const double* __restrict__ mat2tab, const uint64_t tile_base, const uint32_t tid,
                         const bool first, const bool last) {
    const int c = sw.c;
    int rp[R];
#pragma unroll
    for (int k = 0; k < R; ++k) rp[k] = rd->regpos[k];
    uint32_t tl = tid;
#pragma unroll
    for (int k = 0; k < R; ++k) tl = ins0_32(tl, rp[k]);
    uint64_t g = tile_base | (uint64_t)(tl & ((1u << c) - 1u));
    for (int i = c; i < TILE_BITS; ++i) g |= (uint64_t)((tl >> i) & 1u) << sw.tileq[i];

    double2 a[1 << R];
    if (first) {
        uint64_t gs[R];
#pragma unroll
        for (int k = 0; k < R; ++k) gs[k] = 1ull << sw.tileq[rp[k]];
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) {
            uint64_t off = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) if (j >> k & 1) off += gs[k];
            a[j] = src[g + off];
        }
    } else {
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) {
            uint32_t off = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) if (j >> k & 1) off += 1u << rp[k];
            a[j] = tile_smem[swz(tl + off)];
        }
    }

    const int ob = rd->op_begin, oe = rd->op_end;
    double2 pend = make_double2(1.0, 0.0);
    for (int o = ob; o < oe; ++o) apply_op<R>(a, ops + o, g, mat2tab, pend);
    if (pend.x != 1.0 || pend.y != 0.0) {
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) a[j] = cmul(a[j], pend);
    }

    if (last) {
        uint64_t gs[R];
#pragma unroll
        for (int k = 0; k < R; ++k) gs[k] = 1ull << sw.tileq[rp[k]];
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) {
            uint64_t off = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) if (j >> k & 1) off += gs[k];
            dst[g + off] = a[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) {
            uint32_t off = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) if (j >> k & 1) off += 1u << rp[k];
            tile_smem[swz(tl + off)] = a[j];
        }
    }
}

template <int R>
__global__ void __launch_bounds__(SWEEP_THREADS, 2)
sv_sweep_kernel(const double2* __restrict__ src, double2* __restrict__ dst, const DevSweep sw,
                const DevRound* __restrict__ rounds, const DevOp* __restrict__ ops,
                const double* __restrict__ mat2tab, const uint32_t ntiles) {
    extern __shared__ __align__(16) double2 tile_smem[];
    static_assert(R == REG_BITS, "register bits");
    const int nr = sw.round_end - sw.round_begin;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t tile_base = sweep_tile_base(sw, tile);
        for (int r = 0; r < nr; ++r) {
            sweep_round<R>(src, dst, tile_smem, sw, rounds + sw.round_begin + r, ops, mat2tab, tile_base,
                           threadIdx.x, r == 0, r == nr - 1);
            if (r != nr - 1) __syncthreads();
        }
        if (nr > 1) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// small path: n <= SMALL_MAX_QUBITS, one CTA, state in shared memory
// ---------------------------------------------------------------------------------------------
// One op of the small path for thread `tid` of `nthreads` (state `st` in shared memory).
B200_HD void small_apply_op(double2* st, const uint32_t dim, const DevOp* __restrict__ op,
                            const double* __restrict__ mat2tab, const uint32_t tid, const uint32_t nthreads) {
    const int kind = op->kind;
    if (kind == K_DIAG) {
        const double2* ph = reinterpret_cast<const double2*>(op->m);
        const int d0 = op->dq0, d1 = op->dq1;
        for (uint32_t i = tid; i < dim; i += nthreads) {
            const int b0 = (i >> d0) & 1, b1 = d1 >= 0 ? (i >> d1) & 1 : 0;
            st[i] = cmul(st[i], ph[b0 + 2 * b1]);
        }
    } else if (kind == K_MAT1 || kind == K_X) {
        const int t = op->treg0, cq = op->cq;
        const double* m = op->m;
        const double2 m00 = make_double2(m[0], m[1]), m01 = make_double2(m[2], m[3]);
        const double2 m10 = make_double2(m[4], m[5]), m11 = make_double2(m[6], m[7]);
        for (uint32_t k = tid; k < dim / 2; k += nthreads) {
            const uint32_t i0 = ins0_32(k, t), i1 = i0 | (1u << t);
            if (cq >= 0 && !((i0 >> cq) & 1)) continue;
            const double2 x = st[i0], yv = st[i1];
            if (kind == K_X) { st[i0] = yv; st[i1] = x; }
            else {
                st[i0] = cfma(m01, yv, cmul(m00, x));
                st[i1] = cfma(m11, yv, cmul(m10, x));
            }
        }
    } else {  // K_MAT2: index = bit(t0) + 2*bit(t1)
        const int t0 = op->treg0, t1 = op->treg1;
        const int lo = t0 < t1 ? t0 : t1, hi = t0 < t1 ? t1 : t0;
        const double2* mm = reinterpret_cast<const double2*>(mat2tab + op->mat2);
        for (uint32_t k = tid; k < dim / 4; k += nthreads) {
            const uint32_t b = ins0_32(ins0_32(k, lo), hi);
            const uint32_t idx[4] = {b, b | (1u << t0), b | (1u << t1), b | (1u << t0) | (1u << t1)};
            double2 v[4], w[4];
            for (int j = 0; j < 4; ++j) v[j] = st[idx[j]];
            for (int r = 0; r < 4; ++r) {
                double2 acc = cmul(mm[4 * r], v[0]);
                for (int cc = 1; cc < 4; ++cc) acc = cfma(mm[4 * r + cc], v[cc], acc);
                w[r] = acc;
            }
            for (int j = 0; j < 4; ++j) st[idx[j]] = w[j];
        }
    }
}

__global__ void __launch_bounds__(256)
sv_small_kernel(const double2* __restrict__ src, double2* __restrict__ dst, const int n,
                const DevOp* __restrict__ ops, const int nops, const double* __restrict__ mat2tab,
                const int src_is_zero) {
    extern __shared__ __align__(16) double2 st[];
    const uint32_t dim = 1u << n;
    for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x)
        st[i] = src_is_zero ? make_double2(i == 0 ? 1.0 : 0.0, 0.0) : src[i];
    __syncthreads();
    for (int o = 0; o < nops; ++o) {
        small_apply_op(st, dim, ops + o, mat2tab, threadIdx.x, blockDim.x);
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) dst[i] = st[i];
}

__global__ void sv_init_zero_kernel(double2* __restrict__ psi, const uint64_t dim) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride)
        psi[i] = make_double2(i == 0 ? 1.0 : 0.0, 0.0);
}

// ---------------------------------------------------------------------------------------------
// deterministic reductions: per-block partials, then one fixed-order final sum
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// Sums `width` per-thread values over the block (fixed order) and writes them to out[0..width).
template <int WIDTH>
__device__ __forceinline__ void block_sum_store(const double (&v)[WIDTH], double* __restrict__ out) {
    __shared__ double sh[RED_THREADS / 32][WIDTH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < WIDTH; ++k) {
        const double s = warp_sum(v[k]);
        if (lane == 0) sh[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < WIDTH) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < RED_THREADS / 32; ++w) s += sh[w][threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// out[w] = sum_b partial[b*width + w].  One warp per output column: lane l adds blocks l, l+32, ...
// in ascending order, then a fixed shuffle tree -- deterministic for a given grid size.
__global__ void __launch_bounds__(32)
reduce_partials_kernel(const double* __restrict__ partial, const int nblocks, const int width,
                       double* __restrict__ out) {
    const int w = blockIdx.x;
    if (w >= width) return;
    double s = 0;
    for (int b = threadIdx.x; b < nblocks; b += 32) s += partial[(size_t)b * width + w];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[w] = s;
}

// K4.  Total threads = 2^tb (tb >= 8); thread gt owns amplitudes (k << tb) | gt, k < 2^kb, kb >= 3.
// Block partial layout: [0] total, [1..8] S_q for the 8 in-block thread bits, [9..9+kb) S_q for
// the k bits.  (S_q = probability of bit q being 1; block-index bits are resolved in the final
// stage from [0].)
__global__ void __launch_bounds__(RED_THREADS)
sv_expz_kernel(const double2* __restrict__ psi, const int tb, const int kb, double* __restrict__ partial) {
    const uint32_t gt = blockIdx.x * RED_THREADS + threadIdx.x;
    double v[EXPZ_WIDTH];
#pragma unroll
    for (int k = 0; k < EXPZ_WIDTH; ++k) v[k] = 0.0;
    const uint64_t n8 = 1ull << (kb - 3);
    for (uint64_t k8 = 0; k8 < n8; ++k8) {
        double p[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double2 x = psi[(((k8 << 3) | (uint64_t)j) << tb) | gt];
            p[j] = fma(x.x, x.x, x.y * x.y);
        }
        const double s01 = p[0] + p[1], s23 = p[2] + p[3], s45 = p[4] + p[5], s67 = p[6] + p[7];
        const double s = (s01 + s23) + (s45 + s67);
        v[0] += s;
        v[9] += (p[1] + p[3]) + (p[5] + p[7]);
        v[10] += s23 + s67;
        v[11] += s45 + s67;
#pragma unroll
        for (int q = 3; q < EXPZ_WIDTH - 9; ++q)
            if (q < kb && ((k8 >> (q - 3)) & 1ull)) v[9 + q] += s;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) v[1 + q] = ((threadIdx.x >> q) & 1u) ? v[0] : 0.0;
    block_sum_store<EXPZ_WIDTH>(v, partial + (size_t)blockIdx.x * EXPZ_WIDTH);
}

// out[q] = <Z_q> (q < n), out[n] = norm^2
__global__ void sv_expz_final_kernel(const double* __restrict__ partial, const int nblocks, const int n,
                                     const int tb, double* __restrict__ out) {
    const int q = threadIdx.x;
    if (q > n) return;
    double total = 0, sq = 0;
    for (int b = 0; b < nblocks; ++b) {
        const double* pb = partial + (size_t)b * EXPZ_WIDTH;
        total += pb[0];
        if (q < 8) sq += pb[1 + q];
        else if (q < tb) sq += ((b >> (q - 8)) & 1) ? pb[0] : 0.0;
        else if (q < n) sq += pb[9 + (q - tb)];
    }
    out[q] = q < n ? total - 2.0 * sq : total;
}

// generic small/odd-size fallback: one block, thread q walks the whole state (n <= 14)
__global__ void sv_expz_small_kernel(const double2* __restrict__ psi, const int n, double* __restrict__ out) {
    const int q = threadIdx.x;
    if (q > n) return;
    double total = 0, sq = 0;
    for (uint32_t i = 0; i < (1u << n); ++i) {
        const double2 x = psi[i];
        const double p = fma(x.x, x.x, x.y * x.y);
        total += p;
        if (q < n && ((i >> q) & 1u)) sq += p;
    }
    out[q] = q < n ? total - 2.0 * sq : total;
}

// K5.  Accumulates the 4x4 RDM of a quad v[0..3] (index = bit_lo + 2*bit_hi) into 16 doubles:
// [0..3] diagonal, then (r,c) = (1,0),(2,0),(3,0),(2,1),(3,1),(3,2) as re,im of v_r conj(v_c).
__device__ __forceinline__ void rdm_acc(double* __restrict__ acc, const double2 v0, const double2 v1,
                                        const double2 v2, const double2 v3) {
    acc[0] = fma(v0.x, v0.x, fma(v0.y, v0.y, acc[0]));
    acc[1] = fma(v1.x, v1.x, fma(v1.y, v1.y, acc[1]));
    acc[2] = fma(v2.x, v2.x, fma(v2.y, v2.y, acc[2]));
    acc[3] = fma(v3.x, v3.x, fma(v3.y, v3.y, acc[3]));
#define RDM_OFF(k, vr, vc)                                        \
    acc[4 + 2 * (k)] = fma(vr.x, vc.x, fma(vr.y, vc.y, acc[4 + 2 * (k)])); \
    acc[5 + 2 * (k)] = fma(vr.y, vc.x, fma(-vr.x, vc.y, acc[5 + 2 * (k)]));
    RDM_OFF(0, v1, v0) RDM_OFF(1, v2, v0) RDM_OFF(2, v3, v0)
    RDM_OFF(3, v2, v1) RDM_OFF(4, v3, v1) RDM_OFF(5, v3, v2)
#undef RDM_OFF
}

// qubits x < y < z; partial layout per block: [pair xy | pair xz | pair yz] x 16 doubles
__global__ void __launch_bounds__(RED_THREADS)
sv_rdm3_kernel(const double2* __restrict__ psi, const int n, const int x, const int y, const int z,
               double* __restrict__ partial) {
    double acc[RDM3_WIDTH];
#pragma unroll
    for (int k = 0; k < RDM3_WIDTH; ++k) acc[k] = 0.0;
    const uint64_t groups = 1ull << (n - 3);
    const uint64_t bx = 1ull << x, by = 1ull << y, bz = 1ull << z;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < groups; k += stride) {
        const uint64_t b = ins0_64(ins0_64(ins0_64(k, x), y), z);
        double2 a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            a[j] = psi[b + ((j & 1) ? bx : 0) + ((j & 2) ? by : 0) + ((j & 4) ? bz : 0)];
        rdm_acc(acc, a[0], a[1], a[2], a[3]);        // (x,y), z = 0
        rdm_acc(acc, a[4], a[5], a[6], a[7]);        // (x,y), z = 1
        rdm_acc(acc + 16, a[0], a[1], a[4], a[5]);   // (x,z), y = 0
        rdm_acc(acc + 16, a[2], a[3], a[6], a[7]);   // (x,z), y = 1
        rdm_acc(acc + 32, a[0], a[2], a[4], a[6]);   // (y,z), x = 0
        rdm_acc(acc + 32, a[1], a[3], a[5], a[7]);   // (y,z), x = 1
    }
    block_sum_store<RDM3_WIDTH>(acc, partial + (size_t)blockIdx.x * RDM3_WIDTH);
}

// single pair x < y (used when n == 2 or a pair cannot be completed to a triple)
__global__ void __launch_bounds__(RED_THREADS)
sv_rdm2_kernel(const double2* __restrict__ psi, const int n, const int x, const int y,
               double* __restrict__ partial) {
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    const uint64_t groups = 1ull << (n - 2);
    const uint64_t bx = 1ull << x, by = 1ull << y;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < groups; k += stride) {
        const uint64_t b = ins0_64(ins0_64(k, x), y);
        rdm_acc(acc, psi[b], psi[b + bx], psi[b + by], psi[b + bx + by]);
    }
    block_sum_store<16>(acc, partial + (size_t)blockIdx.x * 16);
}

// K6.  M[i][j] = sum_rest conj(L[i,rest]) R[j,rest]; q < 0: M[0][0] = <L|R>.
__global__ void __launch_bounds__(RED_THREADS)
sv_inner_kernel(const double2* __restrict__ L, const double2* __restrict__ Rv, const int n, const int q,
                double* __restrict__ partial) {
    double2 m00 = make_double2(0, 0), m01 = m00, m10 = m00, m11 = m00;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    if (q < 0) {
        const uint64_t dim = 1ull << n;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride)
            m00 = cjfma(L[i], Rv[i], m00);
    } else {
        const uint64_t pairs = 1ull << (n - 1), bq = 1ull << q;
        for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < pairs; k += stride) {
            const uint64_t i0 = ins0_64(k, q), i1 = i0 | bq;
            const double2 l0 = L[i0], l1 = L[i1], r0 = Rv[i0], r1 = Rv[i1];
            m00 = cjfma(l0, r0, m00);
            m01 = cjfma(l0, r1, m01);
            m10 = cjfma(l1, r0, m10);
            m11 = cjfma(l1, r1, m11);
        }
    }
    const double v[INNER_WIDTH] = {m00.x, m00.y, m01.x, m01.y, m10.x, m10.y, m11.x, m11.y};
    block_sum_store<INNER_WIDTH>(v, partial + (size_t)blockIdx.x * INNER_WIDTH);
}

// K6b.  Two-qubit transfer matrix T[i][j] = sum_rest conj(L[i,rest]) R[j,rest], i,j = bit(qa) + 2 bit(qb),
// qa < qb.  <L| O |R> = sum_ij O[i][j] T[i][j] for ANY operator O supported on (qa, qb): one read
// pass over L and R serves every Rotosolve / Rotoselect evaluation of a whole ansatz layer.
__global__ void __launch_bounds__(RED_THREADS)
sv_inner2_kernel(const double2* __restrict__ L, const double2* __restrict__ Rv, const int n, const int qa,
                 const int qb, double* __restrict__ partial) {
    double2 t[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) t[k] = make_double2(0.0, 0.0);
    const uint64_t groups = 1ull << (n - 2), ba = 1ull << qa, bb = 1ull << qb;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < groups; k += stride) {
        const uint64_t b = ins0_64(ins0_64(k, qa), qb);
        double2 l[4], r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t idx = b + ((j & 1) ? ba : 0) + ((j & 2) ? bb : 0);
            l[j] = L[idx];
            r[j] = Rv[idx];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) t[4 * i + j] = cjfma(l[i], r[j], t[4 * i + j]);
    }
    double v[INNER2_WIDTH];
#pragma unroll
    for (int k = 0; k < 16; ++k) { v[2 * k] = t[k].x; v[2 * k + 1] = t[k].y; }
    block_sum_store<INNER2_WIDTH>(v, partial + (size_t)blockIdx.x * INNER2_WIDTH);
}

// K6c.  Transfer matrix against a COMPACT bra.  <L| = suffix^+ applied to <0..0| is supported only on
// the K qubits the suffix touches: L[x] = ell[c] when x = deposit(c) over qmap, 0 elsewhere.  Then
// T[i][j] = sum_c conj(ell[c_i]) R[deposit(c)_j] needs a GATHER of 2^K amplitudes of R instead of a
// pass over all 2^n (for the newest ADAPT layer K = 2: four amplitudes).  ca < cb: compact positions
// of the open qubits (qmap[ca] = qa, qmap[cb] = qb).
struct QMap { int32_t q[40]; };
__global__ void __launch_bounds__(RED_THREADS)
sv_inner2_gather_kernel(const double2* __restrict__ ell, const int K, const double2* __restrict__ Rv, const QMap qm,
                        const int ca, const int cb, double* __restrict__ partial) {
    double2 t[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) t[k] = make_double2(0.0, 0.0);
    const uint64_t groups = 1ull << (K - 2);
    const uint64_t la = 1ull << ca, lb = 1ull << cb, ra = 1ull << qm.q[ca], rb = 1ull << qm.q[cb];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        const uint64_t c = ins0_64(ins0_64(g, ca), cb);
        uint64_t x = 0;
        for (int b = 0; b < K; ++b) x |= ((c >> b) & 1ull) << qm.q[b];
        double2 l[4], r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            l[j] = ell[c + ((j & 1) ? la : 0) + ((j & 2) ? lb : 0)];
            r[j] = Rv[x + ((j & 1) ? ra : 0) + ((j & 2) ? rb : 0)];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) t[4 * i + j] = cjfma(l[i], r[j], t[4 * i + j]);
    }
    double v[INNER2_WIDTH];
#pragma unroll
    for (int k = 0; k < 16; ++k) { v[2 * k] = t[k].x; v[2 * k + 1] = t[k].y; }
    block_sum_store<INNER2_WIDTH>(v, partial + (size_t)blockIdx.x * INNER2_WIDTH);
}

}  // namespace b200
