// sv_kernels.cuh -- sm_100a statevector kernels (complex128, little-endian qubit order).
//
//   sv_sweep_kernel<R>  K1/K2: fused gate sweep.  One CTA = one tile of 2^12 amplitudes
//                       (the 5 lowest qubits + 7 arbitrary qubits); 2^R amplitudes per thread live in
//                       registers for a whole round of gates; rounds exchange through swizzled
//                       shared memory; the first round loads from HBM (128-bit accesses, whole 128 B
//                       lines per warp access) and the last round stores the same way; gates on the
//                       three lane qubits of those rounds go through warp shuffles.  The sweep's
//                       program is a kernel parameter (constant bank, uniform decode).
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
namespace b200 {

constexpr int SWEEP_THREADS = 1 << (TILE_BITS - REG_BITS);  // 256
constexpr size_t FUSED_SMEM_BYTES = (((size_t)1 << TILE_BITS) + 4 * SWEEP_THREADS) * 16;   // tile + 4 accumulators per thread
constexpr int RED_THREADS = 256;
constexpr int EXPZ_WIDTH = 32;   // doubles per block partial
constexpr int RDM3_WIDTH = 48;
constexpr int RDM4_WIDTH = 96;
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
B200_HD double2 ld_c(const double* m, int k) { return make_double2(m[2 * k], m[2 * k + 1]); }
// runtime selection among register-resident constants WITHOUT dynamic indexing (which would force local memory)
B200_HD double2 sel2(const double* m, const int base, const int u) {
    const double2 lo = ld_c(m, base), hi = ld_c(m, base + 1);
    return make_double2(u ? hi.x : lo.x, u ? hi.y : lo.y);
}
B200_HD double2 sel4(const double* m, const int u0, const int u1) {
    const double2 lo = sel2(m, 0, u0), hi = sel2(m, 2, u0);
    return make_double2(u1 ? hi.x : lo.x, u1 ? hi.y : lo.y);
}

// ---------------------------------------------------------------------------------------------
// register-resident op bodies: every register index is a template parameter, so the 2^R amplitudes
// stay in named registers (no local memory) and the loops unroll to straight-line code
// ---------------------------------------------------------------------------------------------
template <int R, int TB>
B200_HD void op_mat1(double2 (&a)[1 << R], const double* __restrict__ m) {
    const double2 m00 = ld_c(m, 0), m01 = ld_c(m, 1), m10 = ld_c(m, 2), m11 = ld_c(m, 3);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        if (j & (1 << TB)) continue;
        const double2 x = a[j], y = a[j | (1 << TB)];
        a[j] = cfma(m01, y, cmul(m00, x));
        a[j | (1 << TB)] = cfma(m11, y, cmul(m10, x));
    }
}

// In-place exchange.  On the device the exchange is an opaque asm block with read-write operands: written as
// plain assignments it turns the ops switch into a register PERMUTATION at the loop back-edge, which the
// compiler resolved by copying all 2^R amplitudes (64 moves) on every op of every round (ncu source page,
// profiles/prof_pipe_r01m: 54 % of the executed instructions were IMAD.MOV/MOV).
B200_HD void swap_amp(double2& x, double2& y) {
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .f64 t;\n\tmov.f64 t, %0;\n\tmov.f64 %0, %1;\n\tmov.f64 %1, t;\n\t}" : "+d"(x.x), "+d"(y.x));
    asm("{\n\t.reg .f64 t;\n\tmov.f64 t, %0;\n\tmov.f64 %0, %1;\n\tmov.f64 %1, t;\n\t}" : "+d"(x.y), "+d"(y.y));
#else
    const double2 t = x; x = y; y = t;
#endif
}

template <int R, int TB>
B200_HD void op_x(double2 (&a)[1 << R]) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        if (j & (1 << TB)) continue;
        swap_amp(a[j], a[j | (1 << TB)]);
    }
}

template <int R, int TB, int CB>
B200_HD void op_cx(double2 (&a)[1 << R]) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        if ((j & (1 << TB)) || !(j & (1 << CB))) continue;
        swap_amp(a[j], a[j | (1 << TB)]);
    }
}

template <int R, int TB0, int TB1>
B200_HD void op_mat2(double2 (&a)[1 << R], const double* __restrict__ mm) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        if (j & ((1 << TB0) | (1 << TB1))) continue;
        const double2 v0 = a[j], v1 = a[j | (1 << TB0)], v2 = a[j | (1 << TB1)],
                      v3 = a[j | (1 << TB0) | (1 << TB1)];
        double2 o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double2 acc = cmul(ld_c(mm, 4 * r), v0);
            acc = cfma(ld_c(mm, 4 * r + 1), v1, acc);
            acc = cfma(ld_c(mm, 4 * r + 2), v2, acc);
            acc = cfma(ld_c(mm, 4 * r + 3), v3, acc);
            o[r] = acc;
        }
        a[j] = o[0];
        a[j | (1 << TB0)] = o[1];
        a[j | (1 << TB1)] = o[2];
        a[j | (1 << TB0) | (1 << TB1)] = o[3];
    }
}

// amplitudes with register bit TB set are multiplied by `ratio` (the base phase went into `pend`)
template <int R, int TB>
B200_HD void op_diag1(double2 (&a)[1 << R], const double2 ratio) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j)
        if (j & (1 << TB)) a[j] = cmul(a[j], ratio);
}

template <int R, int TB0, int TB1>
B200_HD void op_diag2(double2 (&a)[1 << R], const double2 r10, const double2 r01, const double2 r11) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        const bool b0 = j & (1 << TB0), b1 = j & (1 << TB1);
        if (b0 && b1) a[j] = cmul(a[j], r11);
        else if (b0) a[j] = cmul(a[j], r10);
        else if (b1) a[j] = cmul(a[j], r01);
    }
}

// generic diagonal (non-unitary input with a zero phase entry): phase index = bit(q0) + 2 bit(q1), each
// bit either a register bit (r >= 0) or a bit of the thread's base index g
template <int R, class OP>
B200_HD void op_diagraw_impl(double2 (&a)[1 << R], const OP& op, const uint64_t g) {
    const int u0 = op.dq0 >= 0 ? (int)((g >> op.dq0) & 1ull) : 0;
    const int u1 = op.dq1 >= 0 ? (int)((g >> op.dq1) & 1ull) : 0;
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        const int b0 = op.r0 >= 0 ? ((j >> op.r0) & 1) : u0;
        const int b1 = op.r1 >= 0 ? ((j >> op.r1) & 1) : u1;
        a[j] = cmul(a[j], sel4(op.m, b0, b1));
    }
}

// ---------------------------------------------------------------------------------------------
// lane ops (HBM rounds): the mixing target is one of the warp-lane bits 0..COAL_BITS-1, partners are
// exchanged through EX (device: __shfl_xor_sync; emulator: a snapshot of the warp's registers)
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
struct ShflExchange {
    __device__ __forceinline__ double2 operator()(const double2 v, const int /*j*/, const int lane_mask) const {
        return make_double2(__shfl_xor_sync(0xffffffffu, v.x, lane_mask), __shfl_xor_sync(0xffffffffu, v.y, lane_mask));
    }
};
#endif

// X on lane bit `lb`; control: none, register bit rc, or thread-level (ctl = the control's value for this thread;
// the partner has the same value because the control is not the target)
template <int R, class EX>
B200_HD void op_xlane(double2 (&a)[1 << R], const int lb, const int rc, const bool ctl, const EX& ex) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        if (rc >= 0 && !((j >> rc) & 1)) continue;
        const double2 v = ex(a[j], j, 1 << lb);
        if (ctl) a[j] = v;
    }
}

// dense 2x2 on lane bit `lb`: every thread computes its own output row
template <int R, class EX>
B200_HD void op_mat1lane(double2 (&a)[1 << R], const double* __restrict__ m, const int lb, const uint32_t lane,
                         const EX& ex) {
    const bool hi = (lane >> lb) & 1u;
    const double2 cs = hi ? ld_c(m, 3) : ld_c(m, 0);   // coefficient of the thread's own amplitude
    const double2 co = hi ? ld_c(m, 2) : ld_c(m, 1);   // coefficient of the partner's
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        const double2 p = ex(a[j], j, 1 << lb);
        a[j] = cfma(co, p, cmul(cs, a[j]));
    }
}

// ---------------------------------------------------------------------------------------------
// K1/K2: fused sweep
// ---------------------------------------------------------------------------------------------
// Base index of tile `tile`: zero bits inserted at the tile's high qubits.
B200_HD uint64_t sweep_tile_base(const SweepProg& sp, const uint32_t tile) {
    uint64_t y = tile;
    for (int i = sp.c; i < TILE_BITS; ++i) y = ins0_64(y, sp.tileq[i] - sp.c);
    return y << sp.c;
}

// Thread `tid`'s tile-local index with zeros at the round's register positions, and its global index.
template <int R>
B200_HD void round_index(const SweepProg& sp, const PRound& rd, const uint64_t tile_base, const uint32_t tid,
                         uint32_t& tl, uint64_t& g) {
    tl = tid;
#pragma unroll
    for (int k = 0; k < R; ++k) tl = ins0_32(tl, rd.regpos[k]);
    const int c = sp.c;
    g = tile_base | (uint64_t)(tl & ((1u << c) - 1u));
    for (int i = c; i < TILE_BITS; ++i) g |= (uint64_t)((tl >> i) & 1u) << sp.tileq[i];
}

// Per-thread address masks of the X-type ops folded into a round's load / store (sv_plan.h, PFold): the control is a
// bit of the thread's base index g (or absent).  Uniform loop, uniform constant-bank loads, <= MAX_FOLD iterations.
B200_HD uint32_t fold_smask(const PFold* f, const int n, const uint64_t g) {
    uint32_t m = 0;
    for (int i = 0; i < n; ++i) {
        const bool ctl = f[i].cq < 0 || ((g >> f[i].cq) & 1ull);
        m ^= ctl ? f[i].smask : 0u;
    }
    return m;
}
// HBM rounds: returns the thread's base index WITH the flips applied.  A folded op may flip a lane qubit that a
// neighbouring folded op is controlled by, so the ops are evaluated one after the other on the running index: the source
// index of a load is f_1(f_2(...f_m(y))) (every X-type op is its own inverse: REVERSE order), the destination of a store is
// f_m(...f_1(y)) (forward order).
template <bool REVERSE>
B200_HD uint64_t fold_gindex(const PFold* f, const int n, const uint64_t g) {
    uint64_t x = g;
    for (int k = 0; k < n; ++k) {
        const int i = REVERSE ? n - 1 - k : k;
        const bool ctl = f[i].cq < 0 || ((x >> f[i].cq) & 1ull);
        x ^= ctl ? f[i].gmask : 0ull;
    }
    return x;
}

// An EMBEDDED source: the state is phi (2^K amplitudes, a smaller engine's slot) on the qubits q[0..K) and |0> on every
// other qubit -- amplitude x is phi[extract(x)] when x has no bit outside those qubits, else 0.  A sweep that starts from
// it has no read pass over the register (and needs no zero fill + scatter pass before it): 16 * 2^n bytes instead of 48.
struct EmbedSrc {
    const double2* phi;
    uint64_t outside;     // bits of the global index that must be zero
    int32_t K, pad;
    int32_t q[40];
    // extract() is GF(2)-linear in the index, so the compact index of register amplitude j of a thread is
    // extract(thread index) ^ coff[j]: ONE bit-gather loop per thread and tile instead of 16 (embed_prepare)
    uint64_t coff[1 << REG_BITS];
    uint64_t regspan;     // OR of the first round's register offsets
    uint64_t outside_nontile;   // bits of `outside` that are not tile qubits: a tile whose base has one set is all zero
    uint64_t skipmask;          // tile qubits | outside_nontile: the index bits a NON-ZERO tile's number does not cover
    int32_t nq, n_skip;         // register size; popcount(outside_nontile)
};
// number t of a non-zero tile -> its base index: the bits of t go to the positions outside `skipmask`, ascending
B200_HD uint64_t embed_tile_base(const EmbedSrc& es, const uint32_t t) {
    uint64_t out = 0;
    int k = 0;
    for (int b = 0; b < es.nq; ++b)
        if (!((es.skipmask >> b) & 1ull)) { out |= (uint64_t)((t >> k) & 1u) << b; ++k; }
    return out;
}
// The mirror image on the STORE side: only the amplitudes with no bit outside q[0..K) are kept, phi[extract(x)] = value --
// the projection of the swept state onto |0> of every other qubit (sv_gather_kernel) without writing the swept state.
struct ProjectDst {
    double2* phi;
    uint64_t outside;
    int32_t K, pad;
    int32_t q[40];
    uint64_t coff[1 << REG_BITS];   // extract(goff_st[j]) of the last round
    uint64_t regspan;
};
B200_HD uint64_t embed_extract(const EmbedSrc& es, const uint64_t x) {
    uint64_t c = 0;
    for (int b = 0; b < es.K; ++b) c |= ((x >> es.q[b]) & 1ull) << b;
    return c;
}
// host: tables of the sweep's first round
inline void embed_prepare(EmbedSrc& es, const SweepProg& sp) {
    uint64_t tilemask = 0;
    for (int i = 0; i < TILE_BITS; ++i) tilemask |= 1ull << sp.tileq[i];
    es.outside_nontile = es.outside & ~tilemask;
    es.skipmask = tilemask | es.outside_nontile;
    es.n_skip = 0;
    for (uint64_t m = es.outside_nontile; m; m &= m - 1) ++es.n_skip;
    es.regspan = 0;
    for (int j = 0; j < (1 << REG_BITS); ++j) {
        es.coff[j] = embed_extract(es, sp.rounds[0].goff_ld[j]);
        es.regspan |= sp.rounds[0].goff_ld[j];
    }
}
inline void project_prepare(ProjectDst& pd, const SweepProg& sp) {
    const PRound& rd = sp.rounds[sp.nrounds - 1];
    pd.regspan = 0;
    for (int j = 0; j < (1 << REG_BITS); ++j) {
        uint64_t c = 0;
        for (int b = 0; b < pd.K; ++b) c |= ((rd.goff_st[j] >> pd.q[b]) & 1ull) << b;
        pd.coff[j] = c;
        pd.regspan |= rd.goff_st[j];
    }
}
template <int R>
B200_HD void project_store(const double2 (&a)[1 << R], const ProjectDst& pd, const PRound& rd, const uint64_t gm) {
    if (gm & pd.outside & ~pd.regspan) return;      // none of the thread's amplitudes survives the projection
    uint64_t cg = 0;
    for (int b = 0; b < pd.K; ++b) cg |= ((gm >> pd.q[b]) & 1ull) << b;
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        const uint64_t x = gm ^ rd.goff_st[j];
        if (!(x & pd.outside)) pd.phi[cg ^ pd.coff[j]] = a[j];
    }
}

template <int R>
B200_HD void embed_load(double2 (&a)[1 << R], const EmbedSrc& es, const PRound& rd, const uint64_t gm) {
    if (gm & es.outside & ~es.regspan) {     // a bit outside the embedded qubits is set whatever j is: all zeros
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) a[j] = make_double2(0.0, 0.0);
        return;
    }
    const uint64_t cg = embed_extract(es, gm);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        const uint64_t x = gm ^ rd.goff_ld[j];
        a[j] = (x & es.outside) ? make_double2(0.0, 0.0) : es.phi[cg ^ es.coff[j]];
    }
}

// src == nullptr: the source is |0..0> (no read pass, no separate fill pass), or the embedded state `es` if given
template <int R>
B200_HD void round_load_hbm(double2 (&a)[1 << R], const double2* __restrict__ src, const SweepProg& sp,
                            const PRound& rd, const uint64_t g, const EmbedSrc* es = nullptr) {
    const uint64_t gm = fold_gindex<true>(rd.lead, rd.n_lead, g);     // (register positions are zero in g: ^ == |)
    if (es != nullptr) { embed_load<R>(a, *es, rd, gm); return; }
    if (src == nullptr) {
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) a[j] = make_double2(gm == rd.goff_ld[j] ? 1.0 : 0.0, 0.0);
        return;
    }
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) a[j] = src[gm ^ rd.goff_ld[j]];
}

template <int R>
B200_HD void round_store_hbm(const double2 (&a)[1 << R], double2* __restrict__ dst, const SweepProg& sp,
                             const PRound& rd, const uint64_t g, const ProjectDst* pd = nullptr) {
    const uint64_t gm = fold_gindex<false>(rd.trail, rd.n_trail, g);
    if (pd != nullptr) { project_store<R>(a, *pd, rd, gm); return; }
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) dst[gm ^ rd.goff_st[j]] = a[j];
}

// tls: the thread's swizzled tile index (zeros at the register positions) -- the folded flips are applied here
template <int R>
B200_HD void round_load_smem(double2 (&a)[1 << R], const double2* tile_smem, const PRound& rd, const uint32_t tls,
                             const uint64_t g) {
    const uint32_t t = tls ^ fold_smask(rd.lead, rd.n_lead, g);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) a[j] = tile_smem[t ^ (uint32_t)rd.soff[j]];
}

template <int R>
B200_HD void round_store_smem(const double2 (&a)[1 << R], double2* tile_smem, const PRound& rd, const uint32_t tls,
                              const uint64_t g) {
    const uint32_t t = tls ^ fold_smask(rd.trail, rd.n_trail, g);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) tile_smem[t ^ (uint32_t)rd.soff_st[j]] = a[j];
}

// Linear (unswizzled) tile access: the layout the bulk copies of the pipelined kernel land / pick up.
template <int R>
B200_HD void round_load_lin(double2 (&a)[1 << R], const double2* tile_smem, const PRound& rd, const uint32_t tl) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        uint32_t off = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) if (j >> k & 1) off |= 1u << rd.regpos[k];
        a[j] = tile_smem[tl | off];
    }
}

template <int R>
B200_HD void round_store_lin(const double2 (&a)[1 << R], double2* tile_smem, const PRound& rd, const uint32_t tl) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
        uint32_t off = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) if (j >> k & 1) off |= 1u << rd.regpos[k];
        tile_smem[tl | off] = a[j];
    }
}

// Global amplitude index of row `row` of a tile: a row is the 2^c contiguous amplitudes of the tile's
// leading contiguous qubits, the row number spreads over the tile's remaining (scattered) qubits.
B200_HD uint64_t sweep_row_index(const SweepProg& sp, const uint64_t tile_base, const uint32_t row) {
    uint64_t g = tile_base;
    for (int i = sp.c; i < TILE_BITS; ++i) g |= (uint64_t)((row >> (i - sp.c)) & 1u) << sp.tileq[i];
    return g;
}

template <int R>
B200_HD void apply_pend(double2 (&a)[1 << R], const double2 pend) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) a[j] = cmul(a[j], pend);
}

// Warp-uniform header of an op, read one op AHEAD of its use so that the constant-bank latency of the
// decode (and of the 8 doubles of phases / matrix entries) hides behind the previous op's arithmetic.
struct OpHead {
    int32_t code, flush, r0, r1, cq, dq0, dq1, mat2;
    double m[8];
};
B200_HD OpHead load_head(const POp& op) {
    OpHead h;
    h.code = (op.flush >> 8) & 0xff; h.flush = op.flush & 1;
    h.r0 = op.r0; h.r1 = op.r1; h.cq = op.cq; h.dq0 = op.dq0; h.dq1 = op.dq1; h.mat2 = op.mat2;
#pragma unroll
    for (int k = 0; k < 8; ++k) h.m[k] = op.m[k];
    return h;
}

// The same header by reference: fields are read from the constant bank where they are used (no registers
// held across the switch) -- for the direct kernel and the emulator.
struct OpRef {
    int32_t code, flush;
    const int32_t &r0, &r1, &cq, &dq0, &dq1, &mat2;
    const double (&m)[8];
    B200_HD explicit OpRef(const POp& op)
        : code((op.flush >> 8) & 0xff), flush(op.flush & 1), r0(op.r0), r1(op.r1), cq(op.cq), dq0(op.dq0),
          dq1(op.dq1), mat2(op.mat2), m(op.m) {}
};

// One op of a round on the thread's registers.  `pend` collects the phases that multiply all 2^R
// amplitudes of the thread alike: they commute with every register op of the round and are applied once at
// the end of the round (or before a lane exchange, which they do not commute with: h.flush).
// ONE warp-uniform switch selects a fully unrolled, statically indexed body.
template <int R, class HEAD, class EX>
B200_HD void apply_decoded(double2 (&a)[1 << R], const SweepProg& sp, const HEAD& h, const uint64_t g,
                           const uint32_t lane, double2& pend, const EX& ex) {
#define B200_DIAG1(TB) { const int u = h.dq1 >= 0 ? (int)((g >> h.dq1) & 1ull) : 0;                 \
                         pend = cmul(pend, sel2(h.m, 0, u)); op_diag1<R, TB>(a, sel2(h.m, 2, u)); } break
#define B200_DIAG2(T0, T1) { pend = cmul(pend, ld_c(h.m, 0));                                         \
                             op_diag2<R, T0, T1>(a, ld_c(h.m, 1), ld_c(h.m, 2), ld_c(h.m, 3)); } break
#define B200_XREG(TB) if (h.cq < 0 || ((g >> h.cq) & 1ull)) op_x<R, TB>(a); break
    switch (h.code) {
    case 0: {
        const int u0 = h.dq0 >= 0 ? (int)((g >> h.dq0) & 1ull) : 0;
        const int u1 = h.dq1 >= 0 ? (int)((g >> h.dq1) & 1ull) : 0;
        pend = cmul(pend, sel4(h.m, u0, u1));
        break;
    }
    case 1: B200_DIAG1(0);
    case 2: B200_DIAG1(1);
    case 3: B200_DIAG1(2);
    case 4: B200_DIAG1(3);
    case 5: B200_DIAG2(0, 1);
    case 6: B200_DIAG2(0, 2);
    case 7: B200_DIAG2(0, 3);
    case 8: B200_DIAG2(1, 2);
    case 9: B200_DIAG2(1, 3);
    case 10: B200_DIAG2(2, 3);
    case 11: op_diagraw_impl<R>(a, h, g); break;
    case 12: B200_XREG(0);
    case 13: B200_XREG(1);
    case 14: B200_XREG(2);
    case 15: B200_XREG(3);
    case 16: op_cx<R, 0, 1>(a); break;
    case 17: op_cx<R, 0, 2>(a); break;
    case 18: op_cx<R, 0, 3>(a); break;
    case 19: op_cx<R, 1, 0>(a); break;
    case 20: op_cx<R, 1, 2>(a); break;
    case 21: op_cx<R, 1, 3>(a); break;
    case 22: op_cx<R, 2, 0>(a); break;
    case 23: op_cx<R, 2, 1>(a); break;
    case 24: op_cx<R, 2, 3>(a); break;
    case 25: op_cx<R, 3, 0>(a); break;
    case 26: op_cx<R, 3, 1>(a); break;
    case 27: op_cx<R, 3, 2>(a); break;
    case 28: op_mat1<R, 0>(a, h.m); break;
    case 29: op_mat1<R, 1>(a, h.m); break;
    case 30: op_mat1<R, 2>(a, h.m); break;
    case 31: op_mat1<R, 3>(a, h.m); break;
    case 32: op_mat2<R, 0, 1>(a, sp.mat2[h.mat2]); break;
    case 33: op_mat2<R, 0, 2>(a, sp.mat2[h.mat2]); break;
    case 34: op_mat2<R, 0, 3>(a, sp.mat2[h.mat2]); break;
    case 35: op_mat2<R, 1, 2>(a, sp.mat2[h.mat2]); break;
    case 36: op_mat2<R, 1, 3>(a, sp.mat2[h.mat2]); break;
    case 37: op_mat2<R, 2, 3>(a, sp.mat2[h.mat2]); break;
    case 38: {
        if (h.flush) { apply_pend<R>(a, pend); pend = make_double2(1.0, 0.0); }
        const bool ctl = h.cq < 0 || ((g >> h.cq) & 1ull);
        op_xlane<R>(a, h.r0, h.r1, ctl, ex);
        break;
    }
    default:   // 39: dense 2x2 on a lane qubit
        if (h.flush) { apply_pend<R>(a, pend); pend = make_double2(1.0, 0.0); }
        op_mat1lane<R>(a, h.m, h.r0, lane, ex);
        break;
    }
#undef B200_DIAG1
#undef B200_DIAG2
#undef B200_XREG
}

// emulator entry: one op, decoded on the spot
template <int R, class EX>
B200_HD void apply_op(double2 (&a)[1 << R], const SweepProg& sp, const POp& op, const uint64_t g,
                      const uint32_t lane, double2& pend, const EX& ex) {
    const OpRef h(op);
    apply_decoded<R>(a, sp, h, g, lane, pend, ex);
}

// all ops of a round; PREFETCH: software-pipelined decode (costs ~24 registers: the pipelined kernel has
// them, the direct kernel with its 128-register cap does not)
template <int R, bool PREFETCH, class EX>
B200_HD void round_ops(double2 (&a)[1 << R], const SweepProg& sp, const PRound& rd, const uint64_t g,
                       const uint32_t lane, const EX& ex) {
    double2 pend = make_double2(1.0, 0.0);
    int o = rd.op_begin;
    const int oe = rd.op_end;
    if (!PREFETCH) {
        for (; o < oe; ++o) apply_op<R>(a, sp, sp.ops[o], g, lane, pend, ex);
    } else if (o < oe) {
        OpHead h = load_head(sp.ops[o]);
        for (; o < oe; ++o) {
            const OpHead nh = load_head(sp.ops[o + 1]);   // ops[] has one slot of slack
            apply_decoded<R>(a, sp, h, g, lane, pend, ex);
            h = nh;
        }
    }
    if (rd.has_pend) apply_pend<R>(a, pend);
}

// ---------------------------------------------------------------------------------------------
// K1 + K6b fused: epilogue of a sweep whose result is contracted straight away with a second state into the 4x4
// transfer matrix of an open pair,  T[i][j] = sum_rest conj(S[i,rest]) O[j,rest]  (S = the swept state, O = `other`).
// The last round leaves the tile in shared memory; thread `tid` then owns the 16 amplitudes EpiProg assigns to it: it
// loads them from O (coalesced: thread bits 0..4 are qubits 0..4), reads the 4 pair-partners of each from the tile, stores
// its own amplitude of S (optional), and accumulates column j = its own pair value.  Algorithmic bytes: read S, read O,
// write S = 48 * 2^n, against 64 * 2^n for the sweep followed by the separate transfer pass (sv_inner2_kernel).
// ---------------------------------------------------------------------------------------------
struct EpiIdx {
    uint32_t ls[4];   // swizzled tile index of pair value i with the thread's other bits (amplitude bits zero)
    uint64_t g0;      // global index offset of the thread's bits (amplitude bits zero)
    int j;            // the thread's pair value = column of T
};

B200_HD EpiIdx epi_index(const SweepProg& sp, const EpiProg& ep, const uint32_t tid) {
    uint32_t tl = 0;
    for (int b = 0; b < TILE_BITS - REG_BITS; ++b) tl |= ((tid >> b) & 1u) << ep.tpos[b];
    EpiIdx ix;
    ix.j = (int)(((tid >> ep.ja) & 1u) | (((tid >> ep.jb) & 1u) << 1));
    const uint32_t rest = tl & ~((1u << ep.pa) | (1u << ep.pb));
#pragma unroll
    for (int i = 0; i < 4; ++i) ix.ls[i] = swz(rest | ((uint32_t)(i & 1) << ep.pa) | ((uint32_t)(i >> 1) << ep.pb));
    uint64_t g = 0;
    for (int i = 0; i < TILE_BITS; ++i) g |= (uint64_t)((tl >> i) & 1u) << sp.tileq[i];
    ix.g0 = g;
    return ix;
}

template <int R>
B200_HD void epi_tile(const double2* tile_smem, const double2* __restrict__ other, double2* __restrict__ dst,
                      const EpiProg& ep, const EpiIdx& ix, const uint64_t tile_base, double2 (&t)[4]) {
    const uint64_t G = tile_base | ix.g0;
    constexpr int NB = 2, HALF = (1 << R) / NB;   // batches of 8 loads in flight (all 16 at once spill at the 128-register cap)
#pragma unroll
    for (int h = 0; h < NB; ++h) {
        double2 r[HALF];
#pragma unroll
        for (int m = 0; m < HALF; ++m) r[m] = other[G | ep.moff_g[h * HALF + m]];
#pragma unroll
        for (int m = 0; m < HALF; ++m) {
            const uint32_t mo = ep.moff_sw[h * HALF + m];
            const double2 l0 = tile_smem[ix.ls[0] ^ mo], l1 = tile_smem[ix.ls[1] ^ mo], l2 = tile_smem[ix.ls[2] ^ mo],
                          l3 = tile_smem[ix.ls[3] ^ mo];
            if (dst != nullptr) {
                const double2 lo = (ix.j & 1) ? l1 : l0, hi = (ix.j & 1) ? l3 : l2;
                dst[G | ep.moff_g[h * HALF + m]] = (ix.j & 2) ? hi : lo;
            }
            t[0] = cjfma(l0, r[m], t[0]);
            t[1] = cjfma(l1, r[m], t[1]);
            t[2] = cjfma(l2, r[m], t[2]);
            t[3] = cjfma(l3, r[m], t[3]);
        }
    }
}

// zeros for the thread's 16 amplitudes of a tile (tiles of an embedded source that are zero as a whole)
template <int R>
B200_HD void epi_zero_tile(double2* __restrict__ dst, const EpiProg& ep, const EpiIdx& ix, const uint64_t tile_base) {
    const uint64_t G = tile_base | ix.g0;
#pragma unroll
    for (int m = 0; m < (1 << R); ++m) dst[G | ep.moff_g[m]] = make_double2(0.0, 0.0);
}

#ifdef __CUDACC__
// The whole program of the sweep is a kernel parameter: op decode is uniform constant-bank loads and
// uniform branches.  One CTA owns one tile at a time (persistent grid-stride over tiles).
// sv_sweep_project_kernel: the same sweep whose last round keeps only the projected amplitudes (ProjectDst).
template <int R>
__global__ void __launch_bounds__(SWEEP_THREADS, 2)
sv_sweep_project_kernel(const double2* __restrict__ src, const __grid_constant__ SweepProg sp, const uint32_t ntiles,
                        const __grid_constant__ ProjectDst pd) {
    extern __shared__ __align__(16) double2 tile_smem[];
    const int nr = sp.nrounds;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const ShflExchange ex;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t tile_base = sweep_tile_base(sp, tile);
        for (int r = 0; r < nr; ++r) {
            const PRound& rd = sp.rounds[r];
            uint32_t tl; uint64_t g;
            round_index<R>(sp, rd, tile_base, tid, tl, g);
            const uint32_t tls = swz(tl);
            double2 a[1 << R];
            if (r == 0) round_load_hbm<R>(a, src, sp, rd, g);
            else round_load_smem<R>(a, tile_smem, rd, tls, g);
            round_ops<R, false>(a, sp, rd, g, lane, ex);
            if (r == nr - 1) round_store_hbm<R>(a, nullptr, sp, rd, g, &pd);
            else { round_store_smem<R>(a, tile_smem, rd, tls, g); __syncthreads(); }
        }
        if (nr > 1) __syncthreads();
    }
}

template <int R, bool EMBED = false>
__global__ void __launch_bounds__(SWEEP_THREADS, 2)
sv_sweep_kernel(const double2* __restrict__ src, double2* __restrict__ dst, const __grid_constant__ SweepProg sp,
                const uint32_t ntiles, const __grid_constant__ EmbedSrc es) {
    extern __shared__ __align__(16) double2 tile_smem[];
    static_assert(R == REG_BITS, "register bits");
    const int nr = sp.nrounds;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const ShflExchange ex;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t tile_base = sweep_tile_base(sp, tile);
        for (int r = 0; r < nr; ++r) {
            const PRound& rd = sp.rounds[r];
            uint32_t tl; uint64_t g;
            round_index<R>(sp, rd, tile_base, tid, tl, g);
            const uint32_t tls = swz(tl);
            double2 a[1 << R];
            if (r == 0) round_load_hbm<R>(a, src, sp, rd, g, EMBED ? &es : nullptr);
            else round_load_smem<R>(a, tile_smem, rd, tls, g);
            round_ops<R, false>(a, sp, rd, g, lane, ex);
            if (r == nr - 1) round_store_hbm<R>(a, dst, sp, rd, g);
            else { round_store_smem<R>(a, tile_smem, rd, tls, g); __syncthreads(); }
        }
        if (nr > 1) __syncthreads();
    }
}

// Fused sweep + transfer pass.  partial[blockIdx.x * 32 + 2 * (4 * i + j) + {0, 1}] = this CTA's share of T[i][j].
template <int R, bool EMBED = false>
__global__ void __launch_bounds__(SWEEP_THREADS, 2)
sv_sweep_inner2_kernel(const double2* __restrict__ src, double2* __restrict__ dst, const double2* __restrict__ other,
                       const __grid_constant__ SweepProg sp, const __grid_constant__ EpiProg ep, const uint32_t ntiles,
                       double* __restrict__ partial, const __grid_constant__ EmbedSrc es) {
    extern __shared__ __align__(16) double2 tile_smem[];
    static_assert(R == REG_BITS, "register bits");
    const int nr = sp.nrounds;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const ShflExchange ex;
    // the thread's 4 accumulators wait in shared memory while the rounds of the next tile need every register
    double2* acc = tile_smem + ((size_t)1 << TILE_BITS);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i * SWEEP_THREADS + tid] = make_double2(0.0, 0.0);
    // T only from an embedded source: the tiles that are zero as a whole are not even enumerated (skipping them inside a
    // round-robin loop left the CTAs that drew them idle: the lowest zero qubit is the lowest bit of the tile number)
    const bool compact = EMBED && dst == nullptr && es.n_skip > 0;
    const uint32_t nt = compact ? (ntiles >> es.n_skip) : ntiles;
    for (uint32_t tile = blockIdx.x; tile < nt; tile += gridDim.x) {
        const uint64_t tile_base = compact ? embed_tile_base(es, tile) : sweep_tile_base(sp, tile);
        if (EMBED && (tile_base & es.outside_nontile)) {
            // the whole source tile is zero and the gates keep it zero: nothing for T, nothing to read (uniform per CTA)
            if (dst != nullptr) epi_zero_tile<R>(dst, ep, epi_index(sp, ep, tid), tile_base);
            continue;
        }
        for (int r = 0; r < nr; ++r) {
            const PRound& rd = sp.rounds[r];
            uint32_t tl; uint64_t g;
            round_index<R>(sp, rd, tile_base, tid, tl, g);
            const uint32_t tls = swz(tl);
            double2 a[1 << R];
            if (r == 0) round_load_hbm<R>(a, src, sp, rd, g, EMBED ? &es : nullptr);
            else round_load_smem<R>(a, tile_smem, rd, tls, g);
            round_ops<R, false>(a, sp, rd, g, lane, ex);
            round_store_smem<R>(a, tile_smem, rd, tls, g);
            __syncthreads();
        }
        const EpiIdx ix = epi_index(sp, ep, tid);
        double2 t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = acc[i * SWEEP_THREADS + tid];
        epi_tile<R>(tile_smem, other, dst, ep, ix, tile_base, t);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i * SWEEP_THREADS + tid] = t[i];
        __syncthreads();                              // the tile has been read: the next one may land
    }
    // ---- CTA reduction: 64 threads per column j, fixed order ----
    if (tid < 32) {
        const int k = (int)tid >> 1, part = (int)tid & 1, i = k >> 2, j = k & 3;
        const double* sh = reinterpret_cast<const double*>(acc + i * SWEEP_THREADS);
        double s = 0.0;
        for (uint32_t u = 0; u < (uint32_t)SWEEP_THREADS; ++u) {
            const int ju = (int)(((u >> ep.ja) & 1u) | (((u >> ep.jb) & 1u) << 1));
            if (ju == j) s += sh[2 * u + part];
        }
        partial[(size_t)blockIdx.x * INNER2_WIDTH + tid] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// K1/K2, pipelined: one persistent CTA per SM, three 64 KB tile buffers in shared memory.
//   producer warp : bulk async copies (TMA unit, cp.async.bulk -> SASS UBLKCP) HBM -> buffer for tile k+3
//                   and buffer -> HBM for tile k, one copy per tile row (2^c contiguous amplitudes,
//                   >= 512 B), completion through mbarriers / bulk groups;
//   8 compute warps: all rounds of tile k in place in its buffer (first round reads the linear layout
//                   the copies land, middle rounds use the swizzled layout, last round writes linear).
// HBM traffic is fully asynchronous to the rounds: the loads of the next two tiles and the store of the
// previous one are in flight while a tile is being computed.
// ---------------------------------------------------------------------------------------------
constexpr int PIPE_STAGES = 3;
constexpr int PIPE_THREADS = SWEEP_THREADS + 32;
constexpr uint32_t TILE_BYTES = (1u << TILE_BITS) * 16u;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, const uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, const uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, const uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, const uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, const uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// barrier among the 8 compute warps only (the producer warp never joins it)
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, %0;" ::"n"(SWEEP_THREADS) : "memory"); }

template <int R>
__global__ void __launch_bounds__(PIPE_THREADS, 1)
sv_sweep_pipe_kernel(const double2* __restrict__ src, double2* __restrict__ dst, const __grid_constant__ SweepProg sp,
                     const uint32_t ntiles) {
    extern __shared__ __align__(128) double2 pipe_smem[];   // PIPE_STAGES tile buffers
    __shared__ __align__(8) uint64_t full_bar[PIPE_STAGES], done_bar[PIPE_STAGES];
    static_assert(R == REG_BITS, "register bits");
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int b = 0; b < PIPE_STAGES; ++b) { mbar_init(&full_bar[b], 1); mbar_init(&done_bar[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= SWEEP_THREADS) {
        // ---------------- producer warp ----------------
        const int c = sp.c;
        const uint32_t rows = 1u << (TILE_BITS - c), row_bytes = 16u << c;
        auto issue_load = [&](const uint32_t k) {
            const uint32_t b = k % PIPE_STAGES;
            const uint64_t base = sweep_tile_base(sp, blockIdx.x + k * gridDim.x);
            if (lane == 0) mbar_arrive_expect_tx(&full_bar[b], TILE_BYTES);
            __syncwarp();
            double2* buf = pipe_smem + ((size_t)b << TILE_BITS);
            for (uint32_t row = lane; row < rows; row += 32)
                bulk_g2s(buf + ((size_t)row << c), src + sweep_row_index(sp, base, row), row_bytes, &full_bar[b]);
        };
        for (uint32_t k = 0; k < my_tiles && k < (uint32_t)PIPE_STAGES; ++k) issue_load(k);
        for (uint32_t k = 0; k < my_tiles; ++k) {
            const uint32_t b = k % PIPE_STAGES;
            mbar_wait(&done_bar[b], (k / PIPE_STAGES) & 1u);
            const uint64_t base = sweep_tile_base(sp, blockIdx.x + k * gridDim.x);
            double2* buf = pipe_smem + ((size_t)b << TILE_BITS);
            for (uint32_t row = lane; row < rows; row += 32)
                bulk_s2g(dst + sweep_row_index(sp, base, row), buf + ((size_t)row << c), row_bytes);
            bulk_commit();
            if (k + PIPE_STAGES < my_tiles) {
                bulk_wait_read0();       // this lane's rows have left the buffer ...
                __syncwarp();            // ... and so have every other lane's
                issue_load(k + PIPE_STAGES);
            }
        }
        bulk_wait0();
        return;
    }

    // ---------------- compute warps ----------------
    const ShflExchange ex;
    const int nr = sp.nrounds;
    for (uint32_t k = 0; k < my_tiles; ++k) {
        const uint32_t b = k % PIPE_STAGES;
        double2* buf = pipe_smem + ((size_t)b << TILE_BITS);
        const uint64_t tile_base = sweep_tile_base(sp, blockIdx.x + k * gridDim.x);
        mbar_wait(&full_bar[b], (k / PIPE_STAGES) & 1u);
        for (int r = 0; r < nr; ++r) {
            const PRound& rd = sp.rounds[r];
            uint32_t tl; uint64_t g;
            round_index<R>(sp, rd, tile_base, tid, tl, g);
            const uint32_t tls = swz(tl);
            const bool lin_in = r == 0, lin_out = r == nr - 1;
            double2 a[1 << R];
            if (lin_in) round_load_lin<R>(a, buf, rd, tl);
            else round_load_smem<R>(a, buf, rd, tls, g);
            round_ops<R, true>(a, sp, rd, g, lane, ex);
            // a round that changes the layout in place must have finished all its reads before any write
            if (lin_in != lin_out) compute_bar();
            if (lin_out) round_store_lin<R>(a, buf, rd, tl);
            else { round_store_smem<R>(a, buf, rd, tls, g); compute_bar(); }
        }
        fence_proxy_async();             // generic-proxy writes -> visible to the bulk copy engine
        compute_bar();
        if (tid == 0) mbar_arrive(&done_bar[b]);
    }
}

// ---------------------------------------------------------------------------------------------
// K7: global-qubit exchange over NVLink peer memory.  The local slice is `world` chunks; chunk[p] of this
// rank and chunk[rank] of rank p trade places, for every p at once, IN PLACE and without staging: every
// amplitude pair is owned by exactly one thread in the whole system (the lower rank of a pair takes the
// first half of the chunk, the higher rank the second half), which loads both sides (the remote one over
// NVLink), and stores them swapped (the remote store again over NVLink).  Each GPU thus reads and writes
// half a chunk per peer: both NVLink directions of every rank carry the same load.  Replaces NCCL
// send/recv through a staging buffer plus a copy-back.  Ranks synchronise before and after (host side).
// ---------------------------------------------------------------------------------------------
constexpr int PEER_MAX_WORLD = 16;
struct PeerTable { double2* p[PEER_MAX_WORLD]; };

__global__ void __launch_bounds__(512)
sv_peer_swap_kernel(double2* __restrict__ local, const PeerTable peers, const int world, const int rank,
                    const uint64_t chunk) {
    const uint64_t half = chunk >> 1;
    const uint64_t total = (uint64_t)(world - 1) * half;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    constexpr int U = 4;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * U) {
        double2 a[U], b[U];
        double2* mine[U];
        double2* theirs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t idx = base + (uint64_t)u * stride;
            mine[u] = nullptr;
            if (idx < total) {
                const int s = (int)(idx / half) + 1;
                const int peer = rank ^ s;
                const uint64_t i = idx % half + (rank < peer ? 0 : half);
                mine[u] = local + (uint64_t)peer * chunk + i;
                theirs[u] = peers.p[peer] + (uint64_t)rank * chunk + i;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (mine[u]) { a[u] = *mine[u]; b[u] = __ldcg(theirs[u]); }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (mine[u]) { *mine[u] = b[u]; __stcg(theirs[u], a[u]); }
    }
}
// The same exchange without first moving the outgoing qubits to the top of the local index: "chunk p" is the set of
// amplitudes whose bits at the g victim positions spell p (a strided set), traded with the peer's set that spells this
// rank.  Saves the SWAP-localisation sweep (one full read+write pass over the slice) that used to precede every
// exchange.  `sub` = amplitudes per chunk = 2^(nl - g); pos_sorted: victim positions ascending (for the zero
// insertion); pattern[r]: the bits of rank r deposited at the victim positions (rank bit j -> the j-th victim).
struct PeerStride { int32_t g; int32_t pos_sorted[4]; uint64_t pattern[PEER_MAX_WORLD]; };

__global__ void __launch_bounds__(512)
sv_peer_swap_strided_kernel(double2* __restrict__ local, const PeerTable peers, const int world, const int rank,
                            const uint64_t sub, const PeerStride ps) {
    const uint64_t half = sub >> 1;
    const uint64_t total = (uint64_t)(world - 1) * half;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    constexpr int U = 4;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * U) {
        double2 a[U], b[U];
        double2* mine[U];
        double2* theirs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t idx = base + (uint64_t)u * stride;
            mine[u] = nullptr;
            if (idx < total) {
                const int s = (int)(idx / half) + 1;
                const int peer = rank ^ s;
                uint64_t i = idx % half + (rank < peer ? 0 : half);
                for (int k = 0; k < ps.g; ++k) i = ins0_64(i, ps.pos_sorted[k]);
                mine[u] = local + (i | ps.pattern[peer]);
                theirs[u] = peers.p[peer] + (i | ps.pattern[rank]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (mine[u]) { a[u] = *mine[u]; b[u] = __ldcg(theirs[u]); }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (mine[u]) { *mine[u] = b[u]; __stcg(theirs[u], a[u]); }
    }
}
#endif  // __CUDACC__

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

// out[q] = <Z_q> (q < n), out[n] = norm^2.  One CTA per output: 256 threads add the block partials in a
// strided, fixed order and meet in a fixed shuffle / shared-memory tree (deterministic for a given grid).
// (One thread per output walking all ~16k block partials took 0.5 ms -- half of the read pass itself.)
__global__ void __launch_bounds__(RED_THREADS)
sv_expz_final_kernel(const double* __restrict__ partial, const int nblocks, const int n, const int tb,
                     double* __restrict__ out) {
    __shared__ double sh[2][RED_THREADS / 32];
    const int q = blockIdx.x;
    if (q > n) return;
    double total = 0, sq = 0;
    for (int b = threadIdx.x; b < nblocks; b += RED_THREADS) {
        const double* pb = partial + (size_t)b * EXPZ_WIDTH;
        total += pb[0];
        if (q < 8) sq += pb[1 + q];
        else if (q < tb) sq += ((b >> (q - 8)) & 1) ? pb[0] : 0.0;
        else if (q < n) sq += pb[9 + (q - tb)];
    }
    total = warp_sum(total);
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = total; sh[1][threadIdx.x >> 5] = sq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0, s2 = 0;
        for (int w = 0; w < RED_THREADS / 32; ++w) { t += sh[0][w]; s2 += sh[1][w]; }
        out[q] = q < n ? t - 2.0 * s2 : t;
    }
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

// Four qubits q0 < q1 < q2 < q3 per read pass: SIX pair-RDMs from 16 amplitudes per thread in registers.
// The pass is then FP64-FMA bound (48 FMA per amplitude vs 16 B of HBM traffic), i.e. the cost per pair
// drops from 1/3 of an HBM pass to 1/6 of a (slightly longer) FMA-bound pass.
// partial layout per block: pairs (0,1) (0,2) (0,3) (1,2) (1,3) (2,3) x 16 doubles.
__global__ void __launch_bounds__(RED_THREADS)
sv_rdm4_kernel(const double2* __restrict__ psi, const int n, const int q0, const int q1, const int q2, const int q3,
               double* __restrict__ partial) {
    double acc[RDM4_WIDTH];
#pragma unroll
    for (int k = 0; k < RDM4_WIDTH; ++k) acc[k] = 0.0;
    const uint64_t groups = 1ull << (n - 4);
    const uint64_t b0 = 1ull << q0, b1 = 1ull << q1, b2 = 1ull << q2, b3 = 1ull << q3;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < groups; k += stride) {
        const uint64_t b = ins0_64(ins0_64(ins0_64(ins0_64(k, q0), q1), q2), q3);
        double2 a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            a[j] = psi[b + ((j & 1) ? b0 : 0) + ((j & 2) ? b1 : 0) + ((j & 4) ? b2 : 0) + ((j & 8) ? b3 : 0)];
        int pair = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jq = i + 1; jq < 4; ++jq) {
                const int mi = 1 << i, mj = 1 << jq;
#pragma unroll
                for (int rest = 0; rest < 16; ++rest) {
                    if (rest & (mi | mj)) continue;
                    rdm_acc(acc + 16 * pair, a[rest], a[rest | mi], a[rest | mj], a[rest | mi | mj]);
                }
                ++pair;
            }
    }
    block_sum_store<RDM4_WIDTH>(acc, partial + (size_t)blockIdx.x * RDM4_WIDTH);
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

// K6d.  Projection of a state onto |0> of every qubit OUTSIDE qmap: phi[c] = psi[deposit(c)], c < 2^K.
// When every remaining gate of the window acts only on the K qubits of qmap, <0..0| W |psi> equals
// <0_K| W_K |phi>: the rest of the optimisation runs on a 2^K-amplitude state (sv_engine "projected tail").
// Sharded registers (dist_sv): entries of qmap >= n_local name RANK bits.  This rank contributes the
// amplitudes whose rank bits equal its own (`rank_want` on the covered bits; `contributes` = its other
// rank bits are all zero) and zeros elsewhere; the ranks' results are summed afterwards.
__global__ void __launch_bounds__(RED_THREADS)
sv_gather_kernel(const double2* __restrict__ psi, const QMap qm, const int K, double2* __restrict__ phi,
                 const int n_local, const uint32_t rank_want, const int contributes) {
    const uint64_t dim = 1ull << K;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < dim; c += stride) {
        uint64_t x = 0;
        uint32_t rb = 0;
        for (int b = 0; b < K; ++b) {
            const uint64_t bit = (c >> b) & 1ull;
            const int q = qm.q[b];
            if (q < n_local) x |= bit << q;
            else rb |= (uint32_t)bit << (q - n_local);
        }
        phi[c] = (contributes && rb == rank_want) ? psi[x] : make_double2(0.0, 0.0);
    }
}

// Inverse of the projection: psi[x] = phi[extract(x)] where every qubit outside qmap is 0, and 0 elsewhere.
// (<L| = suffix^+ <0| whose tail part was built on the K-qubit engine is embedded into the register.)
__global__ void sv_zero_kernel(double2* __restrict__ psi, const uint64_t dim) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride)
        psi[i] = make_double2(0.0, 0.0);
}

// (after sv_zero_kernel) psi[deposit(c)] = phi[c]: 2^K scattered 16-byte stores
__global__ void __launch_bounds__(RED_THREADS)
sv_scatter_kernel(double2* __restrict__ psi, const QMap qm, const int K, const double2* __restrict__ phi) {
    const uint64_t dim = 1ull << K;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < dim; c += stride) {
        uint64_t x = 0;
        for (int b = 0; b < K; ++b) x |= ((c >> b) & 1ull) << qm.q[b];
        psi[x] = phi[c];
    }
}

}  // namespace b200
