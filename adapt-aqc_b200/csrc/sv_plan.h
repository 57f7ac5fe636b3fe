// sv_plan.h -- host-side planner for the statevector gate stream.
//
// A gate stream (b200_gate records, include/b200aqc.h) is canonicalised into four op kinds,
// 1-qubit runs are fused, and the result is partitioned into SWEEPS (one read+write pass over
// the 2^n amplitudes each) made of ROUNDS (register-resident gate groups inside one tile):
//
//   tile   : 2^12 amplitudes = the 5 lowest qubits (coalescing) + up to 7 arbitrary "mixing"
//            qubits (+ padding).  One CTA owns one tile at a time.
//   round  : every thread keeps 2^R (R=4) amplitudes in registers -- the R "register qubits" of
//            the round -- and applies all ops whose mixing targets are register qubits.
//            Between rounds the tile is exchanged through shared memory; the first round loads
//            straight from HBM and the last stores straight to HBM.
//   roles  : only MIXING targets (non-diagonal action) must be register qubits.  Control
//            qubits and qubits a gate acts on diagonally (rz, cz, u1, ...) can be anywhere in
//            the 2^n index: they only predicate / phase the thread's amplitudes.
//
// This replaces what qiskit-aer's statevector simulator + fusion pass do for the reference at
// adaptaqc/backends/aer_sv_backend.py:42-47.
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

#include "b200aqc.h"

#ifdef __CUDACC__
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

using cplx = std::complex<double>;

// 16-byte-slot swizzle of the 12-bit tile index: the slot-in-row bits (low 3) are XOR-folded with every
// higher 3-bit group, so a quarter-warp whose lanes vary ANY three consecutive tile bits (the three
// lowest non-register positions of a round) lands on the 8 distinct 16 B slots of a 128 B row.
// XOR-linear: swz(a ^ b) == swz(a) ^ swz(b).
B200_HD uint32_t swz(uint32_t i) { return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7u); }

enum OpKind : int32_t { K_MAT1 = 0, K_X = 1, K_DIAG = 2, K_MAT2 = 3 };

constexpr int TILE_BITS = 12;   // tiled path: 2^12 amplitudes (64 KB) per tile
constexpr int REG_BITS = 4;     // 16 amplitudes per thread
constexpr int LANE_BITS = 5;    // lowest 5 qubits always in the tile: 512 B contiguous per tile row
constexpr int COAL_BITS = 3;    // HBM rounds keep the 3 lowest qubits on the lanes: every warp access moves whole
                                // 128 B lines; qubits 3,4 may be register qubits there (4 lines per warp access)
constexpr int MAX_HIGH = TILE_BITS - LANE_BITS;
constexpr int SMALL_MAX_QUBITS = 11;  // n <= 11: whole state in one CTA's shared memory

// Canonical op (host).
struct COp {
    int kind = K_MAT1;
    int t0 = -1, t1 = -1;  // mixing targets (global qubit)
    int c = -1;            // control qubit
    int d0 = -1, d1 = -1;  // qubits acted on diagonally (K_DIAG)
    cplx m[16];            // MAT1: 2x2; DIAG: ph[b0 + 2*b1]; MAT2: 4x4 (index = bit(t0) + 2 bit(t1))
};

// Device op of the SMALL path (n <= SMALL_MAX_QUBITS, program in global memory), 112 bytes.
// treg0/treg1/cq/dq0/dq1 hold global qubit numbers.
struct alignas(16) DevOp {
    int32_t kind;
    int32_t treg0, treg1;    // mixing targets
    int32_t cmask;           // unused in the small path
    int32_t cq;              // control qubit, -1 if none
    int32_t dmask0, dmask1;  // unused in the small path
    int32_t dq0, dq1;        // diagonal qubits, -1 if absent
    int32_t mat2;            // offset (doubles) into the mat2 table
    int32_t pad0, pad1;
    double m[8];             // MAT1: 2x2 complex row-major; DIAG: 4 phases
};
static_assert(sizeof(DevOp) == 112, "DevOp layout");

// ---- tiled path: the whole program of one sweep travels as a KERNEL PARAMETER (constant bank) ----
// Every field is warp-uniform, so the kernel decodes ops with uniform loads / uniform branches and the
// register indices select fully unrolled, statically indexed bodies.
//
// Where a qubit of an op lives in a round:
//   register qubit  : one of the round's REG_BITS positions -> index 0..3 into the thread's 2^R amplitudes
//   lane qubit      : qubits 0..COAL_BITS-1 in a round that loads from / stores to HBM (they are then thread
//                     bits 0..2 = warp-lane bits): a MIXING target there is served by warp shuffles
//   thread-level    : anything else; its value is a bit of the thread's base index g
enum PKind : int32_t {
    P_PEND = 0,      // diagonal, no register qubit: one phase per thread, folded into the round's `pend`
    P_DIAG1 = 1,     // diagonal, one register qubit r0 (+ optional thread-level qubit dq1)
    P_DIAG2 = 2,     // diagonal, two register qubits r0 < r1
    P_DIAGRAW = 3,   // diagonal with a zero entry (non-unitary input): generic slow path
    P_XREG = 4,      // X on register r0, optional thread-level control cq
    P_CXREG = 5,     // X on register r0 controlled by register r1
    P_MAT1 = 6,      // dense 2x2 on register r0
    P_MAT2 = 7,      // dense 4x4 on registers r0 < r1 (matrix index `mat2`)
    P_XLANE = 8,     // X on lane bit r0, control: none / thread-level cq / register r1
    P_MAT1LANE = 9,  // dense 2x2 on lane bit r0
};

// Dispatch code of a tiled op: kind and register selectors folded into ONE warp-uniform switch index.
//   0 PEND | 1..4 DIAG1(r0) | 5..10 DIAG2(r0<r1) | 11 DIAGRAW | 12..15 XREG(r0) | 16..27 CXREG(t=r0,c=r1)
//   28..31 MAT1(r0) | 32..37 MAT2(r0<r1) | 38 XLANE | 39 MAT1LANE
inline int pair_index(int r0, int r1) {   // r0 < r1 < 4 -> 0..5 in the order 01 02 03 12 13 23
    return r0 == 0 ? r1 - 1 : (r0 == 1 ? r1 + 1 : 5);
}
inline int cx_index(int t, int c) { return t * 3 + (c > t ? c - 1 : c); }   // t != c -> 0..11

struct alignas(16) POp {   // 96 bytes
    int32_t kind;
    int32_t r0, r1;        // register indices (lane index in r0 for the *LANE kinds); -1 if unused
    int32_t cq;            // thread-level control qubit, -1 if none
    int32_t dq0, dq1;      // thread-level diagonal qubits, -1 if absent
    int32_t mat2;          // P_MAT2: index into SweepProg::mat2
    int32_t flush;         // bit 0 (lane kinds): phases are pending (they do not commute with a lane exchange):
                           // apply first.  bits 8..15: dispatch code (see above)
    // P_PEND: ph[u0 + 2 u1] (4 complex).  P_DIAG1: base[u], ratio[u] (u = thread-level bit, 4 complex):
    // amplitudes with the register bit set are multiplied by ratio[u], base[u] goes into `pend`.
    // P_DIAG2: base, ratio01, ratio10, ratio11.  P_DIAGRAW: ph[4].  P_MAT1 / P_MAT1LANE: 2x2 row-major.
    double m[8];
};
static_assert(sizeof(POp) == 96, "POp layout");

constexpr int MAX_SWEEP_ROUNDS = 16;
constexpr int MAX_SWEEP_OPS = 96;
constexpr int MAX_SWEEP_MAT2 = 16;
constexpr int MAX_LANE_OPS = 2;   // shuffle-served ops per HBM round before a shared-memory round is cheaper
                                  // (64 SHFL per thread each vs 32 LDS/STS.128 + barrier for a round trip)

// An X / CX whose target is a register qubit only PERMUTES the thread's 2^R amplitudes.  When it commutes its way to the
// front (back) of its round it is not executed at all: it is folded into WHERE the round loads (stores) --
//   * control = another register qubit, or none with a fixed pattern: a GF(2)-linear relabelling of the register
//     index j, resolved on the host into the round's load / store offset tables;
//   * control = any other qubit (a bit of the thread's base index g), or an unconditional X: a per-thread XOR mask on
//     the address (shared-memory index and global index are both XOR-linear in the register index).
// ncu (profiles/prof_direct_r02x.txt): executed as in-place register exchanges these ops were ~29 % of the instructions
// of a thin-layer sweep (96 register moves per conditional X).
constexpr int MAX_FOLD = 8;
struct PFold {
    int32_t cq;        // control qubit read from the thread's base index g; -1: unconditional
    uint32_t smask;    // XOR mask on the swizzled shared-memory index
    uint64_t gmask;    // XOR mask on the global amplitude index
};
static_assert(sizeof(PFold) == 16, "PFold layout");

struct PRound {
    int32_t op_begin, op_end;
    int32_t regpos[REG_BITS];        // tile-local bit positions of the register qubits, ascending
    int32_t soff[1 << REG_BITS];     // LOAD: swizzled shared-memory offset register amplitude j comes from (swz is XOR-linear)
    int32_t has_pend, pad;           // some op of the round multiplies into the per-thread pending phase
    int32_t soff_st[1 << REG_BITS];  // STORE: swizzled shared-memory offset register amplitude j goes to
    uint64_t goff_ld[1 << REG_BITS]; // the same for rounds that load from / store to HBM: offsets in the global index
    uint64_t goff_st[1 << REG_BITS];
    int32_t n_lead, n_trail;         // folded X-type ops with a per-thread condition (see PFold)
    PFold lead[MAX_FOLD], trail[MAX_FOLD];
};

struct alignas(16) SweepProg {
    int32_t nrounds, nops, nmat2;
    int32_t c;                  // number of leading contiguous low qubits in tileq
    int32_t tileq[TILE_BITS];   // global qubit of each tile-local bit, ascending
    PRound rounds[MAX_SWEEP_ROUNDS];
    POp ops[MAX_SWEEP_OPS + 1];   // +1: the kernel prefetches one op ahead
    double mat2[MAX_SWEEP_MAT2][32];
};
static_assert(sizeof(SweepProg) <= 32000, "SweepProg must fit the 32 KB kernel-parameter space");

// Epilogue of a FUSED sweep + transfer pass (sv_sweep_inner2_kernel): after the last round of the sweep the tile sits in
// shared memory; every thread then owns 16 amplitudes of the tile -- thread bit b at tile position tpos[b], amplitude
// m at the 4 tile positions outside tpos -- chosen so that the values of the open pair (qa, qb) are THREAD bits: a thread
// sees one fixed column j of the 4x4 transfer matrix and keeps only 4 complex accumulators.
struct EpiProg {
    int32_t pa, pb;          // tile-local positions of the open pair (pa < pb)
    int32_t ja, jb;          // thread bits that sit at pa / pb
    int32_t tpos[TILE_BITS - REG_BITS];
    uint32_t moff_sw[1 << REG_BITS];   // amplitude m of a thread: swizzled tile offset ...
    uint64_t moff_g[1 << REG_BITS];    // ... and offset in the global index
};

struct Plan {
    int num_qubits = 0;
    bool small = false;              // single-CTA shared-memory path
    std::vector<SweepProg> sweeps;   // tiled path
    std::vector<DevOp> ops;          // small path: flat program
    std::vector<double> mat2;        // small path: 32 doubles per dense 2-qubit op
    int n_gates_in = 0;
    int n_rounds() const { int r = 0; for (const auto& s : sweeps) r += s.nrounds; return r; }
    int n_ops() const { int r = (int)ops.size(); for (const auto& s : sweeps) r += s.nops; return r; }
};

// Returns empty string on success, else an error message.
std::string canonicalize(int num_qubits, const b200_gate* gates, int n_gates, const double* mats,
                         int n_mats, bool inverse, std::vector<COp>& out);
void fuse_single_qubit_runs(std::vector<COp>& ops);
// Pushes 1-/2-qubit diagonals forward through the CX gates they meet and merges them: a thinly dressed
// CNOT layer (rz rz cx rz rz) becomes ONE cx + ONE two-qubit diagonal.
void fuse_diagonals(std::vector<COp>& ops);
// fold_perm: fold leading / trailing X-type ops of every round into its load / store addressing (the direct sweep
// kernel; the pipelined variant keeps executing them).
// pair_a / pair_b >= 0: plan for the fused sweep + transfer pass -- both qubits are tile qubits of EVERY sweep (so of the
// last one), there is at least one sweep, and the last round of the last sweep stores to shared memory (no lane-qubit
// folds on its store side).
void build_plan(int num_qubits, const std::vector<COp>& ops, Plan& plan, bool fold_perm = true, int pair_a = -1,
                int pair_b = -1, uint64_t pad_avoid = 0);
// Epilogue tables for the pair (qa, qb) on the tile of `sp`; false if one of them is not a tile qubit.
bool make_epilogue(const SweepProg& sp, int qa, int qb, EpiProg& ep);

}  // namespace b200
