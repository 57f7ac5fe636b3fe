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

namespace b200 {

using cplx = std::complex<double>;

enum OpKind : int32_t { K_MAT1 = 0, K_X = 1, K_DIAG = 2, K_MAT2 = 3 };

constexpr int TILE_BITS = 12;   // tiled path: 2^12 amplitudes (64 KB) per tile
constexpr int REG_BITS = 4;     // 16 amplitudes per thread
constexpr int LANE_BITS = 5;    // lowest 5 qubits always in the tile: 512 B contiguous per warp
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

// Device op, 112 bytes.  In the tiled kernel the *mask fields are REGISTER masks (1 << register
// index); in the small kernel tmask0/tmask1/cq/dq0/dq1 hold global qubit numbers.
struct alignas(16) DevOp {
    int32_t kind;
    int32_t treg0, treg1;    // register index of mixing targets (tiled) / qubit (small)
    int32_t cmask;           // register mask of the control, 0 if none or thread-uniform
    int32_t cq;              // qubit of a thread-uniform control, -1 if none
    int32_t dmask0, dmask1;  // register masks of diagonal qubits, 0 if thread-uniform / absent
    int32_t dq0, dq1;        // qubits of thread-uniform diagonal bits, -1 if absent / register
    int32_t mat2;            // offset (doubles) into the mat2 table
    int32_t pad0, pad1;
    double m[8];             // MAT1: 2x2 complex row-major; DIAG: 4 phases
};
static_assert(sizeof(DevOp) == 112, "DevOp layout");

struct DevRound {
    int32_t op_begin, op_end;
    int32_t regpos[REG_BITS];  // tile-local bit positions of the register qubits, ascending
};

struct DevSweep {
    int32_t round_begin, round_end;
    int32_t t, c;               // tile bits; number of leading contiguous low qubits
    int32_t tileq[TILE_BITS];   // global qubit of each tile-local bit, ascending
};

struct Plan {
    int num_qubits = 0;
    bool small = false;            // single-CTA shared-memory path
    std::vector<DevSweep> sweeps;  // tiled path
    std::vector<DevRound> rounds;
    std::vector<DevOp> ops;        // tiled: grouped by round; small: flat program
    std::vector<double> mat2;      // 32 doubles per dense 2-qubit op
    int n_gates_in = 0;
};

// Returns empty string on success, else an error message.
std::string canonicalize(int num_qubits, const b200_gate* gates, int n_gates, const double* mats,
                         int n_mats, bool inverse, std::vector<COp>& out);
void fuse_single_qubit_runs(std::vector<COp>& ops);
void build_plan(int num_qubits, const std::vector<COp>& ops, Plan& plan);

}  // namespace b200
