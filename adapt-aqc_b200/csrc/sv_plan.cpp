// sv_plan.cpp -- gate canonicalisation, 1-qubit fusion and sweep/round scheduling (host).
#include "sv_plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace b200 {

namespace {

const cplx I_(0.0, 1.0);

inline cplx expi(double x) { return cplx(std::cos(x), std::sin(x)); }

void set_mat1(COp& o, int q, cplx a, cplx b, cplx c, cplx d) {
    o.kind = K_MAT1;
    o.t0 = q;
    o.m[0] = a; o.m[1] = b; o.m[2] = c; o.m[3] = d;
}

void set_diag1(COp& o, int q, cplx p0, cplx p1) {
    o.kind = K_DIAG;
    o.d0 = q;
    o.m[0] = p0; o.m[1] = p1; o.m[2] = p0; o.m[3] = p1;
}

void set_x(COp& o, int control, int target) {
    o.kind = K_X;
    o.t0 = target;
    o.c = control;
}

inline bool single_uncontrolled(const COp& o, int& q) {
    if (o.kind == K_MAT1 && o.c < 0) { q = o.t0; return true; }
    if (o.kind == K_X && o.c < 0) { q = o.t0; return true; }
    if (o.kind == K_DIAG && o.d1 < 0) { q = o.d0; return true; }
    return false;
}

void as_matrix(const COp& o, cplx m[4]) {
    if (o.kind == K_MAT1) { for (int k = 0; k < 4; ++k) m[k] = o.m[k]; }
    else if (o.kind == K_X) { m[0] = 0; m[1] = 1; m[2] = 1; m[3] = 0; }
    else { m[0] = o.m[0]; m[1] = 0; m[2] = 0; m[3] = o.m[1]; }
}

void invert(COp& o) {
    if (o.kind == K_MAT1) {
        cplx a = std::conj(o.m[0]), b = std::conj(o.m[2]), c = std::conj(o.m[1]), d = std::conj(o.m[3]);
        o.m[0] = a; o.m[1] = b; o.m[2] = c; o.m[3] = d;
    } else if (o.kind == K_DIAG) {
        for (int k = 0; k < 4; ++k) o.m[k] = std::conj(o.m[k]);
    } else if (o.kind == K_MAT2) {
        cplx t[16];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) t[4 * r + c] = std::conj(o.m[4 * c + r]);
        for (int k = 0; k < 16; ++k) o.m[k] = t[k];
    }
}

inline uint64_t qbit(int q) { return q >= 0 ? (1ull << q) : 0ull; }
inline uint64_t mix_mask(const COp& o) { return qbit(o.t0) | qbit(o.t1); }
inline uint64_t diag_mask(const COp& o) { return qbit(o.c) | qbit(o.d0) | qbit(o.d1); }

}  // namespace

std::string canonicalize(int nq, const b200_gate* gates, int n_gates, const double* mats,
                         int n_mats, bool inverse, std::vector<COp>& out) {
    out.clear();
    out.reserve(n_gates + 8);
    const double r2 = 0.70710678118654752440;
    for (int k = 0; k < n_gates; ++k) {
        const b200_gate& g = gates[k];
        const int q0 = g.q0, q1 = g.q1;
        if (q0 < 0 || q0 >= nq) return "gate " + std::to_string(k) + ": qubit q0 out of range";
        const double th = g.p[0], ph = g.p[1], lm = g.p[2];
        COp o;
        bool two = false;
        switch (g.op) {
        case B200_OP_ID: continue;
        case B200_OP_X: set_x(o, -1, q0); break;
        case B200_OP_Y: set_mat1(o, q0, 0, -I_, I_, 0); break;
        case B200_OP_Z: set_diag1(o, q0, 1, -1); break;
        case B200_OP_H: set_mat1(o, q0, r2, r2, r2, -r2); break;
        case B200_OP_S: set_diag1(o, q0, 1, I_); break;
        case B200_OP_SDG: set_diag1(o, q0, 1, -I_); break;
        case B200_OP_T: set_diag1(o, q0, 1, expi(M_PI / 4)); break;
        case B200_OP_TDG: set_diag1(o, q0, 1, expi(-M_PI / 4)); break;
        case B200_OP_SX:
            set_mat1(o, q0, cplx(0.5, 0.5), cplx(0.5, -0.5), cplx(0.5, -0.5), cplx(0.5, 0.5)); break;
        case B200_OP_RX: {
            const double c = std::cos(th / 2), s = std::sin(th / 2);
            set_mat1(o, q0, c, cplx(0, -s), cplx(0, -s), c); break;
        }
        case B200_OP_RY: {
            const double c = std::cos(th / 2), s = std::sin(th / 2);
            set_mat1(o, q0, c, -s, s, c); break;
        }
        case B200_OP_RZ: set_diag1(o, q0, expi(-th / 2), expi(th / 2)); break;
        case B200_OP_U1: set_diag1(o, q0, 1, expi(th)); break;
        case B200_OP_U2: {
            const double phi = g.p[0], lam = g.p[1];
            set_mat1(o, q0, r2, -expi(lam) * r2, expi(phi) * r2, expi(phi + lam) * r2); break;
        }
        case B200_OP_U3: {
            const double c = std::cos(th / 2), s = std::sin(th / 2);
            set_mat1(o, q0, c, -expi(lm) * s, expi(ph) * s, expi(ph + lm) * c); break;
        }
        case B200_OP_MAT1: {
            if (!mats || g.aux < 0 || g.aux + 8 > n_mats)
                return "gate " + std::to_string(k) + ": mat1 offset outside mats[]";
            const double* d = mats + g.aux;
            set_mat1(o, q0, cplx(d[0], d[1]), cplx(d[2], d[3]), cplx(d[4], d[5]), cplx(d[6], d[7]));
            break;
        }
        case B200_OP_CX: two = true; set_x(o, q0, q1); break;
        case B200_OP_CZ:
            two = true; o.kind = K_DIAG; o.d0 = q0; o.d1 = q1;
            o.m[0] = 1; o.m[1] = 1; o.m[2] = 1; o.m[3] = -1; break;
        case B200_OP_SWAP: two = true; break;
        case B200_OP_MAT2: {
            two = true;
            if (!mats || g.aux < 0 || g.aux + 32 > n_mats)
                return "gate " + std::to_string(k) + ": mat2 offset outside mats[]";
            const double* d = mats + g.aux;
            o.kind = K_MAT2; o.t0 = q0; o.t1 = q1;
            bool diag = true;
            for (int j = 0; j < 16; ++j) {
                o.m[j] = cplx(d[2 * j], d[2 * j + 1]);
                if ((j / 4) != (j % 4) && o.m[j] != cplx(0, 0)) diag = false;
            }
            if (diag) {
                COp dg; dg.kind = K_DIAG; dg.d0 = q0; dg.d1 = q1;
                for (int j = 0; j < 4; ++j) dg.m[j] = o.m[5 * j];
                o = dg;
            }
            break;
        }
        default: return "gate " + std::to_string(k) + ": unknown opcode " + std::to_string(g.op);
        }
        if (two) {
            if (q1 < 0 || q1 >= nq || q1 == q0)
                return "gate " + std::to_string(k) + ": qubit q1 out of range";
        }
        if (g.op == B200_OP_SWAP) {
            COp a, b, c;
            set_x(a, q0, q1); set_x(b, q1, q0); set_x(c, q0, q1);
            out.push_back(a); out.push_back(b); out.push_back(c);
            continue;
        }
        // a dense 1-qubit matrix that happens to be exactly diagonal is cheaper as a phase op
        if (o.kind == K_MAT1 && o.m[1] == cplx(0, 0) && o.m[2] == cplx(0, 0)) {
            COp dg; set_diag1(dg, o.t0, o.m[0], o.m[3]); o = dg;
        }
        out.push_back(o);
    }
    if (inverse) {
        std::reverse(out.begin(), out.end());
        for (auto& o : out) invert(o);
    }
    return "";
}

// How op `o` acts on qubit q: 0 = not at all, 1 = diagonally (phase / control), 2 = mixing.
static int action_on(const COp& o, int q) {
    if (o.t0 == q || o.t1 == q) return 2;
    if (o.c == q || o.d0 == q || o.d1 == q) return 1;
    return 0;
}

void fuse_single_qubit_runs(std::vector<COp>& ops) {
    std::vector<COp> out;
    out.reserve(ops.size());
    // per qubit: index (in `out`) of the latest 1-qubit uncontrolled op that can still absorb
    //   any_on[q]  : a following 1-qubit op of ANY kind (nothing touched q since)
    //   diag_on[q] : a following DIAGONAL 1-qubit op (only ops acting diagonally on q came since: a phase
    //                on q commutes with them, e.g. rz on the control of a cx, or across a cz)
    int any_on[64], diag_on[64];
    for (int q = 0; q < 64; ++q) any_on[q] = diag_on[q] = -1;
    for (const COp& op : ops) {
        int q = -1;
        if (single_uncontrolled(op, q)) {
            const int j = op.kind == K_DIAG ? diag_on[q] : any_on[q];
            if (j >= 0) {
                COp& prev = out[j];
                if (prev.kind == K_DIAG && op.kind == K_DIAG) {
                    const cplx p0 = op.m[0] * prev.m[0], p1 = op.m[1] * prev.m[1];
                    set_diag1(prev, q, p0, p1);
                } else {
                    cplx a[4], b[4];
                    as_matrix(op, a);
                    as_matrix(prev, b);
                    COp f;
                    set_mat1(f, q, a[0] * b[0] + a[1] * b[2], a[0] * b[1] + a[1] * b[3],
                             a[2] * b[0] + a[3] * b[2], a[2] * b[1] + a[3] * b[3]);
                    prev = f;
                }
                continue;
            }
            const int idx = (int)out.size();
            out.push_back(op);
            any_on[q] = diag_on[q] = idx;
            continue;
        }
        out.push_back(op);
        const int qs[5] = {op.t0, op.t1, op.c, op.d0, op.d1};
        for (int x : qs) {
            if (x < 0) continue;
            any_on[x] = -1;
            if (action_on(op, x) == 2) diag_on[x] = -1;
        }
    }
    ops.swap(out);
}

namespace {

// Greedy commuting-aware selection: walk `pending` in order and pick every op that (a) commutes
// with all ops skipped so far and (b) `accept` admits.  Picked ops are returned in order.
template <class Accept>
std::vector<int> greedy_pick(const std::vector<COp>& ops, const std::vector<int>& pending,
                             std::vector<char>& taken, Accept accept) {
    std::vector<int> picked;
    uint64_t blocked_any = 0, blocked_mix = 0;
    for (int idx : pending) {
        if (taken[idx]) continue;
        const COp& o = ops[idx];
        const uint64_t mm = mix_mask(o), dm = diag_mask(o);
        const bool conflict = (mm & blocked_any) || (dm & blocked_mix);
        if (!conflict && accept(o)) {
            picked.push_back(idx);
            taken[idx] = 1;
        } else {
            blocked_any |= mm | dm;
            blocked_mix |= mm;
        }
    }
    return picked;
}

void fill_small_op(const COp& o, Plan& plan, DevOp& d) {
    std::memset(&d, 0, sizeof d);
    d.kind = o.kind;
    d.treg0 = o.t0; d.treg1 = o.t1;
    d.cq = o.c;
    d.dq0 = o.d0; d.dq1 = o.d1;
    d.mat2 = -1;
    if (o.kind == K_MAT1) {
        for (int k = 0; k < 4; ++k) { d.m[2 * k] = o.m[k].real(); d.m[2 * k + 1] = o.m[k].imag(); }
    } else if (o.kind == K_DIAG) {
        for (int k = 0; k < 4; ++k) { d.m[2 * k] = o.m[k].real(); d.m[2 * k + 1] = o.m[k].imag(); }
    } else if (o.kind == K_MAT2) {
        d.mat2 = (int32_t)plan.mat2.size();
        for (int k = 0; k < 16; ++k) { plan.mat2.push_back(o.m[k].real()); plan.mat2.push_back(o.m[k].imag()); }
    }
}

}  // namespace

void build_plan(int nq, const std::vector<COp>& ops, Plan& plan) {
    plan = Plan();
    plan.num_qubits = nq;
    if (nq <= SMALL_MAX_QUBITS) {
        plan.small = true;
        plan.ops.resize(ops.size());
        for (size_t k = 0; k < ops.size(); ++k) fill_small_op(ops[k], plan, plan.ops[k]);
        return;
    }

    const int nops = (int)ops.size();
    std::vector<char> taken(nops, 0);
    std::vector<int> all(nops);
    for (int k = 0; k < nops; ++k) all[k] = k;
    int n_taken = 0;
    const uint64_t low_mask = (1ull << LANE_BITS) - 1;

    while (n_taken < nops) {
        // ---- choose the ops of this sweep and its set H of high mixing qubits ----
        uint64_t H = 0;
        std::vector<int> sw_ops = greedy_pick(ops, all, taken, [&](const COp& o) {
            const uint64_t need = mix_mask(o) & ~low_mask & ~H;
            if (__builtin_popcountll(H) + __builtin_popcountll(need) > MAX_HIGH) return false;
            H |= need;
            return true;
        });
        n_taken += (int)sw_ops.size();

        // ---- tile qubit list: lanes, H, padding with the lowest unused qubits ----
        uint64_t tile = low_mask | H;
        for (int q = LANE_BITS; q < nq && __builtin_popcountll(tile) < TILE_BITS; ++q) tile |= 1ull << q;
        DevSweep sw;
        std::memset(&sw, 0, sizeof sw);
        sw.t = TILE_BITS;
        for (int q = 0, i = 0; q < nq; ++q) {
            if (tile >> q & 1) { sw.tileq[i] = q; ++i; }
        }
        int c = 0;
        while (c < TILE_BITS && sw.tileq[c] == c) ++c;
        sw.c = c;

        // ---- split the sweep's ops into rounds of <= REG_BITS register (mixing) qubits ----
        struct RoundTmp { uint64_t regs; std::vector<int> ops; };
        std::vector<RoundTmp> rounds;
        std::vector<char> rtaken(nops, 1);
        for (int idx : sw_ops) rtaken[idx] = 0;
        int left = (int)sw_ops.size();
        bool first = true;
        while (left > 0) {
            uint64_t regs = 0;
            const bool no_low = first;  // the first round loads from HBM: register qubits >= 5
            std::vector<int> r_ops = greedy_pick(ops, sw_ops, rtaken, [&](const COp& o) {
                const uint64_t mm = mix_mask(o);
                if (no_low && (mm & low_mask)) return false;
                const uint64_t need = mm & ~regs;
                if (__builtin_popcountll(regs) + __builtin_popcountll(need) > REG_BITS) return false;
                regs |= need;
                return true;
            });
            left -= (int)r_ops.size();
            if (!(first && r_ops.empty())) rounds.push_back({regs, r_ops});
            else rounds.push_back({0, {}});  // pure load round
            first = false;
        }
        if (rounds.empty()) rounds.push_back({0, {}});
        if (rounds.back().regs & low_mask) rounds.push_back({0, {}});  // pure store round

        sw.round_begin = (int32_t)plan.rounds.size();
        for (auto& rt : rounds) {
            // pad the register set with the highest free tile positions >= LANE_BITS
            uint64_t regs = rt.regs;
            for (int i = TILE_BITS - 1; i >= LANE_BITS && __builtin_popcountll(regs) < REG_BITS; --i)
                if (!(regs >> sw.tileq[i] & 1)) regs |= 1ull << sw.tileq[i];
            DevRound dr;
            int reg_of[64];
            for (int q = 0; q < 64; ++q) reg_of[q] = -1;
            int k = 0;
            for (int i = 0; i < TILE_BITS; ++i)
                if (regs >> sw.tileq[i] & 1) { dr.regpos[k] = i; reg_of[sw.tileq[i]] = k; ++k; }
            dr.op_begin = (int32_t)plan.ops.size();
            for (int idx : rt.ops) {
                const COp& o = ops[idx];
                DevOp d;
                std::memset(&d, 0, sizeof d);
                d.kind = o.kind;
                d.treg0 = d.treg1 = -1;
                d.cq = d.dq0 = d.dq1 = -1;
                d.mat2 = -1;
                if (o.t0 >= 0) d.treg0 = reg_of[o.t0];
                if (o.t1 >= 0) d.treg1 = reg_of[o.t1];
                if (o.c >= 0) { if (reg_of[o.c] >= 0) d.cmask = 1 << reg_of[o.c]; else d.cq = o.c; }
                if (o.d0 >= 0) { if (reg_of[o.d0] >= 0) d.dmask0 = 1 << reg_of[o.d0]; else d.dq0 = o.d0; }
                if (o.d1 >= 0) { if (reg_of[o.d1] >= 0) d.dmask1 = 1 << reg_of[o.d1]; else d.dq1 = o.d1; }
                if (o.kind == K_MAT1 || o.kind == K_DIAG) {
                    for (int j = 0; j < 4; ++j) { d.m[2 * j] = o.m[j].real(); d.m[2 * j + 1] = o.m[j].imag(); }
                } else if (o.kind == K_MAT2) {
                    cplx mm[16];
                    if (d.treg0 < d.treg1) {
                        for (int j = 0; j < 16; ++j) mm[j] = o.m[j];
                    } else {  // kernel wants treg0 < treg1: swap the two index bits
                        auto sw2 = [](int i) { return ((i & 1) << 1) | (i >> 1); };
                        for (int r = 0; r < 4; ++r)
                            for (int cc = 0; cc < 4; ++cc) mm[4 * sw2(r) + sw2(cc)] = o.m[4 * r + cc];
                        std::swap(d.treg0, d.treg1);
                    }
                    d.mat2 = (int32_t)plan.mat2.size();
                    for (int j = 0; j < 16; ++j) { plan.mat2.push_back(mm[j].real()); plan.mat2.push_back(mm[j].imag()); }
                }
                plan.ops.push_back(d);
            }
            dr.op_end = (int32_t)plan.ops.size();
            plan.rounds.push_back(dr);
        }
        sw.round_end = (int32_t)plan.rounds.size();
        plan.sweeps.push_back(sw);
    }
}

}  // namespace b200
