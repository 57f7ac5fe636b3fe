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

// ---------------------------------------------------------------------------------------------
// diagonal fusion: push diagonals forward through CX, merge what meets
// ---------------------------------------------------------------------------------------------
namespace {

// phase of diagonal op `e` for bit values (ba on qubit a, bb on qubit b); a qubit `e` does not act on is ignored
inline cplx diag_phase(const COp& e, int a, int ba, int b, int bb) {
    const int b0 = e.d0 == a ? ba : (e.d0 == b ? bb : 0);
    const int b1 = e.d1 < 0 ? 0 : (e.d1 == a ? ba : (e.d1 == b ? bb : 0));
    return e.m[b0 + 2 * b1];
}

// multiply the 1-qubit diagonal (p0, p1) on qubit q into diagonal op `e` (which acts on q)
inline void mul_diag1_into(COp& e, int q, cplx p0, cplx p1) {
    for (int k = 0; k < 4; ++k) {
        const int bit = e.d0 == q ? (k & 1) : (k >> 1);
        e.m[k] *= bit ? p1 : p0;
    }
}

inline bool same_pair(const COp& e, int a, int b) {
    return e.d1 >= 0 && ((e.d0 == a && e.d1 == b) || (e.d0 == b && e.d1 == a));
}

}  // namespace

void fuse_diagonals(std::vector<COp>& ops) {
    std::vector<COp> out;
    std::vector<char> dead;
    out.reserve(ops.size() + 8);
    int last_mix[64], dcont[64];
    for (int q = 0; q < 64; ++q) last_mix[q] = dcont[q] = -1;
    auto live = [&](int e, int q) { return e >= 0 && !dead[e] && last_mix[q] < e; };
    auto push = [&](const COp& o) { out.push_back(o); dead.push_back(0); return (int)out.size() - 1; };

    for (const COp& op : ops) {
        if (op.kind == K_DIAG && op.d1 < 0) {
            const int q = op.d0, e = dcont[q];
            if (live(e, q)) { mul_diag1_into(out[e], q, op.m[0], op.m[1]); continue; }
            dcont[q] = push(op);
        } else if (op.kind == K_DIAG) {
            const int a = op.d0, b = op.d1;
            const int e = dcont[a];
            if (live(e, a) && same_pair(out[e], a, b) && live(e, b)) {
                for (int k = 0; k < 4; ++k) {
                    const int ba = out[e].d0 == a ? (k & 1) : (k >> 1), bb = out[e].d0 == a ? (k >> 1) : (k & 1);
                    out[e].m[k] *= diag_phase(op, a, ba, b, bb);
                }
                continue;
            }
            COp d = op;
            for (int q : {a, b}) {
                const int e1 = dcont[q];
                if (live(e1, q) && out[e1].d1 < 0) {
                    mul_diag1_into(d, q, out[e1].m[0], out[e1].m[1]);
                    dead[e1] = 1;
                }
            }
            dcont[a] = dcont[b] = push(d);
        } else if (op.kind == K_X && op.c >= 0) {
            const int c = op.c, t = op.t0;
            const int e = dcont[t];
            COp d;
            bool have = false;
            if (live(e, t) && out[e].d1 < 0) {                       // D_t . then CX  ==  CX . then parity phase
                d.kind = K_DIAG; d.d0 = t; d.d1 = c;
                for (int k = 0; k < 4; ++k) d.m[k] = out[e].m[(k & 1) ^ (k >> 1)];
                dead[e] = 1; have = true;
            } else if (live(e, t) && same_pair(out[e], c, t) && live(e, c)) {
                d.kind = K_DIAG; d.d0 = t; d.d1 = c;
                for (int k = 0; k < 4; ++k) {
                    const int bt = k & 1, bc = k >> 1;
                    d.m[k] = diag_phase(out[e], t, bt ^ bc, c, bc);
                }
                dead[e] = 1; have = true;
            }
            const int ix = push(op);
            last_mix[t] = ix;
            if (have) {
                const int ec = dcont[c];
                if (live(ec, c) && out[ec].d1 < 0) {                  // commutes with the control: absorb
                    mul_diag1_into(d, c, out[ec].m[0], out[ec].m[1]);
                    dead[ec] = 1;
                }
                dcont[t] = dcont[c] = push(d);
            }
        } else {
            const int ix = push(op);
            if (op.t0 >= 0) last_mix[op.t0] = ix;
            if (op.t1 >= 0) last_mix[op.t1] = ix;
        }
    }
    ops.clear();
    for (size_t k = 0; k < out.size(); ++k) {
        if (dead[k]) continue;
        const COp& o = out[k];
        if (o.kind == K_DIAG) {       // drop exact identities
            bool ident = true;
            for (int j = 0; j < 4; ++j) ident = ident && o.m[j] == cplx(1, 0);
            if (ident) continue;
        }
        ops.push_back(o);
    }
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

inline void put(double* m, int k, cplx z) { m[2 * k] = z.real(); m[2 * k + 1] = z.imag(); }

// Translate canonical op `o` for a round whose register qubits are given by reg_of[] (-1 = not a register).
// Mixing targets that are not registers are lane qubits (only scheduled in HBM rounds, qubit < COAL_BITS).
void fill_tiled_op(const COp& o, const int* reg_of, SweepProg& sp, POp& d) {
    std::memset(&d, 0, sizeof d);
    d.r0 = d.r1 = d.cq = d.dq0 = d.dq1 = d.mat2 = -1;
    if (o.kind == K_DIAG) {
        const int ra = reg_of[o.d0], rb = o.d1 >= 0 ? reg_of[o.d1] : -1;
        bool zero = false;
        for (int k = 0; k < 4; ++k) zero = zero || std::abs(o.m[k]) < 1e-300;
        if (ra < 0 && rb < 0) {
            d.kind = P_PEND; d.dq0 = o.d0; d.dq1 = o.d1;
            for (int k = 0; k < 4; ++k) put(d.m, k, o.m[k]);
        } else if (zero) {
            d.kind = P_DIAGRAW; d.r0 = ra; d.r1 = rb;
            d.dq0 = ra < 0 ? o.d0 : -1; d.dq1 = (rb < 0 && o.d1 >= 0) ? o.d1 : -1;
            for (int k = 0; k < 4; ++k) put(d.m, k, o.m[k]);
        } else if (ra >= 0 && rb >= 0) {
            d.kind = P_DIAG2;
            const bool sw = ra > rb;   // kernel wants r0 < r1: swap the two index bits
            d.r0 = sw ? rb : ra; d.r1 = sw ? ra : rb;
            const cplx p00 = o.m[0], p10 = sw ? o.m[2] : o.m[1], p01 = sw ? o.m[1] : o.m[2], p11 = o.m[3];
            put(d.m, 0, p00); put(d.m, 1, p10 / p00); put(d.m, 2, p01 / p00); put(d.m, 3, p11 / p00);
        } else {
            d.kind = P_DIAG1;
            if (ra >= 0) {            // register = d0, thread-level = d1 (may be absent)
                d.r0 = ra; d.dq1 = o.d1;
                put(d.m, 0, o.m[0]); put(d.m, 1, o.m[2]);
                put(d.m, 2, o.m[1] / o.m[0]); put(d.m, 3, o.m[3] / o.m[2]);
            } else {                  // register = d1, thread-level = d0
                d.r0 = rb; d.dq1 = o.d0;
                put(d.m, 0, o.m[0]); put(d.m, 1, o.m[1]);
                put(d.m, 2, o.m[2] / o.m[0]); put(d.m, 3, o.m[3] / o.m[1]);
            }
        }
    } else if (o.kind == K_X) {
        const int rt = reg_of[o.t0], rc = o.c >= 0 ? reg_of[o.c] : -1;
        if (rt >= 0 && rc >= 0) { d.kind = P_CXREG; d.r0 = rt; d.r1 = rc; }
        else if (rt >= 0) { d.kind = P_XREG; d.r0 = rt; d.cq = o.c; }
        else { d.kind = P_XLANE; d.r0 = o.t0; d.r1 = rc; d.cq = rc >= 0 ? -1 : o.c; }
    } else if (o.kind == K_MAT1) {
        const int rt = reg_of[o.t0];
        d.kind = rt >= 0 ? P_MAT1 : P_MAT1LANE;
        d.r0 = rt >= 0 ? rt : o.t0;
        for (int j = 0; j < 4; ++j) put(d.m, j, o.m[j]);
    } else {  // K_MAT2
        d.kind = P_MAT2;
        d.r0 = reg_of[o.t0]; d.r1 = reg_of[o.t1];
        cplx mm[16];
        if (d.r0 < d.r1) {
            for (int j = 0; j < 16; ++j) mm[j] = o.m[j];
        } else {  // kernel wants r0 < r1: swap the two index bits
            auto sw2 = [](int i) { return ((i & 1) << 1) | (i >> 1); };
            for (int r = 0; r < 4; ++r)
                for (int cc = 0; cc < 4; ++cc) mm[4 * sw2(r) + sw2(cc)] = o.m[4 * r + cc];
            std::swap(d.r0, d.r1);
        }
        d.mat2 = sp.nmat2++;
        for (int j = 0; j < 16; ++j) put(sp.mat2[d.mat2], j, mm[j]);
    }
}

}  // namespace

namespace {

// does the X-type op `x` (target t, optional control c) commute with `o`?
bool x_commutes_with(const COp& x, const COp& o) {
    const int t = x.t0, c = x.c;
    if (o.kind == K_DIAG) return o.d0 != t && o.d1 != t;            // a diagonal may sit on the control
    if (o.kind == K_X) return t != o.c && o.t0 != c;                // targets may coincide
    for (int q : {o.t0, o.t1, o.c, o.d0, o.d1})
        if (q >= 0 && (q == t || q == c)) return false;
    return true;
}

}  // namespace

bool make_epilogue(const SweepProg& sp, int qa, int qb, EpiProg& ep) {
    std::memset(&ep, 0, sizeof ep);
    int pa = -1, pb = -1;
    for (int i = 0; i < TILE_BITS; ++i) {
        if (sp.tileq[i] == qa) pa = i;
        if (sp.tileq[i] == qb) pb = i;
    }
    if (pa < 0 || pb < 0 || pa == pb) return false;
    if (pa > pb) std::swap(pa, pb);
    ep.pa = pa; ep.pb = pb;
    // amplitude bits: the REG_BITS highest tile positions that are not the pair; thread bits: the rest, ascending
    uint32_t mmask = 0;
    for (int i = TILE_BITS - 1, k = 0; i >= 0 && k < REG_BITS; --i)
        if (i != pa && i != pb) { mmask |= 1u << i; ++k; }
    int mpos[REG_BITS];
    for (int i = 0, k = 0, b = 0; i < TILE_BITS; ++i) {
        if (mmask >> i & 1) { mpos[k++] = i; continue; }
        if (i == pa) ep.ja = b;
        if (i == pb) ep.jb = b;
        ep.tpos[b++] = i;
    }
    for (uint32_t m = 0; m < (1u << REG_BITS); ++m) {
        uint32_t off = 0;
        uint64_t g = 0;
        for (int k = 0; k < REG_BITS; ++k)
            if (m >> k & 1) { off |= 1u << mpos[k]; g |= 1ull << sp.tileq[mpos[k]]; }
        ep.moff_sw[m] = swz(off);
        ep.moff_g[m] = g;
    }
    return true;
}

void build_plan(int nq, const std::vector<COp>& ops, Plan& plan, bool fold_perm, int pair_a, int pair_b, uint64_t pad_avoid) {
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
    const uint64_t coal_mask = (1ull << COAL_BITS) - 1;

    const bool fused = pair_a >= 0 && pair_b >= 0;
    uint64_t H_forced = 0;
    if (fused) H_forced = ((1ull << pair_a) | (1ull << pair_b)) & ~low_mask;
    bool need_one = (fused || (pair_a == -2 && pair_b == -2)) && nops == 0;    // (-2, -2): at least one (possibly empty) sweep

    while (n_taken < nops || need_one) {
        need_one = false;
        // ---- choose the ops of this sweep and its set H of high mixing qubits ----
        uint64_t H = H_forced;
        int n_sw = 0, n_m2 = 0;
        std::vector<int> sw_ops = greedy_pick(ops, all, taken, [&](const COp& o) {
            if (n_sw >= MAX_SWEEP_OPS || (o.kind == K_MAT2 && n_m2 >= MAX_SWEEP_MAT2)) return false;
            const uint64_t need = mix_mask(o) & ~low_mask & ~H;
            if (__builtin_popcountll(H) + __builtin_popcountll(need) > MAX_HIGH) return false;
            H |= need;
            ++n_sw;
            if (o.kind == K_MAT2) ++n_m2;
            return true;
        });

        // ---- tile qubit list: lanes, H, padding with the lowest unused qubits ----
        uint64_t tile = low_mask | H;
        // (pad_avoid: qubits better left OUT of the tile -- an embedded source is zero wherever one of them is set, and a
        // tile whose base index has such a bit is skipped whole, sv_sweep_inner2_kernel)
        for (int q = LANE_BITS; q < nq && __builtin_popcountll(tile) < TILE_BITS; ++q)
            if (!(pad_avoid >> q & 1)) tile |= 1ull << q;
        for (int q = LANE_BITS; q < nq && __builtin_popcountll(tile) < TILE_BITS; ++q) tile |= 1ull << q;
        plan.sweeps.emplace_back();
        SweepProg& sp = plan.sweeps.back();
        std::memset(&sp, 0, sizeof sp);
        for (int q = 0, i = 0; q < nq; ++q) {
            if (tile >> q & 1) { sp.tileq[i] = q; ++i; }
        }
        int c = 0;
        while (c < TILE_BITS && sp.tileq[c] == c) ++c;
        sp.c = c;

        // ---- split the sweep's ops into rounds ----
        // HBM round (first / last): register qubits among the tile qubits >= COAL_BITS, up to MAX_LANE_OPS
        // mixing ops on the lane qubits 0..COAL_BITS-1 served by shuffles, no dense 2-qubit op on a lane qubit.
        // Middle round (shared memory on both sides): any REG_BITS tile qubits as registers.
        struct RoundTmp { uint64_t regs; std::vector<int> ops; bool hbm; };
        std::vector<RoundTmp> rounds;
        std::vector<char> rtaken(nops, 1);
        for (int idx : sw_ops) rtaken[idx] = 0;
        int left = (int)sw_ops.size();
        bool first = true;
        while (left > 0 && (int)rounds.size() < MAX_SWEEP_ROUNDS - 1) {
            uint64_t regs = 0;
            int lanes = 0;
            std::vector<char> tk = rtaken;
            std::vector<int> r_ops = greedy_pick(ops, sw_ops, tk, [&](const COp& o) {
                const uint64_t mm = mix_mask(o), low = mm & coal_mask;
                if (low && o.kind == K_MAT2) return false;
                const int nl = __builtin_popcountll(low);
                if (lanes + nl > MAX_LANE_OPS) return false;
                const uint64_t need = mm & ~coal_mask & ~regs;
                if (__builtin_popcountll(regs) + __builtin_popcountll(need) > REG_BITS) return false;
                regs |= need; lanes += nl;
                return true;
            });
            bool hbm = true;
            if (!first && (int)r_ops.size() != left) {   // cannot finish here: a shared-memory round instead
                hbm = false;
                regs = 0;
                tk = rtaken;
                r_ops = greedy_pick(ops, sw_ops, tk, [&](const COp& o) {
                    const uint64_t need = mix_mask(o) & ~regs;
                    if (__builtin_popcountll(regs) + __builtin_popcountll(need) > REG_BITS) return false;
                    regs |= need;
                    return true;
                });
            }
            rtaken.swap(tk);
            left -= (int)r_ops.size();
            rounds.push_back({regs, r_ops, hbm});
            first = false;
        }
        if (rounds.empty()) rounds.push_back({0, {}, true});
        if (!rounds.back().hbm) rounds.push_back({0, {}, true});  // pure store round
        // ops that did not fit the round budget go back to the pool (they commute with everything that
        // was scheduled ahead of them, see greedy_pick)
        for (int idx : sw_ops) if (!rtaken[idx]) taken[idx] = 0;
        n_taken += (int)sw_ops.size() - left;
        const bool tail_in_smem = fused && n_taken == nops;   // the fused epilogue picks the tile up from shared memory

        sp.nrounds = (int32_t)rounds.size();
        for (size_t r = 0; r < rounds.size(); ++r) {
            RoundTmp& rt = rounds[r];
            // pad the register set with the highest free tile positions >= LANE_BITS
            uint64_t regs = rt.regs;
            for (int i = TILE_BITS - 1; i >= LANE_BITS && __builtin_popcountll(regs) < REG_BITS; --i)
                if (!(regs >> sp.tileq[i] & 1)) regs |= 1ull << sp.tileq[i];
            PRound& dr = sp.rounds[r];
            int reg_of[64];
            for (int q = 0; q < 64; ++q) reg_of[q] = -1;
            int k = 0;
            for (int i = 0; i < TILE_BITS; ++i)
                if (regs >> sp.tileq[i] & 1) { dr.regpos[k] = i; reg_of[sp.tileq[i]] = k; ++k; }
            // ---- X-type ops that commute to the front / back of the round are folded into its addressing ----
            std::vector<int> lead, trail, middle;
            {
                // An X-type op on a LANE qubit (0..COAL_BITS-1, served by warp shuffles in HBM rounds: 64 SHFL + 64 selects,
                // the hottest spot of a thin-layer sweep in profiles/r2d_sweep_ansatz.md) folds into the GLOBAL address of
                // the round that loads from (lead) / stores to (trail) HBM: the thread simply loads its partner's
                // amplitude -- same 128 B lines per warp access.  Its control must be a plain thread bit (the kernel
                // evaluates the folded ops of an HBM round one after the other on the running index, so a folded op may
                // be controlled by a lane qubit that another folded op flips: fold_gindex).
                const bool loads_hbm = r == 0, stores_hbm = r + 1 == rounds.size() && !tail_in_smem;
                auto foldable_side = [&](const COp& o, bool hbm_side) {
                    if (!fold_perm || o.kind != K_X) return false;
                    if (reg_of[o.t0] >= 0) return true;
                    return hbm_side && o.t0 < COAL_BITS && (o.c < 0 || reg_of[o.c] < 0);
                };
                auto foldable = [&](const COp& o) { return foldable_side(o, loads_hbm); };
                std::vector<char> is_lead(rt.ops.size(), 0), is_trail(rt.ops.size(), 0);
                std::vector<int> before;
                for (size_t i = 0; i < rt.ops.size(); ++i) {
                    const COp& o = ops[rt.ops[i]];
                    bool ok = foldable(o) && (int)lead.size() < MAX_FOLD;
                    for (size_t b = 0; ok && b < before.size(); ++b) ok = x_commutes_with(o, ops[before[b]]);
                    if (ok) { is_lead[i] = 1; lead.push_back(rt.ops[i]); }
                    else before.push_back(rt.ops[i]);
                }
                std::vector<int> after;
                for (size_t i = rt.ops.size(); i-- > 0;) {
                    if (is_lead[i]) continue;
                    const COp& o = ops[rt.ops[i]];
                    bool ok = foldable_side(o, stores_hbm) && (int)trail.size() < MAX_FOLD;
                    for (size_t b = 0; ok && b < after.size(); ++b) ok = x_commutes_with(o, ops[after[b]]);
                    if (ok) { is_trail[i] = 1; trail.push_back(rt.ops[i]); }
                    else after.push_back(rt.ops[i]);
                }
                std::reverse(trail.begin(), trail.end());
                for (size_t i = 0; i < rt.ops.size(); ++i)
                    if (!is_lead[i] && !is_trail[i]) middle.push_back(rt.ops[i]);
            }
            auto expand = [&](uint32_t v) { uint32_t off = 0; for (int b = 0; b < REG_BITS; ++b) if (v >> b & 1) off |= 1u << dr.regpos[b]; return off; };
            auto gexpand = [&](uint32_t v) { uint64_t off = 0; for (int b = 0; b < REG_BITS; ++b) if (v >> b & 1) off |= 1ull << sp.tileq[dr.regpos[b]]; return off; };
            // index-space action of a folded op on the register index j: CX between registers is linear (j ^= j_rc << rt);
            // anything else is a (conditional) flip of bit rt, kept as a per-thread mask
            auto is_linear = [&](const COp& o) { return o.c >= 0 && reg_of[o.c] >= 0; };     // (never true for a lane target)
            auto is_lane = [&](const COp& o) { return reg_of[o.t0] < 0; };
            auto lin_apply = [&](const COp& o, uint32_t j) { return j ^ (((j >> reg_of[o.c]) & 1u) << reg_of[o.t0]); };
            {   // load side: register k receives the amplitude that sat at A^-1 k (^ the thread's flips, pulled in front of A)
                uint32_t fwd[1 << REG_BITS], inv[1 << REG_BITS];
                for (uint32_t j = 0; j < (1u << REG_BITS); ++j) fwd[j] = j;
                dr.n_lead = 0;
                for (int idx : lead) {
                    const COp& o = ops[idx];
                    if (is_linear(o)) { for (uint32_t j = 0; j < (1u << REG_BITS); ++j) fwd[j] = lin_apply(o, fwd[j]); continue; }
                    PFold& f = dr.lead[dr.n_lead++];
                    if (is_lane(o)) { f.cq = o.c; f.smask = 0; f.gmask = 1ull << o.t0; continue; }   // HBM round: lane bit = qubit
                    // flip of e_b applied AFTER the linear ops so far (fwd = B): pulled to the front it flips B^-1 e_b
                    uint32_t u = 0;
                    for (uint32_t j = 0; j < (1u << REG_BITS); ++j) if (fwd[j] == (1u << reg_of[o.t0])) u = j;
                    f.cq = o.c; f.smask = swz(expand(u)); f.gmask = gexpand(u);
                }
                for (uint32_t j = 0; j < (1u << REG_BITS); ++j) inv[fwd[j]] = j;
                for (uint32_t k = 0; k < (1u << REG_BITS); ++k) { dr.soff[k] = (int32_t)swz(expand(inv[k])); dr.goff_ld[k] = gexpand(inv[k]); }
            }
            {   // store side: register j goes to A j (^ the thread's flips, pushed behind A)
                uint32_t fwd[1 << REG_BITS];
                for (uint32_t j = 0; j < (1u << REG_BITS); ++j) fwd[j] = j;
                dr.n_trail = 0;
                for (size_t i = 0; i < trail.size(); ++i) {
                    const COp& o = ops[trail[i]];
                    if (is_linear(o)) { for (uint32_t j = 0; j < (1u << REG_BITS); ++j) fwd[j] = lin_apply(o, fwd[j]); continue; }
                    PFold& f = dr.trail[dr.n_trail++];
                    if (is_lane(o)) { f.cq = o.c; f.smask = 0; f.gmask = 1ull << o.t0; continue; }
                    uint32_t w = 1u << reg_of[o.t0];               // pushed through the linear ops that follow
                    for (size_t l = i + 1; l < trail.size(); ++l) if (is_linear(ops[trail[l]])) w = lin_apply(ops[trail[l]], w);
                    f.cq = o.c; f.smask = swz(expand(w)); f.gmask = gexpand(w);
                }
                for (uint32_t j = 0; j < (1u << REG_BITS); ++j) { dr.soff_st[j] = (int32_t)swz(expand(fwd[j])); dr.goff_st[j] = gexpand(fwd[j]); }
            }
            dr.op_begin = sp.nops;
            bool pending = false;
            for (int idx : middle) {
                POp& po = sp.ops[sp.nops++];
                fill_tiled_op(ops[idx], reg_of, sp, po);
                if (po.kind == P_PEND || po.kind == P_DIAG1 || po.kind == P_DIAG2) { dr.has_pend = 1; pending = true; }
                if (po.kind == P_XLANE || po.kind == P_MAT1LANE) { po.flush = pending ? 1 : 0; pending = false; }
                int code = 0;
                switch (po.kind) {
                case P_PEND: code = 0; break;
                case P_DIAG1: code = 1 + po.r0; break;
                case P_DIAG2: code = 5 + pair_index(po.r0, po.r1); break;
                case P_DIAGRAW: code = 11; break;
                case P_XREG: code = 12 + po.r0; break;
                case P_CXREG: code = 16 + cx_index(po.r0, po.r1); break;
                case P_MAT1: code = 28 + po.r0; break;
                case P_MAT2: code = 32 + pair_index(po.r0, po.r1); break;
                case P_XLANE: code = 38; break;
                default: code = 39; break;
                }
                po.flush |= code << 8;
            }
            dr.op_end = sp.nops;
        }
    }
}

}  // namespace b200
