/*
 * sv_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the statevector arithmetic that the reference's
 * AerSVBackend path executes.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (adapt-aqc_b200/) never links, imports or calls it.
 *
 * The arithmetic itself lives in third-party packages that are absent from
 * /root/reference (qiskit-aer ~=0.16.0 statevector simulator, qiskit ~=1.3.1
 * quantum_info.partial_trace), so this file restates their published
 * behaviour and is pinned on the reference's own call sites and known-answer
 * tests (see tests/test_oracle_kats.py):
 *
 *   - full re-simulation of every gate of compiler.full_circuit from |0..0>
 *     on every evaluation            adaptaqc/backends/aer_sv_backend.py:37-47
 *   - global cost 1-|sv[0]|^2         adaptaqc/backends/aer_sv_backend.py:23-30
 *   - <Z_i> = p0-p1 from one probabilities([i]) pass per qubit
 *                                     adaptaqc/backends/aer_sv_backend.py:49-59
 *   - local cost 0.5*(1-mean<Z>)      adaptaqc/backends/aer_sv_backend.py:32-35
 *   - pair RDM = partial trace over all qubits but (a,b), little-endian
 *                       adaptaqc/utils/entanglement_measures.py:325-340
 *   - gate matrices: standard qiskit definitions of the gates that survive
 *     unroll_to_basis_gates
 *       adaptaqc/utils/circuit_operations/circuit_operations_basic.py:204-205
 *       adaptaqc/utils/circuit_operations/circuit_operations_full_circuit.py:318-326
 *
 * Little-endian: basis index i = sum_q bit_q << q, qubit 0 = LSB.
 * One pass over the 2^n amplitudes per gate (Aer's gate-fusion pass is a
 * performance feature, not restated; results are identical up to rounding).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

/* Same 40-byte wire format as include/b200aqc.h (b200_gate). */
typedef struct {
    int32_t op, q0, q1, aux;
    double p[3];
} orc_gate;

enum {
    OP_ID = 0, OP_X, OP_Y, OP_Z, OP_H, OP_RX, OP_RY, OP_RZ, OP_U1, OP_U2, OP_U3,
    OP_CX, OP_CZ, OP_MAT1, OP_MAT2, OP_S, OP_SDG, OP_T, OP_TDG, OP_SX, OP_SWAP
};

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* 2x2 matrix (row-major m00 m01 m10 m11) of a 1-qubit opcode; returns 0 if op is
 * not a 1-qubit gate. */
static int mat_of_1q(const orc_gate *g, const double *mats, cplx m[4]) {
    const double t = g->p[0], ph = g->p[1], lm = g->p[2];
    const double r = 0.70710678118654752440;
    switch (g->op) {
    case OP_ID: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 1; return 1;
    case OP_X: m[0] = 0; m[1] = 1; m[2] = 1; m[3] = 0; return 1;
    case OP_Y: m[0] = 0; m[1] = -I; m[2] = I; m[3] = 0; return 1;
    case OP_Z: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = -1; return 1;
    case OP_H: m[0] = r; m[1] = r; m[2] = r; m[3] = -r; return 1;
    case OP_S: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = I; return 1;
    case OP_SDG: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = -I; return 1;
    case OP_T: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = cexp(I * M_PI / 4); return 1;
    case OP_TDG: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = cexp(-I * M_PI / 4); return 1;
    case OP_SX:
        m[0] = 0.5 + 0.5 * I; m[1] = 0.5 - 0.5 * I;
        m[2] = 0.5 - 0.5 * I; m[3] = 0.5 + 0.5 * I; return 1;
    case OP_RX: /* exp(-i t X/2) */
        m[0] = cos(t / 2); m[1] = -I * sin(t / 2);
        m[2] = -I * sin(t / 2); m[3] = cos(t / 2); return 1;
    case OP_RY:
        m[0] = cos(t / 2); m[1] = -sin(t / 2);
        m[2] = sin(t / 2); m[3] = cos(t / 2); return 1;
    case OP_RZ:
        m[0] = cexp(-I * t / 2); m[1] = 0; m[2] = 0; m[3] = cexp(I * t / 2); return 1;
    case OP_U1: /* p[0] = lambda */
        m[0] = 1; m[1] = 0; m[2] = 0; m[3] = cexp(I * t); return 1;
    case OP_U2: { /* p[0]=phi p[1]=lambda ; u3(pi/2, phi, lambda) */
        const double phi = g->p[0], lam = g->p[1];
        m[0] = r; m[1] = -cexp(I * lam) * r;
        m[2] = cexp(I * phi) * r; m[3] = cexp(I * (phi + lam)) * r; return 1;
    }
    case OP_U3:
        m[0] = cos(t / 2); m[1] = -cexp(I * lm) * sin(t / 2);
        m[2] = cexp(I * ph) * sin(t / 2); m[3] = cexp(I * (ph + lm)) * cos(t / 2);
        return 1;
    case OP_MAT1: {
        const double *d = mats + g->aux;
        for (int k = 0; k < 4; ++k) m[k] = d[2 * k] + I * d[2 * k + 1];
        return 1;
    }
    default: return 0;
    }
}

static inline uint64_t insert_zero_bit(uint64_t x, int pos) {
    const uint64_t lo = x & ((1ull << pos) - 1);
    return ((x >> pos) << (pos + 1)) | lo;
}

static void apply_1q(int nq, cplx *psi, int q, const cplx m[4]) {
    const uint64_t half = 1ull << (nq - 1), bit = 1ull << q;
    const cplx m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma omp parallel for schedule(static)
    for (uint64_t k = 0; k < half; ++k) {
        const uint64_t i0 = insert_zero_bit(k, q), i1 = i0 | bit;
        const cplx a0 = psi[i0], a1 = psi[i1];
        psi[i0] = m00 * a0 + m01 * a1;
        psi[i1] = m10 * a0 + m11 * a1;
    }
}

/* 4x4 row-major matrix, basis index = bit(q0) + 2*bit(q1) (qiskit convention:
 * first qubit argument is the least significant). */
static void apply_2q(int nq, cplx *psi, int q0, int q1, const cplx m[16]) {
    const uint64_t quarter = 1ull << (nq - 2), b0 = 1ull << q0, b1 = 1ull << q1;
    const int lo = q0 < q1 ? q0 : q1, hi = q0 < q1 ? q1 : q0;
#pragma omp parallel for schedule(static)
    for (uint64_t k = 0; k < quarter; ++k) {
        const uint64_t base = insert_zero_bit(insert_zero_bit(k, lo), hi);
        const uint64_t idx[4] = {base, base | b0, base | b1, base | b0 | b1};
        cplx a[4], o[4];
        for (int j = 0; j < 4; ++j) a[j] = psi[idx[j]];
        for (int r = 0; r < 4; ++r) {
            o[r] = 0;
            for (int c = 0; c < 4; ++c) o[r] += m[4 * r + c] * a[c];
        }
        for (int j = 0; j < 4; ++j) psi[idx[j]] = o[j];
    }
}

/* Apply gates in place.  Returns 0, or -(index+1) of the first bad gate. */
int orc_sv_apply(int nq, double *psi_ri, const orc_gate *g, int ng, const double *mats) {
    cplx *psi = (cplx *)psi_ri;
    for (int k = 0; k < ng; ++k) {
        cplx m[16];
        const int q0 = g[k].q0, q1 = g[k].q1;
        if (q0 < 0 || q0 >= nq) return -(k + 1);
        if (mat_of_1q(&g[k], mats, m)) {
            if (g[k].op == OP_ID) continue;
            apply_1q(nq, psi, q0, m);
            continue;
        }
        if (q1 < 0 || q1 >= nq || q1 == q0) return -(k + 1);
        memset(m, 0, sizeof m);
        switch (g[k].op) {
        case OP_CX: /* control q0, target q1: flips bit q1 where bit q0 = 1 */
            m[0 * 4 + 0] = 1; m[1 * 4 + 3] = 1; m[2 * 4 + 2] = 1; m[3 * 4 + 1] = 1; break;
        case OP_CZ:
            m[0] = 1; m[5] = 1; m[10] = 1; m[15] = -1; break;
        case OP_SWAP:
            m[0] = 1; m[1 * 4 + 2] = 1; m[2 * 4 + 1] = 1; m[15] = 1; break;
        case OP_MAT2: {
            const double *d = mats + g[k].aux;
            for (int j = 0; j < 16; ++j) m[j] = d[2 * j] + I * d[2 * j + 1];
            break;
        }
        default: return -(k + 1);
        }
        apply_2q(nq, psi, q0, q1, m);
    }
    return 0;
}

/* |0..0> then all gates: what one AerSVBackend.evaluate_circuit call does. */
int orc_sv_simulate(int nq, double *psi_ri, const orc_gate *g, int ng, const double *mats) {
    const uint64_t dim = 1ull << nq;
    cplx *psi = (cplx *)psi_ri;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < dim; ++i) psi[i] = 0;
    psi[0] = 1;
    return orc_sv_apply(nq, psi_ri, g, ng, mats);
}

/* sv.probabilities([q]) -> [p0, p1]   (aer_sv_backend.py:56) */
void orc_probabilities(int nq, const double *psi_ri, int q, double out[2]) {
    const cplx *psi = (const cplx *)psi_ri;
    const uint64_t dim = 1ull << nq;
    double p0 = 0, p1 = 0;
#pragma omp parallel for schedule(static) reduction(+ : p0, p1)
    for (uint64_t i = 0; i < dim; ++i) {
        const double a = creal(psi[i]) * creal(psi[i]) + cimag(psi[i]) * cimag(psi[i]);
        if ((i >> q) & 1) p1 += a; else p0 += a;
    }
    out[0] = p0; out[1] = p1;
}

/* rho_ab[(ia + 2 ib), (ja + 2 jb)] = sum_rest psi[ia,ib,rest] conj(psi[ja,jb,rest])
 * with a<b after sorting: qubit a (the lower one) is the least significant
 * index of the 4x4 matrix, as qiskit.quantum_info.partial_trace returns it.
 * out = 16 complex, row-major, interleaved re/im. */
void orc_partial_trace_pair(int nq, const double *psi_ri, int a, int b, double out[32]) {
    const cplx *psi = (const cplx *)psi_ri;
    if (a > b) { int t = a; a = b; b = t; }
    const uint64_t ba = 1ull << a, bb = 1ull << b;
    double acc[32];
    for (int k = 0; k < 32; ++k) acc[k] = 0;
    if (nq == 2) { /* entanglement_measures.py:333-334 */
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                const cplx v = psi[r] * conj(psi[c]);
                out[2 * (4 * r + c)] = creal(v);
                out[2 * (4 * r + c) + 1] = cimag(v);
            }
        return;
    }
    const uint64_t quarter = 1ull << (nq - 2);
#pragma omp parallel for schedule(static) reduction(+ : acc[:32])
    for (uint64_t k = 0; k < quarter; ++k) {
        const uint64_t base = insert_zero_bit(insert_zero_bit(k, a), b);
        const cplx v[4] = {psi[base], psi[base | ba], psi[base | bb], psi[base | ba | bb]};
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                const cplx x = v[r] * conj(v[c]);
                acc[2 * (4 * r + c)] += creal(x);
                acc[2 * (4 * r + c) + 1] += cimag(x);
            }
    }
    for (int k = 0; k < 32; ++k) out[k] = acc[k];
}

/* <L|R> , conjugate on the first argument. */
void orc_vdot(int nq, const double *l_ri, const double *r_ri, double out[2]) {
    const cplx *l = (const cplx *)l_ri, *r = (const cplx *)r_ri;
    const uint64_t dim = 1ull << nq;
    double re = 0, im = 0;
#pragma omp parallel for schedule(static) reduction(+ : re, im)
    for (uint64_t i = 0; i < dim; ++i) {
        const cplx v = conj(l[i]) * r[i];
        re += creal(v); im += cimag(v);
    }
    out[0] = re; out[1] = im;
}
