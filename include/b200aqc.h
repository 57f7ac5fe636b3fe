/*
 * b200aqc.h -- C-ABI of libb200aqc.so, the B200 (sm_100a) simulation library behind
 * ADAPT-AQC's backend interface.
 *
 * The reference (qiskit-community/adapt-aqc) is pure Python; its simulator "FFI" is the set
 * of Python calls it makes into qiskit-aer / aqc_research.  Every entry point below replaces
 * one of those call sites (cited as file:line under /root/reference).  INTEGRATION.md shows
 * the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; b200_last_error() gives the text
 *     (thread-local).  The Python host raises on any non-zero return: there is no CPU
 *     fallback.
 *   - opaque context handle; one context = one GPU + one CUDA stream; not thread-safe per ctx.
 *   - caller-owned host buffers; complex numbers are interleaved (re, im) doubles.
 *   - little-endian qubit order: basis index i = sum_q bit_q << q (qubit 0 = LSB), as
 *     qiskit/Aer (adaptaqc/utils/utilityfunctions.py:145-149,220).
 *   - a statevector context owns `n_slots` device buffers of 2^n complex128 ("slots");
 *     slot ids are small integers chosen by the caller (work state, cached target state
 *     U|0>, bra/ket halves of the Rotosolve evaluation ...).
 */
#ifndef B200AQC_H
#define B200AQC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200AQC_ABI_VERSION 1

/* Gate record, 40 bytes.  Gate names are those that reach the simulator after
 * unroll_to_basis_gates (adaptaqc/utils/circuit_operations/circuit_operations_basic.py:204-205,
 * circuit_operations_full_circuit.py:318-326): u3 cx cz rx ry rz x y z h (+ u1 u2 id from the
 * DEFAULT_GATES unroll of the starting circuit), plus dense fall-backs for custom ansatz gates. */
typedef struct {
    int32_t op;   /* B200_OP_*                                             */
    int32_t q0;   /* qubit (1q gates); control (cx); first qubit (2q)      */
    int32_t q1;   /* target (cx); second qubit (2q); -1 for 1q gates       */
    int32_t aux;  /* MAT1/MAT2: offset (in doubles) into the `mats` array  */
    double p[3];  /* rx/ry/rz: theta; u1: lambda; u2: phi,lambda; u3: theta,phi,lambda */
} b200_gate;

enum {
    B200_OP_ID = 0, B200_OP_X = 1, B200_OP_Y = 2, B200_OP_Z = 3, B200_OP_H = 4,
    B200_OP_RX = 5, B200_OP_RY = 6, B200_OP_RZ = 7,
    B200_OP_U1 = 8, B200_OP_U2 = 9, B200_OP_U3 = 10,
    B200_OP_CX = 11, B200_OP_CZ = 12,
    B200_OP_MAT1 = 13, /* dense 2x2, row-major, 8 doubles at mats[aux]                        */
    B200_OP_MAT2 = 14, /* dense 4x4, row-major, index = bit(q0) + 2*bit(q1), 32 doubles       */
    B200_OP_S = 15, B200_OP_SDG = 16, B200_OP_T = 17, B200_OP_TDG = 18, B200_OP_SX = 19,
    B200_OP_SWAP = 20
};

typedef struct b200_ctx b200_ctx;

/* ---- library / context ------------------------------------------------------------------ */
int b200_abi_version(void);
const char *b200_last_error(void);
int b200_device_count(int *count);
/* Replaces `Aer.get_backend("statevector_simulator")` / `AerSimulator(method=...)` object
 * creation (adaptaqc/backends/aer_sv_backend.py:20, aer_mps_backend.py:37-42). */
int b200_ctx_create(int device, b200_ctx **out);
int b200_ctx_destroy(b200_ctx *ctx);
int b200_ctx_sync(b200_ctx *ctx);
/* Counters since ctx creation: [0] kernels launched, [1] sweeps, [2] gates applied,
 * [3] bytes of algorithmic statevector traffic (32*2^n per sweep, 16*2^n per read-only pass),
 * [4] host->device bytes copied, [5] device->host bytes copied, [6] C-ABI compute calls,
 * [7] real flops issued to the FP64 tensor cores (MPS GEMMs). */
int b200_ctx_counters(b200_ctx *ctx, uint64_t out[8]);
/* Bench brackets: record CUDA event `which` (0 = start, 1 = stop) on the context's stream;
 * elapsed_ms synchronises on the stop event and returns stop - start. */
int b200_ctx_mark(b200_ctx *ctx, int which);
int b200_ctx_elapsed_ms(b200_ctx *ctx, double *ms);
/* Per-kernel-class device times: while enabled, every kernel launch is bracketed by its own CUDA
 * event pair on the context's stream.  enable != 0 also resets the accumulators.  profile_read
 * synchronises the stream and returns, per class, the summed milliseconds and launch count. */
#define B200_PROF_CLASSES 15
enum {
    B200_PROF_SWEEP = 0,  /* sv_sweep_kernel (fused gate sweep, tiled path)  */
    B200_PROF_SMALL = 1,  /* sv_small_kernel (n <= 11, one CTA)              */
    B200_PROF_EXPZ = 2,   /* all-qubit <Z> pass                              */
    B200_PROF_RDM = 3,    /* pair-RDM passes                                 */
    B200_PROF_INNER = 4,  /* <L|.|R> transfer-matrix pass                    */
    B200_PROF_FILL = 5,   /* |0..0> fill / device copies                     */
    B200_PROF_REDUCE = 6, /* final fixed-order reductions                    */
    B200_PROF_MPS = 7,    /* MPS kernels other than the two below            */
    B200_PROF_SVD = 8,    /* on-device Jacobi SVD (jacobi_*_kernel)          */
    B200_PROF_GEMM = 9,   /* complex GEMM on the FP64 tensor cores (DMMA)    */
    B200_PROF_FUSED = 10, /* sv_sweep_inner2_kernel (sweep + transfer pass)  */
    B200_PROF_FUSED_EMBED = 11, /* the same from an embedded source: read `other`, write dst (32 bytes per amplitude) */
    B200_PROF_FUSED_READ = 12,  /* the same, T only: read src and `other`, no write (32 bytes per amplitude) */
    B200_PROF_PROJECT = 13,     /* sv_sweep_project_kernel: sweep that keeps only the projected amplitudes (16 bytes per amplitude) */
    B200_PROF_FUSED_EMBED_READ = 14 /* fused pass from an embedded source, T only: read `other` (16 bytes per amplitude) */
};
int b200_ctx_profile(b200_ctx *ctx, int enable);
int b200_ctx_profile_read(b200_ctx *ctx, double ms[B200_PROF_CLASSES], uint64_t launches[B200_PROF_CLASSES]);
/* Per-launch device times (ms, launch order) of the sweep kernel since b200_ctx_profile(ctx, 1): *n = number
 * of launches recorded, the first min(*n, max_entries) go to ms[].  Roofline accounting per sweep. */
int b200_ctx_profile_sweeps(b200_ctx *ctx, double *ms, int max_entries, int *n);
/* CUDA-event time (ms) of the kernels launched by the most recent b200_sv_* / b200_mps_* call. */
int b200_ctx_last_ms(b200_ctx *ctx, double *ms);
int b200_ctx_set_timing(b200_ctx *ctx, int enable);

/* ---- statevector path ------------------------------------------------------------------- */
/* (Re)allocate `n_slots` device statevectors of `num_qubits` qubits. */
int b200_sv_alloc(b200_ctx *ctx, int num_qubits, int n_slots);
/* Declare `n_slots` empty slots of `num_qubits` qubits without allocating: every slot must be
 * given memory with b200_sv_attach before use (the sharded multi-GPU path attaches buffers that are
 * also registered with the NCCL communicator). */
int b200_sv_reserve_slots(b200_ctx *ctx, int num_qubits, int n_slots);
/* Use caller-provided device memory (e.g. a torch CUDA tensor's data_ptr()) for one slot. */
int b200_sv_attach(b200_ctx *ctx, int slot, void *device_ptr);
int b200_sv_device_ptr(b200_ctx *ctx, int slot, void **out);
/* Multi-GPU statevector sharded by global qubits (one process per GPU): peer access to another rank's
 * slot through CUDA IPC, and the in-place exchange of global with local qubits over NVLink peer memory
 * (replaces the NCCL send/recv + staging + copy-back of dist_sv.TorchComm.exchange_chunks; the reference
 * has no multi-device path -- SURVEY 8e).  `handle` is a cudaIpcMemHandle_t (64 bytes) of a library-owned
 * slot; peer_ptrs[p] = pointer opened from rank p's handle (entry `rank` ignored).  peer_swap trades
 * chunk[p] of this rank's slot with chunk[rank] of rank p's, for all p, in one kernel; the caller
 * synchronises the ranks before and after. */
int b200_sv_ipc_export(b200_ctx *ctx, int slot, unsigned char handle[64]);
int b200_sv_ipc_open(b200_ctx *ctx, const unsigned char handle[64], void **peer_ptr);
int b200_sv_ipc_close(b200_ctx *ctx, void *peer_ptr);
int b200_sv_peer_swap(b200_ctx *ctx, int slot, void *const *peer_ptrs, int world, int rank);
/* The same exchange for outgoing qubits that sit ANYWHERE in the local index: positions[j] (j < log2(world)) is the
 * local bit position that trades places with rank bit j.  Chunk p is then the strided set of amplitudes whose bits at
 * those positions spell p.  Saves the SWAP-localisation sweep (one read+write pass over the slice) that must otherwise
 * precede every exchange to bring the outgoing qubits to the top of the local index. */
int b200_sv_peer_swap_strided(b200_ctx *ctx, int slot, void *const *peer_ptrs, int world, int rank,
                              const int32_t *positions);
int b200_sv_num_qubits(b200_ctx *ctx, int *out);

/* slot <- |0...0> */
int b200_sv_init_zero(b200_ctx *ctx, int slot);
/* dst <- src (device copy) */
int b200_sv_copy(b200_ctx *ctx, int dst_slot, int src_slot);

/* dst <- G_{n-1} ... G_1 G_0 src.  src_slot = -1 means |0...0>.  dst may equal src.
 * This is the replacement for `simulator.run(full_circuit).result().get_statevector()`
 * (adaptaqc/backends/aer_sv_backend.py:42-47; circuit_operations_running.py:58-63): the gate
 * stream is planned into fused sweeps on the host and executed by the sm_100a gate kernels;
 * the state stays resident in HBM.  `mats` may be NULL when no MAT1/MAT2 op is present. */
int b200_sv_run(b200_ctx *ctx, int dst_slot, int src_slot, const b200_gate *gates, int n_gates,
                const double *mats, int n_mats);
/* Same, applying the inverse circuit (gates reversed, each inverted): dst <- G_0^+ ... G_{n-1}^+ src */
int b200_sv_run_inverse(b200_ctx *ctx, int dst_slot, int src_slot, const b200_gate *gates,
                        int n_gates, const double *mats, int n_mats);

/* out = amplitude <index|psi>  -> `sv[0]` in 1-|sv[0]|^2 (aer_sv_backend.py:29). */
int b200_sv_amp(b200_ctx *ctx, int slot, uint64_t index, double out[2]);
/* out[q] = <Z_q> for every qubit in one pass; out[n] = <psi|psi>.  Replaces the n separate
 * `sv.probabilities([i])` passes (aer_sv_backend.py:49-59). */
int b200_sv_expz(b200_ctx *ctx, int slot, double *out /* n+1 */);
/* out[p] = 4x4 reduced density matrix (row-major, 32 doubles) of pairs[2p], pairs[2p+1]; the
 * lower-numbered qubit of the pair is the least-significant matrix index, as
 * qiskit.quantum_info.partial_trace returns it (adaptaqc/utils/entanglement_measures.py:325-340).
 * All pairs from a few read passes instead of one re-simulation + host trace per pair
 * (adaptaqc/compilers/adapt/adapt_compiler.py:964-975). */
int b200_sv_pair_rdm(b200_ctx *ctx, int slot, const int32_t *pairs, int n_pairs, double *out);
/* One share of the same work, for ranks that hold REPLICAS of the state ("pair-RDM kernel: split the
 * pair list across GPUs, allgather"): the pass list is the one b200_sv_pair_rdm builds for `pairs`, only the passes
 * k = part (mod n_parts) are launched, pairs owned by other parts are returned as zeros.  Summing `out` over the parts
 * (one small all-reduce) reproduces the undivided call bit for bit, so the pair chosen by
 * adaptaqc/compilers/adapt/adapt_compiler.py:858-921 does not depend on the number of GPUs. */
int b200_sv_pair_rdm_part(b200_ctx *ctx, int slot, const int32_t *pairs, int n_pairs, int part, int n_parts,
                          double *out);
/* out = 2x2 complex M[i][j] = sum_rest conj(L[i,rest]) R[j,rest] for qubit q (row-major,
 * 8 doubles).  <L|G_q|R> = sum_ij G[i][j] M[i][j] then gives the cost for ANY 1-qubit gate G
 * on q, i.e. all Rotosolve / Rotoselect shift evaluations of one gate
 * (adaptaqc/utils/cost_minimiser.py:318-368) from one launch.  q = -1: out[0..1] = <L|R>. */
int b200_sv_inner(b200_ctx *ctx, int l_slot, int r_slot, int q, double out[8]);

/* out = 4x4 complex T[i][j] = sum_rest conj(L[i,rest]) R[j,rest], i,j = bit(qa) + 2 bit(qb)
 * (row-major, 32 doubles).  <L|O|R> = sum_ij O[i][j] T[i][j] for ANY operator O supported on
 * (qa, qb): a whole ansatz layer (rz rz cx rz rz on one pair,
 * adaptaqc/utils/circuit_operations/circuit_operations_basic.py:135-189) is such an operator, so
 * one read pass over L and R serves every Rotoselect / Rotosolve evaluation of that layer
 * (adaptaqc/utils/cost_minimiser.py:267-368), for any number of optimiser cycles. */
int b200_sv_inner2(b200_ctx *ctx, int l_slot, int r_slot, int qa, int qb, double out[32]);

/* b200_sv_run (or b200_sv_run_inverse: `inverse` != 0) followed by b200_sv_inner2 with the swept state as the bra, in ONE
 * pass: dst <- gates applied to src, out = T[i][j] = sum_rest conj(dst[i,rest]) other[j,rest].  The last sweep of the
 * program keeps each tile in shared memory and contracts it with `other` before storing it, so the pair costs
 * 48 * 2^n bytes of HBM traffic instead of 64 * 2^n.  The optimiser's walk from one ansatz block to the next
 * (adaptaqc/utils/cost_minimiser.py:267-316: one small edit of the bra, then a fresh transfer matrix) is exactly this
 * pair.  dst may equal src; src_slot = -1: the source is |0..0> as in b200_sv_run (no read pass); `other` must differ from
 * dst.  Registers of fewer than 12 qubits: error (call the two functions).  `stored` (may be NULL = always store): *stored = 0 on entry asks for T only -- when one sweep carries the
 * whole program the swept state is then NOT written (32 * 2^n bytes: the two reads) and dst keeps its contents; on return
 * *stored says whether dst was written (a program of several sweeps needs dst for its intermediate states). */
int b200_sv_run_inner2(b200_ctx *ctx, int dst_slot, int src_slot, const b200_gate *gates, int n_gates, const double *mats,
                       int n_mats, int inverse, int other_slot, int qa, int qb, double out[32], int *stored);

/* dst <- gates applied to the EMBEDDED state: compact_state (2^K amplitudes on the device, e.g. a slot of a K-qubit context
 * that the caller has synchronised) on the qubits qmap[0..K), |0> on every other qubit.  Equivalent to b200_sv_scatter
 * followed by b200_sv_run, without the zero-fill pass and without the sweep's read pass: the first sweep reads its tiles
 * from the compact array (16 * 2^n bytes instead of 16 + 32).  The bra (suffix)^+ |0..0> of the optimiser's walk
 * (adaptaqc/utils/cost_minimiser.py:267-316) is such a state: its tail is built on a small register and only the head gates
 * run at full size.  b200_sv_run_embedded_inner2: the same followed by the transfer pass of b200_sv_run_inner2, in one pass
 * (read `other`, write dst: 32 * 2^n bytes against 16 + 32 + 32); `stored` as in b200_sv_run_inner2: with *stored = 0 a
 * one-sweep program does not write dst either -- the whole bra exists only tile by tile in shared memory and the pass is
 * ONE read of `other` (16 * 2^n bytes) for a transfer matrix that the reference obtains from 3-7 full re-simulations. */
/* b200_sv_run followed by b200_sv_gather without storing the swept state: compact_dst[c] = (gates applied to src)[deposit(c,
 * qmap)], i.e. the swept state projected onto |0> of every qubit outside qmap -- the input of the "projected tail"
 * (DESIGN 2.3).  The last sweep stores only the 2^K surviving amplitudes (16 * 2^n + 16 * 2^K bytes instead of 32 * 2^n plus
 * the gather).  A program of several sweeps uses `scratch_slot` for its intermediate states (*scratch_used = 1). */
int b200_sv_run_project(b200_ctx *ctx, int scratch_slot, int src_slot, const b200_gate *gates, int n_gates, const double *mats,
                        int n_mats, int inverse, void *compact_dst, int K, const int32_t *qmap, int *scratch_used);
int b200_sv_run_embedded(b200_ctx *ctx, int dst_slot, const void *compact_state, int K, const int32_t *qmap,
                         const b200_gate *gates, int n_gates, const double *mats, int n_mats, int inverse);
int b200_sv_run_embedded_inner2(b200_ctx *ctx, int dst_slot, const void *compact_state, int K, const int32_t *qmap,
                                const b200_gate *gates, int n_gates, const double *mats, int n_mats, int inverse,
                                int other_slot, int qa, int qb, double out[32], int *stored);

/* Same T as b200_sv_inner2, for a bra that is given COMPACTLY: <L| = (suffix)^+ <0..0| is supported
 * only on the K qubits the suffix touches, so it is stored as a 2^K-amplitude device array
 * `compact_state` (e.g. slot memory of a second, K-qubit context on the same device) with
 * qmap[b] = the full-register qubit of compact bit b.  The kernel gathers the 2^K needed amplitudes
 * of R instead of reading all 2^n: for the newest ADAPT layer (empty suffix) that is 4 amplitudes.
 * The caller must have synchronised the context that produced `compact_state`. */
int b200_sv_inner2_gather(b200_ctx *ctx, int r_slot, const void *compact_state, int K, const int32_t *qmap,
                          int qa, int qb, double out[32]);

/* Projection onto |0> of every qubit outside qmap: dst[c] = slot[deposit(c, qmap)], c < 2^K (dst = device
 * memory of 2^K amplitudes on the same device, e.g. a slot of a K-qubit context).  Once the remaining gates
 * of the window (utils/cost_minimiser.py:267-368 walks them layer by layer) act only on these K qubits, every
 * further cost evaluation <0|W|psi> = <0_K|W_K|dst> runs on the 2^K-amplitude state. */
int b200_sv_gather(b200_ctx *ctx, int slot, const int32_t *qmap, int K, void *dst);
/* The inverse embedding: slot[x] = src[extract(x, qmap)] where every qubit outside qmap is 0, and 0 elsewhere
 * (a bra <L| = suffix^+ <0| whose tail was built on the K-qubit context is placed into the register). */
int b200_sv_scatter(b200_ctx *ctx, int slot, const int32_t *qmap, int K, const void *src);
/* Same for one slice of a register sharded by global qubits: qmap entries >= num_qubits name rank bit
 * (entry - num_qubits).  The rank writes the amplitudes it owns and zeros elsewhere; summing the ranks'
 * results (all-reduce) gives the projection of the whole register. */
int b200_sv_gather_ranked(b200_ctx *ctx, int slot, const int32_t *qmap, int K, int rank_bits, int rank, void *dst);
/* Host <-> device transfer of `count` amplitudes starting at `offset` (tests, small n, target
 * upload).  Replaces the Statevector object's `.data`. */
int b200_sv_download(b200_ctx *ctx, int slot, uint64_t offset, uint64_t count, double *host);
int b200_sv_upload(b200_ctx *ctx, int slot, uint64_t offset, uint64_t count, const double *host);

/* ---- matrix-product-state path ------------------------------------------------------------ */
/* A b200_mps is a Vidal-form MPS resident in HBM: Gamma_i as [2][chi_{i-1}][chi_i] complex128 and
 * lambda_i as chi_i doubles -- the content of the reference's QiskitMPS wire format
 * (adaptaqc/utils/constants.py:17).  Host-side layouts below: `gammas` = the n site tensors back to
 * back, each [2][chi_l][chi_r] interleaved complex; `lambdas` = the n-1 bond vectors back to back;
 * `bond_dims[i]` = chi_i, i < n-1 (chi_{-1} = chi_{n-1} = 1). */
typedef struct b200_mps b200_mps;

/* Replaces `AerSimulator(method="matrix_product_state", matrix_product_state_truncation_threshold,
 * matrix_product_state_max_bond_dimension)` (adaptaqc/backends/aer_mps_backend.py:27-42).  The new
 * MPS is |0...0>.  max_bond_dimension <= 0 means unlimited. */
int b200_mps_create(b200_ctx *ctx, int num_qubits, double truncation_threshold, int max_bond_dimension,
                    b200_mps **out);
int b200_mps_destroy(b200_mps *mps);
int b200_mps_set_truncation(b200_mps *mps, double truncation_threshold, int max_bond_dimension);
int b200_mps_num_qubits(b200_mps *mps, int *out);
int b200_mps_init_zero(b200_mps *mps);
/* `set_matrix_product_state` (adaptaqc/compilers/approximate_compiler.py:196-204): loads (Gamma,
 * lambda) verbatim. */
int b200_mps_set(b200_mps *mps, const int32_t *bond_dims, const double *gammas, const double *lambdas);
/* `save_matrix_product_state` as used by aqc_research.mps_from_circuit
 * (adaptaqc/backends/aer_mps_backend.py:76-78): bond dimensions first, then the tensors. */
int b200_mps_bond_dims(b200_mps *mps, int32_t *out /* n-1 */);
int b200_mps_get(b200_mps *mps, double *gammas, double *lambdas);
int b200_mps_copy(b200_mps *dst, b200_mps *src);
/* Applies the gate stream: 1-qubit gates contract the physical index; 2-qubit gates contract the
 * two sites (complex GEMM on FP64 tensor cores), run the on-device Jacobi SVD and truncate with
 * Aer's rule (see b200_mps_set_chop_rule: chop, cap at max bond, drop smallest while the sum of squares
 * stays below the threshold, renormalise if anything was dropped); non-neighbours are swapped together and
 * back.  Replaces the Aer MPS run inside mps_from_circuit (aer_mps_backend.py:78;
 * adaptaqc/compilers/adapt/adapt_compiler.py:1129-1131). */
int b200_mps_apply(b200_mps *mps, const b200_gate *gates, int n_gates, const double *mats, int n_mats);
int b200_mps_apply_inverse(b200_mps *mps, const b200_gate *gates, int n_gates, const double *mats, int n_mats);
/* out = T[i][j] = <a| (|i><j| on `qubits`) |b>: i = bra, j = ket physical index, bit(qubits[0]) +
 * 2 bit(qubits[1]); n_open = 0 (<a|b>, 2 doubles), 1 (2x2) or 2 (4x4, 32 doubles), row-major.
 * The MPS counterpart of b200_sv_inner / b200_sv_inner2: with |b> = prefix applied to the target
 * and <a| = suffix applied to <0|, sum_ij O[i][j] T[i][j] is <0|psi> for ANY operator O on the open
 * qubits, i.e. every Rotoselect / Rotosolve value of a layer (adaptaqc/utils/cost_minimiser.py:318-368)
 * from one environment sweep instead of one Aer MPS run + mps_dot per value. */
int b200_mps_transfer(b200_mps *a, b200_mps *b, const int32_t *qubits, int n_open, double *out);
/* out[t] = <bitstrings[t] | psi> (little-endian integers, n <= 64), all in one launch: `mps_dot`
 * with |0..0> (aer_mps_backend.py:54) is bitstring 0; the Hamming-weight-one overlaps are
 * bitstrings 2^i (aer_mps_backend.py:88-93, aqc_research.extract_amplitude). */
int b200_mps_amps(b200_mps *mps, const uint64_t *bitstrings, int count, double *out /* 2*count */);
/* out = <a|b> (conjugate on a): aqc_research.mps_dot (adaptaqc/utils/gradients.py:77,94,110). */
int b200_mps_dot(b200_mps *a, b200_mps *b, double out[2]);
/* out[q] = <Z_q> for all q, out[n] = <psi|psi>: aqc_research.mps_expectation per qubit
 * (aer_mps_backend.py:80-86), all qubits from one right + one left sweep. */
int b200_mps_expz(b200_mps *mps, double *out /* n+1 */);
/* 4x4 reduced density matrices, same conventions as b200_sv_pair_rdm: aqc_research.partial_trace
 * (adaptaqc/utils/entanglement_measures.py:76-79). */
int b200_mps_pair_rdm(b200_mps *mps, const int32_t *pairs, int n_pairs, double *out /* 32*n_pairs */);
/* out[p] = T_p[i][j] = <a| (|i><j| on pairs[2p], pairs[2p+1]) |b>: i = bra, j = ket, index = bit(pairs[2p]) +
 * 2 bit(pairs[2p+1]); 32 doubles per pair, row-major.  ALL pairs from one left + one right environment sweep of
 * <a|b> (pairs sharing their lower qubit share the open environments).  Replaces the P x (#generators + 1) Aer MPS
 * runs + mps_dot calls of general_grad_of_pairs (adaptaqc/utils/gradients.py:23-124, called at
 * adaptaqc/compilers/adapt/adapt_compiler.py:839-856): with <a| = <s| (the starting state) and |b> = |psi>,
 * <s|G|psi> = sum_ij G[i][j] T_p[i][j] for every generator G on the pair. */
int b200_mps_pair_transfer(b200_mps *a, b200_mps *b, const int32_t *pairs, int n_pairs, double *out /* 32*n_pairs */);
/* Truncation rule of the 2-qubit gates = qiskit-aer's reduce_zeros (svd.cpp, called from the Aer MPS run the
 * reference starts at adaptaqc/backends/aer_mps_backend.py:37-42,78).  The Aer source is not part of the reference
 * tree; its rule is restated and the two ambiguous details are selectable so that both readings stay testable:
 *   B200_CHOP_AER   (default) count values with sigma^2 > 1e-16 (std::norm of a real), and leave the count unchanged
 *                   when the tail-drop loop runs out without a break;
 *   B200_CHOP_SIGMA count values with sigma > 1e-16, a loop that runs out keeps one value (round-1 behaviour).
 * Process-wide.  b200_mps_reduce_zeros applies the current rule to a descending vector on the host (no GPU):
 * *n_kept values, renormalised if anything was dropped, are written to kept_out (capacity n). */
#define B200_CHOP_AER 0
#define B200_CHOP_SIGMA 1
int b200_mps_set_chop_rule(int rule);
int b200_mps_reduce_zeros(const double *s_desc, int n, int max_bond_dimension, double truncation_threshold,
                          int *n_kept, double *kept_out);
/* out = {SVDs run, Jacobi sweeps run, current max bond dimension, algorithmic flop of those sweeps:
 * (40 p + 28 q) flop per column pair per sweep of a p x q matrix}. */
int b200_mps_stats(b200_mps *mps, uint64_t out[4]);

/* Planner introspection (no GPU needed): how many sweeps / rounds / fused ops the gate stream
 * compiles to for an n-qubit state.  out = {sweeps, rounds, ops, small_path}. */
int b200_sv_plan_stats(int num_qubits, const b200_gate *gates, int n_gates, const double *mats,
                       int n_mats, int32_t out[4]);

/* Per-sweep detail of the plan: out[4k..4k+3] = {rounds, fused ops, dense (FP64-heavy) ops, number of
 * leading contiguous low qubits of the tile} for sweep k < max_sweeps; *n_sweeps = total sweeps. */
int b200_sv_plan_detail(int num_qubits, const b200_gate *gates, int n_gates, const double *mats, int n_mats,
                        int32_t *out, int max_sweeps, int32_t *n_sweeps);

#ifdef __cplusplus
}
#endif
#endif /* B200AQC_H */
