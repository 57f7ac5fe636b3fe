"""CPU ORACLE -- matrix-product-state path (test infrastructure, NOT product code).

numpy restatement of the two third-party pieces the reference's AerMPSBackend path executes.
Neither is vendored in /root/reference nor installable offline, so their published algorithms
are restated here and pinned on the reference's call sites, known-answer tests and fixtures
(tests/test_mps_oracle.py).  **Parity status: truncation behaviour is "parity unpinned"** -- no
test of the reference fixes Aer's truncation numerically; exact (untruncated) behaviour is
pinned against the statevector oracle and the 54 paper/random_mps fixtures.

(1) qiskit-aer ~=0.16.0, `AerSimulator(method="matrix_product_state")`
    (adaptaqc/backends/aer_mps_backend.py:27-42): Vidal-form MPS (Gamma_i as a pair of
    chi_{i-1} x chi_i matrices, lambda_i real); 1-qubit gate = contraction of the physical index;
    2-qubit gate on neighbours = contract two sites with the surrounding lambdas, apply the 4x4
    matrix, SVD of the (2 chi_l x 2 chi_r) matrix, truncate, divide the outer lambdas back out;
    non-neighbours are brought together with swap gates.  Truncation ("reduce_zeros"): keep
    singular values with sigma^2 > 1e-16 (see set_chop_rule), cap at max_bond_dimension, then drop the smallest while the running
    sum of their squares stays below truncation_threshold; if anything was dropped renormalise
    the kept values to unit 2-norm.
(2) aqc_research.mps_operations (unpinned git dependency, setup.py:22): mps_from_circuit,
    _preprocess_mps, mps_dot, mps_expectation, extract_amplitude, partial_trace, check_mps,
    mps_to_vector -- call sites: adaptaqc/backends/aer_mps_backend.py:14-19,54,78,83,90;
    adaptaqc/utils/entanglement_measures.py:76-79; adaptaqc/utils/gradients.py:60-111;
    adaptaqc/compilers/approximate_compiler.py:133-135,198-200.

Wire format `QiskitMPS = (gammas, lambdas)`: adaptaqc/utils/constants.py:17.  "Preprocessed" =
list of n arrays (2, chi_l, chi_r) with lambda_i multiplied into the right bond of site i.
Qubit i = site i; little-endian integers for bitstrings (aer_mps_backend.py:88-93).

Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
import numpy as np

CHOP_THRESHOLD = 1e-16

_R2 = 1 / np.sqrt(2)
_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def gate_matrix(name, params):
    """Standard qiskit gate matrices; 2-qubit matrices are indexed bit(first qubit) + 2 bit(second)."""
    p = list(params) + [0.0, 0.0, 0.0]
    t = p[0]
    c, s = np.cos(t / 2), np.sin(t / 2)
    table = {
        "id": lambda: np.eye(2), "x": lambda: np.array([[0, 1], [1, 0]]),
        "y": lambda: np.array([[0, -1j], [1j, 0]]), "z": lambda: np.diag([1, -1]),
        "h": lambda: _R2 * np.array([[1, 1], [1, -1]]),
        "s": lambda: np.diag([1, 1j]), "sdg": lambda: np.diag([1, -1j]),
        "t": lambda: np.diag([1, np.exp(0.25j * np.pi)]), "tdg": lambda: np.diag([1, np.exp(-0.25j * np.pi)]),
        "sx": lambda: 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]),
        "rx": lambda: np.array([[c, -1j * s], [-1j * s, c]]),
        "ry": lambda: np.array([[c, -s], [s, c]]),
        "rz": lambda: np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)]),
        "u1": lambda: np.diag([1, np.exp(1j * t)]), "p": lambda: np.diag([1, np.exp(1j * t)]),
        "u2": lambda: _R2 * np.array([[1, -np.exp(1j * p[1])], [np.exp(1j * p[0]), np.exp(1j * (p[0] + p[1]))]]),
        "u3": lambda: np.array([[c, -np.exp(1j * p[2]) * s], [np.exp(1j * p[1]) * s, np.exp(1j * (p[1] + p[2])) * c]]),
        "cx": lambda: np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]]),
        "cz": lambda: np.diag([1, 1, 1, -1]),
        "swap": lambda: _SWAP,
    }
    table["u"] = table["u3"]
    return np.asarray(table[name](), dtype=np.complex128)


# ---------------------------------------------------------------------------------------------
# (1) Aer's MPS simulator
# ---------------------------------------------------------------------------------------------
CHOP_AER, CHOP_SIGMA = "aer", "sigma"
CHOP_RULE = CHOP_AER


def set_chop_rule(rule):
    """Select how Aer's reduce_zeros is read (the Aer source is not vendored in the reference tree):

    "aer"   (default) num_of_SV counts ``std::norm(S[i]) > CHOP_THRESHOLD`` -- the norm of a real number is its
            square, so sigma^2 > 1e-16 -- and the tail-drop loop lowers the count only when it breaks;
    "sigma" round-1 reading: sigma > 1e-16, and a loop that runs out keeps one value.
    Mirrors ``b200_mps_set_chop_rule`` of the product so that both readings can be parity-tested."""
    global CHOP_RULE
    if rule not in (CHOP_AER, CHOP_SIGMA):
        raise ValueError(rule)
    CHOP_RULE = rule


def reduce_zeros(S, max_bond_dimension, truncation_threshold):
    """Number of singular values kept + the (possibly renormalised) values.  S descending.
    Restates reduce_zeros / num_of_SV of qiskit-aer 0.16 svd.cpp (reached from aer_mps_backend.py:37-42,78)."""
    S = np.asarray(S, dtype=np.float64)
    aer = CHOP_RULE == CHOP_AER
    sv_num = int(np.count_nonzero((S * S if aer else S) > CHOP_THRESHOLD))
    new_num = sv_num
    if max_bond_dimension is not None and max_bond_dimension < sv_num:
        new_num = int(max_bond_dimension)
    sum_squares = 0.0
    i = new_num - 1
    broke = False
    while i > 0:
        if sum_squares + S[i] ** 2 < truncation_threshold:
            sum_squares += S[i] ** 2
            i -= 1
        else:
            broke = True
            break
    if broke or not aer:
        new_num = i + 1
    new_num = max(1, new_num)
    kept = np.array(S[:new_num], dtype=np.float64)
    if new_num < sv_num:
        kept = kept / np.sqrt(np.sum(kept ** 2))
    return new_num, kept


class AerMPSState:
    """Vidal-form MPS: gam[i] has shape (2, chi_{i-1}, chi_i); lam[i] (i < n-1) real, length chi_i."""

    def __init__(self, n, truncation_threshold=1e-16, max_bond_dimension=None):
        self.n = n
        self.thr = truncation_threshold
        self.max_chi = max_bond_dimension
        self.gam = [np.array([[[1.0 + 0j]], [[0.0 + 0j]]]) for _ in range(n)]
        self.lam = [np.ones(1) for _ in range(n - 1)]
        self.svd_count = 0

    def set_mps(self, mps):
        gammas, lambdas = mps
        self.gam = [np.stack([np.asarray(a0, dtype=np.complex128), np.asarray(a1, dtype=np.complex128)])
                    for a0, a1 in gammas]
        self.lam = [np.asarray(l, dtype=np.float64).reshape(-1) for l in lambdas]

    def get_mps(self):
        return ([(g[0].copy(), g[1].copy()) for g in self.gam], [l.copy() for l in self.lam])

    def _lam(self, i):
        return self.lam[i] if 0 <= i < self.n - 1 else np.ones(1)

    def apply_1q(self, q, m):
        self.gam[q] = np.einsum("ab,bxy->axy", m, self.gam[q])

    def _apply_adjacent(self, i, m4):
        """m4 indexed bit(site i) + 2 bit(site i+1)."""
        ll, lm, lr = self._lam(i - 1), self._lam(i), self._lam(i + 1)
        A = self.gam[i] * ll[None, :, None] * lm[None, None, :]
        B = self.gam[i + 1] * lr[None, None, :]
        theta = np.einsum("axb,cby->caxy", A, B)          # [b_{i+1}, b_i, chi_l, chi_r]
        u = m4.reshape(2, 2, 2, 2)                          # [b'_{i+1}, b'_i, b_{i+1}, b_i]
        theta = np.einsum("pqca,caxy->pqxy", u, theta)
        chi_l, chi_r = theta.shape[2], theta.shape[3]
        M = theta.transpose(1, 2, 0, 3).reshape(2 * chi_l, 2 * chi_r)   # rows (b_i, chi_l), cols (b_{i+1}, chi_r)
        U, S, Vh = np.linalg.svd(M, full_matrices=False)
        self.svd_count += 1
        k, kept = reduce_zeros(S, self.max_chi, self.thr)
        U = U[:, :k].reshape(2, chi_l, k)
        Vh = Vh[:k, :].reshape(k, 2, chi_r).transpose(1, 0, 2)
        self.gam[i] = U / ll[None, :, None]
        self.gam[i + 1] = Vh / lr[None, None, :]
        self.lam[i] = kept

    def apply_2q(self, q0, q1, m4):
        """m4 indexed bit(q0) + 2 bit(q1)."""
        if q0 > q1:   # re-index the matrix so that its first index bit belongs to the lower site
            perm = [0, 2, 1, 3]
            m4 = m4[np.ix_(perm, perm)]
            q0, q1 = q1, q0
        for j in range(q1 - 1, q0, -1):      # bring q1 next to q0
            self._apply_adjacent(j, _SWAP)
        self._apply_adjacent(q0, m4)
        for j in range(q0 + 1, q1):          # and back to sorted order
            self._apply_adjacent(j, _SWAP)

    def apply_gates(self, gates):
        """gates: [(name, qubits, params)]; mat1/mat2 carry the matrix as params."""
        for name, qubits, params in gates:
            if name in ("mat1", "mat2"):
                m = np.asarray(params, dtype=np.complex128)
            elif name in ("id", "barrier"):
                continue
            else:
                m = gate_matrix(name, params)
            if len(qubits) == 1:
                self.apply_1q(qubits[0], m)
            else:
                self.apply_2q(qubits[0], qubits[1], m)


class OracleMPSSimulator:
    """`AerSimulator(method="matrix_product_state", ...)` stand-in (aer_mps_backend.py:37-42)."""

    class _Options:
        def __init__(self, thr, max_chi):
            self.matrix_product_state_truncation_threshold = thr
            self.matrix_product_state_max_bond_dimension = max_chi

    def __init__(self, mps_truncation_threshold=1e-16, max_chi=None):
        self.options = self._Options(mps_truncation_threshold, max_chi)
        self.runs = 0

    def simulate(self, circuit):
        """Runs a QuantumCircuit-shaped object whose first instruction may be
        set_matrix_product_state; returns the final AerMPSState."""
        self.runs += 1
        st = AerMPSState(circuit.num_qubits, self.options.matrix_product_state_truncation_threshold,
                         self.options.matrix_product_state_max_bond_dimension)
        gates = []
        for inst in circuit.data:
            op = inst.operation
            if op.name == "set_matrix_product_state":
                st.set_mps(op.params[0])
                continue
            if op.name in ("save_matrix_product_state", "barrier"):
                continue
            qs = [int(q) for q in inst.qubits]
            try:
                gate_matrix(op.name, op.params)
                gates.append((op.name, qs, [float(x) for x in op.params]))
            except KeyError:
                gates.append(("mat2" if len(qs) == 2 else "mat1", qs, np.asarray(op.to_matrix())))
        st.apply_gates(gates)
        return st


# ---------------------------------------------------------------------------------------------
# (2) aqc_research.mps_operations
# ---------------------------------------------------------------------------------------------
def check_mps(x):
    """True iff x is a QiskitMPS tuple (approximate_compiler.py:121,181)."""
    return (isinstance(x, tuple) and len(x) == 2 and isinstance(x[0], list) and isinstance(x[1], list)
            and len(x[0]) == len(x[1]) + 1 and all(isinstance(g, tuple) and len(g) == 2 for g in x[0]))


def _preprocess_mps(mps):
    gammas, lambdas = mps
    n = len(gammas)
    out = []
    for i, (a0, a1) in enumerate(gammas):
        g = np.stack([np.asarray(a0, dtype=np.complex128), np.asarray(a1, dtype=np.complex128)])
        if i < n - 1:
            g = g * np.asarray(lambdas[i], dtype=np.float64).reshape(1, 1, -1)
        out.append(g)
    return out


def _pp(mps, already_preprocessed):
    return mps if already_preprocessed else _preprocess_mps(mps)


def mps_from_circuit(qc, trunc_thr=1e-16, print_log_data=False, return_preprocessed=False, sim=None):
    """Appends a save instruction to `qc` IN PLACE (pinned by test_utilityfunctions.py:186-193:
    callers pass copies), runs it, returns QiskitMPS or the preprocessed list."""
    if sim is None:
        sim = OracleMPSSimulator(trunc_thr)
    if hasattr(qc, "save_matrix_product_state"):
        qc.save_matrix_product_state()
    st = sim.simulate(qc)
    mps = st.get_mps()
    return _preprocess_mps(mps) if return_preprocessed else mps


def mps_dot(mps1, mps2, already_preprocessed=False):
    """<mps1|mps2> (conjugate on the first argument)."""
    a, b = _pp(mps1, already_preprocessed), _pp(mps2, already_preprocessed)
    env = np.ones((1, 1), dtype=np.complex128)
    for ga, gb in zip(a, b):
        P = _site_products(env, ga, gb)
        env = P[0, 0] + P[1, 1]
    return complex(env[0, 0])


def extract_amplitude(mps, bitstring, already_preprocessed=False):
    """<b|psi>, little-endian integer b (aer_mps_backend.py:88-93)."""
    a = _pp(mps, already_preprocessed)
    v = np.ones((1,), dtype=np.complex128)
    for i, g in enumerate(a):
        v = v @ g[(bitstring >> i) & 1]
    return complex(v[0])


_PAULI = {"X": np.array([[0, 1], [1, 0]], dtype=np.complex128), "Y": np.array([[0, -1j], [1j, 0]]),
          "Z": np.diag([1.0 + 0j, -1.0]), "I": np.eye(2, dtype=np.complex128)}


def _site_products(env, g_bra, g_ket):
    """P[..., s, t, a, c] = sum_xy env[..., x, y] conj(g_bra[s, x, a]) g_ket[t, y, c]  -- two batched matrix products
    (BLAS zgemm, O(chi^3) per site; a plain three-operand einsum loops over all index tuples, O(chi^4))."""
    T = np.matmul(env[..., None, :, :], g_ket)                                  # [..., t, x, c]
    return np.matmul(g_bra.conj().transpose(0, 2, 1)[:, None], T[..., None, :, :, :])   # [..., s, t, a, c]


def mps_expectation(mps, pauli, qubit, already_preprocessed=False):
    """<psi| P_qubit |psi> (real)."""
    a = _pp(mps, already_preprocessed)
    env = np.ones((1, 1), dtype=np.complex128)
    for i, g in enumerate(a):
        h = np.einsum("st,txy->sxy", _PAULI[pauli], g) if i == qubit else g
        P = _site_products(env, g, h)
        env = P[0, 0] + P[1, 1]
    return float(np.real(env[0, 0]))


def partial_trace(mps, qubits, already_preprocessed=False):
    """4x4 reduced density matrix of the two qubits in `qubits`; the lower-numbered qubit is the
    least-significant matrix index (same convention as the statevector path)."""
    a = _pp(mps, already_preprocessed)
    lo, hi = min(qubits), max(qubits)
    n = len(a)
    env = np.ones((1, 1, 1, 1), dtype=np.complex128)     # [ket idx, bra idx, bra bond, ket bond]
    for i in range(n):
        g = a[i]
        P = _site_products(env, g, g)                      # [k, b, s(bra), t(ket), a, c]
        if i == lo or i == hi:
            # open the physical legs; the new bit becomes the most significant index bit
            K, B = env.shape[0], env.shape[1]
            env = P.transpose(3, 0, 2, 1, 4, 5).reshape(2 * K, 2 * B, P.shape[4], P.shape[5])
        else:
            env = P[:, :, 0, 0] + P[:, :, 1, 1]
    return env[:, :, 0, 0]


def mps_to_vector(mps, already_preprocessed=False):
    """Dense little-endian statevector."""
    a = _pp(mps, already_preprocessed)
    v = np.ones((1, 1), dtype=np.complex128)          # [index, bond]
    for i, g in enumerate(a):
        v = np.einsum("kx,sxy->sky", v, g).reshape(-1, g.shape[2])   # new bit is the most significant so far
    return v[:, 0]


def zero_mps(n):
    return _preprocess_mps(([(np.array([[1.0 + 0j]]), np.array([[0.0 + 0j]])) for _ in range(n)],
                            [np.ones(1) for _ in range(n - 1)]))
