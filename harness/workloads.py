"""The BASELINE.json configurations as seeded generators: C3 / C5 brickwork targets + thinly dressed
CNOT ansatz layers in brickwall order, C4 random Vidal-form MPS.  Shared by bench.py and the tests so that the parity
tests check exactly the circuits the bench times."""
import numpy as np

from .circuit import Circuit


def brickwork(n, depth, seed):
    """SURVEY 8d C3/C5 target: layer l acts on pairs (i,i+1), i = l mod 2; brick = u3 on both
    qubits then cx(i,i+1); angles uniform(-pi,pi) drawn in (layer, pair, qubit, param) order."""
    rng = np.random.default_rng(seed)
    c = Circuit(n)
    for layer in range(depth):
        for i in range(layer % 2, n - 1, 2):
            for q in (i, i + 1):
                t, p, l = rng.uniform(-np.pi, np.pi, 3)
                c.u3(t, p, l, q)
            c.cx(i, i + 1)
    return c, rng


def brickwall_pairs(n, layers):
    """Pair order of AdaptConfig(method='brickwall') (adapt_compiler.py:803-825)."""
    pairs = []
    for _ in range(layers):
        if not pairs or n == 2:
            pairs.append((0, 1))
            continue
        prev = pairs[-1]
        nxt = (prev[0] + 2, prev[1] + 2)
        n_odd = n % 2
        if nxt == (n, n + 1):
            nxt = (1 - n_odd, 2 - n_odd)
        elif nxt == (n - 1, n):
            nxt = (0 + n_odd, 1 + n_odd)
        pairs.append(nxt)
    return pairs


def thin_ansatz(n, layers, rng):
    """`layers` thinly-dressed-CNOT layers (rz rz cx rz rz, basic.py:135-189) in brickwall order."""
    c = Circuit(n)
    for (a, b) in brickwall_pairs(n, layers):
        th = rng.uniform(-np.pi, np.pi, 4)
        c.rz(th[0], a, label="rz"); c.rz(th[1], b, label="rz")
        c.cx(a, b)
        c.rz(th[2], a, label="rz"); c.rz(th[3], b, label="rz")
    return c


def random_vidal_mps(n, chi, seed):
    """Random canonical Vidal-form MPS (QiskitMPS tuple, constants.py:17) with bond dimensions
    chi_i = min(2^(i+1), 2^(n-1-i), chi) -- SURVEY 8d config C4: complex standard-normal tensors,
    right-canonicalised by QR, then a left-to-right SVD sweep."""
    rng = np.random.default_rng(seed)
    dims = [1] + [int(min(2 ** min(i + 1, n - 1 - i, 30), chi)) for i in range(n - 1)] + [1]
    B = [rng.normal(size=(2, dims[i], dims[i + 1])) + 1j * rng.normal(size=(2, dims[i], dims[i + 1])) for i in range(n)]
    for i in range(n - 1, 0, -1):                      # right-canonicalise
        cl, cr = dims[i], dims[i + 1]
        M = B[i].transpose(1, 0, 2).reshape(cl, 2 * cr)
        Q, R = np.linalg.qr(M.conj().T)                # M^H = Q R  ->  M = R^H Q^H
        B[i] = Q.conj().T.reshape(cl, 2, cr).transpose(1, 0, 2)
        B[i - 1] = np.einsum("sab,bc->sac", B[i - 1], R.conj().T)
    B[0] = B[0] / np.linalg.norm(B[0])
    gammas, lambdas = [], []
    M = B[0]
    prev = np.ones(1)
    for i in range(n - 1):
        cl, cr = dims[i], dims[i + 1]
        U, S, Vh = np.linalg.svd(M.reshape(2 * cl, cr), full_matrices=False)
        A = U.reshape(2, cl, cr)
        g = A / prev.reshape(1, -1, 1)
        gammas.append((g[0].copy(), g[1].copy()))
        lambdas.append(S.copy())
        M = np.einsum("ab,sbc->sac", S[:, None] * Vh, B[i + 1])
        prev = S
    g = M / prev.reshape(1, -1, 1)
    gammas.append((g[0].copy(), g[1].copy()))
    return (gammas, lambdas)


def build_workload(n, depth, layers, seed=1234):
    """C3 / C5: (target circuit, ansatz circuit)."""
    target, rng = brickwork(n, depth, seed)
    return target, thin_ansatz(n, layers, rng)


def compilable_target(n, layers, seed=1234, lo=0.15, hi=0.45):
    """A target ADAPT-AQC actually compiles to the reference's sufficient cost (1e-2) at 28 qubits: a product layer
    ry(theta_q), theta_q in [lo, hi], followed by `layers` thinly dressed CNOT layers in brickwall order.  The global
    cost 1 - |<0|psi>|^2 of a GENERIC 28-qubit state is 1 - O(2^-28): the reference's stopping rule
    (has_stopped_improving on the last 10 layers, adapt_compiler.py:368-393) then ends the run before any progress is
    visible -- measured on the C3 brickwork target and on large-angle product layers.  Small angles keep the initial
    overlap at O(1) (prod cos^2(theta_q / 2) ~ 0.5), so the run CONVERGES and compile wall-time means something."""
    rng = np.random.default_rng(seed)
    c = Circuit(n)
    for q in range(n):
        c.ry(float(rng.uniform(lo, hi)), q)
    c.data.extend(thin_ansatz(n, layers, rng).data)
    return c


def build_mps_workload(n, chi, layers, seed=1):
    """C4: (random Vidal MPS target at bond dimension chi, `layers` un-absorbed thin layers around the middle bond)."""
    target = random_vidal_mps(n, chi, seed)
    rng = np.random.default_rng(seed)
    ansatz = Circuit(n)
    mid = n // 2 - 1
    for k in range(layers):
        a, b = mid + (k % 2), mid + (k % 2) + 1
        th = rng.uniform(-np.pi, np.pi, 4)
        ansatz.rz(th[0], a, label="rz"); ansatz.rz(th[1], b, label="rz")
        ansatz.cx(a, b)
        ansatz.rz(th[2], a, label="rz"); ansatz.rz(th[3], b, label="rz")
    return target, ansatz
