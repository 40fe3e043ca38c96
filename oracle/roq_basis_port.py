"""
TEST INFRASTRUCTURE -- NumPy restatement of the ROQ basis construction the reference carries as COMMENTED-OUT code
(bajes/pipe/utils/roq.py:16-662, "internal bajes implementation of an ROQ algorithm").  **Parity unpinned**: that code
cannot be imported or run (it is a comment), the reference has no test or fixture for it, and the run-time side of ROQ
(the only part that executes, roq.py:808-872) consumes bases built elsewhere.  The functions below follow the commented
lines cited at each of them; they are what ``tests/test_roq_basis*.py`` holds the GPU construction against.
Only tests may import this module.
"""
import numpy as np


HERMITIAN = True     # see inner_product


def inner_product(weights, func_1, func_2, df):
    """roq.py:654-662:  (a, b) = 4 df Re sum_i w_i conj(a_i) b_i.  As commented there it keeps the REAL part only; under a
    real inner product e and i e are orthogonal, the greedy admits both, and the complex DEIM / interpolant (:614-652)
    then work on columns that are linearly dependent over C (``test_real_inner_product_of_the_commented_code_breaks_deim``).
    REPAIR (HERMITIAN = True, what the construction is tested with): the Hermitian product of the cited algorithm
    (Antil et al., arXiv:1210.0577)."""
    s = 4. * df * (weights * np.conj(func_1) * func_2).sum()
    return s if HERMITIAN else np.real(s)


def norm(weights, func, df):
    """roq.py:475-478"""
    return np.sqrt(np.abs(inner_product(weights, func, func, df)))


def normalize(weights, func, df):
    """roq.py:480-483"""
    return func / norm(weights, func, df)


def project(func, basis, weights, df):
    """roq.py:496-503"""
    projection = np.zeros_like(func, dtype=func.dtype)
    for e in basis:
        projection += inner_product(weights, e, func, df) * e
    return projection


def rb_greedy(training_space, weights, df, epsilon, prebuilt_basis=(), max_basis=None):
    """roq.py:505-612 (init_rb_greedy, update_projection, rb_greedy) without the reallocation of well-approximated rows
    (:598-610: an optimisation, rows below epsilon can never be the argmax before the loop ends).
    Returns (basis list, picked row indices, sigma before every pick + the sigma at which the loop ended)."""
    training_space = np.asarray(training_space, dtype=np.complex128)
    if len(prebuilt_basis) == 0:
        e = [normalize(weights, training_space[0], df)]        # :508-511 "arbitrary choice"
        added, sigmas_out = [0], [norm(weights, training_space[0], df)]
    else:
        e = list(prebuilt_basis)
        added, sigmas_out = [], []
    projected = np.zeros_like(training_space)
    for element in e:                                          # :519-520
        for f in range(len(training_space)):                   # update_projection :524-527
            projected[f] += inner_product(weights, element, training_space[f], df) * element
    picks = list(added)
    while True:
        sig = np.array([norm(weights, training_space[i] - projected[i], df) for i in range(len(training_space))])   # :571
        index_max = int(np.argmax(sig))
        sigma = sig[index_max]
        if sigma <= epsilon or index_max in added or (max_basis is not None and len(e) >= max_basis):              # :575-584
            sigmas_out.append(sigma)
            break
        added.append(index_max)
        picks.append(index_max)
        sigmas_out.append(sigma)
        new = normalize(weights, training_space[index_max] - projected[index_max], df)                           # :587-588
        e.append(new)
        for f in range(len(training_space)):                                                                       # :589
            projected[f] += inner_product(weights, new, training_space[f], df) * new
    return e, picks, np.array(sigmas_out)


def deim(V, quad_points):
    """roq.py:614-652, Algorithm 5 of Antil et al.: V [M x m] with independent columns -> (node frequencies, node indices)"""
    V = np.asarray(V)
    M, m = V.shape
    j = int(np.argmax(np.abs(V[:, 0])))
    U = np.zeros((M, m), dtype=np.complex128)
    p_ind = np.zeros(m, dtype=np.int_)
    U[:, 0] = V[:, 0]
    p_ind[0] = j
    for i in range(1, m):
        # P^T selects the rows p_ind[:i]
        c = np.linalg.solve(U[p_ind[:i], :i], V[p_ind[:i], i])
        r = V[:, i] - U[:, :i] @ c
        j = int(np.argmax(np.abs(r)))
        U[:, i] = r
        p_ind[i] = j
    return np.asarray(quad_points)[p_ind], p_ind


def interpolant(V, p_ind):
    """empirical interpolant of the basis at the DEIM nodes: h(f) ~ sum_j B[f, j] h(F_j),  B = V (V[p, :])^-1
    (the ``basis_interpolant_*`` arrays the run-time side loads, roq.py:846-853)"""
    return V @ np.linalg.inv(V[p_ind, :])
