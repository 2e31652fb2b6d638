"""numpy restatement of the factorised cross-term formula the CUDA path uses, driven
only by reference-produced tables.  TEST INFRASTRUCTURE: an independent, slow check of kernel K2/K3's
mathematics, and the generator of tests/golden/fit_cases.npz.

    X_k[q] = const_k[q] + 2 Re sum_{m=-L..L} w^(m a2) sum_{l>=|m|} At^c[m,l,g1] St^c'[m,l,g2]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

M_PI = 3.14159265358


def cross_terms(idx, A, B, q, zvals, L, tables):
    """A, B: complex [3][Q][(L+1)^2]; tables = (dsymb[l][m+L][l1][p], dwig[b][l][m+L][m1+L], bessel[z][q][p])"""
    dsymb, dw, bes_all = tables
    nb, N, Q = L + 1, 2 * L + 1, len(q)
    lm = lambda l, m: l * (l + 1) + m
    step = 2 * M_PI / N
    w = np.cos(np.arange(N) * step) - 1j * np.sin(np.arange(N) * step)
    W = lambda k: w[k % N]
    ip = np.array([1, 1j, -1, -1j])

    def tmat(zi):
        bes = bes_all[zi]
        T = np.zeros((Q, nb, nb, nb), complex)
        for m in range(nb):
            for l in range(m, nb):
                for l1 in range(l, nb):
                    ps = np.arange(abs(l - l1), l + l1 + 1)
                    v = ((-1) ** m) * (dsymb[l, m + L, l1, ps][None, :] * bes[:, ps] * ip[ps % 4][None, :]).sum(1)
                    T[:, m, l, l1] = v
                    T[:, m, l1, l] = v
        return T

    def rot(coef, b, g, conj):
        out = np.zeros((3, Q, N, nb), complex)
        for m in range(-L, L + 1):
            for l in range(abs(m), nb):
                m1 = np.arange(-l, l + 1)
                c = coef[:, :, lm(l, 0) + m1]
                if conj:
                    c = np.conj(c)
                out[:, :, m + L, l] = (c * (dw[b, l, m + L, m1 + L] * W(m1 * g))[None, None, :]).sum(-1)
        return out

    def translate(T, bt):
        out = np.zeros((3, Q, N, nb), complex)
        for m in range(-L, L + 1):
            am = abs(m)
            out[:, :, m + L, am:] = np.einsum("qab,cqb->cqa", np.conj(T[:, am, am:, am:]), bt[:, :, m + L, am:])
        return out

    pairs = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    cA = lambda c1, c2: (np.conj(A[c1]) * A[c2]).real.sum(-1) + (np.conj(B[c1]) * B[c2]).real.sum(-1)
    const = np.array([cA(0, 0), 2 * cA(0, 1), 2 * cA(0, 2), cA(1, 1), 2 * cA(1, 2), cA(2, 2)])
    X = np.zeros((len(idx), 6, Q))
    Tcache = {}
    for n, i in enumerate(idx):
        i = int(i)
        g2 = i % N; i //= N
        g1 = i % N; i //= N
        a2 = i % N; i //= N
        b2 = i % nb; i //= nb
        b1 = i % nb
        zi = i // nb
        if zi not in Tcache:
            Tcache[zi] = tmat(zi)
        At = rot(A, b1, g1, False)
        St = translate(Tcache[zi], rot(B, b2, g2, True))
        ph = W(np.arange(-L, L + 1) * a2)
        for k, (c1, c2) in enumerate(pairs):
            Cm = (At[c1] * St[c2]).sum(-1)
            if c1 != c2:
                Cm = Cm + (At[c2] * St[c1]).sum(-1)
            X[n, k] = const[k] + 2 * (Cm * ph[None, :]).sum(-1).real
    return X


def reference_tables(L, q, zvals):
    import refso
    nb, N = L + 1, 2 * L + 1
    dsymb = np.zeros((nb, N, nb, N))
    for l in range(nb):
        for l1 in range(nb):
            k2 = np.sqrt((2 * l1 + 1) * (2 * l + 1))
            for p in range(abs(l - l1), l + l1 + 1):
                k3 = (2 * p + 1) * k2 * refso.wigner_3j(l, p, l1, 0, 0, 0)
                for m in range(-min(l, l1), min(l, l1) + 1):
                    dsymb[l, m + L, l1, p] = k3 * refso.wigner_3j(l, p, l1, -m, 0, m)
    dw = np.array([refso.wigner_d(L, i * (M_PI / L)) for i in range(nb)])
    bes = np.array([[[refso.sbessel(p, z * qq) for p in range(N)] for qq in q] for z in zvals])
    return dsymb, dw, bes


def perturbed_family(X, n, seed):
    """seeded smooth multiplicative perturbations of golden cross terms: more fit test cases"""
    rng = np.random.default_rng(seed)
    out = []
    Q = X.shape[-1]
    for _ in range(n):
        i = rng.integers(0, len(X))
        amp = 10 ** rng.uniform(-4, -0.5)
        smooth = 1 + amp * np.cos(np.outer(rng.uniform(0, 6, 6), np.arange(Q) / Q * np.pi) + rng.uniform(0, 6, (6, 1))) \
            * rng.uniform(-1, 1, (6, 1))
        out.append(X[i] * smooth)
    return np.array(out)


if __name__ == "__main__":
    import refso
    G = np.load(os.path.join(HERE, "golden_4g9s.npz"))
    L = int(G["L"])
    q = G["qvals"]
    A = G["rec_coef"][..., 0] + 1j * G["rec_coef"][..., 1]
    B = G["lig_coef"][..., 0] + 1j * G["lig_coef"][..., 1]
    tabs = reference_tables(L, q, [40.0])
    X = cross_terms(G["z40_index"], A, B, q, [40.0], L, tabs)
    fit52 = refso.fit(X, G["a"], G["scal"], q, True)
    assert np.abs(fit52[:, 0] / G["z40_scores"] - 1).max() < 1e-9
    fam = perturbed_family(X, 4000, 1)
    fit_fam = refso.fit(fam, G["a"], G["scal"], q, True)
    np.savez_compressed(os.path.join(HERE, "fit_cases.npz"), X52=X, fit52=fit52, fit_family=fit_fam)
    print("fit_cases.npz written; nfg histogram", np.bincount(fit_fam[:, 3].astype(int))[:20])
