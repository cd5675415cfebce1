"""NumPy model of the algebra the CUDA kernels implement (K1 sums -> K2 solve -> K3 combine),
with the same incremental Gram bookkeeping (physical-slot indexing, one column folded per
accepted pair).  Host-logic test aid: lets the CPU suite check the compact form, the ring
order and the pending-column protocol against the oracle's two-loop recursion without a GPU.
Mirrors stochqn_b200/csrc/kernels.cuh (k1_dots / k2_solve / k3_combine) and kernels_adaqn.cuh
(ka1_dots / ka_solve_u / ka2_wdots / ka_solve_a / ka3_combine).
"""
import numpy as np


class CompactModel:
    def __init__(self, m):
        self.m = m
        self.SY = np.zeros((m, m))
        self.YY = np.zeros((m, m))
        self.SS = np.zeros(m)
        self.pending = -1

    def fold(self, S, Y, used):
        c = self.pending
        if c < 0:
            return
        for j in range(used):
            self.SY[j, c] = S[j] @ Y[c]
            self.YY[j, c] = self.YY[c, j] = Y[j] @ Y[c]
        self.SS[c] = S[c] @ S[c]
        self.pending = -1

    def zero_slot(self, c):
        self.SY[:, c] = 0; self.SY[c, :] = 0; self.YY[:, c] = 0; self.YY[c, :] = 0; self.SS[c] = 0

    def _order(self, used, st):
        oldest = 0 if st == used else st
        return [(oldest + i) % self.m for i in range(used)]

    def direction(self, g, S, Y, used, st, h0=0.0):
        """H*g for scalar H0 (oLBFGS / SQN).  Returns (d, U bound)."""
        self.fold(S, Y, used)
        if used == 0:
            return g.copy(), np.sqrt(g @ g)
        ph = self._order(used, st)
        p = np.array([S[s] @ g for s in ph])
        q0 = np.array([Y[s] @ g for s in ph])
        R = np.triu(self.SY[np.ix_(ph, ph)])
        YYl = self.YY[np.ix_(ph, ph)]
        with np.errstate(all="ignore"):
            gamma = h0 if h0 > 0 else self.SY[ph[-1], ph[-1]] / self.YY[ph[-1], ph[-1]]
            u = np.zeros(used)
            for i in range(used - 1, -1, -1):
                u[i] = (p[i] - R[i, i + 1:] @ u[i + 1:]) / R[i, i]
            w = np.diag(R) * u + gamma * (YYl @ u) - gamma * q0
            a = np.zeros(used)
            for i in range(used):
                a[i] = (w[i] - R[:i, i] @ a[:i]) / R[i, i]
            b = -u
            d = gamma * g
            for i, s in enumerate(ph):
                d = d + a[i] * S[s] + (gamma * b[i]) * Y[s]
            U = abs(gamma) * np.sqrt(g @ g) + sum(abs(a[i]) * np.sqrt(self.SS[s]) + abs(gamma * b[i]) * np.sqrt(self.YY[s, s])
                                                 for i, s in enumerate(ph))
        return d, U

    def direction_diag(self, g, h, S, Y, used, st):
        """H*g for H0 = diag(h) (adaQN, quirk Q2: h = g / sqrt(G + eps)).  The weighted Gram matrix
        W = Y' diag(h) Y is never formed: (W u - Y'(h.g)) = Y' [h . (Y u - g)]  (kernels_adaqn.cuh: KA1 -> KAu ->
        KA2 -> KAa -> KA3)."""
        self.fold(S, Y, used)
        if used == 0:
            return h.copy(), np.sqrt(h @ h)
        ph = self._order(used, st)
        p = np.array([S[s] @ g for s in ph])                       # KA1
        R = np.triu(self.SY[np.ix_(ph, ph)])
        with np.errstate(all="ignore"):
            u = np.zeros(used)                                     # KAu
            for i in range(used - 1, -1, -1):
                u[i] = (p[i] - R[i, i + 1:] @ u[i + 1:]) / R[i, i]
            t = -g.copy()                                          # KA2: t = Y u - g, w = Y'(h.t)
            for i, s in enumerate(ph):
                t = t + u[i] * Y[s]
            ht = h * t
            w2 = np.array([Y[s] @ ht for s in ph])
            w = np.diag(R) * u + w2                                # KAa
            a = np.zeros(used)
            for i in range(used):
                a[i] = (w[i] - R[:i, i] @ a[:i]) / R[i, i]
            b = -u
            t3 = g.copy()                                          # KA3
            acc = np.zeros_like(g)
            for i, s in enumerate(ph):
                t3 = t3 + b[i] * Y[s]
                acc = acc + a[i] * S[s]
            d = h * t3 + acc
            U = np.sqrt(ht @ ht) + sum(abs(a[i]) * np.sqrt(self.SS[s]) for i, s in enumerate(ph))
        return d, U
