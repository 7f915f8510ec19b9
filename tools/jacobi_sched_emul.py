"""Pair schedules of the blocked Jacobi sweeps whose steps split into C independent groups (design aid, not product code).

A sweep must visit every pair of column blocks once.  The circle method couples every pair of step t+1 to its two neighbours of
step t, so groups of pairs on separate streams run in lockstep.  The recursive schedule below only couples groups at a handful of
phase boundaries:  RR(S) = RR(S1) || RR(S2), then BIP(S1, S2);  BIP(A, B) = BIP(A1, B1) || BIP(A2, B2), then BIP(A1, B2) || BIP(A2, B1).
Usage: python tools/jacobi_sched_emul.py n b C [kind]   -- sweeps to convergence with the circle order and with the split order."""
import sys
import numpy as np


def rr_steps(S):
    """round robin inside the block list S: len(S)-1 steps (len(S) steps with a bye when odd)"""
    S = list(S)
    if len(S) < 2:
        return []
    if len(S) % 2:
        S = S + [None]
    n = len(S)
    steps = []
    for st in range(n - 1):
        prs = []
        for k in range(n // 2):
            if n == 2: a, b = 0, 1
            elif k == 0: a, b = n - 1, st
            else: a, b = (st + k) % (n - 1), (st - k + n - 1) % (n - 1)
            if S[a] is not None and S[b] is not None:
                prs.append((min(S[a], S[b]), max(S[a], S[b])))
        steps.append(prs)
    return steps


def bip_steps(A, B):
    A, B = list(A), list(B)
    if len(A) > len(B):
        A, B = B, A
    if not A:
        return []
    return [[(min(A[i], B[(i + s) % len(B)]), max(A[i], B[(i + s) % len(B)])) for i in range(len(A))] for s in range(len(B))]


def expand(task, c):
    """-> list of phases; a phase is a list of concurrent leaf tasks; a leaf task is a list of steps (lists of pairs)"""
    kind = task[0]
    if c <= 1 or (kind == 'rr' and len(task[1]) < 4) or (kind == 'bip' and min(len(task[1]), len(task[2])) < 2):
        steps = rr_steps(task[1]) if kind == 'rr' else bip_steps(task[1], task[2])
        return [[steps]] if steps else []
    def par(x, y):
        out = []
        for i in range(max(len(x), len(y))):
            out.append((x[i] if i < len(x) else []) + (y[i] if i < len(y) else []))
        return out
    if kind == 'rr':
        S = task[1]; h = (len(S) + 1) // 2
        S1, S2 = S[:h], S[h:]
        return par(expand(('rr', S1), c // 2), expand(('rr', S2), c // 2)) + expand(('bip', S1, S2), c)
    A, B = task[1], task[2]
    ha, hb = (len(A) + 1) // 2, (len(B) + 1) // 2
    A1, A2, B1, B2 = A[:ha], A[ha:], B[:hb], B[hb:]
    return par(expand(('bip', A1, B1), c // 2), expand(('bip', A2, B2), c // 2)) + par(expand(('bip', A1, B2), c // 2), expand(('bip', A2, B1), c // 2))


def schedule(nb, c):
    return expand(('rr', list(range(nb))), c)


def check(nb, c):
    ph = schedule(nb, c)
    seen = set()
    depth = 0
    for phase in ph:
        used_phase = []
        for task in phase:
            blocks = set()
            for step in task:
                in_step = set()
                for p, q in step:
                    assert p != q and p not in in_step and q not in in_step
                    in_step |= {p, q}
                    assert (p, q) not in seen
                    seen.add((p, q))
                blocks |= in_step
            used_phase.append(blocks)
        for i in range(len(used_phase)):
            for j in range(i):
                assert not (used_phase[i] & used_phase[j]), "tasks of a phase share a block"
        depth += max(len(t) for t in phase)
    assert len(seen) == nb * (nb - 1) // 2, (len(seen), nb)
    return len(ph), depth


def flat_steps(nb, c):
    """the same schedule merged into synchronous steps (what the convergence emulation needs)"""
    out = []
    for phase in schedule(nb, c):
        for i in range(max(len(t) for t in phase)):
            st = []
            for t in phase:
                if i < len(t): st += t[i]
            out.append(st)
    return out


if __name__ == "__main__":
    for nb in (2, 4, 6, 8, 10, 16, 18, 32, 34, 64, 128, 256):
        for c in (1, 2, 4, 8):
            print(nb, c, check(nb, c))
    if len(sys.argv) > 3:
        sys.path.insert(0, __import__('os').path.dirname(__file__))
        import jacobi_emul as je
        n, b, c = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]); kind = sys.argv[4] if len(sys.argv) > 4 else "graded"
        rng = np.random.default_rng(0)
        if kind == "randn":
            A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        else:
            u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
            v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
            A = (u * np.exp(-np.arange(n) * 30.0 / n)) @ v.conj().T
        # QR preconditioning as the product does: A = Q1 R1, R1^H = Q2 R2, Jacobi on X = R2^H
        R1 = np.linalg.qr(A)[1]; R2 = np.linalg.qr(R1.conj().T)[1]; X = R2.conj().T
        nb = n // b
        for cc in (1, c):
            steps = flat_steps(nb, cc)
            W = X.copy(); tol = 3 * np.sqrt(n) * 2.2e-16
            for sweep in range(30):
                offmax = 0.0
                for st in steps:
                    for p, q in st:
                        cols = np.r_[p * b:(p + 1) * b, q * b:(q + 1) * b]
                        P = W[:, cols]; G = P.conj().T @ P; G = (G + G.conj().T) / 2
                        dg = np.sqrt(np.abs(np.diag(G).real)); R = np.abs(G) / np.maximum(np.outer(dg, dg), 1e-300); np.fill_diagonal(R, 0)
                        offmax = max(offmax, R.max())
                        if R.max() <= tol: continue
                        J, _ = je.evd_jacobi(G, tol, 1)
                        W[:, cols] = P @ J
                print("C =", cc, "sweep", sweep + 1, "offmax %.3e" % offmax, flush=True)
                if offmax <= tol or offmax <= 1e-9: break
            s = np.sort(np.linalg.norm(W, axis=0))[::-1]
            print("C =", cc, "sweeps", sweep + 1, "max sigma err %.2e" % np.max(np.abs(s - np.linalg.svd(A, compute_uv=False))))
