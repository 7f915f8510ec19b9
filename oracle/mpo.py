"""Automatic MPO construction -- restates
/root/reference/src/structures/mps/mpo.jl:323-472 (``MPO(st, H)`` + ``expand``).

MPO site tensor layout (w_l, out, in, w_r).  The builder starts from the
2x2 "identity ladder", adds on-site terms to the top-right corner, routes every
longer-range term through fresh bond channels, then compresses with a
right-going and a left-going truncated-SVD sweep (cutoff 1e-15)."""
import numpy as np
from .tensors import svd
from .gmps import GMPS


def _expand(O, D1, D2):
    """mpo.jl:463-471."""
    d0, d1, d2, d3 = O.shape
    new = np.zeros((d0 + D1, d1, d2, d3 + D2), dtype=np.complex128)
    k = 1 if d0 == 1 else d0 - 1
    new[:k, :, :, :d3 - 1] = O[:k, :, :, :d3 - 1]
    new[:k, :, :, -1] = O[:k, :, :, -1]
    new[d0 + D1 - 1, :, :, d3 + D2 - 1] = O[d0 - 1, :, :, d3 - 1]
    return new


def MPO(st, H, cutoff=1e-15, maxdim=0, mindim=1):
    N, d = len(H), st.dim
    ident = st.op("id")
    ten = np.zeros((2, d, d, 2), dtype=np.complex128)
    ten[0, :, :, 0] = ident
    ten[1, :, :, 1] = ident
    O = GMPS.zeros(2, d, N)
    O[1] = ten[0:1, :, :, 0:2].copy()
    for i in range(2, N):
        O[i] = ten.copy()
    O[N] = ten[0:2, :, :, 1:2].copy()

    maxrng = H.siterange()
    rngs = [[] for _ in range(maxrng)]
    for i in range(len(H.ops)):
        rngs[H.sites[i][-1] - H.sites[i][0]].append(i)

    for rng in range(1, maxrng + 1):  # mpo.jl:354
        nextterms = [[] for _ in range(rng)]
        coeffs = [[] for _ in range(rng)]
        ingoings = [[] for _ in range(rng)]
        outgoings = [[] for _ in range(rng)]
        for site in range(1, N + 1):
            idxs = [j for j in rngs[rng - 1] if H.sites[j][0] == site]
            if rng == 1:  # mpo.jl:370-374
                for idx in idxs:
                    O[site][0, :, :, -1] += H.coeffs[idx] * st.op(H.ops[idx][0])
                continue
            for idx in idxs:  # mpo.jl:377-403
                ops, sites = H.ops[idx], H.sites[idx]
                outgoing = 0
                for k in range(1, rng + 1):
                    ingoing = outgoing
                    for l in range(1, len(outgoings[k - 1]) + 2):
                        outgoing = l
                        if outgoing not in outgoings[k - 1]:
                            break
                    if k == rng:
                        outgoing = 0
                    s = site + k - 1
                    opname = ops[sites.index(s)] if s in sites else "id"
                    nextterms[k - 1].append(opname)
                    coeffs[k - 1].append(H.coeffs[idx] if k == 1 else 1)
                    ingoings[k - 1].append(ingoing)
                    outgoings[k - 1].append(outgoing)
            terms, ins, outs, cos = nextterms[0], ingoings[0], outgoings[0], coeffs[0]  # mpo.jl:406-419
            nextterms = nextterms[1:] + [[]]
            ingoings = ingoings[1:] + [[]]
            outgoings = outgoings[1:] + [[]]
            coeffs = coeffs[1:] + [[]]
            if terms:  # mpo.jl:422-437
                ingoinglen = sum(1 for x in ins if x != 0)
                outgoinglen = sum(1 for x in outs if x != 0)
                ingoingsrt = O[site].shape[0] - 1
                outgoingsrt = O[site].shape[3] - 1
                O[site] = _expand(O[site], ingoinglen, outgoinglen)
                for j in range(len(terms)):
                    x = 1 if ins[j] == 0 else ingoingsrt + ins[j]
                    y = outgoingsrt + 1 + outgoinglen if outs[j] == 0 else outgoingsrt + outs[j]
                    O[site][x - 1, :, :, y - 1] = cos[j] * st.op(terms[j])

    for site in range(1, N):  # mpo.jl:443-449
        U, S, V = svd(O[site], 4, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
        O[site] = np.tensordot(U, S, axes=([3], [0]))
        O[site + 1] = np.tensordot(V, O[site + 1], axes=([1], [0]))
    for site in range(N, 1, -1):  # mpo.jl:451-457
        U, S, V = svd(O[site], 1, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
        O[site] = np.tensordot(S, U, axes=([1], [0]))
        O[site - 1] = np.tensordot(O[site - 1], V, axes=([3], [1]))
    return O


def adjoint(O):
    """mpo.jl:32-39: Hermitian conjugate of an MPO (swap the physical indices, conjugate)."""
    if O.rank != 2:
        raise ValueError("The generalised MPS must be of rank 2 (an MPO).")
    return GMPS(2, O.dim, [np.conj(np.transpose(t, (0, 2, 1, 3))) for t in O.tensors], O.center)


def trace(*args):
    """mpo.jl:229-252: trace of a product of MPOs, tr(O1 O2 ... On) (the measurement at the end of examples/thermal.jl)."""
    if len(args) < 1:
        raise ValueError("There must be atleast 1 MPO arguments (rank 2).")
    if any(a.dim != args[0].dim for a in args) or any(len(a) != len(args[0]) for a in args):
        raise ValueError("GMPS must share the same physical dim and length.")
    if any(a.rank != 2 for a in args):
        raise ValueError("Arguments must be GMPS of rank 2 (MPO).")
    n = len(args)
    prod = np.ones((1,) * n, dtype=np.complex128)            # one left bond per argument
    for site in range(1, len(args[0]) + 1):
        # prod(b1..bn) A1(b1,o,i,b1') A2(b2,i,i2,b2') ... with the first out index traced against the last in index
        t = np.tensordot(prod, args[0][site], axes=([0], [0]))            # (b2..bn, o, i, b1')
        for j in range(1, n):
            t = np.tensordot(t, args[j][site], axes=([0, t.ndim - 2], [0, 1]))
        # now (o, b1', ..., i_n, bn'): trace o against i_n (tensors.jl:75-78)
        t = np.trace(t, axis1=0, axis2=t.ndim - 2)
        prod = t
    return prod.reshape(-1)[0]


def applyMPO(O, psi, **kw):
    """mpo.jl:105-143 for an MPO acting on an MPS (``O * psi``): exact site products (bonds w * chi, MPO bond fastest), a
    right-going gauge sweep while the sites are built, then movecenter!(phi, 1; kwargs...) with the truncation arguments."""
    if O.rank != 2 or psi.rank != 1:
        raise ValueError("Unallowed combinations of MPS ranks.")
    if O.dim != psi.dim or len(O) != len(psi):
        raise ValueError("GMPS must share the same physical dims and length.")
    N = len(psi)
    phi = GMPS.zeros(1, psi.dim, N)
    for i in range(1, N + 1):
        M, A = O[i], psi[i]
        B = np.einsum('wstx,atb->wasxb', M, A)                       # (w, chi, s, w', chi')
        w, a, s, x, b = B.shape
        phi[i] = np.reshape(B, (w * a, s, x * b), order='F')         # fused bonds, MPO index fastest (mpo.jl:133-135)
        if i > 1:
            phi.moveright(i - 1)
    phi.movecenter(1, **kw)
    return phi
