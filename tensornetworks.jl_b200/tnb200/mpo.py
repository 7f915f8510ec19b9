"""MPO(st, H) for operator lists of any range (/root/reference/src/structures/mps/mpo.jl:323-459): a finite-state-machine
MPO is assembled on the host (a few d x d blocks per term) and its bonds are compressed on the device by the two truncated-
SVD sweeps of mpo.jl:443-457 (tn_mpo_compress).

The assembly below is this package's own construction, not the reference's channel bookkeeping: every term of range > 1 gets
a private channel on each bond it crosses, so the uncompressed bond dimension is 2 + (number of terms crossing the bond); the
compression sweeps then reduce it (the J1-J2 cylinder of BASELINE.json's config 5 goes from ~140 to ~20).  The operator is
identical to the reference's; the gauge of the compressed tensors need not be."""
import numpy as np

from . import _lib
from ._lib import check
from .api import GMPS, Trunc


def fsm_tensors(N, d, terms):
    """Uncompressed MPO site tensors (w_l, out, in, w_r) of sum_t coeff_t prod_k O_{t,k}(site_{t,k}).
    ``terms``: iterable of (ops, sites, coeff) with ``ops`` d x d matrices and ``sites`` 1-based (any order, no repeats).
    Row 0 / column 0 carry "nothing applied yet", the last row / column "term finished"."""
    ident = np.eye(d, dtype=np.complex128)
    norm_terms = []
    for ops, sites, coeff in terms:
        order = np.argsort(np.asarray(sites))
        st = [int(sites[j]) for j in order]
        if len(set(st)) != len(st) or st[0] < 1 or st[-1] > N:
            raise _lib.TNError("MPO: operator sites must be distinct and lie between 1 and N")
        norm_terms.append(([np.asarray(ops[j], dtype=np.complex128) for j in order], st, complex(coeff)))
    # channel numbers: bond b (between sites b and b+1) carries 0 = start, 1..n_b = private channels, n_b + 1 = done
    nchan = [0] * (N + 1)
    chan = []
    for ops, st, _ in norm_terms:
        c = {}
        for b in range(st[0], st[-1]):
            nchan[b] += 1
            c[b] = nchan[b]
        chan.append(c)

    def wdim(b):                      # bond b in 0..N (0 and N are the closed ends)
        return 1 if b in (0, N) else nchan[b] + 2
    T = [np.zeros((wdim(i - 1), d, d, wdim(i)), dtype=np.complex128) for i in range(1, N + 1)]

    def row(i, x):                    # index of state x on the left bond of site i; x: 'start', 'done' or a channel number
        b = i - 1
        if b == 0:
            return 0
        return 0 if x == 'start' else (nchan[b] + 1 if x == 'done' else x)

    def col(i, x):
        b = i
        if b == N:
            return 0
        return 0 if x == 'start' else (nchan[b] + 1 if x == 'done' else x)
    for i in range(1, N + 1):          # identities that carry "not started" and "finished" along the chain
        if i < N:
            T[i - 1][row(i, 'start'), :, :, col(i, 'start')] = ident
        if i > 1:
            T[i - 1][row(i, 'done'), :, :, col(i, 'done')] = ident
    for (ops, st, coeff), c in zip(norm_terms, chan):
        first, last = st[0], st[-1]
        for q in range(first, last + 1):
            o = ops[st.index(q)] if q in st else ident
            if q == first:
                o = coeff * o
            r = row(q, 'start') if q == first else row(q, c[q - 1])
            cc = col(q, 'done') if q == last else col(q, c[q])
            T[q - 1][r, :, :, cc] += o
    return T


def MPO(N, d, terms, cutoff=1e-15, maxdim=0, mindim=1, ctx=None, compress=True):
    """MPO(st, H; cutoff=1e-15, maxdim=0, mindim=1): mpo.jl:323-459 with H given as ``terms`` = [(ops, sites, coeff)]
    (the flattened OpList: op(st, name) matrices, sites, coefficient).  Returns a device-resident rank-2 GMPS."""
    O = GMPS(2, d, fsm_tensors(N, d, terms), 0, ctx)
    if compress:
        check(O.lib.tn_mpo_compress(O.h, Trunc(cutoff, maxdim, mindim)))
    return O
