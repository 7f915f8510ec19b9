"""Tensor primitives -- restates /root/reference/src/tensors.jl.

Index numbers in this module are 1-based, exactly as in the reference, so the
call sites in the other oracle modules read like the reference's.  Arrays are
NumPy complex128; "first index fastest" (Julia column-major) is reproduced with
``order='F'`` reshapes, so results are element-for-element what Julia returns.
"""
import numpy as np
import scipy.linalg as sla


def _as_list(i):
    return [i] if np.isscalar(i) else list(i)


def contract(x, y, idxs1, idxs2, conjx=False, conjy=False):
    """tensors.jl:9-18.  Pairwise contraction; output indices are the open
    indices of ``x`` (original order) followed by the open indices of ``y``.
    TensorOperations.tensorcontract == TTGT (permute, zgemm, permute), which is
    what ``numpy.tensordot`` does."""
    a1 = [i - 1 for i in _as_list(idxs1)]
    a2 = [i - 1 for i in _as_list(idxs2)]
    if len(a1) != len(a2):
        raise ValueError("The length of contracting indexs differ.")
    xx = np.conj(x) if conjx else x
    yy = np.conj(y) if conjy else y
    return np.tensordot(xx, yy, axes=(a1, a2))


def moveidx(x, currentidx, newidx):
    """tensors.jl:135-160.  Move one index to a new position (always a copy)."""
    nd = x.ndim
    if newidx == -1:
        newidx = nd
    if currentidx == newidx:
        return x.copy()
    return np.ascontiguousarray(np.moveaxis(x, currentidx - 1, newidx - 1))


def combineidxs(x, idxs):
    """tensors.jl:87-103.  Move ``idxs`` to the end (in the order listed) and
    fuse them into one trailing index, first listed index fastest."""
    idxs = list(idxs)
    dims = tuple(x.shape[i - 1] for i in idxs)
    rest = [i for i in range(1, x.ndim + 1) if i not in idxs]
    y = np.transpose(x, [i - 1 for i in rest] + [i - 1 for i in idxs])
    newshape = tuple(x.shape[i - 1] for i in rest) + (int(np.prod(dims)),)
    return np.reshape(y, newshape, order='F'), (idxs, dims)


def uncombineidxs(x, cmb):
    """tensors.jl:111-127.  Inverse of :func:`combineidxs`."""
    idxs, dims = cmb
    offset = x.ndim
    y = np.reshape(x, tuple(x.shape[:offset - 1]) + tuple(dims), order='F')
    for i in range(1, len(idxs) + 1):
        y = np.moveaxis(y, offset - 1 + i - 1, idxs[i - 1] - 1)
    return np.ascontiguousarray(y)


def trace(x, idx1, idx2):
    """tensors.jl:75-78."""
    return np.trace(x, axis1=idx1 - 1, axis2=idx2 - 1)


def truncation_rank(S, cutoff=0.0, maxdim=0, mindim=1):
    """The rank-selection rule of tensors.jl:201-215, reproduced exactly.

    Note tensors.jl:205 (``findfirst(S == 0)``) compares a Vector with a scalar
    and is therefore always ``nothing``: exact zeros are kept unless ``cutoff``
    removes them."""
    n = len(S)
    mindim = min(mindim, n)
    maxdim = n if (maxdim == 0 or maxdim > n) else maxdim
    maxdim = 1 if maxdim == 0 else maxdim
    if cutoff != 0:
        S2 = np.asarray(S, dtype=np.float64) ** 2
        S2cum = np.cumsum(S2[::-1])[::-1] / np.sum(S2)
        above = np.nonzero(S2cum > cutoff)[0]
        keep = 1 if len(above) == 0 else int(above[-1]) + 1
        maxdim = min(maxdim, keep)
    return max(maxdim, mindim)


def svd(x, idx, cutoff=0.0, maxdim=0, mindim=1):
    """tensors.jl:168-227.  Returns ``U, S, V`` with ``S`` a dense k x k
    diagonal matrix and ``V`` = V^H (k x dim(idx)); in ``U`` the new bond takes
    the position of ``idx``."""
    nd = x.ndim
    if idx == -1:
        idx = nd
    rest = [i for i in range(1, nd + 1) if i != idx]
    rest_shape = tuple(x.shape[i - 1] for i in rest)
    y = np.transpose(x, [i - 1 for i in rest] + [idx - 1])
    y = np.reshape(y, (int(np.prod(rest_shape)), x.shape[idx - 1]), order='F')
    try:
        U, S, Vt = sla.svd(y, full_matrices=False, lapack_driver='gesdd')
    except sla.LinAlgError:  # tensors.jl:193-195 fallback
        U, S, Vt = sla.svd(y, full_matrices=False, lapack_driver='gesvd')
    vals = truncation_rank(S, cutoff, maxdim, mindim)
    U = U[:, :vals]
    Sm = np.diag(S[:vals]).astype(np.complex128)
    V = Vt[:vals, :]
    U = np.reshape(U, rest_shape + (vals,), order='F')
    U = np.ascontiguousarray(np.moveaxis(U, -1, idx - 1))
    return U, Sm, V


def tensor_exp(x, outeridxs):
    """tensors.jl:319-338.  Matrix exponential of a tensor viewed as a matrix
    whose rows are the remaining indices and whose columns are ``outeridxs``
    (both fused first-fastest)."""
    outeridxs = _as_list(outeridxs)
    y, cmb1 = combineidxs(x, outeridxs)
    y, cmb2 = combineidxs(y, list(range(1, y.ndim)))
    y = moveidx(y, 2, 1)
    y = sla.expm(y)
    y = moveidx(y, 2, 1)
    y = uncombineidxs(y, cmb2)
    y = uncombineidxs(y, cmb1)
    return y
