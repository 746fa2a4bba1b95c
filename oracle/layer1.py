"""NumPy restatement of the reference's array math (TEST INFRASTRUCTURE).

Follows ``src/layer1.jl``: ``contract_tensors`` (:85-92, i.e.
``TensorOperations.tensorcontract`` = TTGT: permute to [open|contracted] /
[contracted|open], one BLAS gemm, result axes = A-open then B-open),
``reshape_tensor`` (:100-103), ``permute_tensor`` (:111-114, Julia
``permutedims``), ``tensor_view`` (:191-194).

Julia arrays are column-major; every array here is kept Fortran-contiguous so
that linear layouts coincide with the reference's and ``permutedims(A, p)`` is
``np.transpose(A, p-1)``.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

def _asf(a):
    """Fortran-contiguous view/copy that keeps 0-d arrays 0-d."""
    return np.asarray(a, order="F")



def _prod(xs) -> int:
    out = 1
    for x in xs:
        out *= int(x)
    return out


def _fuse_for_transpose(shape, perm):
    """Merge runs of axes that stay adjacent and in order under ``perm`` so that
    the NumPy copy loop sees a handful of long axes instead of dozens of
    length-2 ones (the same trick Strided.jl plays for the reference)."""
    shape = list(shape)
    groups = []  # runs of source axes, in destination order
    for p in perm:
        if groups and groups[-1][-1] + 1 == p:
            groups[-1].append(p)
        else:
            groups.append([p])
    src_order = sorted(range(len(groups)), key=lambda g: groups[g][0])
    fused_shape = [_prod(shape[a] for a in groups[g]) for g in src_order]
    pos_in_src = {g: i for i, g in enumerate(src_order)}
    fused_perm = [pos_in_src[g] for g in range(len(groups))]
    return fused_shape, fused_perm


def permute_tensor(tensor: np.ndarray, dims: Sequence[int]) -> np.ndarray:
    """``src/layer1.jl:111-114``: ``permutedims(tensor, dims)`` with 1-based
    ``dims``; ``size(out, k) == size(in, dims[k])``.  Out-of-place copy."""
    perm = [int(d) - 1 for d in dims]
    if sorted(perm) != list(range(tensor.ndim)):
        raise ValueError("not a permutation: %r" % (dims,))
    out_shape = tuple(tensor.shape[p] for p in perm)
    if tensor.size == 0 or tensor.ndim <= 1:
        return _asf(tensor.copy())
    fshape, fperm = _fuse_for_transpose(tensor.shape, perm)
    src = np.reshape(tensor, fshape, order="F")
    out = _asf(np.transpose(src, fperm))
    if out is src or np.shares_memory(out, src):
        out = out.copy(order="F")
    return np.reshape(out, out_shape, order="F")


def reshape_tensor(tensor: np.ndarray, dims) -> np.ndarray:
    """``src/layer1.jl:100-103``: column-major reshape (metadata only)."""
    if isinstance(dims, (int, np.integer)):
        dims = [int(dims)]
    return np.reshape(tensor, tuple(int(d) for d in dims), order="F")


def tensor_view(node_data: np.ndarray, bond_idx: int, bond_range: Sequence[int]) -> np.ndarray:
    """``src/layer1.jl:191-194``: copy of the sub-block with axis ``bond_idx``
    (1-based) restricted to ``bond_range`` (1-based values); the axis is kept."""
    idx = [slice(None)] * node_data.ndim
    idx[bond_idx - 1] = [int(i) - 1 for i in bond_range]
    return _asf(node_data[tuple(idx)])


def transpose_tensor(tensor, index_permutation):
    """``src/layer1.jl:122-125``."""
    return permute_tensor(tensor, index_permutation)


def conjugate_tensor(tensor):
    """``src/layer1.jl:132-134``."""
    return np.conj(tensor)


def classify_indices(a_idx: Sequence[int], b_idx: Sequence[int]):
    """Positions (0-based) of open / contracted axes for ``tensorcontract`` with
    output ``symdiff(a_idx, b_idx)``: contracted labels are those present in
    both; B's contracted axes are listed in the order of A's."""
    a_idx = [int(x) for x in a_idx]
    b_idx = [int(x) for x in b_idx]
    if len(set(a_idx)) != len(a_idx) or len(set(b_idx)) != len(b_idx):
        raise ValueError("repeated label inside one tensor (partial trace) is not supported")
    b_pos = {lab: i for i, lab in enumerate(b_idx)}
    a_open = [i for i, lab in enumerate(a_idx) if lab not in b_pos]
    a_con = [i for i, lab in enumerate(a_idx) if lab in b_pos]
    b_con = [b_pos[a_idx[i]] for i in a_con]
    a_set = set(a_idx)
    b_open = [i for i, lab in enumerate(b_idx) if lab not in a_set]
    return a_open, a_con, b_con, b_open


def contract_tensors(tensors_to_contract: Tuple[np.ndarray, np.ndarray],
                     tensor_indices: Tuple[Sequence[int], Sequence[int]]) -> np.ndarray:
    """``src/layer1.jl:85-92``.  C[A-open..., B-open...] = sum over shared labels
    of A*B, computed as TTGT with one gemm.  Column-major C(M,N) = A'(M,K)B'(K,N)
    is evaluated as the C-order product C^T(N,M) = B'^T(N,K) A'^T(K,M) so that no
    extra layout copies are made."""
    A, B = tensors_to_contract
    a_idx, b_idx = tensor_indices
    if A.ndim != len(a_idx) or B.ndim != len(b_idx):
        raise ValueError("index list length does not match tensor rank")
    a_open, a_con, b_con, b_open = classify_indices(a_idx, b_idx)
    for ia, ib in zip(a_con, b_con):
        if A.shape[ia] != B.shape[ib]:
            raise ValueError("DimensionMismatch on contracted axis: %d vs %d"
                             % (A.shape[ia], B.shape[ib]))
    M = _prod(A.shape[i] for i in a_open)
    K = _prod(A.shape[i] for i in a_con)
    N = _prod(B.shape[i] for i in b_open)
    out_shape = tuple(A.shape[i] for i in a_open) + tuple(B.shape[i] for i in b_open)
    dtype = np.result_type(A.dtype, B.dtype)

    Ap = permute_tensor(A, [i + 1 for i in a_open + a_con])      # [open | contracted]
    Bp = permute_tensor(B, [i + 1 for i in b_con + b_open])      # [contracted | open]
    # F-order (M,K) memory == C-order (K,M); F-order (K,N) memory == C-order (N,K)
    At = np.reshape(Ap, (M, K), order="F").T                      # (K, M) C-contiguous
    Bt = np.reshape(Bp, (K, N), order="F").T                      # (N, K) C-contiguous
    Ct = np.matmul(Bt, At).astype(dtype, copy=False)              # (N, M) C-contiguous
    C = Ct.T                                                       # (M, N) F-contiguous
    return np.reshape(C, out_shape, order="F")


def decompose_tensor(tensor: np.ndarray, left_positions: Sequence[int],
                     right_positions: Sequence[int], threshold: float = 1e-13,
                     max_rank: int = 0):
    """``src/layer1.jl:146-184``.  ``permutedims`` to [left | right] (1-based
    positions), reshape to (prod left, prod right), LAPACK SVD, relative
    threshold ``max(threshold, sqrt(eps(real(T))))`` on ``S / norm(S)``,
    optional ``max_rank``; returns ``(U sqrt(S), sqrt(S) V^H, chi)`` reshaped to
    (left..., chi) and (chi, right...)."""
    dims = tensor.shape
    left_dims = [dims[x - 1] for x in left_positions]
    right_dims = [dims[x - 1] for x in right_positions]
    A = permute_tensor(tensor, list(left_positions) + list(right_positions))
    A = np.reshape(A, (_prod(left_dims), _prod(right_dims)), order="F")
    U, S, Vh = np.linalg.svd(A, full_matrices=False)
    real_t = np.finfo(tensor.dtype).eps
    threshold = max(threshold, float(np.sqrt(real_t)))
    s_norm = np.sqrt(np.sum(S.astype(np.float64) ** 2))
    with np.errstate(invalid="ignore", divide="ignore"):
        chi = int(np.sum(S / s_norm > threshold))
    if max_rank > 0:
        chi = min(max_rank, chi)
    s_sqrt = np.sqrt(S[:chi]).astype(S.dtype)
    B = np.reshape(np.asarray(U[:, :chi] * s_sqrt[None, :], order="F"),
                   tuple(left_dims) + (chi,), order="F")
    C = np.reshape(np.asarray(s_sqrt[:, None] * Vh[:chi, :], order="F"),
                   (chi,) + tuple(right_dims), order="F")
    return B.astype(tensor.dtype, copy=False), C.astype(tensor.dtype, copy=False), chi
