"""CPU oracle for the Voltrix-SpMM hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package, and only as
the checker / reported baseline.  Nothing under ``voltrix-spmm_b200/`` imports
it; the product path fails loudly without its CUDA extension instead of
falling back to anything in here.

Three layers, all restating or wrapping the reference (paths relative to
/root/reference):

* ``c``   -- ctypes view of ``oracle/_build/libvoltrix_oracle.so``
             (``oracle/voltrix_oracle.c``, plain C restatement).
* ``np_*`` functions -- independent numpy restatement of the same algorithm
             (``voltrix/include/voltrix/bmat_kernels.cuh:264-320, 66-110,
             169-192``), used to cross-check the C port.
* ``ref`` -- ctypes view of ``oracle/_ref/libvoltrix_ref.so``: the UNMODIFIED
             reference sources compiled in place (``oracle/ref_harness.cu``).
             ``ref_preprocess`` runs on the CPU; the other three entry points
             are the reference's CUDA kernels and need a GPU.

Parity status: PINNED -- see the header of ``voltrix_oracle.c``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BLK_H = 16  # voltrix/include/voltrix/traits.h:6
BLK_W = 8   # voltrix/include/voltrix/traits.h:7

_i32p = ctypes.POINTER(ctypes.c_int32)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_f32p = ctypes.POINTER(ctypes.c_float)


def _ptr(a: np.ndarray, ty):
    return a.ctypes.data_as(ty)


def build(force: bool = False) -> None:
    """Compile the C restatement (and oracle/_ref when the reference tree is here)."""
    args = ["make", "-C", _HERE, "all"]
    if force:
        args.insert(1, "-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)


def _load(path: str, rebuild_target: Optional[str]) -> ctypes.CDLL:
    if not os.path.exists(path) and rebuild_target is not None:
        subprocess.check_call(["make", "-C", _HERE, rebuild_target], stdout=subprocess.DEVNULL)
    try:
        return ctypes.CDLL(path)
    except OSError:
        if rebuild_target is None:
            raise
        subprocess.check_call(["make", "-B", "-C", _HERE, rebuild_target], stdout=subprocess.DEVNULL)
        return ctypes.CDLL(path)


class _COracle:
    """ctypes wrapper over voltrix_oracle.c."""

    def __init__(self) -> None:
        self.lib = _load(os.path.join(_HERE, "_build", "libvoltrix_oracle.so"), "oracle")
        L = self.lib
        L.vo_preprocess.restype = ctypes.c_int64
        L.vo_preprocess.argtypes = [_i32p, _i32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                    _i32p, _i32p, _i32p, _i32p]
        L.vo_hmat.restype = None
        L.vo_hmat.argtypes = [_i32p, _i32p, _i32p, _i32p, _i32p, _i32p, ctypes.c_int32,
                              ctypes.c_int32, _f32p, _i32p]
        L.vo_pack_swizzle.restype = None
        L.vo_pack_swizzle.argtypes = [ctypes.c_int64, _f32p, _u32p]
        L.vo_pack_plain.restype = None
        L.vo_pack_plain.argtypes = [ctypes.c_int64, _f32p, _u32p]
        L.vo_spmm_tiles.restype = None
        L.vo_spmm_tiles.argtypes = [_i32p, _u32p, _i32p, ctypes.c_int32, ctypes.c_int32, _f32p,
                                    _f32p, ctypes.c_int, ctypes.c_int]
        L.vo_spmm_csr.restype = None
        L.vo_spmm_csr.argtypes = [_i32p, _i32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                  _f32p, _f32p, ctypes.c_int]
        L.vo_spmm_csr_acc64.restype = None
        L.vo_spmm_csr_acc64.argtypes = L.vo_spmm_csr.argtypes
        L.vo_num_threads.restype = ctypes.c_int
        L.vo_set_num_threads.restype = None
        L.vo_set_num_threads.argtypes = [ctypes.c_int]

    # -- a2: voltrix::preprocess (bmat_kernels.cuh:264-320) ------------------
    def preprocess(self, indptr: np.ndarray, indices: np.ndarray):
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        M = indptr.size - 1
        W = (M + BLK_H - 1) // BLK_H
        nnz = indices.size
        bp = np.zeros(W, np.int32)
        e2c = np.zeros(max(nnz, 1), np.int32)
        e2r = np.zeros(max(nnz, 1), np.int32)
        p1 = np.zeros(W + 1, np.int32)
        self.lib.vo_preprocess(_ptr(indices, _i32p), _ptr(indptr, _i32p), M, BLK_H, BLK_W,
                               _ptr(bp, _i32p), _ptr(e2c, _i32p), _ptr(e2r, _i32p), _ptr(p1, _i32p))
        return bp, e2c[:nnz], e2r[:nnz], p1

    # -- a3: hmat_cuda_kernel (bmat_kernels.cuh:21-111) ----------------------
    def hmat(self, indptr, indices, bp, e2c, e2r, p1, extra_blocks: int = 0, fill: float = 0.0):
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        M = indptr.size - 1
        W = bp.size
        tcb = int(p1[-1]) + extra_blocks
        hspa = np.full(tcb * BLK_H * BLK_W, fill, np.float32)
        hind = np.full(tcb * BLK_W, int(fill), np.int32)
        e2c = np.ascontiguousarray(e2c, np.int32)
        e2r = np.ascontiguousarray(e2r, np.int32)
        if e2c.size == 0:
            e2c = np.zeros(1, np.int32)
            e2r = np.zeros(1, np.int32)
            indices = np.zeros(1, np.int32)
        self.lib.vo_hmat(_ptr(indptr, _i32p), _ptr(indices, _i32p), _ptr(bp, _i32p),
                         _ptr(e2c, _i32p), _ptr(e2r, _i32p), _ptr(p1, _i32p), W, M,
                         _ptr(hspa, _f32p), _ptr(hind, _i32p))
        return hspa, hind

    # -- a4: hmat_convert_uint32_swizzle_cuda_kernel (bmat_kernels.cuh:151-193)
    def pack_swizzle(self, hspa: np.ndarray, total_blocks: int) -> np.ndarray:
        hspa = np.ascontiguousarray(hspa, np.float32)
        out = np.zeros(max(total_blocks * 4, 1), np.uint32)
        self.lib.vo_pack_swizzle(total_blocks, _ptr(hspa, _f32p), _ptr(out, _u32p))
        return out[: total_blocks * 4]

    def pack_plain(self, hspa: np.ndarray, total_blocks: int) -> np.ndarray:
        hspa = np.ascontiguousarray(hspa, np.float32)
        out = np.zeros(max(total_blocks * 4, 1), np.uint32)
        self.lib.vo_pack_plain(total_blocks, _ptr(hspa, _f32p), _ptr(out, _u32p))
        return out[: total_blocks * 4]

    def csr_to_tiles(self, indptr, indices) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """csr_preprocess (voltrix/spmm/spmm.py:16-89): (blk_offsets, hspa_packed, hind)."""
        bp, e2c, e2r, p1 = self.preprocess(indptr, indices)
        hspa, hind = self.hmat(indptr, indices, bp, e2c, e2r, p1)
        packed = self.pack_swizzle(hspa, int(p1[-1]))
        return p1, packed, hind

    # -- a8: spmm_mma161616_spa_swizzle_d/_dd (spmm_kernels.cuh:1458-2001) ---
    def spmm_tiles(self, p1, packed, hind, num_nodes: int, B: np.ndarray, round_tf32: bool = False,
                   all_windows: bool = True) -> np.ndarray:
        B = np.ascontiguousarray(B, np.float32)
        N = B.shape[1]
        out = np.zeros((num_nodes, N), np.float32)
        p1 = np.ascontiguousarray(p1, np.int32)
        packed = np.ascontiguousarray(packed, np.uint32)
        hind = np.ascontiguousarray(hind, np.int32)
        self.lib.vo_spmm_tiles(_ptr(p1, _i32p), _ptr(packed, _u32p), _ptr(hind, _i32p), num_nodes, N,
                               _ptr(B, _f32p), _ptr(out, _f32p), int(round_tf32), int(all_windows))
        return out

    def spmm_csr(self, indptr, indices, B: np.ndarray, row_begin: int = 0, row_end: Optional[int] = None,
                 assume_coalesced: bool = False, out: Optional[np.ndarray] = None, acc64: bool = False) -> np.ndarray:
        """``acc64``: accumulate in double, round once (for rows that sum millions of terms, where a sequential fp32
        sum has a visible rounding error of its own)."""
        indptr = np.ascontiguousarray(indptr, np.int32)
        indices = np.ascontiguousarray(indices, np.int32)
        B = np.ascontiguousarray(B, np.float32)
        N = B.shape[1]
        if row_end is None:
            row_end = indptr.size - 1
        if out is None:
            out = np.empty((row_end - row_begin, N), np.float32)
        if indices.size == 0:
            indices = np.zeros(1, np.int32)
        fn = self.lib.vo_spmm_csr_acc64 if acc64 else self.lib.vo_spmm_csr
        fn(_ptr(indptr, _i32p), _ptr(indices, _i32p), row_begin, row_end, N, _ptr(B, _f32p), _ptr(out, _f32p),
           int(assume_coalesced))
        return out

    def num_threads(self) -> int:
        return int(self.lib.vo_num_threads())

    def set_num_threads(self, n: int) -> None:
        self.lib.vo_set_num_threads(int(n))


class _RefLib:
    """ctypes wrapper over the unmodified reference compiled by oracle/Makefile."""

    path = os.path.join(_HERE, "_ref", "libvoltrix_ref.so")

    def __init__(self) -> None:
        if not os.path.exists(self.path):
            if os.path.isdir(os.environ.get("VOLTRIX_REF", "/root/reference")):
                subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
            else:
                raise FileNotFoundError(
                    f"{self.path} is missing and the reference tree is not mounted; run "
                    "`make -C oracle` where /root/reference exists (the .so travels with gpurun)")
        self.lib = ctypes.CDLL(self.path)
        L = self.lib
        L.ref_preprocess.restype = None
        L.ref_preprocess.argtypes = [_i32p, _i32p, ctypes.c_int, _i32p, _i32p, _i32p, _i32p]
        vp = ctypes.c_void_p
        L.ref_hmat.restype = ctypes.c_int
        L.ref_hmat.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]
        L.ref_hmat_packed_swizzle.restype = ctypes.c_int
        L.ref_hmat_packed_swizzle.argtypes = [ctypes.c_int, vp, vp, vp]
        L.ref_spmm.restype = ctypes.c_int
        L.ref_spmm.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, vp]

    def preprocess(self, indptr: np.ndarray, indices: np.ndarray):
        """voltrix::preprocess itself (CPU).  Prints its TC_Blocks line to stdout like the reference."""
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        M = indptr.size - 1
        W = (M + BLK_H - 1) // BLK_H
        nnz = indices.size
        bp = np.zeros(W, np.int32)
        e2c = np.zeros(max(nnz, 1), np.int32)
        e2r = np.zeros(max(nnz, 1), np.int32)
        p1 = np.zeros(W + 1, np.int32)
        # the reference dereferences neighbor_window[0] even for an empty window (malloc(0)); give
        # it a non-empty edge list so that read stays inside an allocation we own.
        idx = indices if nnz else np.zeros(1, np.int32)
        self.lib.ref_preprocess(_ptr(idx, _i32p), _ptr(indptr, _i32p), M, _ptr(bp, _i32p),
                                _ptr(e2c, _i32p), _ptr(e2r, _i32p), _ptr(p1, _i32p))
        return bp, e2c[:nnz], e2r[:nnz], p1


_c: Optional[_COracle] = None
_ref: Optional[_RefLib] = None


def c() -> _COracle:
    global _c
    if _c is None:
        _c = _COracle()
    return _c


def ref() -> _RefLib:
    global _ref
    if _ref is None:
        _ref = _RefLib()
    return _ref


def have_ref() -> bool:
    return os.path.exists(_RefLib.path) or os.path.isdir(os.environ.get("VOLTRIX_REF", "/root/reference"))


# ----------------------------------------------------------------------------------------------
# numpy restatement (independent of the C port; small/medium sizes)
# ----------------------------------------------------------------------------------------------
def np_preprocess(indptr: np.ndarray, indices: np.ndarray):
    """bmat_kernels.cuh:264-320 with np.unique per 16-row window."""
    indptr = np.asarray(indptr, np.int64)
    indices = np.asarray(indices, np.int64)
    M = indptr.size - 1
    W = (M + BLK_H - 1) // BLK_H
    bp = np.zeros(W, np.int32)
    e2c = np.zeros(indices.size, np.int32)
    e2r = np.repeat(np.arange(M, dtype=np.int32), np.diff(indptr))
    for w in range(W):
        lo, hi = indptr[w * BLK_H], indptr[min(w * BLK_H + BLK_H, M)]
        if hi == lo:
            bp[w] = 1  # edgeless window still owns one all-zero block (bmat_kernels.cuh:250-252)
            continue
        u, inv = np.unique(indices[lo:hi], return_inverse=True)
        bp[w] = (u.size + BLK_W - 1) // BLK_W
        e2c[lo:hi] = inv
    p1 = np.zeros(W + 1, np.int32)
    np.cumsum(bp, out=p1[1:])
    return bp, e2c, e2r, p1


def np_tiles(indptr, indices, bp, e2c, e2r, p1):
    """hmat_cuda_kernel + swizzle pack (bmat_kernels.cuh:66-110, 169-192) -> (hspa, hind, packed)."""
    M = np.asarray(indptr).size - 1
    tcb = int(p1[-1])
    hspa = np.zeros((tcb, BLK_H, BLK_W), np.float32)
    hind = np.zeros((tcb, BLK_W), np.int32)
    if np.asarray(indices).size:
        w = np.asarray(e2r, np.int64) // BLK_H
        b = np.asarray(p1, np.int64)[w] + np.asarray(e2c, np.int64) // BLK_W
        r = np.asarray(e2r, np.int64) % BLK_H
        cc = np.asarray(e2c, np.int64) % BLK_W
        hspa[b, r, cc] = 1.0
        hind[b, cc] = np.asarray(indices, np.int32)
    packed = np.zeros((tcb, 4), np.uint32)
    for idx in range(4):
        for bit in range(32):
            row = (bit >> 2) + 8 * (idx % 2)
            col = (bit % 4) + 4 * (idx // 2)
            packed[:, idx] |= (hspa[:, row, col] != 0).astype(np.uint32) << np.uint32(bit)
    return hspa.reshape(-1), hind.reshape(-1), packed.reshape(-1)


def np_spmm_binary(indptr, indices, B: np.ndarray) -> np.ndarray:
    """fp64-accumulated binary-adjacency SpMM (duplicates count once) via scipy; the fp32 'truth'."""
    import scipy.sparse as sp

    M = np.asarray(indptr).size - 1
    A = sp.csr_matrix((np.ones(np.asarray(indices).size, np.float64), np.asarray(indices), np.asarray(indptr)),
                      shape=(M, B.shape[0]))
    A.sum_duplicates()
    A.data[:] = 1.0
    return np.asarray(A @ B.astype(np.float64))
