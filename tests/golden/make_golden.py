"""Generates tests/golden/*.npz.  Run HERE (the container with /root/reference mounted):

    python tests/golden/make_golden.py

For every case the four preprocessing outputs (block_partition, edge_to_column, edge_to_row, pointer1)
come from the REFERENCE ITSELF -- voltrix::preprocess (voltrix/include/voltrix/bmat_kernels.cuh:264-320),
compiled unmodified into oracle/_ref/libvoltrix_ref.so by oracle/Makefile.  hind / hspa_packed come from the
C restatement of the reference's two CUDA kernels (no GPU in this container); tests/test_ref_gpu.py
re-derives them from the reference kernels on the GPU box.  Large cases store sha256 digests instead of arrays.
"""
import hashlib
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import oracle  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rand_csr(M, density, seed, sort=True, dup=0.0):
    rng = np.random.default_rng(seed)
    A = sp.random(M, M, density=density, format="csr", random_state=rng)
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    rows = np.repeat(np.arange(M), np.diff(indptr))
    cols = indices
    if dup > 0 and cols.size:
        pick = rng.random(cols.size) < dup
        rows = np.concatenate([rows, rows[pick]])
        cols = np.concatenate([cols, cols[pick]])
    key = rng.random(rows.size) if not sort else cols
    order = np.lexsort((key, rows))
    rows, cols = rows[order], cols[order]
    indptr = np.zeros(M + 1, np.int64)
    np.add.at(indptr, rows + 1, 1)
    return np.cumsum(indptr).astype(np.int32), cols.astype(np.int32)


def drop_rows(indptr, indices, r0, r1):
    """empty rows [r0, r1)"""
    lo, hi = indptr[r0], indptr[r1]
    indices = np.concatenate([indices[:lo], indices[hi:]])
    indptr = indptr.copy()
    indptr[r0 + 1:r1] = lo
    indptr[r1:] -= hi - lo
    return indptr, indices


def cases():
    yield "tiny_37_unsorted_dups", rand_csr(37, 0.3, 2, sort=False, dup=0.3), True
    ip, ix = rand_csr(100, 0.05, 1, sort=False)
    yield "m100_empty_window", drop_rows(ip, ix, 16, 32), True
    yield "m64_dense", rand_csr(64, 0.9, 5), True
    yield "m1000_sparse", rand_csr(1000, 0.01, 3), True
    ip, ix = rand_csr(257, 0.02, 7)
    yield "m257_tail1_empty_tail", drop_rows(ip, ix, 250, 257), True
    yield "m48_all_empty", (np.zeros(49, np.int32), np.zeros(0, np.int32)), True
    # the reference's own test generators (tests/test_spmm_kernel.py:166-169, tests/test_spmm.py:16-22):
    np.random.seed(20)
    A = sp.random(8192, 8192, density=0.01, format="csr")
    yield "ref_test_spmm_kernel_seed20_d0.01", (A.indptr.astype(np.int32), A.indices.astype(np.int32)), False
    # BASELINE config C1
    A = sp.random(16384, 16384, density=1e6 / 16384 ** 2, format="csr", random_state=np.random.default_rng(0))
    yield "c1_uniform_16384", (A.indptr.astype(np.int32), A.indices.astype(np.int32)), False


def main():
    ref, c = oracle.ref(), oracle.c()
    for name, (indptr, indices), full in cases():
        bp, e2c, e2r, p1 = ref.preprocess(indptr, indices)            # the reference itself
        cbp, ce2c, ce2r, cp1 = c.preprocess(indptr, indices)          # restatement must agree
        assert all(np.array_equal(a, b) for a, b in ((bp, cbp), (e2c, ce2c), (e2r, ce2r), (p1, cp1))), name
        hspa, hind = c.hmat(indptr, indices, bp, e2c, e2r, p1)
        packed = c.pack_swizzle(hspa, int(p1[-1]))
        out = dict(indptr=indptr, indices=indices, source=np.array("reference voltrix::preprocess via oracle/_ref"))
        arrays = dict(block_partition=bp, edge_to_column=e2c, edge_to_row=e2r, pointer1=p1, hind=hind,
                      hspa_packed=packed)
        if full:
            out.update(arrays)
        else:
            out.pop("indices")
            out.pop("indptr")
            out.update({k + "_sha256": np.array(digest(v)) for k, v in arrays.items()})
            out.update(indptr_sha256=np.array(digest(indptr)), indices_sha256=np.array(digest(indices)),
                       pointer1=p1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: M={indptr.size - 1} nnz={indices.size} TCB={int(p1[-1])} full={full}")


if __name__ == "__main__":
    main()
