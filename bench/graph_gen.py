"""Generate benchmark inputs in the reference's file formats (reference: bench/graph_gen.py).

Same CLI (--seed --num_feats --base_folder --data_name --only_dense --reorder) and the same outputs in the
current directory:
  indices.csv / indptr.csv   one integer per line                        (graph_gen.py:60-61)
  feat.csv                   RAW little-endian fp32, row-major, despite the name (:75-80)
  output_base.csv            raw fp32 cuSPARSE result  A @ B              (:103-121)
  block_offsets.csv          even split of 16x16 output blocks over 114 "threads" (:84-101; only ./tcgnn reads it)
  data.mtx                   MatrixMarket COO                             (:132-142; skipped above --mtx_max_nnz)
What differs: the reference loads `<base_folder>/<data_name>.npz` through TC-GNN's dataset class; there are no
datasets (and no network) here, so `--data_name` selects a synthetic graph of the same (M, nnz) shape from
voltrix.graphs (`reddit`, `ddi`, `products`, `rmat<scale>`, `uniform:<M>:<nnz>` ...).  If
`<base_folder>/<data_name>.npz` does exist it is loaded (keys `src_li`/`dst_li`/`num_nodes`, the TC-GNN layout).
`--seed` is honoured (the reference parses it and ignores it).  `--reorder` relabels the nodes with
voltrix.reorder.cluster_reorder (window-aware agglomerative clustering on the GPU -- the role DTC-SpMM's offline
TCA_reorder.py plays for the reference).
"""
import argparse
import os
import os.path as osp
import sys

import numpy as np
import torch

sys.path.insert(0, osp.abspath(osp.join(osp.dirname(__file__), "..", "voltrix-spmm_b200")))


def load_graph(args, device):
    from voltrix import graphs
    name = args.data_name
    path = osp.join(args.base_folder or ".", name + (".reorder" if args.reorder else "") + ".npz")
    if args.base_folder and osp.exists(path):
        z = np.load(path)
        src, dst, M = z["src_li"], z["dst_li"], int(z["num_nodes"])
        keys = torch.from_numpy((src.astype(np.int64) << 32) | dst.astype(np.int64)).to(device)
        return graphs._csr_from_keys(keys, M, 32)
    if name == "reddit":
        ip, ix = graphs.reddit_shaped(seed=args.seed, device=device)
    elif name == "products":
        ip, ix = graphs.products_shaped(seed=args.seed, device=device)
    elif name.startswith("rmat"):
        ip, ix = graphs.rmat_csr(int(name[4:] or 20), 32, seed=args.seed, device=device)
    elif name.startswith("uniform:"):
        _, M, nnz = name.split(":")
        ip, ix = graphs.uniform_csr(int(M), int(nnz), seed=args.seed, device=device)
    else:
        ip, ix = graphs.suite_graph(name, seed=args.seed, device=device)
    if args.reorder:
        from voltrix import reorder
        ip, ix = reorder.permute_graph(ip, ix, reorder.cluster_reorder(ip, ix, seed=args.seed))
    return ip, ix


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Generate data with specified seed.")
    parser.add_argument("--seed", type=int, default=20, help="Random seed value")
    parser.add_argument("--num_feats", type=int, default=1024, help="Feature dimension")
    parser.add_argument("--base_folder", type=str, default=os.getenv("DATASET_PATH"), help="Base datasets folder")
    parser.add_argument("--data_name", type=str, default="ddi", help="Data name")
    parser.add_argument("--only_dense", action="store_true", help="Only generate dense data")
    parser.add_argument("--reorder", action="store_true", help="Use acc reorder")
    parser.add_argument("--mtx_max_nnz", type=int, default=20_000_000, help="skip data.mtx above this many nnz")
    args = parser.parse_args()
    num_feats = args.num_feats
    assert ".npz" not in args.data_name and ".mtx" not in args.data_name, \
        "Please do not contain suffix .npz or .mtx in the data name"
    assert torch.cuda.is_available(), "graph_gen.py writes the cuSPARSE result as output_base.csv and needs a GPU"

    indptr, indices = load_graph(args, "cuda")
    num_nodes, num_edges = indptr.numel() - 1, indices.numel()
    column_index, row_pointers = indices.cpu().numpy(), indptr.cpu().numpy()
    print("Indices:", column_index)
    print("Indptr:", row_pointers)
    if not args.only_dense:
        np.savetxt("indices.csv", column_index, delimiter=",", fmt="%d")
        np.savetxt("indptr.csv", row_pointers, delimiter=",", fmt="%d")

    B = np.random.default_rng(args.seed).random((num_nodes, num_feats), dtype=np.float32)   # uniform [0,1) like :66

    def save_to_file(data, filename):
        with open(filename, "wb") as out_file:
            out_file.write(np.ascontiguousarray(data, dtype=np.float32).tobytes())

    save_to_file(B, "feat.csv")

    BLK_H, num_thd = 16, 114
    assert num_feats % BLK_H == 0, "num_feats should be multiple of BLK_H"
    padded = num_nodes + (BLK_H - num_nodes % BLK_H) % BLK_H
    base_size = padded * num_feats // BLK_H ** 2 // num_thd
    rem_size = padded * num_feats // BLK_H ** 2 % num_thd
    block_offsets = np.full(num_thd + 1, base_size, dtype=np.int32)
    block_offsets[0] = 0
    block_offsets[-1] = base_size + rem_size
    block_offsets = np.cumsum(block_offsets)
    if not args.only_dense:
        np.savetxt("block_offsets.csv", block_offsets, delimiter=",", fmt="%d")
    print("Block offsets:", block_offsets)

    csr = torch.sparse_csr_tensor(indptr, indices, values=torch.ones(num_edges, device="cuda"),
                                  size=(num_nodes, num_nodes))
    o_base = (csr @ torch.from_numpy(B).cuda()).cpu().numpy()
    save_to_file(o_base, "output_base.csv")

    if not args.only_dense and num_edges <= args.mtx_max_nnz:
        from scipy.io import mmwrite
        from scipy.sparse import csr_matrix
        A = csr_matrix((np.ones(num_edges), column_index, row_pointers), shape=(num_nodes, num_nodes))
        mmwrite("data.mtx", A.tocoo())
        print("Save to data.mtx")
    print("Data generated successfully.")
