"""Host-side data plumbing the hot path is fed by (reference: dataloaders.py, utilities.py:174-235).
Pure numpy/scipy; kept bit-compatible with the reference's seeded splits (random_seed=123) so the
end-to-end Recall@10 check evaluates exactly the same held-out items.
"""
import math
import os
import pickle

import numpy as np
import torch
from scipy.sparse import coo_matrix, csr_matrix, vstack

DATASETS = ("ml-100k", "ml-1m", "adm", "avg", "ami", "alb")


def split_train_test_proportion_from_csr_matrix(csr_data, test_prop=0.2, batch_size=None, random_seed=None,
                                                ignore_zeros=False):
    """Per-user split of the STORED entries into train / test rows of ones (utilities.py:174-235).
    RNG call order is the reference's: np.random.seed(seed) once, then one np.random.choice per kept user."""
    if random_seed:
        np.random.seed(random_seed)
    if type(csr_data) is not csr_matrix:
        raise TypeError("Input data is not of type csr_matrix")
    if ignore_zeros:
        csr_data.eliminate_zeros()
    n_cols = csr_data.shape[1]
    tr_blocks, te_blocks, tr_rows, te_rows = [], [], [], []

    def flush():
        if tr_rows:
            tr_blocks.append(csr_matrix(np.array(tr_rows)))
            te_blocks.append(csr_matrix(np.array(te_rows)))
            tr_rows.clear()
            te_rows.clear()

    indptr, indices = csr_data.indptr, csr_data.indices
    for u in range(csr_data.shape[0]):
        cols = indices[indptr[u]:indptr[u + 1]]
        n_items = cols.shape[0]
        if n_items < 2:
            print(f"Warning: skipping user with {n_items} items rated")
            continue
        pick = np.zeros(n_items, dtype=bool)
        pick[np.random.choice(n_items, size=math.ceil(test_prop * n_items), replace=False).astype("int32")] = True
        tr = np.zeros(n_cols)
        te = np.zeros(n_cols)
        tr[cols[~pick]] = 1
        te[cols[pick]] = 1
        tr_rows.append(tr)
        te_rows.append(te)
        if batch_size and len(tr_rows) >= batch_size:
            flush()
    flush()
    return vstack(tr_blocks), vstack(te_blocks)


def load_data(dataset_name, data_dir_path="./data"):
    """-> (train_test csr [U, I], train_test + visible part of validation users, validation csr)  (dataloaders.py:82-116)"""
    if dataset_name not in DATASETS:
        raise ValueError("Dataset not found")

    def read(kind):
        path = os.path.normpath(os.path.join(data_dir_path, dataset_name, f"{dataset_name}_{kind}.pkl"))
        with open(path, "rb") as fh:
            return pickle.load(fh)

    train_test, valid = read("train_test"), read("valid")
    val_train, _ = split_train_test_proportion_from_csr_matrix(valid, batch_size=1000, random_seed=123, test_prop=0.2)
    return train_test, vstack((train_test, val_train)), valid


class SparseDataset:
    """Index a scipy sparse matrix by a batch of row ids (dataloaders.py:12-43)."""

    def __init__(self, data, targets, transform=None):
        self.data = data.tocsr() if isinstance(data, coo_matrix) else data
        self.targets = targets.tocsr() if isinstance(targets, coo_matrix) else targets
        self.transform = transform

    def __getitem__(self, index):
        return self.data[index], self.targets[index]

    def __len__(self):
        return self.data.shape[0]

    def get_all_data(self):
        return self.data, self.targets


def _to_device(block):
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    if isinstance(block, csr_matrix) or hasattr(block, "tocoo"):
        coo = block.tocoo()
        idx = torch.from_numpy(np.vstack((coo.row, coo.col)).astype(np.int64))
        val = torch.from_numpy(coo.data.astype(np.float32))
        return torch.sparse_coo_tensor(idx, val, torch.Size(coo.shape)).to(dev)
    return torch.as_tensor(block, dtype=torch.float32).to(dev)


def sparse_batch_collate(batch):
    """BatchSampler hands over one (data_rows, target_rows) pair; both become torch sparse COO tensors on the
    device (dataloaders.py:61-79).  train_SDRM densifies them with .to_dense()."""
    data_batch, targets_batch = batch[0]
    return _to_device(data_batch), _to_device(targets_batch)
