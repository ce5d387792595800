"""Downstream consumers of the synthetic rows.  OUT OF SCOPE as kernels: the SVD
evaluator is re-stated here on sklearn because the end-to-end Recall@10 parity check needs it; the Keras MLP
and NeuMF evaluators are loaded lazily from the user's reference checkout if it is importable.
Metrics go through the K3 top-k kernel (sdrm_b200.metrics).
"""
import importlib

import numpy as np

from . import metrics
from .data import split_train_test_proportion_from_csr_matrix

K_LIST = [1, 3, 5, 10, 20, 50]


def compute_mf_results(training_dataset, testing_dataset, synthetic_data=None, nnmf=False, only_synthetic=False):
    """TruncatedSVD(20, n_iter=100) recommender on [train | synthetic] + visible validation part
    (reference: svd_benchmark.compute_mf_results, svd_benchmark.py:17-70; same seeds, same row bookkeeping)."""
    from sklearn.decomposition import NMF, TruncatedSVD
    test_data, valid_data = split_train_test_proportion_from_csr_matrix(testing_dataset, batch_size=1000, random_seed=123)
    synth = None if synthetic_data is None else np.asarray(synthetic_data)
    head = synth if only_synthetic else training_dataset.toarray()
    training_data = np.concatenate([head, test_data.toarray()], axis=0)
    combined = training_data if (only_synthetic and synth is not None) else np.concatenate([training_data, synth], axis=0)
    mf = NMF(n_components=15, max_iter=50) if nnmf else TruncatedSVD(n_components=20, n_iter=100)
    recon = mf.inverse_transform(mf.fit_transform(combined))
    masked = metrics.mask_training_examples(sparse_training_set=training_data, dense_matrix=recon[:training_data.shape[0]].copy())
    lo = head.shape[0]
    block = masked[lo: lo + valid_data.shape[0]]
    both = metrics.recall_ndcg_multi_k(block, valid_data, K_LIST)   # one pass over the scores for all six cut-offs
    recall = [np.round(np.nanmean(both[k][0]), 4) for k in K_LIST]
    ndcg = [np.round(np.nanmean(both[k][1]), 4) for k in K_LIST]
    return np.array(recall), np.array(ndcg)


def _external(module, fn):
    try:
        return getattr(importlib.import_module(module), fn)
    except Exception as exc:  # TensorFlow / the reference checkout is not part of this package
        raise RuntimeError(f"evaluator {module}.{fn} is outside the B200 hot path and is not bundled; put the reference "
                           f"checkout on PYTHONPATH to use it ({exc})")


def compute_mlp_results(*args, **kwargs):
    return _external("mlp_benchmark", "compute_mlp_results")(*args, **kwargs)


def compute_neuralcf_results(*args, **kwargs):
    return _external("neural_cf_benchmark_pt", "compute_neuralcf_results")(*args, **kwargs)
