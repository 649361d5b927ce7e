"""schpf_b200 -- B200-native CAVI engine behind the scHPF estimator API.

    from schpf_b200 import scHPF
    model = scHPF(nfactors=20).fit(X)          # X: scipy.sparse.coo_matrix, cells x genes
    model.cell_score(); model.gene_score(); model.project(X_new); model.transform(X_new)

The compute path is hand-written sm_100a CUDA loaded through a C ABI
(include/schpf_b200.h); importing the package does not need a GPU, using it does.
"""
__version__ = "0.1.0"

from .scHPF_ import (HPF_Gamma, scHPF, load_model, save_model,   # noqa: E402,F401
                     combine_across_cells)
from .trials import run_trials, run_trials_pool                   # noqa: E402,F401
