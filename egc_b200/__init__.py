"""egc_b200 - B200 (sm_100a) implementation of the EGConv message-passing hot path of shyam196/egc.

Public surface (mirrors the reference operator API, /root/reference/experiments/optimized_layers.py):
    EGConv            drop-in layer
    EfficientGraphConv  the paper variant (reference experiments/layers.py) as an adapter over the same kernels
    SparseTensor      minimal `torch_sparse.SparseTensor` container accepted by EGConv.forward
    GraphStructure    prepared device graph (CSR / CSC / symnorm weights / long-row plan)
    egconv            functional form on a prepared graph
    GraphedStep       capture a (multi-layer) forward + backward step into one CUDA graph and replay it
    EGC               the reference's full-graph layer stack (mag/models.py) on the same kernels
    EGCBlock          conv -> BatchNorm -> ReLU -> dropout -> (+ identity) block (arxiv/norm_models.py); fused epilogue in eval mode
    EgcArxivNet       the reference's normalised full-graph model (arxiv/norm_models.py) on the same kernels
    REGConv           heterogeneous layer of the reference's rmag experiments on the same kernels
    collate / Batch / global_{add,mean,max}_pool   device-side mini-batch collation and graph readout
    build / load      compile / load libegc_b200.so (C ABI in include/egc_b200.h)
"""
from ._lib import (BWD_DETERMINISTIC, GEMM_3XTF32, GEMM_AUTO, GEMM_FP32_SIMT, GEMM_TF32, EGCError, build,  # noqa: F401
                   load)
from .batch import (Batch, collate, collate_arrays, global_add_pool, global_max_pool, global_mean_pool, pad_batch,  # noqa: F401
                    segment_ptr)
from .compat import EfficientGraphConv, convert_paper_state_dict, paper_to_egconv_perm  # noqa: F401
from .conv import EGConv  # noqa: F401
from .dist import GradientAllReduce, GraphedStep, OverlappedGradientAllReduce  # noqa: F401
from .functional import aggregate_combine, egconv, make_desc, project  # noqa: F401
from .hetero import REGConv  # noqa: F401
from .stack import EGC, EGCBlock, EgcArxivNet, fold_batchnorm  # noqa: F401
from .graph import GraphStructure, SparseTensor, to_sparse_tensor  # noqa: F401

__version__ = "0.1.0"
