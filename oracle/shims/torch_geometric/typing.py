"""Type aliases imported at `/root/reference/experiments/optimized_layers.py:12`."""
from typing import Optional, Union

from torch import Tensor
from torch_sparse import SparseTensor

Adj = Union[Tensor, SparseTensor]
OptTensor = Optional[Tensor]
