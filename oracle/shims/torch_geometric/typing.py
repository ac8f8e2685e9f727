from typing import Optional, Union
import torch
Adj = Union[torch.Tensor, object]
OptTensor = Optional[torch.Tensor]
