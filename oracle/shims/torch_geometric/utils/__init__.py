import torch
def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    out = torch.zeros(n, dtype=dtype if dtype is not None else torch.get_default_dtype())
    return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=out.dtype))
def get_laplacian(*a, **k):
    raise NotImplementedError
def to_scipy_sparse_matrix(*a, **k):
    raise NotImplementedError
