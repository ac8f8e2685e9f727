"""Pure-torch stand-in for torch-scatter==2.0.5 (requirements.txt:7 of the reference), which is
not installed here.  TEST INFRASTRUCTURE: used only by oracle/gen_golden.py to import the
UNMODIFIED reference in this container.  Semantics restated from the published torch-scatter
API (CPU path): reductions run sequentially in index order; arg-max/min ties resolve to the
FIRST (lowest) source index ("parity unpinned": the reference has no test that pins this)."""
import torch


def _bidx(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1 and src.dim() > 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape).expand_as(src)
    return index, dim


def _size(index, dim_size):
    if dim_size is not None:
        return int(dim_size)
    return int(index.max()) + 1 if index.numel() > 0 else 0


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index, dim = _bidx(index, src, dim)
    n = _size(index, dim_size)
    shape = list(src.shape)
    shape[dim] = n
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return res.scatter_add_(dim, index, src)


scatter_add = scatter_sum


def _count(index, dim, n, like):
    ones = torch.ones_like(like)
    shape = list(like.shape)
    shape[dim] = n
    return torch.zeros(shape, dtype=like.dtype, device=like.device).scatter_add_(dim, index, ones)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    index, dim = _bidx(index, src, dim)
    n = _size(index, dim_size)
    s = scatter_sum(src, index, dim, dim_size=n)
    cnt = _count(index, dim, n, src).clamp_(min=1)
    if s.is_floating_point():
        return s / cnt
    return torch.div(s, cnt, rounding_mode='floor')


def scatter_std(src, index, dim=-1, out=None, dim_size=None, unbiased=True):
    index, dim = _bidx(index, src, dim)
    n = _size(index, dim_size)
    cnt = _count(index, dim, n, src)
    mean = scatter_sum(src, index, dim, dim_size=n) / cnt.clamp(min=1)
    var = (src - mean.gather(dim, index))
    var = scatter_sum(var * var, index, dim, dim_size=n)
    denom = (cnt - 1 if unbiased else cnt).clamp(min=1)
    return (var / (denom + 1e-6)).sqrt()


def _scatter_arg(src, index, dim, dim_size, largest):
    index, dim = _bidx(index, src, dim)
    n = _size(index, dim_size)
    shape = list(src.shape)
    shape[dim] = n
    fill = torch.finfo(src.dtype).min if src.is_floating_point() else torch.iinfo(src.dtype).min
    if not largest:
        fill = torch.finfo(src.dtype).max if src.is_floating_point() else torch.iinfo(src.dtype).max
    out = torch.full(shape, fill, dtype=src.dtype, device=src.device)
    out = out.scatter_reduce(dim, index, src, 'amax' if largest else 'amin', include_self=True)
    # arg: the FIRST source position attaining the extremum
    hit = src == out.gather(dim, index)
    pos_shape = [1] * src.dim()
    pos_shape[dim] = -1
    pos = torch.arange(src.size(dim), device=src.device).view(pos_shape).expand_as(src)
    big = src.size(dim)
    cand = torch.where(hit, pos, torch.full_like(pos, big))
    arg = torch.full(shape, big, dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(dim, index, cand, 'amin', include_self=True)
    cnt = _count(index, dim, n, torch.ones_like(src, dtype=torch.float32))
    out = torch.where(cnt > 0, out, torch.zeros_like(out))
    return out, arg


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    return _scatter_arg(src, index, dim, dim_size, True)


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    return _scatter_arg(src, index, dim, dim_size, False)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
    if reduce in ('sum', 'add'):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == 'mean':
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == 'max':
        return scatter_max(src, index, dim, out, dim_size)[0]
    if reduce == 'min':
        return scatter_min(src, index, dim, out, dim_size)[0]
    raise ValueError(reduce)


def segment_csr(src, indptr, out=None, reduce='sum'):
    """Sum over row segments [indptr[i], indptr[i+1]) of dim 0 (1-D indptr), sequential order."""
    assert reduce in ('sum', 'add')
    indptr = indptr.long()
    n = indptr.numel() - 1
    counts = (indptr[1:] - indptr[:-1])
    seg = torch.repeat_interleave(torch.arange(n), counts)
    src_used = src[int(indptr[0]):int(indptr[-1])]
    shape = [n] + list(src.shape[1:])
    res = torch.zeros(shape, dtype=src.dtype)
    return res.index_add_(0, seg, src_used)


def segment_coo(src, index, out=None, dim_size=None, reduce='sum'):
    assert reduce in ('sum', 'add')
    n = _size(index, dim_size)
    shape = [n] + list(src.shape[1:])
    res = torch.zeros(shape, dtype=src.dtype)
    return res.index_add_(0, index.long(), src)
