"""Name-only stand-in for torch-sparse (imported by the reference's unused GatedGraphConv)."""
class SparseTensor:  # never instantiated on the validate path
    pass
def matmul(*a, **k):
    raise NotImplementedError
