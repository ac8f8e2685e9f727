import torch


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_edges(self):
        return 0 if self.edge_index is None else self.edge_index.size(1)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith('_')]

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class Batch(Data):
    """Collation as published for PyG 1.7: node-level tensors concatenated on dim 0, edge_index
    concatenated on dim 1 with cumulative node offsets, `batch` (graph id per node), `ptr`."""
    def __init__(self, batch=None, ptr=None, **kwargs):
        super().__init__(**kwargs)
        self.batch = batch
        self.ptr = ptr

    @staticmethod
    def from_data_list(data_list):
        out = Batch()
        xs, eis, eas, bs, us = [], [], [], [], []
        ptr = [0]
        off = 0
        for i, d in enumerate(data_list):
            n = int(d.num_nodes)
            xs.append(d.x)
            eis.append(d.edge_index + off)
            eas.append(d.edge_attr)
            bs.append(torch.full((n,), i, dtype=torch.long))
            if getattr(d, 'u', None) is not None:
                us.append(d.u)
            off += n
            ptr.append(off)
        out.x = torch.cat(xs, 0)
        out.edge_index = torch.cat(eis, 1)
        out.edge_attr = torch.cat(eas, 0)
        out.batch = torch.cat(bs, 0)
        out.ptr = torch.tensor(ptr, dtype=torch.long)
        out.u = torch.cat(us, 0) if us else None
        out.num_nodes = off
        out._num_graphs = len(data_list)
        return out

    @property
    def num_graphs(self):
        return self._num_graphs
