"""Canonical ECORD network: shapes, state_dict ingestion and random initialisation.

The canonical network is the one `experiments/train.py:55-96` configures through
`experiments/configs.py:83-209` (reference paths are relative to /root/reference):

  gnn          GNNEmbedder(dim_in=1, dim_embedding=16, num_layers=4, LN)   models/gnn_embedder.py:152-180
  gnn2rnn      Linear(16->16)                                              models/gnn_to_rnn.py:6-18
  node_encoder Linear(3->16)                                               experiments/configs.py:174-192
  rnn          GRUDecoder(dim_msg_in=34, dim_node_embedding=32,
               dim_internal_embedding=1024, project_rnn_to_gnn=True,
               output_type=MLP_ACTION_STATE)                               models/rnn_decoder.py:22-136,313-381

State-dict key names are the reference's (SURVEY.md 8b "Weight ingestion"), so a DQNSolver
checkpoint (`solver.py:582-617`, `actors/ecord.py:234-252`) loads without renaming.
"""
import math
from collections import OrderedDict

import torch

DIM_GNN = 16        # GNN embedding width
DIM_ENC = 16        # node-feature encoder width
DIM_NODE = 32       # dim_node_embedding = gnn2rnn(16) + encoder(16)
DIM_GLOB = 2        # global features appended to the RNN message (add_glob_to_obs)
DIM_MSG_IN = 34     # DIM_NODE + DIM_GLOB
DIM_MSG = 64        # msg_in output / GRU input
DIM_H = 1024        # GRU hidden (dim_internal_embedding)
DIM_P1 = 512        # proj_rnn_to_gnn middle width
DIM_Q = 64          # mlp_out width (= DIM_NODE + d_state)
GNN_LAYERS = 4      # hard-coded by configs.py:99
NODE_FEATS = 3      # STATE, PEEK_NORM, STEPS_STATIC_CLIP

# name -> shape, in the reference's state_dict order
SHAPES = {
    'gnn': OrderedDict([
        ('embed.weight', (16, 1)), ('embed.bias', (16,)),
        ('node_model.msg_snd.0.weight', (16, 16)), ('node_model.msg_snd.0.bias', (16,)),
        ('node_model.msg_rec.0.weight', (32, 32)), ('node_model.msg_rec.0.bias', (32,)),
        ('node_model.rnn.weight_ih', (48, 32)), ('node_model.rnn.weight_hh', (48, 16)),
        ('node_model.rnn.bias_ih', (48,)), ('node_model.rnn.bias_hh', (48,)),
    ]),
    'gnn2rnn': OrderedDict([('node_projection.0.weight', (16, 16)), ('node_projection.0.bias', (16,))]),
    'node_encoder': OrderedDict([('0.weight', (16, 3)), ('0.bias', (16,))]),
    'rnn': OrderedDict([
        ('proj_rnn_to_gnn.1.weight', (512, 1024)), ('proj_rnn_to_gnn.1.bias', (512,)),
        ('proj_rnn_to_gnn.2.weight', (512,)), ('proj_rnn_to_gnn.2.bias', (512,)),
        ('proj_rnn_to_gnn.4.weight', (32, 512)), ('proj_rnn_to_gnn.4.bias', (32,)),
        ('proj_rnn_to_gnn.5.weight', (32,)), ('proj_rnn_to_gnn.5.bias', (32,)),
        ('val_out.1.weight', (32, 32)), ('val_out.1.bias', (32,)),
        ('val_out.3.weight', (1, 32)), ('val_out.3.bias', (1,)),
        ('mlp_out.0.weight', (64, 64)), ('mlp_out.0.bias', (64,)),
        ('mlp_out.1.weight', (64,)), ('mlp_out.1.bias', (64,)),
        ('mlp_out.3.weight', (1, 64)), ('mlp_out.3.bias', (1,)),
        ('msg_in.0.weight', (64, 34)), ('msg_in.0.bias', (64,)),
        ('graph_rnn.0.weight_ih', (3072, 64)), ('graph_rnn.0.weight_hh', (3072, 1024)),
        ('graph_rnn.0.bias_ih', (3072,)), ('graph_rnn.0.bias_hh', (3072,)),
    ]),
}
_LN_KEYS = {'proj_rnn_to_gnn.2', 'proj_rnn_to_gnn.5', 'mlp_out.1'}
_GRU_HIDDEN = {'node_model.rnn': 16, 'graph_rnn.0': 1024}


def random_state_dicts(seed=0, perturb_norms=False):
    """Random-init weights of the canonical architecture with PyTorch's default distributions
    (Linear: U(+-1/sqrt(fan_in)); GRUCell: U(+-1/sqrt(hidden)); LayerNorm: 1/0), drawn from a
    seeded CPU generator so the same seed gives the same weights on any machine.
    perturb_norms=True jitters the LayerNorm affine parameters (gamma~1+-0.2, beta~+-0.2) so that
    parity tests exercise them (the default 1/0 would hide a swapped gamma/beta)."""
    gen = torch.Generator().manual_seed(int(seed))

    def uni(shape, bound):
        return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound

    out = {}
    for mod, shapes in SHAPES.items():
        sd = OrderedDict()
        for name, shape in shapes.items():
            stem, kind = name.rsplit('.', 1)
            if stem in _LN_KEYS:
                base = torch.ones(shape) if kind == 'weight' else torch.zeros(shape)
                sd[name] = base + (uni(shape, 0.2) if perturb_norms else 0)
            elif stem in _GRU_HIDDEN:
                sd[name] = uni(shape, 1.0 / math.sqrt(_GRU_HIDDEN[stem]))
            else:
                fan_in = SHAPES[mod][stem + '.weight'][1]
                sd[name] = uni(shape, 1.0 / math.sqrt(fan_in))
        out[mod] = sd
    return out


def check_state_dicts(sd):
    """Validate names/shapes; returns fp32 CPU copies.  Extra keys (e.g. target nets) are ignored."""
    out = {}
    for mod, shapes in SHAPES.items():
        if mod not in sd:
            raise KeyError(f"missing state_dict for '{mod}'")
        out[mod] = OrderedDict()
        for name, shape in shapes.items():
            if name not in sd[mod]:
                raise KeyError(f"missing parameter {mod}.{name}")
            t = torch.as_tensor(sd[mod][name]).detach().to('cpu', torch.float32)
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{mod}.{name}: expected shape {shape}, got {tuple(t.shape)} "
                                 f"(only the canonical ECORD network of experiments/train.py:55-96 is supported)")
            out[mod][name] = t.contiguous()
    return out


def state_dicts_from_checkpoint(checkpoint):
    """Extract the four state_dicts from a reference DQNSolver checkpoint dict
    (`solver.py:587-595` wraps `actors/ecord.py:234-252` under 'actor:checkpoint')."""
    ck = checkpoint.get('actor:checkpoint', checkpoint)
    need = {'gnn': 'gnn:state_dict', 'rnn': 'rnn:state_dict', 'gnn2rnn': 'gnn2rnn:state_dict',
            'node_encoder': 'node_encoder:state_dict'}
    sd = {}
    for mod, key in need.items():
        if key not in ck:
            raise KeyError(f"checkpoint has no '{key}' (canonical ECORD network expected)")
        sd[mod] = ck[key]
    return check_state_dicts(sd)


def node_classifier_heads(sd):
    """NodeClassifier (models/node_classifier.py:8-30: `heads.<i>.0` = Linear(16 -> 1), one per probability map) as
    (weight [H,16], bias [H]) fp32 CPU tensors.  Accepts the module, its state_dict, or None (-> None)."""
    if sd is None:
        return None
    if hasattr(sd, "state_dict"):
        sd = sd.state_dict()
    ws, bs = [], []
    while f"heads.{len(ws)}.0.weight" in sd:
        i = len(ws)
        w = torch.as_tensor(sd[f"heads.{i}.0.weight"]).detach().to('cpu', torch.float32)
        b = torch.as_tensor(sd[f"heads.{i}.0.bias"]).detach().to('cpu', torch.float32)
        if tuple(w.shape) != (1, DIM_GNN) or tuple(b.shape) != (1,):
            raise ValueError(f"node_classifier head {i}: expected Linear({DIM_GNN} -> 1)")
        ws.append(w.view(-1)); bs.append(b.view(()))
    if not ws:
        raise KeyError("node_classifier state_dict has no 'heads.0.0.weight'")
    return torch.stack(ws).contiguous(), torch.stack(bs).contiguous()
