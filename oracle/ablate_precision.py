"""oracle/ablate_precision.py -- which arithmetic of the recurrent update keeps the reference's best cuts?

TEST INFRASTRUCTURE (CPU only, runs in the build container or anywhere torch runs).  Full-length greedy rollouts
(4*|V| steps, 50 trajectories, tau = 0) of the oracle with the GRU / projection arithmetic replaced by emulations of
what a tensor-core kernel could compute, compared graph by graph with the plain fp32 oracle (== the unmodified reference,
see gen_golden.py).  The products of rounded operands are formed in float64 (exact) and accumulated there, i.e. the
emulation shows the effect of OPERAND rounding alone; "fp64" shows the effect of a different fp32 summation order
(any other correct fp32 implementation differs from the reference by about as much).

    python oracle/ablate_precision.py er200 10          # set, graphs
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np
import torch
import torch.nn.functional as F

import ecord_oracle as orc
from ecord_b200.graphs import synthetic_set
from ecord_b200.network import random_state_dicts


def split(x, dt, scale=1.0):
    """x (fp32) -> (hi, lo) in dtype dt, returned as float64 values; x*scale is what gets split."""
    xs = x.double() * scale
    hi = xs.float().to(dt)
    lo = (xs - hi.double()).float().to(dt)
    return hi.double() / scale, lo.double() / scale


def mm_variant(kind):
    """Returns f(x[M,K] fp32, w[N,K] fp32) -> x @ w.T (fp32) under the named operand arithmetic."""
    if kind == "fp32":
        return lambda x, w: F.linear(x, w)
    if kind == "fp64":
        return lambda x, w: (x.double() @ w.double().t()).float()
    if kind == "bf16":
        return lambda x, w: (x.bfloat16().double() @ w.bfloat16().double().t()).float()
    if kind == "tf32":
        def rt(x):   # round to 10 mantissa bits (nearest even)
            i = x.view(torch.int32)
            i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
            return i.view(torch.float32)
        return lambda x, w: (rt(x.contiguous()).double() @ rt(w.contiguous()).double().t()).float()
    if kind in ("bf16x3", "fp16x3", "fp16x3s", "bf16x2a", "fp16x2a"):
        dt = torch.bfloat16 if kind.startswith("bf16") else torch.float16
        sx, sw = (2.0 ** 4, 2.0 ** 12) if kind.endswith("s") else (1.0, 1.0)

        def f(x, w):
            xh, xl = split(x, dt, sx)
            wh, wl = split(w, dt, sw)
            if kind.endswith("2a"):      # activation split only (weights single rounded)
                return ((xh + xl) @ wh.t()).float()
            return (xh @ wh.t() + xl @ wh.t() + xh @ wl.t()).float()
        return f
    raise ValueError(kind)


class AblatedModel(orc.ModelOracle):
    """ModelOracle with the GRU / PROJ GEMMs and the gate functions swapped."""

    def __init__(self, sd, gru="fp32", proj="fp32", tail="fp32", tanh_ulp=0.0, compat=True):
        super().__init__(sd, compat=compat)
        self.mm_gru, self.mm_proj, self.mm_tail = mm_variant(gru), mm_variant(proj), mm_variant(tail)
        self.tanh_ulp = tanh_ulp
        self.gen = torch.Generator().manual_seed(1)

    def _tanh(self, x):
        y = torch.tanh(x)
        if self.tanh_ulp:   # MUFU.TANH: ~2^-11 relative error; emulate by rounding to 11 bits
            i = y.contiguous().view(torch.int32)
            sh = 23 - int(self.tanh_ulp)
            i = (i + (1 << (sh - 1))) & ~((1 << sh) - 1)
            y = i.view(torch.float32)
        return y

    def _sigmoid(self, x):
        if self.tanh_ulp:
            return 0.5 * self._tanh(0.5 * x) + 0.5
        return torch.sigmoid(x)

    @torch.no_grad()
    def update_internal_state(self, glob):
        x = torch.cat([self.last_sel, glob], -1)
        u = orc._lrelu(F.linear(x, self.msg_w, self.msg_b)).view(-1, 64)
        h = self.hidden.view(-1, 1024)
        c = self.cell
        gi = self.mm_gru(u, c.weight_ih) + c.bias_ih
        gh = self.mm_gru(h, c.weight_hh) + c.bias_hh
        i_r, i_z, i_n = gi.chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        r = self._sigmoid(i_r + h_r)
        z = self._sigmoid(i_z + h_z)
        n = self._tanh(i_n + r * h_n)
        hn = (1 - z) * n + z * h
        self.hidden = hn.view(*self.hidden.shape)
        self.last_sel = None

    @torch.no_grad()
    def project(self):
        p = self._tanh(self.hidden)
        G, T = p.shape[:2]
        y = self.mm_proj(p.view(-1, 1024), self.p1_w) + self.p1_b
        y = orc._lrelu(F.layer_norm(y, (512,), self.ln1_w, self.ln1_b))
        y = self.mm_tail(y, self.p2_w) + self.p2_b
        y = orc._lrelu(F.layer_norm(y, (32,), self.ln2_w, self.ln2_b))
        return y.view(G, T, 32)


def load_set(name, count):
    if name == "er500":
        return synthetic_set("er", 500, count, p=0.15, signed=True)
    if name in ("er2000", "er2000s"):      # BASELINE config 4's family: G(2000, avg degree 10), unit ("s": +-1) weights
        return synthetic_set("er", 2000, count, avg_degree=10, signed=name.endswith("s"))
    import pickle
    f = {"er200": "testing/ER_200spin_p15_50graphs.pkl", "ba500": "testing/BA_500spin_m4_50graphs.pkl",
         "er40": "validation/ER_40spin_p15_100graphs.pkl", "er100": "testing/ER_100spin_p15_50graphs.pkl"}[name]
    with open(os.path.join("/root/reference/graphs/ecodqn", f), "rb") as fh:
        return pickle.load(fh)[:count]


VARIANTS = {
    # name: kwargs
    "fp32 (reference arithmetic)": dict(),
    "fp64 GEMMs (summation-order floor)": dict(gru="fp64", proj="fp64", tail="fp64"),
    "GRU fp16x3 scaled, PROJ/tail fp32": dict(gru="fp16x3s"),
    "GRU+PROJ fp16x3 scaled, tail fp32": dict(gru="fp16x3s", proj="fp16x3s"),
    "GRU fp16x3 unscaled": dict(gru="fp16x3"),
    "GRU bf16x3": dict(gru="bf16x3"),
    "GRU+PROJ bf16x3": dict(gru="bf16x3", proj="bf16x3"),
    "GRU fp32, PROJ bf16, tail bf16": dict(proj="bf16", tail="bf16"),
    "GRU fp32, PROJ bf16x3": dict(proj="bf16x3"),
    "GRU tf32": dict(gru="tf32"),
    "GRU bf16 (activations hi/lo, weights bf16)": dict(gru="bf16x2a"),
    "GRU bf16": dict(gru="bf16"),
    "GRU+PROJ+tail bf16": dict(gru="bf16", proj="bf16", tail="bf16"),
    "fp32 GEMMs, 11-bit tanh/sigmoid": dict(tanh_ulp=11),
    "all bf16 + 11-bit tanh (round-1 throughput mode)": dict(gru="bf16", proj="bf16", tail="bf16", tanh_ulp=11),
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "er200"
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    only = sys.argv[3].split(",") if len(sys.argv) > 3 and sys.argv[3] else None
    num_steps = int(sys.argv[4]) if len(sys.argv) > 4 else -4          # default: the full 4 |V| steps
    T = int(sys.argv[5]) if len(sys.argv) > 5 else 50
    torch.set_num_threads(os.cpu_count())
    graphs = load_set(name, count)
    ob = orc.build_batch(graphs)
    init = np.random.RandomState(7).randint(0, 2, (ob.N, T)).astype(np.int8)
    sd = random_state_dicts(0)
    base = None
    print(f"{name}: {count} graphs x {T} trajectories, {num_steps if num_steps > 0 else 4 * int(ob.num_nodes_list.max())} steps, random-init weights seed 0")
    for i, (vn, kw) in enumerate(VARIANTS.items()):
        if only and i and str(i) not in only:
            continue
        t0 = time.time()
        log = orc.rollout(AblatedModel(sd, **kw), ob, T, num_steps, init, tau=0.0, record=(i == 0))
        if i == 0 and 'q' in log:        # how close the reference's own top two Q-values are, relative to the unit's Q spread
            gaps = []
            for q in log['q']:
                for g in range(ob.G):
                    qq = q[int(ob.ptr[g]):int(ob.ptr[g + 1])]
                    top = qq.topk(2, 0).values
                    gaps.append(((top[0] - top[1]) / (qq.max(0).values - qq.min(0).values)).flatten())
            gaps = torch.cat(gaps)
            print("    top-2 gap of the reference's Q-values / Q spread of the unit, over all units and steps: "
                  + "  ".join(f"P(< {t:g}) = {(gaps < t).float().mean().item():.3f}" for t in (1e-6, 1e-5, 1e-4, 1e-3))
                  + f"  exact ties {(gaps == 0).float().mean().item():.3f}", flush=True)
        s = torch.stack(log['scores'], -1)                 # [G,T,S]
        by_t = s.max(-1).values
        smax = by_t.max(-1).values
        acts = torch.stack(log['actions'])                 # [S,G,T]
        if base is None:
            base = (smax, by_t, acts)
        same_traj = (acts == base[2]).all(0).float().mean().item()
        first_div = (acts != base[2]).float().cumsum(0).eq(0).sum(0).float().mean().item()
        print(f"[{i:2d}] {vn:50s} best cut equal {int((smax == base[0]).sum()):3d}/{count}  "
              f"per-traj best equal {(by_t == base[1]).float().mean().item():.3f}  "
              f"whole trajectory identical {same_traj:.3f}  mean steps before divergence {first_div:7.1f}  "
              f"mean best {smax.mean().item():.2f}  ({time.time() - t0:.0f}s)", flush=True)


if __name__ == "__main__":
    main()
