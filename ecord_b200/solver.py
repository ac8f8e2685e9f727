"""B200 drop-in for the reference's solver seam B1.

`B200Solver` keeps the constructor keywords, methods and log contract of
`ecord.solvers.solver.DQNSolver` as used by the test-time path
(`experiments/validate.py:167-202` -> `ecord/validate/testing.py:146-230`):

    solver = B200Solver(gnn=..., rnn=..., gnn2rnn=..., node_encoder=..., node_features=..., global_features=...)
    solver.load("checkpoint.pth"); solver.test()
    log = solver.rollout(batch, num_tradj, num_steps, tau=0, use_epsilon=False, ...)

but `rollout()` runs entirely on the GPU: one GNN pass, one env-init, then S fused steps as CUDA graphs
(engine.RolloutState), with a single device->host copy of the score log at the end.
`gnn`, `rnn`, `gnn2rnn`, `node_encoder` may be torch modules (anything with .state_dict()) or state dicts.
Training-only arguments are accepted and ignored (the engine is inference-only).
"""
import time
from collections import defaultdict

import numpy as np
import torch

from . import _lib
from .data import as_tg_batch
from .engine import DeviceGraph, DeviceWeights, RolloutState
from .network import check_state_dicts, node_classifier_heads, state_dicts_from_checkpoint
from .observations import (CANONICAL_GLOB_FEATURES, CANONICAL_NODE_FEATURES, ActorState, EnvObservation, check_canonical)


def _sd(x):
    return x.state_dict() if hasattr(x, "state_dict") else x


class _ActorShim:
    """`solver.actor.reset_state()` is called between rollouts by test_solver (testing.py:197)."""

    def __init__(self, solver):
        self._solver = solver

    def reset_state(self):
        self._solver._last_state = None

    def test(self):
        pass

    def train(self):
        pass


class B200Solver:
    def __init__(self, gnn, rnn, gnn2rnn, node_encoder=None, node_classifier=None,
                 add_glob_to_nodes=False, add_glob_to_obs=True, detach_gnn_from_rnn=False,
                 node_features=CANONICAL_NODE_FEATURES, global_features=CANONICAL_GLOB_FEATURES,
                 allow_reversible_actions=True, intermediate_reward=0, revisit_reward=0,
                 use_double_q_networks=False, use_gnn_embeddings_from_buffer=False, use_ecodqn=False,
                 default_actor_idx=None, tau=0, alpha=0, munch_lower_log_clip=-1,
                 initial_epsilon=1, final_epsilon=0.05, final_epsilon_step=1000, device=None,
                 # engine options (not in the reference)
                 gemm="bf16x3", compat=True, steps_per_graph=16, seed=0):
        if use_ecodqn or use_double_q_networks:
            raise NotImplementedError("only the single-Q ECORD actor (Actor_Q, actors/ecord.py:26) is implemented")
        if node_encoder is None or add_glob_to_nodes or not add_glob_to_obs:
            raise NotImplementedError("canonical ECORD wiring expected: node_encoder set, add_glob_to_nodes=False, "
                                      "add_glob_to_obs=True (experiments/train.py:88-90)")
        check_canonical(node_features, global_features)
        if not torch.cuda.is_available():
            raise _lib.EcordError("B200Solver needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        self.lib = _lib.load()          # fail loudly, now, if the extension is missing
        self.tau = tau
        self.gemm, self.compat, self.steps_per_graph, self.seed = gemm, compat, steps_per_graph, seed
        self.node_features, self.global_features = node_features, global_features
        self.is_training = False
        self.trainable = False
        self.intermediate_reward, self.revisit_reward = intermediate_reward, revisit_reward
        self.num_env_steps = self.num_rollouts = self.num_param_steps = 0
        self.initial_epsilon, self.final_epsilon, self.final_epsilon_step = initial_epsilon, final_epsilon, final_epsilon_step
        self.set_weights({'gnn': _sd(gnn), 'rnn': _sd(rnn), 'gnn2rnn': _sd(gnn2rnn), 'node_encoder': _sd(node_encoder)})
        self.set_node_classifier(node_classifier)
        self.actor = _ActorShim(self)
        self._last_state = None
        self._graph_cache = (None, None)
        self._subgraph_cache = {}
        self._state_cache = []
        self._net_init_draws = 0

    # ---- weights ------------------------------------------------------------------------------
    def set_weights(self, state_dicts):
        self.state_dicts = check_state_dicts(state_dicts)
        self.weights = DeviceWeights(self.state_dicts, self.device, pack_bf16=True)
        self._state_cache = []          # cached states hold pointers into the old weights

    def set_node_classifier(self, node_classifier):
        """The optional per-vertex head that proposes initial spins (models/node_classifier.py; experiments/configs.py:
        194-206): module, state_dict or None."""
        heads = node_classifier_heads(node_classifier)
        self.node_classifier = None if heads is None else tuple(t.to(self.device) for t in heads)

    def load(self, fname, quiet=False, allow_pickle=False):
        """Ingest a reference checkpoint written by DQNSolver.save (solver.py:582-617).  Such checkpoints hold only dicts,
        OrderedDicts, tensors and numbers, so they are read with torch's restricted unpickler; allow_pickle=True opts into
        the full one for legacy files that carry other objects (it executes code from the file: trusted sources only)."""
        checkpoint = torch.load(fname, map_location="cpu", weights_only=not allow_pickle)
        self.set_weights(state_dicts_from_checkpoint(checkpoint))
        ck = checkpoint.get('actor:checkpoint', checkpoint)
        if 'node_classifier:state_dict' in ck:          # actors/ecord.py:250-251,269
            self.set_node_classifier(ck['node_classifier:state_dict'])
        for k in ("num_env_steps", "num_rollouts", "num_param_steps", "initial_epsilon", "final_epsilon",
                  "final_epsilon_step"):
            if k in checkpoint:
                setattr(self, k, checkpoint[k])
        if not quiet:
            print(f"Loaded DQNSolver checkpoint from {fname} into the B200 engine.")

    def save(self, fname, quiet=False):
        import os
        if os.path.splitext(fname)[-1] != '.pth':
            fname += '.pth'
        actor_ck = {f'{m}:state_dict': sd for m, sd in self.state_dicts.items()}
        if self.node_classifier is not None:            # actors/ecord.py:250-251
            hw, hb = (t.cpu() for t in self.node_classifier)
            actor_ck['node_classifier:state_dict'] = {k: v for i in range(hw.shape[0]) for k, v in
                                                      ((f'heads.{i}.0.weight', hw[i:i + 1].clone()), (f'heads.{i}.0.bias', hb[i:i + 1].clone()))}
        ck = {'actor:checkpoint': actor_ck,
              'num_env_steps': self.num_env_steps, 'num_rollouts': self.num_rollouts,
              'num_param_steps': self.num_param_steps, 'initial_epsilon': self.initial_epsilon,
              'final_epsilon': self.final_epsilon, 'final_epsilon_step': self.final_epsilon_step}
        torch.save(ck, fname)
        return fname

    def train(self):
        """Training-mode ROLLOUTS (data collection for the reference's Trainer, trainer.py:279-289): rollout() then computes
        rewards, keeps the step / rollout counters and feeds the callbacks (solver.py:329-365).  The engine holds no
        optimiser: gradient steps stay with the reference's torch modules, whose weights come back through set_weights()."""
        if self.intermediate_reward != 0 or self.revisit_reward != 0:
            raise NotImplementedError("intermediate / revisit rewards (environment.py:116-139; the ECO-DQN training setting) are "
                                      "not implemented: canonical ECORD trains with both at 0 (experiments/train.py:104-105)")
        self.is_training = True

    def test(self):
        self.is_training = False

    def set_training(self, is_training):
        self.train() if is_training else self.test()

    # ---- rollout (seam B1) ---------------------------------------------------------------------
    def _device_graph(self, batch):
        tg = batch.tg
        if self._graph_cache[0] is tg:
            return self._graph_cache[1]
        dg = DeviceGraph(as_tg_batch(tg), self.device)     # our TgBatch, or the reference batcher's torch_geometric Batch
        self._graph_cache = (tg, dg)
        self._subgraph_cache = {}
        return dg

    def _device_subgraph(self, batch, graph_subset):
        """(DeviceGraph of the sub-batch, its vertices as indices into the full batch: numpy and device)."""
        key = tuple(int(g) for g in graph_subset)
        hit = self._subgraph_cache.get(key)
        if hit is None:
            tg = as_tg_batch(batch.tg)
            vidx = tg.vertex_index(key)
            hit = (DeviceGraph(tg.select(key), self.device), vidx, torch.from_numpy(vidx).to(self.device))
            self._subgraph_cache = {key: hit}
        return hit

    def get_epsilon(self):
        """The reference's linear schedule over parameter steps (solver.py:184-186)."""
        return self.initial_epsilon - (self.initial_epsilon - self.final_epsilon) * min(
            1, self.num_param_steps / self.final_epsilon_step)

    def _rollout_state(self, dg, T, gemm, tau, S, epsilon=0.0, seed=None):
        """Device state for a rollout.  States are kept (the two most recent) and re-initialised when the same graph
        batch is rolled out again with the same shape -- test_solver's trajectory chunks (testing.py:174-189) and
        repeated validation passes -- which saves the buffer allocations and the CUDA-graph capture."""
        seed = self.seed if seed is None else seed
        # a state created for a rollout of up to two graphs' worth of steps runs it as ONE CUDA graph (no remainder graph, one
        # launch); a cached state is reused with whatever graph length it was built for
        spg = S if self.steps_per_graph < S <= 2 * self.steps_per_graph else self.steps_per_graph
        key = (id(dg), T, gemm, self.compat, tau, seed, self.steps_per_graph, epsilon)
        for i, (k, st) in enumerate(self._state_cache):
            if k == key and st.g is dg and st.max_steps >= S:
                self._state_cache.insert(0, self._state_cache.pop(i))
                st.reset()
                return st
        st = RolloutState(dg, self.weights, T, gemm=gemm, compat=self.compat, tau=tau, seed=seed, max_steps=S,
                          log_scores=True, log_actions=True, steps_per_graph=spg, dump_q=False,
                          epsilon=epsilon)
        self._state_cache.insert(0, (key, st))
        for _, old in self._state_cache[2:]:
            old.close()
        del self._state_cache[2:]
        return st

    @torch.no_grad()
    def rollout(self, batch, num_tradj=1, num_steps=-1, log_stats=False, callbacks=None, use_epsilon=True,
                use_network_initialisation=False, tau=None, max_time=None, pre_solve_with_greedy=False,
                post_solve_with_greedy_from_best=False, init_state=None, gemm=None, traj_slice=None, graph_subset=None):
        """Same contract as DQNSolver.rollout (solver.py:196-377) for the test-time path.
        Extra keywords: init_state (int8 [N, num_tradj]) pins the initial spins (the reference draws them
        from numpy's global RNG, tradjectories.py:64 -- reproduced here when init_state is None);
        traj_slice=(lo, hi) runs only trajectories lo..hi of that initial state (multi-GPU trajectory sharding);
        graph_subset=[g0, g1, ...] rolls out only those graphs of the batch, in that order (multi-GPU graph sharding,
        dist.py): the t=0 GNN still runs over the WHOLE batch, because its LayerNorm statistics are batch-wide (quirk B), and
        init_state stays the [N, num_tradj] array of the whole batch; every [G, ...] entry of the log then has one row per
        graph of the subset.
        use_epsilon=True applies the epsilon-greedy exploration of actors/ecord.py:171-180 with epsilon =
        get_epsilon() (solver.py:184-186,334), drawn on the device; validation passes use_epsilon=False."""
        tau = self.tau if tau is None else tau
        tau = 0.0 if tau is None else float(tau)
        log = defaultdict(list)
        t_all = time.time()
        dg = self._device_graph(batch)
        N, G = dg.N, dg.G
        num_graphs = len(batch.info.num_nodes) if batch.info is not None else G
        assert num_graphs == G
        dg_full, sub_vidx = dg, None
        if graph_subset is not None:
            if use_network_initialisation and self.node_classifier is not None and init_state is None:
                raise NotImplementedError("graph_subset with use_network_initialisation: pass init_state")
            dg, sub_vidx_np, sub_vidx = self._device_subgraph(batch, graph_subset)

        if num_steps < 0:
            S = int(abs(num_steps) * int(np.max(dg_full.num_nodes_list)))   # loop runs to the max (solver.py:306-316)
        else:
            S = int(num_steps)

        # use_network_initialisation (solver.py:256-277): spins drawn from the node classifier on the GNN embedding; without
        # such a head the reference falls back to the random default.  An explicit init_state (engine extra) wins.
        from_network = bool(use_network_initialisation) and self.node_classifier is not None and init_state is None
        if from_network:
            T = num_tradj if traj_slice is None else traj_slice[1] - traj_slice[0]
        else:
            if init_state is None:
                init_state = np.random.randint(0, 2, (N, num_tradj), dtype=np.int8)   # tradjectories.py:64
            else:
                if torch.is_tensor(init_state):
                    init_state = init_state.cpu().numpy()
                assert init_state.shape == (N, num_tradj), \
                    f"init_state must have shape {(N, num_tradj)}.  Got {init_state.shape}."
                init_state = np.asarray(init_state, dtype=np.int8)
            if traj_slice is not None:
                init_state = np.ascontiguousarray(init_state[:, traj_slice[0]:traj_slice[1]])
            if sub_vidx is not None:
                init_state = np.ascontiguousarray(init_state[sub_vidx_np])
                N, G = dg.N, dg.G
            T = init_state.shape[1]

        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        # solver.py:334: epsilon = get_epsilon() whenever use_epsilon is set (test_solver passes use_epsilon=False)
        epsilon = float(self.get_epsilon()) if use_epsilon else 0.0
        # random streams (tau sampling, epsilon-greedy, greedy tie-breaks) are functions of (seed, unit, vertex, step): a
        # trajectory shard gets its own seed so that the shards of one rollout do not repeat each other's draws
        seed = self.seed if traj_slice is None else self.seed + 1000003 * int(traj_slice[0])
        state = self._rollout_state(dg, T, gemm or self.gemm, tau, S, epsilon, seed)
        ev[0].record()
        if sub_vidx is None:
            state.gnn_embed()
        else:
            state.gnn_embed_from(dg_full, sub_vidx)
        ev[1].record()
        if from_network:
            hw, hb = self.node_classifier
            head = int(np.random.randint(0, hw.shape[0]))           # solver.py:267: a random readout head
            init_state = state.network_init_state(hw, hb, head, perturb_epsilon=epsilon,
                                                  tradj_offset=0 if traj_slice is None else traj_slice[0],
                                                  draw=self._net_init_draws)
        state.env_init(init_state)
        if pre_solve_with_greedy:                       # solver.py:294-295: greedy_solve(env, init_from_best=False)
            log['greedy_pre_steps'] = [state.greedy_solve(init_from_best=False, rebase=True)]
        init_score_dev = state.scores().clone()
        init_score = init_score_dev
        ev[2].record()

        per_step = defaultdict(list)
        stats_base = state.clone_env() if log_stats else None      # the spins / peeks the rollout starts from
        state.stamp_time_base()
        if self.is_training:
            S = self._training_rollout(state, dg, S, num_steps, callbacks, max_time, t_used0=None)
        elif max_time is None:
            state.run(S)
        else:
            # solver.py:318-320: stop before a step once the accumulated time (embedding + steps) has reached max_time.  The
            # time is the device clock stamped at the end of every step.  While a whole CUDA graph of k steps fits into the
            # remaining budget (at the measured pace) it is launched; close to the limit the loop goes step by step, so the
            # rollout ends at the same step granularity as the reference's.
            ev[1].synchronize()
            t_used = ev[0].elapsed_time(ev[1]) * 1e-3
            done_steps, dt_est = 0, None
            while done_steps < S:
                if t_used >= max_time:
                    print(f"Max time exceeded, terminating optimisation after {done_steps} steps.")
                    break
                k = self.steps_per_graph
                c = min(k, S - done_steps) if (dt_est is not None and t_used + (k + 1) * dt_est < max_time) else 1
                state.run(c)
                dts = state.step_times_host(done_steps + c)     # (synchronises on the stream)
                t_used = ev[0].elapsed_time(ev[1]) * 1e-3 + float(dts.sum())
                done_steps += c
                dt_est = float(dts[-min(len(dts), 8):].mean())
            S = done_steps
        post_scores = None
        if post_solve_with_greedy_from_best:            # solver.py:367-371: greedy from the best states + one more log entry
            log['greedy_post_steps'] = [state.greedy_solve(init_from_best=True)]
            post_scores = state.scores().clone()
        ev[3].record()

        # Results -> host.  Every device-side reduction and every device->host copy is enqueued behind the rollout (the copies
        # into page-locked buffers of torch's caching host allocator, which the caller then owns: no second host copy), and the
        # host waits ONCE.  Round 1 did a synchronous .cpu() per item and cloned the 4 N T bytes of best_state out of a
        # shared staging buffer.
        def to_host(t):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            return h

        h_scores = to_host(state.score_log[:S]) if S > 0 else None
        h_init = to_host(init_score_dev)
        h_times = to_host(torch.cat([state.time_base, state.step_time[:S]])) if S > 0 else None
        h_opt = to_host(state.t_opt_steps(S, init_score_dev)) if S > 0 else None
        h_best = to_host(state.best_from_log(S)) if S > 0 else None
        h_post = to_host(post_scores) if post_scores is not None else None
        h_acts = to_host(state.action_log[:S]) if (log_stats and S > 0) else None
        h_bstate = to_host(state.best_states_device(true_snapshot=not self.compat))
        ev[4].record()
        ev[4].synchronize()
        t_init_emb = ev[0].elapsed_time(ev[1]) * 1e-3
        t_setup = ev[1].elapsed_time(ev[2]) * 1e-3
        t_rollout = ev[2].elapsed_time(ev[3]) * 1e-3

        if log_stats:
            # per-step peeks / spins (before each flip) and actions, reconstructed from the device-side action log after the
            # rollout has run at full speed: see RolloutState.replay_stats
            acts = h_acts.view(S, G, T).long() if S else torch.zeros(0, G, T, dtype=torch.long)
            for s_, pk, stt in state.replay_stats(stats_base, S):
                per_step['peeks'].append(pk)
                per_step['state'].append(stt)
                per_step['actions'].append(acts[s_])
        scores = h_scores.view(S, G, T) if S > 0 else torch.zeros(0, G, T)
        init_score = h_init
        log['t_init_emb'].append(torch.tensor([t_init_emb]))
        log['t_setup'].append(torch.tensor([t_setup]))
        log['init_score'] = init_score
        # per-step times from the device clock; t_opt (solver.py:241-242): accumulated time (embedding + steps) at the last
        # step where a trajectory's score rose to its running best; trajectories that never improve keep t_init_emb
        if S > 0:
            tt = h_times.double()
            dts = (tt[1:] - tt[:-1]) * 1e-9
            cum = t_init_emb + torch.cumsum(dts, 0)
            opt_step = h_opt.long()
            log['t_opt'] = torch.where(opt_step > 0, cum[(opt_step - 1).clamp(min=0)], torch.full_like(cum[:1], t_init_emb)).float()
            rows = list(scores.unbind(0))
            log['scores'] = rows
            log['scores0'] = [init_score] + rows[:-1]
            log['t_step'] = dts.tolist()
            log['t_inf'] = list(log['t_step'])
        else:
            log['t_opt'] = torch.full((G, T), t_init_emb)
        if post_scores is not None:                     # log_step(env, -ones, ...) of solver.py:369-371
            log['scores'].append(h_post)
            if log_stats:
                per_step['actions'].append(-torch.ones(G, T, dtype=torch.long))
        if log_stats:
            for k, v in per_step.items():
                log[k] = v
        if S > 0:       # engine extra: max over the logged steps per trajectory, reduced on the device (testing.py:201 needs it)
            log['best_score_by_tradj'] = h_best
            if post_scores is not None:
                log['best_score_by_tradj'] = torch.maximum(h_best, h_post)
        log['best_state'] = h_bstate
        log['t_rollout'] = torch.tensor([t_rollout])
        log['t_tot'] = torch.tensor([t_init_emb + t_rollout])
        # engine extras are lists / tensors too, so that the reference's merge_log (testing.py:120-143) accepts the log as it is
        log['kernel_launches'] = [state.kernel_launches]
        self._last_state = state
        self._net_init_draws += int(from_network)
        self.last_wall = time.time() - t_all
        return log

    def _training_rollout(self, state, dg, S, num_steps, callbacks, max_time, t_used0=None):
        """The training-time loop of DQNSolver.rollout (solver.py:316-365), step by step: before each step the actor state
        (GRU state, cached embedding) and the observation are snapshotted for the callbacks, after it the reward
            r = max(score_new - best_before, 0) / n_g          (environment.py:113-141 with the extra rewards at 0)
        is formed on the device; `done` marks the graphs whose own step budget (|num_steps| * n_g) has been used up.
        Callbacks are (fn, env_step_freq, param_step_freq) triples called as fn(solver, actor_state=, obs=, action=, reward=,
        done=, t_step=), exactly as Trainer registers them."""
        G, T, N = dg.G, state.T, dg.N
        dev = self.device
        n_nodes = torch.as_tensor(dg.num_nodes_list, dtype=torch.float32, device=dev).view(G, 1)
        per_graph_steps = (torch.as_tensor(dg.num_nodes_list, dtype=torch.long) * abs(int(num_steps))) if num_steps < 0 \
            else torch.full((G,), int(num_steps), dtype=torch.long)
        batch_idx = torch.as_tensor(dg.tg.batch, device=dev) if not torch.is_tensor(dg.tg.batch) else dg.tg.batch.to(dev)
        batch_ptr = torch.as_tensor(dg.tg.np_ptr, device=dev)
        done = torch.zeros(G, T, device=dev)
        want_cb = bool(callbacks)
        init_emb = state.node_emb.unsqueeze(1).expand(N, T, 16) if want_cb else None
        t = 0
        for t in range(S):
            if max_time is not None and t > 0 and float(state.step_times_host(t).sum()) >= max_time:
                print(f"Max time exceeded, terminating optimisation after {t} steps.")
                return t
            actor_state = obs = None
            if want_cb:
                actor_state = ActorState(hidden_embeddings=state.hidden_state().clone(), init_node_embeddings=init_emb,
                                         last_selected_node_embeddings=state.last_sel[:state.M].view(G, T, 32).clone())
                nf, gf, _ = state.export(normalise=True)
                obs = EnvObservation(node_features=nf, glob_features=gf, edge_index=None, edge_attr=None, batch_idx=batch_idx,
                                     batch_ptr=batch_ptr, degree=None, num_tradj=T)
            best_before = state.best_score.view(G, T).clone()
            state.step()
            reward = (state.score.view(G, T) - best_before).clamp(min=0) / n_nodes
            actions = state.actions.view(G, T).long().clone()
            self.num_env_steps += 1
            done[(per_graph_steps == t + 1).to(dev)] = 1
            if want_cb:
                for fn, env_step_freq, param_step_freq in callbacks:
                    if env_step_freq is not None and self.num_env_steps % env_step_freq == 0:
                        fn(self, actor_state=actor_state, obs=obs, action=actions, reward=reward, done=done, t_step=t)
                    if param_step_freq is not None and self.num_param_steps % param_step_freq == 0:
                        fn(self, actor_state=actor_state, obs=obs, action=actions, reward=reward, done=done, t_step=t)
        self.num_rollouts += G * T
        return S

    @staticmethod
    def _flat2batch(dg, arr, fill):
        G, T = dg.G, arr.shape[1]
        out = fill * torch.ones(G, dg.n_max, T)
        p = dg.tg.np_ptr
        for gi in range(G):
            out[gi, :p[gi + 1] - p[gi]] = arr[p[gi]:p[gi + 1]]
        return out


# the name the reference's scripts construct (experiments/validate.py:187-200)
DQNSolver = B200Solver
