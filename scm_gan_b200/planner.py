"""Model-predictive control on the learned world model: host-side mirror of the reference's `play()` /
`compute_rollout_reward()` (main.py:327-400, 455-489) on the drop-in modules.

The decision rule is the reference's: for every candidate action a, advance the latent state one step with a, then
score it with the best of a beam of A^lookahead plans (every [i, j] action pair followed by a fixed roll-out policy)
simulated `rollout_depth` steps with Transition + RewardPredictor; take argmax over a.  Nothing is decoded.

Everything stays on the device; the only device->host transfer per decision is the vector of A scores (the reference
reads `max()`/`argmax` of python floats, i.e. one transfer per candidate).  Call order and batch shapes follow the
reference exactly by default, because every Transition call advances the spectral-norm power iteration
(spectral_normalization.py:28-31) and, in train mode, draws Bernoulli noise.  `fold_actions=True` simulates the A
candidates as one batch of A * A^lookahead trajectories (A x fewer Transition calls of A x the batch - the shape the
tcgen05 kernels like).  It is the same computation: the power iterations of all 13 A calls of the sequential order
run ahead in one launch (Transition.power_iterations) and every candidate's segment of the batch is normalised with
the sigma of ITS call, so scores agree with the sequential planner to rounding and u, v end in the same state.
"""
import numpy as np
import torch


def onehot(a_idx, num_actions, device):
    """reference main.py:447-452: int -> [1, A]; LongTensor [B] -> [B, A]."""
    eye = torch.eye(num_actions, dtype=torch.float32, device=device)
    if isinstance(a_idx, int):
        return eye[a_idx].unsqueeze(0)
    return eye[a_idx]


def beam_actions(num_actions, lookahead=2, rollout_depth=12, rollout_policy="noop", rng=None):
    """[A^lookahead, rollout_depth] int64 plan table of compute_rollout_reward (main.py:463-473; lookahead is 2 there)."""
    assert lookahead == 2, "the reference enumerates action pairs"
    rows = []
    for i in range(num_actions):
        for j in range(num_actions):
            if rollout_policy == "noop":
                tail = [0] * (rollout_depth - lookahead)
            elif rollout_policy == "random":
                rng = rng or np.random
                tail = [int(rng.randint(num_actions)) for _ in range(rollout_depth - lookahead)]
            else:
                raise ValueError(rollout_policy)
            rows.append([i, j] + tail)
    return torch.as_tensor(np.asarray(rows, dtype=np.int64))


@torch.no_grad()
def rollout_scores(z, transition, reward_predictor, num_actions, lookahead=2, rollout_depth=12,
                   rollout_policy="noop", negative_positive_tradeoff=10.0, actions=None, sigma_for_step=None):
    """Score of every plan of the beam started at each row of z: [Bz * A^lookahead] (plans of one start state are
    contiguous).  With Bz = 1 this is the `cumulative_reward.sum(dim=1)` of main.py:476-486.
    sigma_for_step(t) (optional) -> the `sigma=` rows of Transition.forward for rollout step t (one row per start
    state: the folded planner)."""
    width = num_actions ** lookahead
    if actions is None:
        actions = beam_actions(num_actions, lookahead, rollout_depth, rollout_policy)
    actions = actions.to(z.device)
    bz = z.shape[0]
    z = z.repeat_interleave(width, dim=0) if bz > 1 else z.repeat(width, 1, 1, 1)
    plan = actions.repeat(bz, 1)
    cumulative = reward_predictor(z).clone()
    for t in range(rollout_depth):
        z = transition(z.detach(), onehot(plan[:, t], num_actions, z.device),
                       sigma=None if sigma_for_step is None else sigma_for_step(t))
        cumulative += reward_predictor(z)
    cumulative[:, 0] *= negative_positive_tradeoff  # "caution" about the first (negative) reward channel
    return cumulative.sum(dim=1)


@torch.no_grad()
def compute_rollout_reward(z, transition, reward_predictor, num_actions, selected_action=None, lookahead=2,
                           rollout_depth=12, rollout_policy="noop", negative_positive_tradeoff=10.0):
    """Reference signature (main.py:455-489): best achievable plan score from latent state z [1, L, H, W]
    (a 0-dim device tensor)."""
    return rollout_scores(z, transition, reward_predictor, num_actions, lookahead, rollout_depth, rollout_policy,
                          negative_positive_tradeoff).max(dim=0)[0]


@torch.no_grad()
def choose_action(z, transition, reward_predictor, num_actions, rollout_depth=12, rollout_policy="noop",
                  fold_actions=False):
    """One decision of play() (main.py:356-368).  Returns (best action, [score of every action] as a CPU tensor)."""
    dev = z.device
    if fold_actions:
        # sequential order: for every action a: 1 call on z, then rollout_depth calls on its beam -> call index
        # a * (1 + rollout_depth) + k.  Folded call k serves all actions, one segment each.
        per = 1 + rollout_depth
        sig = transition.power_iterations(num_actions * per)

        def rows(k):
            return torch.stack([sig[a * per + k] for a in range(num_actions)])
        z_all = transition(z.repeat(num_actions, 1, 1, 1),
                           onehot(torch.arange(num_actions, device=dev), num_actions, dev), sigma=rows(0))
        scores = rollout_scores(z_all, transition, reward_predictor, num_actions, 2, rollout_depth, rollout_policy,
                                sigma_for_step=lambda t: rows(1 + t))
        rewards = scores.view(num_actions, -1).max(dim=1)[0]
    else:
        per_action = []
        for a in range(num_actions):
            z_a = transition(z, onehot(a, num_actions, dev))
            per_action.append(compute_rollout_reward(z_a, transition, reward_predictor, num_actions, a,
                                                     rollout_depth=rollout_depth, rollout_policy=rollout_policy))
        rewards = torch.stack(per_action)
    rewards = rewards.cpu()  # the decision itself is taken on the host, like the reference's np.argmax
    return int(torch.argmax(rewards)), rewards


@torch.no_grad()
def play(env, convert_frame, nets, num_actions, no_op=3, max_steps=300, fold_actions=False, on_step=None):
    """The agent loop of main.py:327-400 without the video/file output: env must offer reset() -> state and
    step(a) -> (state, reward, done, info); convert_frame(state) -> (network frame [C,H,W], rgb frame).
    Returns (cumulative reward, list of chosen actions)."""
    enc, tr, rew = nets["encoder"], nets["transition"], nets["reward_predictor"]
    dev = next(enc.parameters()).device
    no_op = min(no_op, num_actions - 1)

    def frames_to_z(frames, action):
        x = torch.as_tensor(np.asarray(frames, dtype=np.float32), device=dev).unsqueeze(0)
        return tr(enc(x), onehot(action, num_actions, dev))

    state = env.reset()
    frames = [convert_frame(state)[0]]
    done = False
    for _ in range(2):  # no-op through the first frames for the initial state estimate (main.py:333-345)
        state, _, done, _ = env.step(no_op)
        frames.append(convert_frame(state)[0])
    z = frames_to_z(frames, no_op)
    total, chosen, t = 0.0, [], 2
    while not done:
        a, scores = choose_action(z.detach(), tr, rew, num_actions, fold_actions=fold_actions)
        state, r, done, info = env.step(a)
        total += float(r)
        chosen.append(a)
        if on_step is not None:
            on_step(t, a, scores, r, info)
        frames = frames[1:] + [convert_frame(state)[0]]
        z = frames_to_z(frames, a)  # re-estimate the state from the real frames (main.py:389-391)
        t += 1
        if t > max_steps:
            break
    return total, chosen


class GraphedPlanner:
    """One decision of play() (main.py:356-368) - and the state re-estimate that precedes it (main.py:389-391) - as ONE
    CUDA graph: encoder + Transition on the three latest frames, then the folded beam (1 + rollout_depth Transition calls
    and 1 + rollout_depth RewardPredictor calls at batch A * A^2), the "caution" weighting, the max over plans and the
    argmax over actions, all on the device.  Per decision the host uploads three frames and one action index and reads
    back A + 1 floats (the scores and the chosen action).

    Spectral-norm bookkeeping is the reference's: 1 + A * (1 + rollout_depth) Transition calls per decision advance
    u, v that many times (one multi-iteration launch each for the re-estimate and the beam).
    Bernoulli noise (train-mode modules, the way main.py's play() runs them) comes from the Transition's device-side
    Philox state, so replays draw fresh noise.
    """

    def __init__(self, nets, num_actions, rollout_depth=12, rollout_policy="noop", negative_positive_tradeoff=10.0):
        assert rollout_policy == "noop", "the captured plan table is fixed: only the deterministic roll-out policy"
        self.nets = nets
        self.A = num_actions
        self.depth = rollout_depth
        self.tradeoff = negative_positive_tradeoff
        self._graph = None
        self._plan = None   # device copy of the plan table (uploaded before capture: no host copies inside the graph)
        self.launches = None

    @torch.no_grad()
    def _decide(self, frames, prev_action):
        enc, tr, rew = self.nets["encoder"], self.nets["transition"], self.nets["reward_predictor"]
        dev, A = frames.device, self.A
        eye = torch.eye(A, dtype=torch.float32, device=dev)
        z = tr(enc(frames), eye[prev_action])                      # main.py:389-391 (prev_action: int64 [1])
        per = 1 + self.depth
        sig = tr.power_iterations(A * per)

        def rows(k):
            return torch.stack([sig[a * per + k] for a in range(A)])
        z_all = tr(z.repeat(A, 1, 1, 1), eye, sigma=rows(0))
        scores = rollout_scores(z_all, tr, rew, A, 2, self.depth, "noop", self.tradeoff, actions=self._plan,
                                sigma_for_step=lambda t: rows(1 + t))
        per_action = scores.view(A, -1).max(dim=1)[0]
        best = torch.argmax(per_action).to(torch.float32).reshape(1)
        return torch.cat([per_action, best]), z

    def decide(self, frames, prev_action, use_graph=True):
        """frames [1, 3, C, H, W] f32 device tensor (the three latest frames), prev_action: int or int64 tensor [1].
        Returns (best action, per-action scores as a CPU tensor, z of the current state)."""
        dev = frames.device
        if not torch.is_tensor(prev_action):
            prev_action = torch.tensor([int(prev_action)], dtype=torch.int64, device=dev)
        if self._plan is None:
            self._plan = beam_actions(self.A, 2, self.depth, "noop").to(dev)
        if not use_graph:
            out, z = self._decide(frames, prev_action)
        else:
            if self._graph is None:
                from . import kernels as K
                static = {"frames": frames.clone(), "prev_action": prev_action.clone()}
                tr = self.nets["transition"]
                snap = [(p, p.detach().clone()) for net in self.nets.values() for n_, p in net.named_parameters()
                        if n_.endswith("weight_u") or n_.endswith("weight_v")]
                rng = tr._rng_state.clone()
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._decide(**static)
                torch.cuda.current_stream().wait_stream(s)
                with torch.no_grad():
                    for p, val in snap:
                        p.copy_(val)
                    tr._rng_state.copy_(rng)
                n0 = K.launch_count()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out, z = self._decide(**static)
                self.launches = K.launch_count() - n0
                self._graph = (graph, static, out, z)
            graph, static, out, z = self._graph
            static["frames"].copy_(frames, non_blocking=True)
            static["prev_action"].copy_(prev_action, non_blocking=True)
            graph.replay()
        host = out.cpu()   # the only device->host transfer of the decision
        return int(host[-1].item()), host[:-1], z
