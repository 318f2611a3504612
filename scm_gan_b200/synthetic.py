"""Synthetic trajectory source honouring the reference's datasource contract (datasource.py:8-120,
envs/minipacman.py:122-164): `get_trajectories(batch_size, timesteps, random_start, training)` returns numpy
`(states [B,T,C,H,W] float, rewards [B,T,R] float, dones [B,T] bool, actions [B,T] int)`.

The reference's environments (gym_minipacman, SC2, ...) cannot be installed offline; this stand-in produces frames of
the same shapes with learnable, action-conditional dynamics on a torus (which is what Transition's circular padding
models): a few "agent" pixels move by the chosen action, "food" pixels are static and are eaten (reward +1) when an
agent lands on them, a "ghost" drifts and costs -1 on contact.
"""
import numpy as np
import torch

MOVES = [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1), (1, 1), (-1, -1)]  # action -> (dy, dx); extra actions reuse entries


class MovingDots:
    def __init__(self, channels=3, height=15, width=19, num_actions=5, num_rewards=2, seed=0, p_done=0.0):
        self.conv_input_channels = channels
        self.conv_output_channels = channels
        self.binary_input_channels = num_actions
        self.scalar_output_channels = num_rewards
        self.height, self.width = height, width
        self.rng = np.random.RandomState(seed)
        self.p_done = p_done

    def make_env(self, **kwargs):
        """A gym-style environment with the same dynamics (reset() / step(a)), for the MPC loop of main.py:327-400."""
        return MovingDotsEnv(self.conv_input_channels, self.height, self.width, self.binary_input_channels,
                             self.scalar_output_channels, seed=int(self.rng.randint(1 << 30)))

    def convert_frame(self, state, **kwargs):
        """-> (network frame [C,H,W] float32, rgb frame [H,W,3] uint8), as datasource.convert_frame does."""
        rgb = (np.transpose(state[:3], (1, 2, 0)) * 255).astype(np.uint8)
        return state.astype(np.float32), rgb

    def get_trajectories(self, batch_size=32, timesteps=10, random_start=True, training=True):
        C, H, W = self.conv_input_channels, self.height, self.width
        A, R = self.binary_input_channels, self.scalar_output_channels
        rng = self.rng
        states = np.zeros((batch_size, timesteps, C, H, W), dtype=np.float32)
        rewards = np.zeros((batch_size, timesteps, R), dtype=np.float32)
        dones = np.zeros((batch_size, timesteps), dtype=bool)
        actions = rng.randint(A, size=(batch_size, timesteps))
        for b in range(batch_size):
            agent = np.array([rng.randint(H), rng.randint(W)])
            ghost = np.array([rng.randint(H), rng.randint(W)])
            gdir = np.array(MOVES[1 + rng.randint(4)])
            food = rng.rand(H, W) < 0.08
            done = False
            for t in range(timesteps):
                a = actions[b, t]
                agent = (agent + np.array(MOVES[a % len(MOVES)])) % (H, W)
                ghost = (ghost + gdir) % (H, W)
                if food[agent[0], agent[1]]:
                    food[agent[0], agent[1]] = False
                    rewards[b, t, 0] = 1.0
                if R > 1 and (agent == ghost).all():
                    rewards[b, t, 1] = -1.0
                states[b, t, 0, agent[0], agent[1]] = 1.0
                states[b, t, 1 % C][food] = 1.0
                states[b, t, 2 % C, ghost[0], ghost[1]] = 1.0
                if C > 3:
                    states[b, t, 3, :, :] = 0.0
                if not done and rng.rand() < self.p_done:
                    done = True
                dones[b, t] = done
        return states, rewards, dones, actions


class MovingDotsEnv:
    """One interactive episode of the MovingDots dynamics: reset() -> frame, step(a) -> (frame, reward, done, info).
    reward is the sum of the reward channels; info carries them separately, like the reference's environments."""

    def __init__(self, channels=3, height=15, width=19, num_actions=5, num_rewards=2, seed=0, episode_length=40):
        self.C, self.H, self.W, self.A, self.R = channels, height, width, num_actions, num_rewards
        self.rng = np.random.RandomState(seed)
        self.episode_length = episode_length
        self.reset()

    def _frame(self):
        f = np.zeros((self.C, self.H, self.W), dtype=np.float32)
        f[0, self.agent[0], self.agent[1]] = 1.0
        f[1 % self.C][self.food] = 1.0
        f[2 % self.C, self.ghost[0], self.ghost[1]] = 1.0
        return f

    def reset(self):
        rng, H, W = self.rng, self.H, self.W
        self.agent = np.array([rng.randint(H), rng.randint(W)])
        self.ghost = np.array([rng.randint(H), rng.randint(W)])
        self.gdir = np.array(MOVES[1 + rng.randint(4)])
        self.food = rng.rand(H, W) < 0.08
        self.t = 0
        return self._frame()

    def step(self, action):
        H, W = self.H, self.W
        self.agent = (self.agent + np.array(MOVES[int(action) % len(MOVES)])) % (H, W)
        self.ghost = (self.ghost + self.gdir) % (H, W)
        info = {}
        if self.food[self.agent[0], self.agent[1]]:
            self.food[self.agent[0], self.agent[1]] = False
            info["food"] = 1.0
        if self.R > 1 and (self.agent == self.ghost).all():
            info["ghost"] = -1.0
        self.t += 1
        done = self.t >= self.episode_length or not self.food.any()
        return self._frame(), float(sum(info.values())), done, info


def synthetic_batch(batch, horizon, channels, height, width, num_actions, num_rewards, seed=1234, p_done=0.0):
    """i.i.d. benchmark trajectories of SURVEY.md section 8d (sparse binary frames, 5 % non-zero rewards): the shapes
    and dtypes of `get_trajectories`, no dynamics - what bench.py feeds the step.  Returns CPU tensors
    (states f32, rewards f32, dones f32) and a numpy int64 action array."""
    g = torch.Generator().manual_seed(seed)
    states = (torch.rand(batch, horizon, channels, height, width, generator=g) < 0.15).float()
    r = torch.rand(batch, horizon, num_rewards, generator=g)
    rewards = torch.where(r < 0.025, -1.0, torch.where(r > 0.975, 1.0, 0.0))
    dones = (torch.rand(batch, horizon, generator=g) < p_done).float()
    actions = torch.randint(num_actions, (batch, horizon), generator=g)
    return states, rewards, dones, actions.numpy()
