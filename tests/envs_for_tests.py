"""Deterministic toy environments used by the EnvPool tests (the reference's known-answer env of
test/unit/test_envpool.py:13-36 restated, plus an Atari-shaped one)."""
import numpy as np
import torch


class ToyEnv:
    """n is a 4x4 float tensor; action 1 doubles it, 2 halves it; done when it sums below 1; reward |sum - 4|."""

    def __init__(self):
        self.n = None

    def reset(self):
        self.n = torch.ones(4, 4)
        self.n[0][0] = 4.0
        self.n[3][1] = 0.5
        self.n[1][2] = 0.25
        return {"n": self.n}

    def step(self, action):
        if action == 1:
            self.n *= 2
        elif action == 2:
            self.n /= 2
        elif action != 0:
            raise RuntimeError("bad action")
        return {"n": self.n}, abs(self.n.sum() - 4), self.n.sum() < 1


class FrameEnv:
    """Atari-shaped: obs u8 [4,84,84] = a counter pattern that depends on the action history; plain-array observation
    (-> key "state"), done every 7th step."""

    def __init__(self):
        self.t = 0
        self.acc = 0

    def _obs(self):
        return (np.arange(4 * 84 * 84, dtype=np.int64).reshape(4, 84, 84) * (self.acc + 1) + self.t).astype(np.uint8)

    def reset(self):
        self.t, self.acc = 0, 0
        return self._obs()

    def step(self, action):
        self.t += 1
        self.acc = (self.acc * 31 + int(action)) % 251
        return self._obs(), float(self.acc) / 7.0, self.t % 7 == 0
