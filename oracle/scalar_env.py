"""An INDEPENDENT scalar restatement of Env.reset / Env.step (pure Python loops, one instance at a time).
TEST INFRASTRUCTURE ONLY.

Written from the reference text (VRPTW/vrptw_utils.py:199-358, VRP/vrp_utils.py:159-255) without looking at
oracle/restatement.py's vectorised code, so that the two can cross-check each other: a transcription slip in one
is unlikely to be repeated in the other.  Arithmetic is numpy float32 scalar arithmetic, one rounding per op,
exactly as TensorFlow executes the graph (Square, Add, Sqrt are separate ops).
"""
import numpy as np

f32 = np.float32


class ScalarEnv:
    def __init__(self, instance, capacity, tw):
        """instance: [N, D] array, depot last; D = 5 (x, y, b_tw, e_tw, demand) if tw else 3 (x, y, demand)."""
        inst = np.asarray(instance).astype(np.float32)
        self.tw = tw
        self.n = inst.shape[0]
        self.depot = self.n - 1
        self.x = [f32(v) for v in inst[:, 0]]
        self.y = [f32(v) for v in inst[:, 1]]
        if tw:
            self.b = [f32(v) for v in inst[:, 2]]
            self.e = [f32(v) for v in inst[:, 3]]
        self.demand0 = [f32(v) for v in inst[:, -1]]
        self.capacity = f32(capacity)

    def reset(self):
        self.demand = list(self.demand0)
        self.load = self.capacity
        self.time = f32(0)
        self.px, self.py = self.x[self.depot], self.y[self.depot]
        self.mask = [f32(1) if (i == self.depot or self.demand[i] == 0) else f32(0) for i in range(self.n)]
        return self.mask

    def _dist(self, ax, ay, bx, by):
        dx, dy = f32(ax - bx), f32(ay - by)
        return f32(np.sqrt(f32(f32(dx * dx) + f32(dy * dy))))

    def step(self, idx):
        d_sat = min(self.demand[idx], self.load)
        self.demand[idx] = f32(self.demand[idx] - d_sat)
        self.load = f32(self.load - d_sat)
        if self.tw:
            t_spent = self._dist(self.px, self.py, self.x[idx], self.y[idx])
            self.px, self.py = self.x[idx], self.y[idx]
            self.time = max(f32(self.time + t_spent), self.b[idx])
        at_depot = idx == self.depot
        if at_depot:
            self.load = self.capacity
            if self.tw:
                self.time = f32(0)
        remaining = f32(0)
        for d in self.demand:
            remaining = f32(remaining + d)
        mask = []
        for i in range(self.n):
            if i == self.depot:
                mask.append(f32(1) if (remaining > 0 and at_depot) else f32(0))
                continue
            m = f32(0)
            if self.demand[i] == 0:
                m = f32(m + 1)
            if self.load == 0:
                m = f32(m + 1)
            if self.tw and f32(self.time + self._dist(self.px, self.py, self.x[i], self.y[i])) > self.e[i]:
                m = f32(m + 1)
            mask.append(m)
        self.mask = mask
        return d_sat, mask


def tour_length(points):
    """reward_func, depot=None branch (vrptw_utils.py:397-414): closed tour over the visited coordinates."""
    total = f32(0)
    prev = points[-1]
    for p in points:
        dx, dy = f32(prev[0] - p[0]), f32(prev[1] - p[1])
        total = f32(total + f32(np.sqrt(f32(f32(dx * dx) + f32(dy * dy)))))
        prev = p
    return total
