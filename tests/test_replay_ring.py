"""ReplayRing == the reference's `Buffer` (models/workers/ddpg.py:11-27) semantics: fixed capacity, overwrite at the
write pointer; sampling is uniform over the filled part.  Runs on CPU tensors (the class is device-agnostic plumbing)."""
import numpy as np
import torch

from pfpn_b200.sac import ReplayRing


class RefBuffer:  # restatement of ddpg.py:11-27 for the test
    def __init__(self, capacity):
        self.capacity, self.data, self.pointer, self.size = capacity, [None] * capacity, 0, 0

    def append(self, item):
        self.data[self.pointer] = item
        self.pointer = (self.pointer + 1) % self.capacity
        self.size = min(self.capacity, self.size + 1)


def test_ring_matches_reference_buffer_semantics():
    S, A, cap = 5, 2, 7
    ring, ref = ReplayRing(cap, S, A, device="cpu", seed=1), RefBuffer(cap)
    rng = np.random.RandomState(0)
    t = 0
    for n in (1, 3, 2, 4, 9, 1):  # single appends, batches, a wrap-around, a batch larger than the capacity
        st, ac, rw, nt, st2 = rng.randn(n, S), rng.randn(n, A), rng.randn(n), (rng.rand(n) > 0.3).astype(np.float64), rng.randn(n, S)
        ring.append(st, ac, rw, nt, st2)
        for i in range(n):
            ref.append(np.concatenate([st[i], ac[i], [rw[i]], [nt[i]], st2[i]]).astype(np.float32))
            t += 1
        assert len(ring) == ref.size and ring.pointer == ref.pointer
        for j in range(ref.size):
            assert np.array_equal(ring.data[j].numpy(), ref.data[j])
    s, a, r, nt, s2 = ring.sample(1000)
    assert s.shape == (1000, S) and a.shape == (1000, A) and r.shape == (1000,) and nt.shape == (1000,) and s2.shape == (1000, S)
    # every sampled row is one of the stored transitions, and all slots get drawn
    stored = {tuple(np.round(row.numpy(), 6)) for row in ring.data[:len(ring)]}
    rows = torch.cat([s, a, r[:, None], nt[:, None], s2], 1)
    seen = {tuple(np.round(row.numpy(), 6)) for row in rows}
    assert seen <= stored and len(seen) == len(stored)
