"""CPU checks of host-side policies that need no device: the engine pool of a module (models._EngineCache) and the
equal-length batching of the downstream validation loop (epochs.length_buckets)."""
import torch

from zeronotesamba_b200 import epochs
from zeronotesamba_b200.models import models as M


class _FakeEngine:
    """Stands in for engine.EncoderEngine (which allocates device workspaces): capacity, seed, set_T."""
    built = 0

    def __init__(self, batch, T, n_br, device, seed=0):
        _FakeEngine.built += 1
        self.B, self.T_cap, self.n_br, self.device, self.seed, self.T = batch, T, n_br, device, seed, T

    def set_T(self, T):
        assert 1 <= T <= self.T_cap
        self.T = T


def _reserve(eng, clock):
    eng._pending, eng._version = True, clock


def test_engine_pool_policy(monkeypatch):
    monkeypatch.setattr(M, "EncoderEngine", _FakeEngine)
    _FakeEngine.built = 0
    cache, dev = M._EngineCache(), torch.device("cuda", 0)
    e0 = cache.get(4, 300, 1, dev)
    assert cache.get(4, 200, 1, dev) is e0 and e0.T == 200 and _FakeEngine.built == 1     # free engine: reused, re-viewed
    _reserve(e0, 1)                                                                        # a forward awaits its backward
    e1 = cache.get(4, 300, 1, dev)
    assert e1 is not e0 and e1.seed != e0.seed                                             # second engine, other dropout seed
    _reserve(e1, 2)
    e2 = cache.get(4, 300, 1, dev)
    _reserve(e2, 3)
    assert len({id(e0), id(e1), id(e2)}) == 3 and _FakeEngine.built == 3 == M._EngineCache.MAX_PENDING
    e3 = cache.get(4, 300, 1, dev)                     # pool exhausted: the OLDEST outstanding forward loses its engine
    assert e3 is e0 and not e0._pending and _FakeEngine.built == 3
    e1._pending = False                                # its backward ran
    assert cache.get(4, 100, 1, dev) is e0             # first free slot wins (e0 was handed out unreserved above)
    # growth keeps the slot, its seed and its version counter; capacity gets head room
    _reserve(e0, 9)
    g = cache.get(4, 301, 1, dev)                      # slot 1 is free but too short: rebuilt in place with 25 % head room
    assert g is not e1 and g.seed == e1.seed and g.T_cap == 375 and g.T == 301
    big = cache.get(4, 1000, 1, dev)
    assert big.seed == e1.seed and big.T_cap == 1000 and big.T == 1000
    # another geometry has its own pool; at most four geometries are kept
    for b in (1, 2, 3, 5, 6):
        cache.get(b, 64, 2, dev)
    assert len(cache._engines) <= 4


def test_length_buckets():
    inputs = {k: torch.zeros(2, 96, T) for k, T in zip("abcdefg", (400, 400, 520, 400, 400, 520, 7))}
    idx = list("abcdefg")
    assert epochs.length_buckets(idx, inputs, 3) == [["a", "b", "d"], ["e"], ["c", "f"], ["g"]]
    assert epochs.length_buckets(idx, inputs, 16) == [["a", "b", "d", "e"], ["c", "f"], ["g"]]
    assert epochs.length_buckets([], inputs) == []
    flat = [w for grp in epochs.length_buckets(idx, inputs, 2) for w in grp]
    assert sorted(flat) == idx                          # every file exactly once
    single = {k: torch.zeros(96, T) for k, T in (("x", 10), ("y", 10))}          # vanilla status: (96, T) inputs
    assert epochs.length_buckets(["x", "y"], single) == [["x", "y"]]
