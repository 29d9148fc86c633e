"""CPU checks of the resampler restatement (oracle/resample.py) against the closed-form invariants
of SURVEY.md section 8c(3): mass conservation of the logit split, untouched particles, ordering."""
import numpy as np
import pytest

from oracle import resample as orr
from oracle.head import tf_multinomial_cpu


def synth(A=6, P=10, H=16, dead_frac=0.3, seed=33406, all_but_one=False):
    rng = np.random.default_rng(seed)
    logits = rng.normal(0, 3, size=(256, A, P))
    probs = np.exp(logits - logits.max(-1, keepdims=True))
    probs /= probs.sum(-1, keepdims=True)
    max_active = probs.max(0).astype(np.float32)
    sum_active = probs.sum(0).astype(np.float32)
    dead = rng.random((A, P)) < dead_frac
    if all_but_one:
        dead[:] = True
        dead[np.arange(A), rng.integers(0, P, A)] = False
    max_active[dead] = 1e-6
    sum_active[dead] *= 1e-4
    loc = np.linspace(-1, 1, P)[None].repeat(A, 0).astype(np.float32) + rng.normal(0, .02, (A, P)).astype(np.float32)
    logstd = (np.log(2 / (P - 1)) + rng.normal(0, .1, (A, P))).astype(np.float32)
    bias = rng.normal(0, 1, A * P).astype(np.float32)
    W = rng.normal(0, .01, (H, A * P)).astype(np.float32)
    draws = dict(cat_u=rng.random((A, P)), noise_u=(rng.random(A * P) * 2 - 1).astype(np.float32),
                 choice=rng.integers(0, 3, A * P).astype(np.int32))
    return dict(max_active=max_active, sum_active=sum_active, loc=loc, logstd=logstd, bias=bias, weight=W), draws, dead


def test_multinomial_matches_definition():
    lg = np.log(np.array([[0.1, 0.2, 0.0, 0.7]], dtype=np.float32))
    u = np.array([[0.0, 0.0999, 0.1001, 0.2999, 0.3001, 0.999999]])
    assert tf_multinomial_cpu(lg, u).tolist() == [[0, 0, 1, 1, 3, 3]]  # the zero-weight class is never drawn


@pytest.mark.parametrize("dead_frac,all_but_one", [(0.0, False), (0.05, False), (0.3, False), (0.5, False), (1.0, True)])
def test_resample_invariants(dead_frac, all_but_one):
    t, d, dead = synth(dead_frac=dead_frac, all_but_one=all_but_one)
    out, ints = orr.resample(**t, resample=-1, **d)
    A, P = t["loc"].shape
    M = ints["M"]
    assert M == int(dead.sum())
    # row-major order of tf.where
    flat = ints["invalid"][:, 0] * P + ints["invalid"][:, 1]
    assert np.all(np.diff(flat) > 0) and np.array_equal(flat, ints["col"])
    if M == 0:
        for k in ("loc", "logstd", "bias", "weight"):
            assert np.array_equal(out[k], t[k])
        return
    # sources are the j-th draw of row a
    assert np.array_equal(ints["src"], ints["cand"][ints["invalid"][:, 0], ints["invalid"][:, 1]])
    # particles that are neither dead nor a source are untouched
    touched = np.zeros(A * P, bool)
    touched[ints["col"]] = True
    touched[ints["tcol"]] = True
    assert np.array_equal(out["bias"][~touched], t["bias"][~touched])
    assert np.array_equal(out["weight"][:, ~touched], t["weight"][:, ~touched])
    assert np.array_equal(out["loc"].ravel()[~dead.ravel()], t["loc"].ravel()[~dead.ravel()])
    # mass conservation: exp(logit) of a LIVE source is split evenly over itself and its copies
    h = np.random.default_rng(1).normal(size=(5, t["weight"].shape[0])).astype(np.float32)
    z0 = orr.mixture_weights(t["bias"], t["weight"], h)
    z1 = orr.mixture_weights(out["bias"], out["weight"], h)
    for u, cnt, dl in zip(ints["uniq"], ints["count"], ints["delta"]):
        copies = ints["col"][ints["tcol"] == u]
        if dl == 0:  # live source keeps a share
            mass = np.exp(z1[:, u]) + np.exp(z1[:, copies]).sum(1)
            assert np.allclose(mass, np.exp(z0[:, u]), rtol=1e-5)
            assert np.allclose(z1[:, copies], z1[:, [u]], atol=1e-5)
    # unique_with_counts bookkeeping
    assert ints["count"].sum() == M and len(set(ints["uniq"].tolist())) == len(ints["uniq"])
    assert np.array_equal(ints["uniq"][ints["idx"]], ints["tcol"])
    # statistics are reset, logstd clipped, noise never zero
    assert not out["max_active"].any() and not out["sum_active"].any()
    assert np.all(out["logstd"] <= 2) and np.all(out["logstd"] >= -20)
    moved = out["loc"].ravel()[ints["col"]] - t["loc"].ravel()[ints["tcol"]]
    assert np.all(np.abs(moved) >= 0.99e-4)


def test_tanh_variant_and_topk_mode():
    t, d, dead = synth(dead_frac=0.3)
    out, ints = orr.resample(**t, resample=-1, tanh=True, **d)
    assert np.all(np.isfinite(out["loc"]))
    out, ints = orr.resample(**t, resample=3, **d)
    assert ints["cand"].shape == (6, 3)
    avg = t["sum_active"] / t["sum_active"].sum(1, keepdims=True)
    assert np.array_equal(ints["cand"], np.argsort(-avg, 1, kind="stable")[:, :3])
    assert np.array_equal(ints["src"], ints["cand"][ints["invalid"][:, 0], d["choice"][:ints["M"]]])


def test_dead_particle_chosen_as_source():
    """delta = 1: the source is itself dead and is overwritten by the copy of ITS source."""
    t, d, dead = synth(A=2, P=6, H=4, dead_frac=0.0)
    t["max_active"][0, 1] = 1e-6
    t["max_active"][0, 4] = 1e-6
    t["sum_active"][0, :] = np.array([0, 1, 0, 0, 0, 0], np.float32) + 1e-12  # every draw lands on particle 1
    out, ints = orr.resample(**t, resample=-1, **d)
    assert ints["M"] == 2 and ints["src"].tolist() == [1, 1] and ints["delta"].tolist() == [1]
    assert ints["count"].tolist() == [2]
    # b -= log(2 + 1 - 1): the two copies share the old mass, nothing is left at a third place
    assert np.allclose(out["bias"][[1, 4]], t["bias"][1] - np.log(2.0), atol=1e-6)
    assert np.array_equal(out["weight"][:, 4], t["weight"][:, 1])
