"""The constants the oracle / kernels hard-code equal the ones in the lowered TF graph of the shipped
DPPO-PFPN-35 checkpoint (tests/golden/graph_constants.json, extracted by oracle/metagraph.py)."""
import json
import math
import os

import numpy as np

from oracle import head as oh

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "graph_constants.json")))
f32 = lambda x: float(np.float32(x))


def test_normal_prob_and_ppo_constants():
    n = G["normal_prob_consts"]
    assert n["global_net/actor/Normal/prob_1/mul/x"] == -0.5
    assert n["global_net/actor/Normal/prob_1/add/x"] == f32(oh.HALF_LOG_2PI) == f32(0.5 * math.log(2 * math.pi))
    c = G["clipped_surrogate_consts"]
    assert c["global_net/policy_loss/clipped_surrogate/clip_by_value/y"] == f32(1 - 0.2)
    assert c["global_net/policy_loss/clipped_surrogate/clip_by_value/Minimum/y"] == f32(1 + 0.2)
    assert G["normalize_advantage_consts"]["global_net/normalize_advantage/add/y"] == f32(1e-8)
    s = G["state_clip_consts"]
    assert s["global_net/state_normalizer/clip_state/clip_range/lb_-5.0"] == -5.0
    assert s["global_net/state_normalizer/clip_state/clip_range/ub_5.0"] == 5.0
    assert s["global_net/state_normalizer/update/Const"] == f32(0.9999)
    assert s["global_net/state_normalizer/update/Maximum/x"] == f32(1e-6)
    assert len(G["softmax_nodes"]) == 2  # Categorical/probs and ExpRelaxedOneHotCategorical/probs


def test_resampler_and_optimizer_constants():
    r = G["resample_cond_consts"]
    assert r["cond/Less/y"] == f32(0.05 / 35)                      # a2c.py:391 at P = 35
    assert r["cond/mul_3/x"] == f32(-1e-4) and r["cond/mul_4/x"] == f32(1e-4)  # a2c.py:442-444
    assert r["cond/clip_by_value/y"] == -20.0 and r["cond/clip_by_value/Minimum/y"] == 2.0  # a2c.py:451
    assert r["cond/random_uniform/min"] == -1.0 and r["cond/random_uniform/max"] == 1.0
    assert r["cond/add_5/y"] == 1.0                                 # log(count + 1 - delta)
    assert G["resample_interval"] == [368.0]
    assert "cond/categorical/Multinomial" in G["multinomial"]
    assert G["adam"]["optimizer/optimizer/lr"] == f32(1e-4)
    assert G["clip_by_global_norm_consts"]["optimizer/clip_by_global_norm/Const"] == 1.0
    assert G["n_nodes"] == 6052
