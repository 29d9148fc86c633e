"""TF-1 tensor-bundle checkpoint interop: index parser against the facts recorded from the shipped
`.index` (tests/golden/graph_constants.json), and a write -> read round trip of the data path.  CPU only."""
import json
import os

import numpy as np
import pytest

from pfpn_b200 import checkpoint as ck

HERE = os.path.dirname(__file__)
REF_INDEX = "/root/reference/ckpt_DeepMimicWalk-v0/deepmimic_dppo_pfpn_particle35/34114/model.ckpt-78000.index"


def test_crc32c_known_answers():
    assert ck.crc32c(b"") == 0
    assert ck.crc32c(b"123456789") == 0xE3069283          # the standard CRC-32C check value
    assert ck.crc32c(bytes(32)) == 0x8A9136AA              # RFC 3720 B.4: 32 zero bytes


def test_bundle_round_trip(tmp_path):
    rng = np.random.RandomState(0)
    tensors = {"global_net/actor/fc1/weight": rng.randn(197, 1024).astype(np.float32),
               "global_net/actor/samples": rng.randn(36, 35).astype(np.float32),
               "global_net/resample/train_flag": np.float32(17.0),
               "step/global_step": np.int64(78000),
               "global_net/episode/episode": np.int32(5)}
    prefix = str(tmp_path / "model.ckpt-1")
    ck.write_bundle(prefix, tensors)
    idx = ck.read_bundle_index(prefix + ".index")
    assert sorted(idx) == sorted(tensors)
    assert idx["global_net/actor/fc1/weight"].shape == (197, 1024) and idx["global_net/actor/fc1/weight"].size == 197 * 1024 * 4
    back = ck.load_bundle(prefix)
    for k, v in tensors.items():
        assert back[k].dtype == np.asarray(v).dtype and np.array_equal(back[k], v)
    # corruption is detected through the per-tensor crc32c
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(100)
        f.write(b"\xff")
    with pytest.raises(ValueError, match="crc32c"):
        ck.load_bundle(prefix)


@pytest.mark.skipif(not os.path.exists(REF_INDEX), reason="the reference tree is only present in the build container")
def test_reader_on_the_shipped_index():
    idx = ck.read_bundle_index(REF_INDEX)
    assert len(idx) == 212
    w = idx["global_net/actor/fc_policy/weight"]
    assert w.shape == (512, 1260) and w.dtype == 1 and w.size == 512 * 1260 * 4
    assert idx["global_net/actor/samples"].shape == (36, 35) and idx["global_net/actor/samples_std"].shape == (36, 35)
    assert idx["global_net/max_active_degree"].shape == (36, 35) and idx["global_net/state_normalizer/mean"].shape == (197,)
    assert idx["step/global_step"].dtype == 9 and idx["step/global_step"].shape == ()
    # offsets tile the single data shard without gaps
    ents = sorted(idx.values(), key=lambda e: e.offset)
    assert ents[0].offset == 0 and all(a.offset + a.size == b.offset for a, b in zip(ents, ents[1:]))
    # the same facts the graph-constants fixture recorded from this file
    G = json.load(open(os.path.join(HERE, "golden", "graph_constants.json")))
    shapes = G.get("checkpoint_index_shapes") or G.get("index_shapes") or {}
    for name, shp in shapes.items():
        if name in idx:
            assert list(idx[name].shape) == list(shp)


@pytest.mark.gpu
def test_network_checkpoint_round_trip_by_reference_names(cuda_dev, tmp_path):
    import torch
    from pfpn_b200.network import ParticleFilteringClipPPONetwork

    def make(seed):
        return ParticleFilteringClipPPONetwork(True, [197], [36], action_lower_bound=[-1.] * 36, action_upper_bound=[1.] * 36,
                                               particles=35, resample=-1, resample_interval=368, device=cuda_dev, seed=seed).init()
    a, b = make(1), make(2)
    a.max_active.uniform_(0, 1); a.sum_active.uniform_(0, 5); a.state_mean.normal_(); a.state_std.uniform_(0.5, 2)
    a.train_flag, a.global_step = 17, 78000
    prefix = str(tmp_path / "model.ckpt-78000")
    ck.save_tf_checkpoint(a, prefix)
    idx = ck.read_bundle_index(prefix + ".index")
    # exactly the variable names / shapes of the shipped checkpoint for everything the network owns
    assert idx["global_net/actor/fc1/weight"].shape == (197, 1024) and idx["global_net/critic/fc3/weight"].shape == (512, 1)
    assert idx["global_net/actor/samples_std"].shape == (36, 35) and idx["global_net/sum_active_degree"].shape == (36, 35)
    assert not torch.equal(a.params, b.params)
    loaded, extra = ck.load_tf_checkpoint(b, prefix)
    assert extra == [] and len(loaded) == len(idx)
    assert torch.equal(a.params, b.params)
    for k in ("max_active", "sum_active", "state_mean", "state_std"):
        assert torch.equal(getattr(a, k), getattr(b, k))
    assert (b.train_flag, b.global_step) == (17, 78000)
