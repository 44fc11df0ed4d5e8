"""Host logic of the attention modules against goldens produced by the REFERENCE classes
(tools/make_golden_modules.py).  No GPU here: the one CUDA call inside the modules
(`msda_apply`) is replaced by the CPU oracle, so what is tested is everything around it --
rebatching, Z-anchor broadcast, softmax layout, queue averaging, scatter-add, count
normalisation, projections, parameter names."""
import os

import numpy as np
import pytest
import torch

from oracle import msda_ref
from tests import module_cases as mc
from vidar_b200.modules import deform_attn
from vidar_b200.registry import ATTENTION, build_attention

GOLD = os.path.join(os.path.dirname(__file__), "golden", "modules.npz")
GOLD_CUSTOM = os.path.join(os.path.dirname(__file__), "golden", "modules_custom.npz")


@pytest.fixture()
def oracle_msda(monkeypatch):
    def apply(value, shapes, lsi, loc, attn, im2col_step):
        return msda_ref.msda_grid_sample(value, shapes, loc, attn)
    monkeypatch.setattr(deform_attn, "msda_apply", apply)


@pytest.mark.parametrize("kind,cfg,case,seed", [("sca", mc.SCA_CFG, mc.sca_case, 10),
                                                ("tsa", mc.TSA_CFG, mc.tsa_case, 11),
                                                ("pred", mc.PRED_CFG, mc.pred_case, 12)])
def test_module_matches_reference_class(oracle_msda, kind, cfg, case, seed):
    g = np.load(GOLD)
    m = build_attention(cfg)
    assert sorted(m.state_dict().keys()) == list(g[f"{kind}_params"])     # checkpoints load unchanged
    m.load_state_dict(mc.seeded_state(m, seed))
    m.eval()
    out, gq, gkv = mc.run_module(m, kind, case())
    np.testing.assert_allclose(out.numpy(), g[f"{kind}_out"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gq.numpy(), g[f"{kind}_gq"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gkv[:, ::6].numpy(), g[f"{kind}_gkv_s6"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("kind,boxes,seed", [("custom", False, 13), ("custom_boxes", True, 14)])
def test_detection_decoder_attention_matches_reference_class(oracle_msda, kind, boxes, seed):
    """CustomMSDeformableAttention (decoder.py:132-345): sequence-first layout, both reference forms."""
    g = np.load(GOLD_CUSTOM)
    m = build_attention(mc.CUSTOM_CFG)
    assert sorted(m.state_dict().keys()) == list(g[f"{kind}_params"])
    assert m.batch_first is False
    m.load_state_dict(mc.seeded_state(m, seed))
    m.eval()
    out, gq, gkv = mc.run_module(m, kind, mc.custom_case(boxes=boxes))
    np.testing.assert_allclose(out.numpy(), g[f"{kind}_out"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gq.numpy(), g[f"{kind}_gq"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gkv.numpy(), g[f"{kind}_gkv"], rtol=1e-4, atol=2e-5)


def test_registry_names_and_default_init():
    for name in ("SpatialCrossAttention", "MSDeformableAttention3D", "TemporalSelfAttention",
                 "PredictionMSDeformableAttention", "CustomMSDeformableAttention"):
        assert name in ATTENTION
    m = build_attention(dict(type="MSDeformableAttention3D", embed_dims=256, num_points=8, num_levels=4))
    b = m.sampling_offsets.bias.view(8, 4, 8, 2)
    assert torch.allclose(b[0, 0, :, 0], torch.arange(1.0, 9.0))           # head 0 looks along +x, radius i+1
    assert float(m.sampling_offsets.weight.abs().sum()) == 0.0
    assert m.output_proj is None
    with pytest.raises(KeyError):
        build_attention(dict(type="NoSuchAttention"))


def test_modules_have_no_cpu_fallback():
    m = build_attention(mc.TSA_CFG)
    c = mc.tsa_case()
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        m(c["query"], None, c["value"], reference_points=c["reference_points"],
          spatial_shapes=c["spatial_shapes"], level_start_index=c["level_start_index"])


def test_visible_lists_equal_reference_nonzero():
    """The rebatch path's lists: per camera the reference's `nonzero()` indices, padded to the longest."""
    c = mc.sca_case()
    mask = c["bev_mask"]
    idx, live, max_len = deform_attn._visible_lists(mask)
    hit = mask[:, 0].sum(-1) > 0
    for cam in range(mask.shape[0]):
        want = hit[cam].nonzero().squeeze(-1)                      # the reference's index_query_per_img
        assert torch.equal(idx[cam][live[cam]], want)
    assert max_len == int(hit.sum(-1).max())


def test_unit_plan_partitions_every_row_exactly_once():
    """Camera sharding (vidar_b200/sca.py): over all ranks the (camera, 64-row block) units are covered
    exactly once, shares are equal, and a rank touches at most two partial cameras."""
    from vidar_b200.sca import SLICE_ROWS, plan_cameras, unit_plan
    cams, Q = 6, 1000
    nblk = -(-Q // SLICE_ROWS)
    for world in (1, 2, 3, 4, 5, 6, 8, 12):
        seen = {}
        sizes = []
        for rank in range(world):
            plan = unit_plan(world, rank, cams)
            units = 0
            for cam0, ncl, S, lo, hi in plan:
                for c in range(cam0, cam0 + ncl):
                    for blk in range(nblk):
                        if lo <= blk % S < hi:
                            assert (c, blk) not in seen
                            seen[(c, blk)] = rank
                    units += (hi - lo) * 1.0 / S
            sizes.append(units)
            assert len([g for g in plan if (g[3], g[4]) != (0, g[2])]) <= 2
            assert plan_cameras(plan) == sorted(plan_cameras(plan))
        assert len(seen) == cams * nblk
        assert max(sizes) - min(sizes) < 1e-9


def test_shared_camera_groups_cannot_deadlock():
    """A camera split over several ranks gets its own sub-communicator (features broadcast, grad_value summed
    inside it).  Ranks enqueue these small collectives in whatever order autograd reaches them, so the scheme is
    only safe if no cyclic wait is possible: the graph "ranks joined by a shared camera" must be a forest and
    two ranks must never share more than one camera."""
    from vidar_b200.sca import camera_ranks, plan_cameras, unit_plan
    cams = 6
    for world in range(1, 13):
        owners = camera_ranks(world, cams)
        assert sorted(owners) == list(range(cams)) and all(owners[c] for c in owners)
        for rank in range(world):                     # camera_ranks is the inverse of the per-rank plans
            assert [c for c in owners if rank in owners[c]] == plan_cameras(unit_plan(world, rank, cams))
        shared = [tuple(m) for m in owners.values() if len(m) > 1]
        pairs = [frozenset((a, b)) for m in shared for i, a in enumerate(m) for b in m[i + 1:]]
        assert len(pairs) == len(set(pairs)), f"world {world}: two ranks share two cameras"
        parent = list(range(world))                   # union-find: an edge inside one component would close a cycle

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x
        for m in shared:
            roots = {find(r) for r in m}
            assert len(roots) == len(m), f"world {world}: cyclic wait possible through camera group {m}"
            for r in m[1:]:
                parent[find(r)] = find(m[0])
