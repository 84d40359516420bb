"""CPU tests (no GPU needed): the oracle against the committed reference goldens, the product's host-side logic
(mask builders bit-exact, module/state_dict layout, gradient-arena and spectral-norm planning, optimizer state
format, gloo gradient averaging) and the C-ABI library (loads, exports every symbol include/*.h declares)."""
import hashlib
import os
import random
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import spyramid_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "semantic_pyramid_for_image_generation_b200"


def _digest(tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().float().numpy().tobytes())
    return h.hexdigest()


@pytest.fixture(scope="module")
def golden_step():
    return torch.load(os.path.join(GOLDEN, "step_cf2_b2.pt"), weights_only=False)


@pytest.fixture(scope="module")
def oracle_step(golden_step):
    cfg = golden_step["config"]
    s = cfg["seeds"]
    g_sd = O.init_generator_state(cfg["channel_factor"], seed=s["g"])
    d_sd = O.init_discriminator_state(cfg["channel_factor"], seed=s["d"])
    v_sd = O.init_vgg_state(s["v"])
    images, labels, masks, z_d, z_g = O.synthetic_batch(cfg["batch"], seed=s["batch"], mask_mode=cfg["mask_mode"])
    out = O.train_step(g_sd, d_sd, v_sd, images, labels, masks, z_d, z_g, {}, {}, lr=cfg["lr"])
    return out, g_sd, d_sd


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def test_oracle_losses_match_reference_golden(golden_step, oracle_step):
    out, _, _ = oracle_step
    for name, ref in golden_step["losses"].items():
        assert out[name] == pytest.approx(ref, rel=2e-4, abs=1e-7), name


def test_oracle_activations_match_reference_golden(golden_step, oracle_step):
    out, _, _ = oracle_step
    assert rel(out["images_fake"][:, :, ::8, ::8], golden_step["images_fake_sub"]) < 1e-4
    assert float(out["images_fake"].norm()) == pytest.approx(golden_step["images_fake_norm"], rel=1e-4)
    for f, sub, nrm in zip(out["features_real"], golden_step["features_real_sub"], golden_step["features_real_norm"]):
        mine = f[:, ::8, ::8, ::8] if f.dim() == 4 else f[:, ::16]
        assert rel(mine, sub) < 1e-4
        assert float(f.norm()) == pytest.approx(nrm, rel=1e-4)


def test_oracle_gradients_and_state_match_reference_golden(golden_step, oracle_step):
    out, g_sd, d_sd = oracle_step
    for grads, norms in ((out["d_grads"], golden_step["d_grad_norms"]), (out["g_grads"], golden_step["g_grad_norms"])):
        for k, ref in norms.items():
            if ref > 1e-6:  # analytically-zero gradients (bias feeding a batch norm, key bias) are rounding noise
                assert float(grads[k].norm()) == pytest.approx(ref, rel=2e-2 if grads[k].numel() == 1 else 2e-3), k
    for k, head in golden_step["d_grad_head"].items():
        assert rel(out["d_grads"][k].flatten()[:64], head) < 5e-3, k
    for k, ref in golden_step["post_step"].items():
        assert torch.allclose(g_sd[k].float(), ref.float(), rtol=1e-3, atol=2e-6), k
    for k, ref in golden_step["post_step_d"].items():
        assert torch.allclose(d_sd[k].float(), ref.float(), rtol=1e-3, atol=2.5e-5), k


def test_masks_bit_exact_against_reference_golden():
    """Integer/boolean path: product misc.py and the oracle must reproduce the reference's masks bit for bit."""
    from semantic_pyramid_for_image_generation_b200 import misc
    gold = torch.load(os.path.join(GOLDEN, "masks.pt"), weights_only=False)
    spatial = 0
    for entry in gold["training"]:
        random.seed(entry["seed"])
        np.random.seed(entry["seed"])
        masks = misc.get_masks_for_training()
        assert [tuple(m.shape) for m in masks] == [tuple(s) for s in misc.PYRAMID_SHAPES]
        assert _digest(masks) == entry["sha256"], entry["seed"]
        assert all(set(m.unique().tolist()) <= {0.0, 1.0} for m in masks)
        spatial += int(any(0.0 < mean < 1.0 for mean in entry["means"]))
        # the oracle's restatement agrees as well (same RNG order, same rasteriser injected)
        random.seed(entry["seed"])
        np.random.seed(entry["seed"])
        oracle_masks = O.masks_for_training(
            lambda hw, min_size: misc._builtin_random_shapes(hw, min_shapes=1, max_shapes=4, min_size=min_size)[0][:, :, 0])
        assert _digest(oracle_masks) == entry["sha256"], entry["seed"]
    assert spatial > 0, "the golden set must exercise the spatially varying branch"
    for entry in gold["inference"]:
        assert _digest(misc.get_masks_for_inference(entry["stage"])) == entry["sha256"]
        assert _digest(O.masks_for_inference(entry["stage"])) == entry["sha256"]
    batched = misc.get_masks_for_inference(2, add_batch_size=True)
    assert tuple(batched[0].shape) == (1, 1, 128, 128) and tuple(batched[6].shape) == (1, 365)


def test_mask_descriptor_roundtrip_and_edge_cases():
    from semantic_pyramid_for_image_generation_b200 import misc
    # every stage, with and without a bitmap; ragged bitmap sizes resize by integer nearest-neighbour
    for stage in range(7):
        masks = misc.expand_mask_descriptor(misc.MaskDescriptor(stage, None))
        assert sum(float(m.sum()) > 0 for m in masks) == 1
    bitmap = np.zeros((32, 32), dtype=np.uint8)
    bitmap[:9, 3:] = 1
    masks = misc.expand_mask_descriptor(misc.MaskDescriptor(3, bitmap))  # stage 3 = pool4; the bitmap lives at pool3 (32x32)
    rev = list(reversed(masks))
    assert float(rev[3].min()) == 1.0 and all(float(rev[i].sum()) == 0.0 for i in range(3))
    assert torch.equal(rev[4][0], torch.as_tensor(bitmap, dtype=torch.float32))
    assert torch.equal(rev[6][0, ::4, ::4], torch.as_tensor(bitmap, dtype=torch.float32))  # 128x128: each bit 4x4 times
    assert float(rev[6].sum()) == 16.0 * float(bitmap.sum())


def test_modules_mirror_reference_state_dict_layout():
    from semantic_pyramid_for_image_generation_b200 import models
    for cf in (1, 2):
        g, d = models.Generator(channels_factor=cf), models.Discriminator(channel_factor=cf)
        for module, ref in ((g, O.init_generator_state(cf)), (d, O.init_discriminator_state(cf))):
            sd = module.state_dict()
            assert list(sd.keys()) == list(ref.keys())
            for k in sd:
                assert tuple(sd[k].shape) == tuple(ref[k].shape), k
            module.load_state_dict(ref)  # strict
    g = models.Generator()
    assert sum(p.numel() for p in g.parameters()) == 29967047
    assert sum(p.numel() for p in models.Discriminator().parameters()) == 16820994
    assert g.latent_dimensions == 128
    # Xavier-uniform weights / zero biases / CBN embedding = [1..1|0..0] / gamma = 1
    w = g.main_path[0].main_block[3].weight_orig
    bound = (6.0 / ((w.shape[0] + w.shape[1]) * 9)) ** 0.5
    assert float(w.abs().max()) <= bound * (1 + 1e-6) and float(w.abs().max()) > 0.9 * bound  # FP32 vs double bound
    assert float(g.main_path[0].main_block[3].bias.abs().max()) == 0.0
    emb = g.main_path[0].main_block[0].embedding.weight
    assert float(emb[:, :512].min()) == 1.0 and float(emb[:, 512:].abs().max()) == 0.0
    assert float(g.main_path[3].gamma) == 1.0
    with pytest.raises(ValueError):
        models.Generator(channels_factor=40)
    v = models.VGG16()
    assert set(v.state_dict().keys()) == set(O.init_vgg_state().keys())
    with pytest.raises(RuntimeError):  # no CPU fallback
        g(torch.zeros(2, 128), [None] * 7, [None] * 7, torch.zeros(2, 365))


def test_gradient_arena_and_spectral_norm_plan():
    from semantic_pyramid_for_image_generation_b200 import _native as N
    from semantic_pyramid_for_image_generation_b200 import models
    import ctypes as C
    d = models.Discriminator()
    ga, sn = d._ga, d._sn
    offs = sorted(ga.offset_list)
    assert offs[0] == 0 and all(o % 64 == 0 for o in offs) and ga.total >= 16820994
    assert sn.n == 28 and sn.gw_floats == sum((s.rows * s.cols + 3) // 4 * 4 for s in sn.specs)
    # first-layer operands are im2col rows; every other conv is [taps][cout][cin]
    assert sn.by_key["layers.0.main_block.0"].pack_mode == 1 and sn.by_key["layers.0.main_block.0"].pack_cin == 32
    assert sn.by_key["layers.0.residual_mapping"].pack_cin == 8
    assert sn.by_key["layers.11"].pack_cin == 0 and sn.by_key["embedding"].rows == 365
    assert all(s.pack_off % 64 == 0 for s in sn.specs)
    g = models.Generator()
    spec = g._sn.by_key["main_path.0.masked_feature_mapping"]
    assert spec.stencil and spec.pack_cin == 512 and spec.cin == 513
    # the host-only planner of the C-ABI
    tab = (N.SnLayer * 2)()
    for i, (rows, cin, taps) in enumerate(((64, 64, 9), (365, 128, 1))):
        tab[i].rows, tab[i].cin, tab[i].taps, tab[i].cols = rows, cin, taps, cin * taps
        tab[i].pack_cin = cin if taps == 9 else 0
    plan = N.SnPlan()
    N.call_nostream("spyr_sn_plan", tab, 2, C.byref(plan))
    # power iteration: (row tile, 256 columns) CTAs store partial sums, sn_tsum adds the row tiles in order
    assert plan.tiles_wtu == 1 * 3 + 6 * 1 and plan.tiles_tsum == 3 + 1 and plan.tiles_wv == 8 + 46
    assert plan.tiles_pack == 64 + 1
    assert plan.saved_floats == (1 + 64 + 576) + (1 + 365 + 128)
    # offsets follow the module's CURRENT parameter objects: a deep copy (an EMA generator, say) keeps working
    import copy
    d2 = copy.deepcopy(models.Discriminator(channel_factor=4))
    p_last = list(d2.parameters())[-1]
    assert d2._ga.offset(p_last) == d2._ga.offset_list[-1] and d2._sn.n == 28
    assert d2._sn.param_offsets(d2.classification.weight_orig) == d2._ga.offset(d2.classification.weight_orig)
    tab[0].cols = 7  # inconsistent shape -> error status + message, no crash
    with pytest.raises(RuntimeError, match="bad shape"):
        N.call_nostream("spyr_sn_plan", tab, 2, C.byref(plan))


def test_c_abi_exports_every_declared_symbol():
    from semantic_pyramid_for_image_generation_b200 import _native as N
    header = open(os.path.join(ROOT, "include", "spyramid_b200.h")).read()
    declared = set(re.findall(r"\b(spyr_[a-z0-9_]+)\s*\(", header))
    declared -= {"spyr_conv_src", "spyr_conv_desc", "spyr_wgrad_desc", "spyr_sn_layer", "spyr_sn_plan_out",
                 "spyr_adam_chunk"}
    assert len(declared) >= 55
    handle = N.lib()
    missing = [name for name in sorted(declared) if not hasattr(handle, name)]
    assert not missing, missing
    assert declared == set(N.EXPORTED_SYMBOLS), declared ^ set(N.EXPORTED_SYMBOLS)
    assert handle.spyr_version() >= 100 and handle.spyr_last_error() is not None
    # descriptor structs must match the header's layout (sizes as the C compiler sees them)
    src = '#include "%s"\n#include <stdio.h>\nint main(){printf("%%zu %%zu %%zu %%zu", sizeof(spyr_conv_desc), ' \
          'sizeof(spyr_wgrad_desc), sizeof(spyr_sn_layer), sizeof(spyr_adam_chunk));return 0;}' % \
          os.path.join(ROOT, "include", "spyramid_b200.h")
    exe = os.path.join(ROOT, "build", "abi_sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-o", exe], input=src.encode(), check=True)
    import ctypes as C
    sizes = [int(v) for v in subprocess.run([exe], capture_output=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(N.ConvDesc), C.sizeof(N.WgradDesc), C.sizeof(N.SnLayer), C.sizeof(N.AdamChunk)]


def test_fused_adam_state_dict_is_torch_adam_compatible():
    from semantic_pyramid_for_image_generation_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(4))
    ref = torch.optim.Adam([p], lr=1e-5)
    p.grad = torch.ones(4)
    ref.step()
    mine = FusedAdam([p], lr=1e-5)
    mine.load_state_dict(ref.state_dict())  # a reference checkpoint's optimizer state loads (model_wrapper.py:215-223)
    sd = mine.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    assert sd["param_groups"][0]["lr"] == 1e-5 and sd["param_groups"][0]["betas"] == (0.9, 0.999)
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):  # CPU parameters: fails loudly instead of falling back
        mine.step()


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from semantic_pyramid_for_image_generation_b200 import distributed
    red = distributed.init_from_env("gloo")
    assert red.active and red.world == world and red.rank == rank

    from semantic_pyramid_for_image_generation_b200.engine import GradArena

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1000))
            self.b = torch.nn.Parameter(torch.zeros(200003))  # ragged against the bucket size
            self._ga = GradArena(self)
            self._last_grad_arena = None

    m = M()
    red.bucket_bytes = 1 << 18
    # (1) gradients that ARE views of the flat arena (what a backward from set_to_none gradients leaves): one bucketed
    # all-reduce sequence over the arena
    arena = torch.full((m._ga.total,), float(rank + 1))
    m._last_grad_arena = arena
    m.w.grad, m.b.grad = m._ga.views(arena, [True, True])
    assert red.grads_alias_arena(m, arena)
    red.average(m)
    assert torch.allclose(arena, torch.full_like(arena, (1 + world) / 2.0))
    assert torch.allclose(m.b.grad, torch.full_like(m.b.grad, (1 + world) / 2.0))
    # (2) gradients that do NOT alias the arena (zero_grad(set_to_none=False), accumulation, hooks): the arena is a dead
    # buffer -- it must be ignored and every .grad averaged on its own (silent divergence otherwise)
    m.w.grad = torch.full((1000,), float(10 * rank))
    m.b.grad = torch.full((200003,), float(rank))
    stale = arena.clone()
    assert not red.grads_alias_arena(m, arena)
    red.average(m)
    assert torch.allclose(m.w.grad, torch.full((1000,), 10.0 * (world - 1) / 2.0))
    assert torch.allclose(m.b.grad, torch.full((200003,), (world - 1) / 2.0))
    assert torch.equal(arena, stale)
    # (3) rank 0's parameters and buffers reach every rank (replaces DataParallel's per-forward replicate)
    with torch.no_grad():
        m.w.fill_(float(rank + 5))
    red.broadcast_module(m)
    assert torch.equal(m.w.detach(), torch.full((1000,), 5.0))
    assert red.max_over_ranks(float(rank), device="cpu") == float(world - 1)
    lo, hi = distributed.shard_range(41, rank, world)
    torch.save((lo, hi), os.path.join(out_dir, "shard_%d.pt" % rank))
    red.barrier()
    dist.destroy_process_group()


def test_gradient_reducer_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    shards = [torch.load(os.path.join(str(tmp_path), "shard_%d.pt" % r)) for r in range(2)]
    assert shards[0] == (0, 21) and shards[1] == (21, 41)  # contiguous, covering, sizes differ by at most one


def test_bench_reference_arm_prints_contract_line():
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                          "1", "--cpu-batch", "2"], capture_output=True, text=True, check=True, cwd=ROOT)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["workload"] and line["steps_timed"] == 1


def test_packed_mask_descriptors_round_trip():
    """Host half of the device input pipeline: descriptors -> flat tensors -> (host re-expansion) == direct expansion."""
    import random
    import numpy as np
    from semantic_pyramid_for_image_generation_b200 import input_pipeline as ip, misc
    random.seed(5)
    np.random.seed(5)
    descs = [misc.draw_mask_descriptor(p_random_mask=0.7) for _ in range(24)]
    packed = ip.PackedDescriptors(len(descs), pin=False).fill(descs)
    assert packed.stage.dtype == torch.int32 and packed.bitmaps.dtype == torch.uint8
    for b, d in enumerate(descs):
        n = int(packed.bitmap_hw[b])
        assert int(packed.stage[b]) == d.stage
        if d.bitmap is None:
            assert n == 0
            rebuilt = misc.MaskDescriptor(d.stage, None)
        else:
            assert n == d.bitmap.shape[0] and 0 < d.stage < 6
            rebuilt = misc.MaskDescriptor(d.stage, packed.bitmaps[b, :n * n].view(n, n).numpy())
        for got, want in zip(misc.expand_mask_descriptor(rebuilt), misc.expand_mask_descriptor(d)):
            assert torch.equal(got, want)
    with pytest.raises(ValueError):
        ip.PackedDescriptors(2, pin=False).fill(descs[:3])


def test_main_py_keeps_the_reference_flags_and_replaces_dataparallel_by_torchrun():
    """main.py mirrors reference main.py:4-42 (twelve flags, same defaults) and `--use_data_parallel` becomes a one-process-
    per-GPU torchrun launch instead of nn.DataParallel (main.py:91-94)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("spyr_main", os.path.join(ROOT, "main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    args = mod.build_parser().parse_args([])
    reference_defaults = dict(train=False, test=False, batch_size=20, lr=1e-05, channel_factor=1.0, device='cuda',
                              gpus_to_use='0', use_data_parallel=False, load_checkpoint=None,
                              load_pretrained_vgg16='pre_trained_models/vgg_places_365_fine_tuned.pt',
                              path_to_places365='places365_standard', epochs=50)
    for k, v in reference_defaults.items():
        assert getattr(args, k) == v, k
    env = dict(os.environ, SPYR_MAIN_DRY_RUN="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "--train", "--use_data_parallel", "--gpus_to_use",
                          "0,1,2,3", "--batch_size", "20"], capture_output=True, text=True, check=True, env=env, cwd=ROOT)
    cmd = out.stdout.strip()
    assert "torch.distributed.run" in cmd and "--nproc-per-node 4" in cmd and "--master-addr 127.0.0.1" in cmd
    assert cmd.endswith("--train --use_data_parallel --gpus_to_use 0,1,2,3 --batch_size 20")
