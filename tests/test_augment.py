"""Fused augment+normalize+patchify: the native host sampler and the CUDA kernel against the numpy oracle
(oracle/augment_oracle.py). The contract is BIT-EXACT (integer crop/flip/erase indices, uint8 pixels, bf16 bits)."""
import numpy as np
import pytest
import torch

from oracle import augment_oracle as A


def smooth_images(B, H, W, seed):
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, (B, H // 8 + 2, W // 8 + 2, 3)).astype(np.float32)
    t = torch.from_numpy(small).permute(0, 3, 1, 2)
    up = torch.nn.functional.interpolate(t, size=(H, W), mode="bicubic", align_corners=False)
    noise = torch.from_numpy(rng.normal(0, 12, (B, 3, H, W)).astype(np.float32))
    return (up + noise).clamp(0, 255).permute(0, 2, 3, 1).contiguous().to(torch.uint8).numpy()


# ---------------------------------------------------------------------------------------------- CPU (no GPU needed)
def test_native_sampler_matches_oracle_sampler(lib):
    from touhouimageclassification_b200.augment import sample_params
    for (seed, first, H, W, recipe) in ((7, 100, 256, 256, "full"), (1, 0, 300, 200, "full"), (5, 12345, 64, 640, "generalization"),
                                        (2, 5, 256, 256, "diversity"), (3, 7, 256, 256, "grey"), (4, 9, 256, 300, "none")):
        ints, floats = sample_params(seed, first, 200, H, W, 224, recipe)
        for b in range(200):
            i2, f2 = A.sample_params(seed, first + b, H, W, 224, recipe)
            assert np.array_equal(ints[b], i2) and np.array_equal(floats[b], f2), (seed, b)
        top, left, h, w = ints[:, 0], ints[:, 1], ints[:, 2], ints[:, 3]
        assert (top >= 0).all() and (left >= 0).all() and (top + h <= H).all() and (left + w <= W).all()
        assert (ints[:, 10] + ints[:, 12] <= 224).all() and (ints[:, 11] + ints[:, 13] <= 224).all()
        assert np.array_equal(np.sort(ints[:, 5:9], axis=1), np.tile(np.arange(4), (200, 1)))  # a permutation


def test_sharded_sample_indices_tile_the_global_batch(lib):
    """GpuAugment.shard(rank, world): rank r's batch of B in step k draws the parameters of global samples
    k * world * B + r * B ... + B -- together the ranks reproduce the one-process stream, and no two ranks share a draw."""
    from touhouimageclassification_b200.augment import GpuAugment, sample_params
    B, world, seed = 6, 4, 13
    whole, _ = sample_params(seed, 0, 3 * world * B, 256, 256, 224, "full")
    for rank in range(world):
        aug = GpuAugment(seed=seed).shard(rank, world)
        for step in range(3):
            first = aug.samples_seen + aug.rank * B            # what _prepare computes for a batch of B
            aug.samples_seen += B * aug.world_size
            ints, _ = sample_params(seed, first, B, 256, 256, 224, "full")
            lo = step * world * B + rank * B
            assert first == lo and np.array_equal(ints, whole[lo:lo + B]), (rank, step)


def test_sampler_statistics_follow_torchvision_rules(lib):
    from touhouimageclassification_b200.augment import sample_params
    ints, floats = sample_params(11, 0, 4000, 256, 256)
    assert abs(ints[:, 4].mean() - 0.5) < 0.03            # flip p = 0.5
    assert abs(ints[:, 9].mean() - 0.2) < 0.03            # grayscale p = 0.2
    assert abs((ints[:, 12] > 0).mean() - 0.5) < 0.03     # erasing p = 0.5
    area = ints[:, 2] * ints[:, 3] / (256.0 * 256.0)
    assert 0.07 < area.min() and area.max() <= 1.0 and abs(area.mean() - 0.5) < 0.06
    ratio = ints[:, 3] / ints[:, 2]
    assert ratio.min() > 0.70 and ratio.max() < 1.40
    assert floats[:, :3].min() >= 0.8 and floats[:, :3].max() <= 1.2 and np.abs(floats[:, 3]).max() <= 0.1


def test_oracle_colour_ops_follow_torchvision_tensor_kernels():
    TF = pytest.importorskip("torchvision.transforms.v2.functional")
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.integers(0, 256, (3, 224, 224), dtype=np.uint8))
    xn = x.permute(1, 2, 0).numpy()

    def diff(ours, ref):
        d = np.abs(ours.astype(int) - ref.permute(1, 2, 0).numpy().astype(int))
        return d.max(), (d > 0).mean()
    f = np.float32
    assert diff(A.adjust_brightness(xn, f(1.13)), TF.adjust_brightness(x, float(f(1.13)))) == (0, 0.0)
    assert diff(A.adjust_contrast(xn, f(0.87)), TF.adjust_contrast(x, float(f(0.87)))) == (0, 0.0)
    mx, frac = diff(A.adjust_saturation(xn, f(1.17)), TF.adjust_saturation(x, float(f(1.17))))
    assert mx <= 1 and frac < 1e-3          # torch may fuse a multiply-add; the oracle never does
    for h in (0.07, -0.09):
        mx, frac = diff(A.adjust_hue(xn, f(h)), TF.adjust_hue(x, float(f(h))))
        assert mx <= 1 and frac < 1e-3
    g = A.gray_floor(xn.astype(np.float32)).astype(np.uint8)
    assert diff(np.repeat(g[..., None], 3, -1), TF.rgb_to_grayscale(x, 3)) == (0, 0.0)
    # antialiased bilinear resize: torchvision uses fixed-point weights on uint8, the oracle float32: +-1 LSB
    img = smooth_images(1, 256, 256, 3)[0]
    t = torch.from_numpy(img).permute(2, 0, 1)
    for (top, left, h, w) in ((10, 20, 200, 180), (0, 0, 256, 256), (30, 40, 70, 90)):
        mx, _ = diff(A.resized_crop(img, top, left, h, w), TF.resized_crop(t, top, left, h, w, [224, 224], antialias=True))
        assert mx <= 1


def test_oracle_identity_and_patch_layout():
    img = smooth_images(1, 224, 224, 1)[0]
    ints = np.array([0, 0, 224, 224, 0, 0, 1, 2, 3, 0, 0, 0, 0, 0, 0, 0], np.int32)   # full crop, nothing else
    pix, tok = A.augment_one(img, ints, np.ones(4, np.float32))
    assert np.array_equal(pix, img)                      # scale 1: the triangle filter is the identity
    x = (img.astype(np.float32) / np.float32(255.0) - A.MEAN) / A.STD
    # token (gy=3, gx=5), channel 1, py=2, px=7 sits at column 1*256 + 2*16 + 7
    assert tok[3 * 14 + 5, 256 + 2 * 16 + 7] == A.f32_to_bf16_bits(x[3 * 16 + 2, 5 * 16 + 7, 1:2])[0]
    flipped, _ = A.augment_one(img, np.array([0, 0, 224, 224, 1, 0, 1, 2, 3, 0, 0, 0, 0, 0, 0, 0], np.int32), np.ones(4, np.float32))
    assert np.array_equal(flipped, img[:, ::-1])
    erased, tok = A.augment_one(img, np.array([0, 0, 224, 224, 0, 0, 1, 2, 3, 0, 10, 20, 30, 40, 0, 0], np.int32), np.ones(4, np.float32))
    assert (erased[10:40, 20:60] == 0).all() and np.array_equal(erased[40:], img[40:])


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("H,W,seed,recipe", [(256, 256, 1, "full"), (256, 256, 2, "full"), (300, 200, 3, "full"),
                                             (224, 224, 4, "generalization"), (96, 128, 5, "full"),
                                             (256, 256, 6, "diversity"), (256, 256, 7, "grey"), (256, 320, 8, "none")])
def test_kernel_is_bit_exact_against_oracle(H, W, seed, recipe):
    from touhouimageclassification_b200.augment import GpuAugment
    B = 12
    imgs = smooth_images(B, H, W, seed)
    aug = GpuAugment(seed=seed, recipe=recipe)
    patches, pixels, (ints, floats) = aug(torch.from_numpy(imgs).cuda(), first_sample=1000, return_pixels=True)
    ref_pix, ref_tok = A.augment_batch(imgs, seed, first_sample=1000, recipe=recipe)
    assert np.array_equal(pixels.cpu().numpy(), ref_pix)
    got = patches.view(torch.int16).cpu().numpy().view(np.uint16)
    assert np.array_equal(got, ref_tok)


@pytest.mark.gpu
def test_kernel_every_jitter_order_and_edge_boxes():
    """Hand-made parameter records: all 24 op orders, 1-pixel-high crops, full-frame crops, max erase boxes."""
    import ctypes
    import itertools
    from touhouimageclassification_b200 import _lib
    from touhouimageclassification_b200.augment import IMAGENET_MEAN, IMAGENET_STD
    H = W = 256
    perms = list(itertools.permutations(range(4)))
    B = len(perms) + 4
    imgs = smooth_images(B, H, W, 9)
    ints = np.zeros((B, 16), np.int32)
    floats = np.zeros((B, 4), np.float32)
    rng = np.random.default_rng(5)
    for b, pm in enumerate(perms):
        ints[b] = [rng.integers(0, 30), rng.integers(0, 30), 200, 220, b & 1, *pm, (b % 5 == 0), 0, 0, 0, 0, 1, 0]
        floats[b] = [rng.uniform(.8, 1.2), rng.uniform(.8, 1.2), rng.uniform(.8, 1.2), rng.uniform(-.1, .1)]
    n = len(perms)
    ints[n + 0] = [0, 0, 256, 256, 0, 0, 1, 2, 3, 0, 0, 0, 223, 223, 1, 0]      # full frame, almost everything erased
    ints[n + 1] = [255, 0, 1, 256, 1, 3, 2, 1, 0, 1, 5, 5, 1, 1, 1, 0]          # 1-row crop, gray
    ints[n + 2] = [0, 255, 256, 1, 0, 1, 0, 3, 2, 0, 0, 0, 0, 0, 1, 0]          # 1-column crop
    ints[n + 3] = [100, 100, 64, 18, 1, 2, 3, 0, 1, 0, 200, 200, 24, 24, 1, 0]  # strong anisotropic upscale
    floats[n:] = [[1.2, 0.8, 1.2, 0.1], [0.8, 1.2, 0.8, -0.1], [1.0, 1.0, 1.0, 0.0], [1.1, 0.9, 1.05, 0.05]]
    dev = "cuda"
    patches = torch.empty((B * 196, 768), dtype=torch.bfloat16, device=dev)
    pixels = torch.empty((B, 224, 224, 3), dtype=torch.uint8, device=dev)
    mean = (ctypes.c_float * 3)(*[float(np.float32(m)) for m in IMAGENET_MEAN])
    std = (ctypes.c_float * 3)(*[float(np.float32(s)) for s in IMAGENET_STD])
    img_d = torch.from_numpy(imgs).to(dev)
    i_d, f_d = torch.from_numpy(ints).to(dev), torch.from_numpy(floats).to(dev)
    _lib.check(_lib.load().tic_augment_patchify(
        ctypes.c_void_p(img_d.data_ptr()), B, H, W, ctypes.c_void_p(i_d.data_ptr()), ctypes.c_void_p(f_d.data_ptr()), 224,
        mean, std, ctypes.c_void_p(patches.data_ptr()), ctypes.c_void_p(pixels.data_ptr()),
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    got_pix = pixels.cpu().numpy()
    got_tok = patches.view(torch.int16).cpu().numpy().view(np.uint16)
    for b in range(B):
        ref_pix, ref_tok = A.augment_one(imgs[b], ints[b], floats[b])
        assert np.array_equal(got_pix[b], ref_pix), b
        assert np.array_equal(got_tok[b * 196:(b + 1) * 196], ref_tok), b


@pytest.mark.gpu
def test_augmented_patches_feed_the_engine():
    """uint8 batch -> fused augmentation -> engine forward from patch rows == engine forward from the same pixels."""
    from touhouimageclassification_b200.augment import GpuAugment, IMAGENET_MEAN, IMAGENET_STD
    from touhouimageclassification_b200.model import ViTConfig, ViTForImageClassification
    m = ViTForImageClassification(ViTConfig()).cuda().eval()
    imgs = torch.from_numpy(smooth_images(4, 256, 256, 2)).cuda()
    patches, pixels, _ = GpuAugment(seed=3)(imgs, first_sample=0, return_pixels=True)
    x = pixels.permute(0, 3, 1, 2).float() / 255.0
    x = (x - torch.tensor(IMAGENET_MEAN, device="cuda").view(1, 3, 1, 1)) / torch.tensor(IMAGENET_STD, device="cuda").view(1, 3, 1, 1)
    with torch.no_grad():
        a = m.engine_forward(patches=patches)
        b = m.engine_forward(x.contiguous())
    assert torch.allclose(a, b, atol=2e-2, rtol=2e-2)
    assert GpuAugment(seed=3).__call__(imgs, first_sample=0).equal(patches)       # deterministic in (seed, index)
    with pytest.raises(ValueError):
        GpuAugment()(imgs.cpu())
    # data-parallel shards: rank r's local batch is the slice [r*B, (r+1)*B) of the global batch, step after step
    whole = GpuAugment(seed=3)
    r0, r1 = GpuAugment(seed=3).shard(0, 2), GpuAugment(seed=3).shard(1, 2)
    for _ in range(2):
        g = whole(imgs)
        assert torch.equal(torch.cat([r0(imgs[:2]), r1(imgs[2:])]), g)


@pytest.mark.gpu
def test_augmented_tensor_is_totensor_normalize_of_the_augmented_pixels():
    """GpuAugment.tensor == ToTensor + Normalize (ntrain.py:110-111) of the uint8 image the same draws produce, bit for
    bit in fp32, and its bf16 patch rows are the ones the patch-row path writes."""
    from touhouimageclassification_b200 import ops
    from touhouimageclassification_b200.augment import GpuAugment, IMAGENET_MEAN, IMAGENET_STD
    imgs = torch.from_numpy(smooth_images(5, 256, 256, 7)).cuda()
    patches, pixels, _ = GpuAugment(seed=11)(imgs, first_sample=40, return_pixels=True)
    t = GpuAugment(seed=11).tensor(imgs, first_sample=40)
    assert t.shape == (5, 3, 224, 224) and t.dtype == torch.float32
    # on the CPU, where the reference's transform runs (torch's CUDA kernels divide by a scalar through its reciprocal)
    x = pixels.cpu().permute(0, 3, 1, 2).float().div(255.0)
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(1, 3, 1, 1)
    assert torch.equal(t.cpu(), x.sub(mean).div(std))
    assert torch.equal(ops.patchify_f32(t), patches)


@pytest.mark.gpu
def test_fused_training_step_takes_uint8_thumbnails():
    """ntrain's step with the train transform on the device: uint8 NHWC in, CutMix / MixUp, engine step, finite loss;
    with mixing off the patch rows come straight from the augmentation kernel."""
    from touhouimageclassification_b200.augment import GpuAugment
    from touhouimageclassification_b200.ntrain import ViTLModule
    torch.manual_seed(0)
    imgs = torch.from_numpy(smooth_images(4, 256, 256, 5)).cuda()
    y = torch.tensor([1, 0, 3, 2], device="cuda")
    for mix in (True, False):
        mod = ViTLModule(5, False, "google/vit-base-patch16-224", 1e-3, 0.01, enable_mixup=mix, fused_optimizer=True).cuda().train()
        opt = mod.configure_optimizers()
        before = mod.vit.classifier.weight.detach().clone()
        loss = mod.fused_training_step((imgs, y), opt, augment=GpuAugment(seed=1))
        assert torch.isfinite(loss) and not torch.equal(before, mod.vit.classifier.weight)
        with pytest.raises(ValueError):
            mod.fused_training_step((imgs, y), opt)


@pytest.mark.gpu
def test_inference_transform_on_gpu_matches_reference_recipe():
    """serve.preprocess_u8 == Resize((224,224)) -> ToTensor -> Normalize(dataset mean/std) (preprocess.py:73-77):
    bit-exact against the oracle's 'none' recipe, and within 1 LSB of torchvision's own uint8 resize."""
    TF = pytest.importorskip("torchvision.transforms.v2.functional")
    from touhouimageclassification_b200 import serve as S
    imgs = smooth_images(3, 256, 256, 11)
    mean, std = [0.61, 0.55, 0.52], [0.31, 0.30, 0.29]
    patches = S.preprocess_u8(torch.from_numpy(imgs).cuda(), mean, std)
    A_MEAN, A_STD = A.MEAN.copy(), A.STD.copy()
    try:
        A.MEAN[:] = np.array(mean, np.float32)
        A.STD[:] = np.array(std, np.float32)
        ref_pix, ref_tok = A.augment_batch(imgs, 0, first_sample=0, recipe="none")
    finally:
        A.MEAN[:] = A_MEAN
        A.STD[:] = A_STD
    assert np.array_equal(patches.view(torch.int16).cpu().numpy().view(np.uint16), ref_tok)
    tv = TF.resize(torch.from_numpy(imgs).permute(0, 3, 1, 2), [224, 224], antialias=True).permute(0, 2, 3, 1).numpy()
    assert np.abs(tv.astype(int) - ref_pix.astype(int)).max() <= 1
