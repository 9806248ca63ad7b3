"""Metrics / image output after the sampler path (main_sampling.py:488-519, 735-742)."""
import io
import os

import numpy as np
import pytest
import torch


def test_png_encoder_roundtrip_and_colormap():
    from nhmc_b200 import metrics
    from PIL import Image
    g = torch.Generator().manual_seed(0)
    img = torch.rand(3, 17, 23, generator=g)
    u8 = metrics.to_uint8_hwc(img.clone()).numpy()
    back = np.asarray(Image.open(io.BytesIO(metrics.encode_png(u8))).convert("RGB"))
    assert np.array_equal(back, u8)
    # same quantisation as torchvision.utils.save_image: mul(255).add_(0.5).clamp_(0, 255).to(uint8)
    assert np.array_equal(u8, img.mul(255).add(0.5).clamp(0, 255).permute(1, 2, 0).to(torch.uint8).numpy())
    grey = (torch.rand(9, 5, generator=g) * 255).to(torch.uint8).numpy()
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(metrics.encode_png(grey)))), grey)
    # matplotlib's 'hot': black -> red -> yellow -> white
    cm = metrics.hot_colormap(torch.tensor([[0.0, 0.365079, 0.746032, 1.0]]))
    np.testing.assert_allclose(cm[:, 0, :].numpy(), [[0, 1, 1, 1], [0, 0, 1, 1], [0, 0, 0, 1]], atol=1e-6)


def test_psnr_and_std_dev_map_follow_the_reference_formulas():
    """main_sampling.py:494-497 (x.std(dim=0), mean over channels, min-max) and :517-519 (10 log10(1/mse) on [0,1] images)."""
    from nhmc_b200 import metrics
    g = torch.Generator().manual_seed(1)
    xs = torch.randn(6, 3, 8, 8, generator=g)
    orig = torch.rand(3, 8, 8, generator=g) * 2 - 1
    inv = lambda t: torch.clamp((t + 1.0) / 2.0, 0.0, 1.0)                         # datasets/__init__.py:214-223
    x = torch.stack([inv(t) for t in xs])
    ref_psnr = [float(10 * torch.log10(1 / torch.mean((x[j] - inv(orig)) ** 2))) for j in range(6)]
    sm = metrics.summarize(xs, orig)
    np.testing.assert_allclose(sm["psnr"], ref_psnr, rtol=1e-6)
    assert abs(sm["psnr_mean"] - np.mean(ref_psnr)) < 1e-5 and abs(sm["psnr_std"] - np.std(ref_psnr, ddof=1)) < 1e-5
    std = x.std(dim=0); std_mean = std.mean(dim=0)
    ref_plot = (std_mean - std_mean.min()) / (std_mean.max() - std_mean.min())
    norm_map, raw = metrics.std_dev_map(xs)
    assert torch.allclose(norm_map, ref_plot, atol=1e-6) and torch.allclose(raw, std_mean, atol=1e-7)


@pytest.mark.gpu
def test_async_image_writer_keeps_the_sampler_stream_free(tmp_path):
    """The writer queues its device -> host copies on its own stream; files appear with the right pixels."""
    from nhmc_b200 import metrics
    from PIL import Image
    w = metrics.AsyncImageWriter("cuda")
    g = torch.Generator().manual_seed(2)
    imgs = [torch.rand(3, 32, 48, generator=g).cuda() for _ in range(12)]
    for i, im in enumerate(imgs):
        w.save(im, str(tmp_path / ("a%d.png" % i)))
        im.add_(0.0)
    w.save(torch.rand(16, 16, generator=g).cuda(), str(tmp_path / "grey.png"))
    assert w.close() == 13
    for i, im in enumerate(imgs):
        back = np.asarray(Image.open(tmp_path / ("a%d.png" % i)).convert("RGB"))
        assert np.array_equal(back, metrics.to_uint8_hwc(im.cpu().clone()).numpy())
    assert Image.open(tmp_path / "grey.png").size == (16, 16)


@pytest.mark.gpu
def test_cli_multi_image_run_writes_metrics_and_images(tmp_path):
    """Two synthetic images x two chains through the command-line driver (one HMC iteration each, full ffhq_10m network):
    (image, chain) work items batched into one sampler call, accepted-epoch PNGs through the async writer, metrics.json,
    std-dev maps only when there are >= 2 samples (none after one iteration)."""
    import json
    from nhmc_b200 import cli
    out = str(tmp_path / "run")
    cli.main(["--dataset", "ffhq", "--algo", "hmc", "--timesteps", "3", "--deg", "sr4", "--sigma_0", "0.05", "-i", out, "--tau", "0.2",
              "--epsilon", "0.05", "--chains", "2", "--n_images", "2", "--synthetic", "--max_iters", "1", "--save_progress", "--ni"])
    files = set(os.listdir(out))
    assert {"orig_synthetic0.png", "orig_synthetic1.png"} <= files
    assert any(f.startswith("y0_") for f in files)
    # one iteration collects nothing (samples start at epoch 80): no metrics, but the run completes and the writer drained
    assert "metrics.json" not in files or json.load(open(os.path.join(out, "metrics.json")))["chains"] == 2


@pytest.mark.gpu
def test_cli_metrics_and_std_dev_map_from_collected_samples(tmp_path, monkeypatch):
    """The part of sample_image() after the sampler (main_sampling.py:488-551) on stubbed sampler output: per-image PSNR
    mean / sample std over all samples of all chains, the final image, the std-dev map (PNG + raw .npy)."""
    import json
    import types
    from nhmc_b200 import cli, metrics
    out = str(tmp_path / "run")

    def fake_hmc(x, n, b, seq, seq_next, algo, opt, y_0, H_funcs, x_orig, **kw):
        g = torch.Generator(device="cuda").manual_seed(int(x_orig.abs().sum() * 1000) % 1000)
        outs = [(x_orig[j:j + 1] + 0.1 * torch.randn(20, *x_orig.shape[1:], device="cuda", generator=g)).clamp(-1, 1) for j in range(n)]
        return types.SimpleNamespace(collected=lambda: outs)
    monkeypatch.setattr(cli.sampler, "hmc", fake_hmc)
    monkeypatch.setattr(cli, "init_model", lambda *a, **k: None)
    cli.main(["--dataset", "ffhq", "--algo", "hmc", "--timesteps", "3", "--deg", "sr4", "--sigma_0", "0.05", "-i", out,
              "--chains", "3", "--n_images", "2", "--batch", "4", "--synthetic"])
    rep = json.load(open(os.path.join(out, "metrics.json")))
    assert rep["chains"] == 3 and rep["samples_per_chain"] == 20 and len(rep["images"]) == 2
    for im in rep["images"]:
        assert im["n_samples"] == 60 and 24.0 < im["psnr_mean"] < 28.0 and 0 < im["psnr_std"] < 0.5      # noise std 0.1 on [-1,1] -> 26 dB
        assert os.path.isfile(os.path.join(out, "%s.png" % im["image"]))
        assert os.path.isfile(os.path.join(out, "std_dev_map_%s.png" % im["image"]))
        raw = np.load(os.path.join(out, "std_dev_map_%s.npy" % im["image"]))
        assert raw.shape == (256, 256) and 0.03 < raw.mean() < 0.06                                    # std 0.1 / 2 in [0,1] units
    assert tuple(torch.load(os.path.join(out, "hmc_samples.pt")).shape) == (6, 20, 3, 256, 256)
