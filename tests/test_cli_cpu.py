"""The g_lis/main.py command line keeps the reference's flags and defaults."""
import importlib.util
import os

import pytest

from conftest import PKG

spec = importlib.util.spec_from_file_location("glis_main", os.path.join(PKG, "g_lis", "main.py"))
glis_main = importlib.util.module_from_spec(spec)
spec.loader.exec_module(glis_main)

REFERENCE_DEFAULTS = dict(  # g_lis/main.py:43-164 of the reference
    lsun_class="bedroom", batch_size=32, image_size=-1, width=-1, height=-1, crop_size=-1, crop_width=-1,
    crop_height=-1, code_size=128, nfeature=64, nlayer=-1, norm="none", save_path=None, load_path=None, lr=1e-4,
    test_interval=10000, test_lr=0.01, test_steps=50, vis_interval=2000, vis_size=10, vis_row=-1, vis_col=-1,
    save_interval=5000, niter=50000, final_test=False, ls=False, output_scale=False, net="best", lambda_r=0.9,
    spatial_dropout_r=0, r_iterations=3, always_train_all=False, load_tolerant=False, nb_cache_total=0,
    nb_cache_lists=1, cache_p_drop=0.1, augment="none", g_upscaling="fractional", d_dropout=0)


def test_flags_and_defaults_match_reference():
    opt = glis_main.build_parser().parse_args([])
    for k, v in REFERENCE_DEFAULTS.items():
        assert getattr(opt, k) == v, k
    assert opt.synthetic is False and opt.seed == 1234


@pytest.mark.parametrize("size,levels", [(32, 3), (64, 4), (80, 4), (160, 5)])
def test_automatic_level_count(size, levels):
    opt = glis_main.resolve_geometry(glis_main.build_parser().parse_args(["--image_size", str(size)]))
    assert (opt.width, opt.height, opt.nlayer) == (size, size, levels)
    assert opt.vis_row == opt.vis_col == 10


def test_rejects_missing_size_and_unsupported_norm():
    with pytest.raises(ValueError):
        glis_main.resolve_geometry(glis_main.build_parser().parse_args([]))
    with pytest.raises(SystemExit):
        glis_main.main(["--image_size", "32", "--norm", "batch", "--synthetic", "--save_path", "/tmp/x"])
