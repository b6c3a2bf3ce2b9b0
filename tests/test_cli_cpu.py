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


def _load(name, *parts):
    spec_ = importlib.util.spec_from_file_location(name, os.path.join(PKG, *parts))
    mod = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mod)
    return mod


def test_r_iterative_flags_and_defaults_match_reference():
    """r_iterative/main.py:27-134: g_lis's flags minus cache / augment / upscaling / D-dropout / tolerant load,
    with its own interval defaults."""
    riter = _load("riter_main", "r_iterative", "main.py")
    opt = riter.build_parser().parse_args([])
    want = dict(REFERENCE_DEFAULTS, test_interval=1000, vis_interval=100, save_interval=2000)
    for absent in ("load_tolerant", "nb_cache_total", "nb_cache_lists", "cache_p_drop", "augment", "g_upscaling", "d_dropout"):
        want.pop(absent)
        assert not hasattr(opt, absent), absent
    for k, v in want.items():
        assert getattr(opt, k) == v, k
    with pytest.raises(SystemExit):
        riter.build_parser().parse_args(["--d_dropout", "0.1"])


def test_latent_codes_differ_across_ranks_and_iterations():
    """Data-parallel ranks must draw DIFFERENT latent codes (N ranks x B == one rank x N*B): the noise streams are
    keyed by the per-rank data seed dp.seed_everything returns, and by the iteration."""
    import sys
    sys.path.insert(0, PKG)
    from glis_b200 import dp
    seeds = [dp.seed_everything(1234, r) for r in range(4)]
    assert len(set(seeds)) == 4
    streams = [glis_main.NoiseSource.streams(s, 5) for s in seeds]
    assert len({st[0] for st in streams}) == 4 and len({st[1] for st in streams}) == 4
    assert all(st[0] != st[1] for st in streams)                        # z_d and z_g are different draws
    a, b = glis_main.NoiseSource.streams(seeds[0], 5), glis_main.NoiseSource.streams(seeds[0], 6)
    assert a[0][0] == b[0][0] and b[0][1] - a[0][1] == glis_main.NoiseSource.STRIDE >= 64 * 256 // 4


def test_side_work_does_not_consume_the_training_depth_draws():
    """Sample grids run on rank 0 only; `lis_depth` consumes draws even for a forced depth (as the reference's
    forward does), so without `private_depth_rng` rank 0's later training depths would shift against the other
    ranks'."""
    import random
    import sys
    sys.path.insert(0, PKG)
    from common.model import GeneratorLearnedInputSpace
    gens = [GeneratorLearnedInputSpace(16, 16, 4, 2, 8, "weight", 3, "fractional") for _ in range(2)]
    for g in gens:
        g.rng = random.Random(77)
    seq = [[], []]
    for it in range(40):
        if it % 5 == 0:                       # "rank 0" draws a sample grid now and then
            with glis_main.private_depth_rng(gens[0]):
                assert gens[0].lis_depth("all") == 3
        for g, s in zip(gens, seq):
            s.append((g.lis_depth(None), g.lis_depth(None)))
    assert seq[0] == seq[1] and len(set(seq[0])) > 3


def test_history_container_round_trips_in_the_reference_format():
    """`*_state.pt` carries the history PICKLED (g_lis/main.py:350-357) as common.plotting.History / LineGroup /
    Line objects with numpy arrays that grow in blocks of 500."""
    import pickle
    import sys
    sys.path.insert(0, PKG)
    from common.plotting import GROWTH_BY, History, Line, LineGroup
    h = glis_main.new_history(2)
    assert h.get_group_names() == ["loss-r-mix", "loss-g-mix", "loss-d-mix"]
    assert h.line_groups["loss-d-mix"].get_line_names() == ["train-d-real", "train-d-fake0", "train-d-fake1", "train-d-fake2"]
    for it in range(1, GROWTH_BY + 20):
        h.add_value("loss-g-mix", "train-g1", it, 0.5 + it)
    h.add_value("loss-g-mix", "train-g1", GROWTH_BY + 19, 0.0, average=True)       # folds into the last point
    s = h.to_string()
    assert isinstance(s, bytes) and pickle.loads(s).__class__ is History
    back = History.from_string(s)
    line = back.line_groups["loss-g-mix"].lines["train-g1"]
    assert isinstance(line, Line) and isinstance(back.line_groups["loss-g-mix"], LineGroup)
    assert line.last_index == GROWTH_BY + 18 and line.xs.shape[0] == 2 * GROWTH_BY
    assert line.get_xs()[-1] == GROWTH_BY + 19 and abs(line.get_ys()[-1] - (0.5 + GROWTH_BY + 19) / 2) < 1e-3
    assert line.counts[line.last_index] == 2 and back.get_max_x() == GROWTH_BY + 19
    assert line.xs.dtype.name == "int32" and line.ys.dtype.name == "float32" and line.counts.dtype.name == "uint16"
    # a state file as written by this repository in round 1 (plain list) still loads
    old = glis_main.history_from_state({"history": [(1, 0.7, 0.6, 0.5, [0.1, 0.2])]}, 2)
    assert old.line_groups["loss-r-mix"].lines["train-r1"].get_ys()[0] == pytest.approx(0.2)
