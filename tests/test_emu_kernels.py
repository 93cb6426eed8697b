"""CPU suite, part 3: the kernel SOURCES (csrc/*.cu), compiled with g++ against the test-only CUDA
emulation in tests/emu/, checked against the oracle on small inputs -- tiling, halos, ring buffers,
batching and reductions are exercised before any GPU time is spent.  The real parity tests are the
`-m gpu` ones; this file only proves the logic, not the sm_100a build."""
import numpy as np
import pytest

import parity_checks as pc
from funscript_flow_b200 import api, runner
from funscript_flow_b200.synth import make_clip


@pytest.mark.parametrize("size", [(150, 101), (64, 48)])
def test_stages(emu_ctx, size):
    pc.check_stages(emu_ctx, *size)


@pytest.mark.parametrize("size", [(256, 256), (333, 217)])
def test_farneback_vs_cv2(emu_ctx, size):
    pc.check_farneback_vs_cv2(emu_ctx, *size)


def test_reduction_known_answers(emu_ctx, golden_dir):
    pc.check_reductions_kat(emu_ctx, golden_dir)


def test_golden_pairs(emu_ctx, golden_dir):
    pc.check_golden_pairs(emu_ctx, golden_dir, names=("b", "c"))


def test_golden_bracket(emu_ctx, golden_dir):
    pc.check_golden_bracket(emu_ctx, golden_dir)


def test_batch_independence(emu_ctx):
    pc.check_batch_independence(emu_ctx, 120, 72, n_frames=12)


def test_edge_brackets(emu_ctx):
    pc.check_edge_brackets(emu_ctx)


def test_process_frames_brackets(emu_ctx):
    """Bracket semantics of F:1145-1153: no pair across brackets, trailing 1-frame bracket dropped."""
    api.set_context(emu_ctx)
    clip = make_clip(96, 64, 11, seed=6, period=7.0, amplitude=0.3)
    prm = {"batch_size": 5, "detrend_window": 2.0, "norm_window": 3.0, "keyframe_reduction": False}
    acts, series = runner.process_frames(clip, 30.0, prm, ctx=emu_ctx, return_series=True)
    # brackets [0,5) [5,10) [10,11) -> 4 + 4 + 0 pairs, stamped with the first frame of each pair
    assert series["frame_indices"].tolist() == [0, 1, 2, 3, 5, 6, 7, 8]
    assert len(acts) == 8
    one = api.process_bracket(clip[5:10], {}, ctx=emu_ctx)
    assert np.array_equal(series["values"][4:], one["scalar"])
