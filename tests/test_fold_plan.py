"""CPU checks of the batch-1 time-folding plan (zeronotesamba_b200.models.models.fold_plan)."""
import pytest

from zeronotesamba_b200.models.models import FOLD_SLOTS, TIME_HALO, fold_plan


def test_halo_is_the_receptive_field():
    kws = [11, 13, 15, 17, 19, 21, 23, 25]          # models.py:16-23
    assert TIME_HALO == sum((k - 1) // 2 for k in kws) == 68


@pytest.mark.parametrize("T", [313, 400, 626, 1250, 1876, 1877, 2500, 5000])
def test_plan_tiles_the_clip(T):
    plan = fold_plan(T)
    assert plan is not None
    w_s, delta, cuts = plan
    assert (FOLD_SLOTS - 1) * delta + w_s == T and w_s - delta >= 2 * TIME_HALO
    assert cuts[0] == 0 and cuts[-1] == T and all(a < b for a, b in zip(cuts, cuts[1:]))
    for s in range(FOLD_SLOTS):
        lo, hi = cuts[s] - s * delta, cuts[s + 1] - s * delta        # in segment coordinates
        assert 0 <= lo < hi <= w_s
        if s > 0:
            assert lo >= TIME_HALO                                     # away from the artificial left edge
        if s < FOLD_SLOTS - 1:
            assert hi <= w_s - TIME_HALO                               # away from the artificial right edge
    assert w_s * FOLD_SLOTS < 2.2 * T + 8 * 2 * TIME_HALO


def test_short_clips_are_not_folded():
    assert fold_plan(100) is None and fold_plan(137) is None
