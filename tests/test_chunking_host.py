"""Host logic of the frame chunking of a DiffTRe pass (no GPU): device-resident frames go out in launch groups of FRAME_CHUNK,
frames streamed from pinned host memory start with one wave and ramp up gently (each chunk's copy must fit behind the
evaluation of the frames before it: a frame's copy takes ~0.75 of its evaluation, so growth stays below ~1.5x)."""

import pytest

from mythos_b200.energy import functional


@pytest.mark.parametrize("n_frames", [1, 100, 148, 300, 1024, 2048, 4096, 8192, 20000])
def test_chunks_cover_the_frames_once_in_order(n_frames):
    for streamed in (False, True):
        ch = functional._chunks(n_frames, functional.CellListPairs, streamed)
        assert ch[0].start == 0 and ch[-1].stop == n_frames
        assert all(a.stop == b.start for a, b in zip(ch, ch[1:]))
        assert all(c.stop > c.start for c in ch)


def test_resident_frames_use_the_regular_launch_group():
    ch = functional._chunks(8192, functional.CellListPairs, False)
    assert [c.stop - c.start for c in ch[:-1]] == [functional.FRAME_CHUNK] * (len(ch) - 1)
    assert functional.FRAME_CHUNK % 148 == 0


def test_streamed_frames_ramp_up_from_one_wave():
    sizes = [c.stop - c.start for c in functional._chunks(8192, functional.CellListPairs, True)]
    assert sizes[0] == functional.STREAM_FIRST_CHUNK
    assert max(sizes) <= functional.STREAM_CHUNK
    before = 0
    for a, b in zip(sizes, sizes[1:-1]):
        before += a
        assert b <= 2 * a  # never more than the doubling of the earlier scheme ...
        assert b <= 1.5 * before + functional.STREAM_FIRST_CHUNK  # ... and at most ~1.5x the frames already under way
    # a rank's whole block smaller than one regular chunk (8 GPUs: 1024 frames each) still starts small and leaves no sliver
    small = [c.stop - c.start for c in functional._chunks(1024, functional.CellListPairs, True)]
    assert small[0] == functional.STREAM_FIRST_CHUNK and min(small) >= functional.STREAM_FIRST_CHUNK // 2 and sum(small) == 1024
