"""
Frame-time enumeration, re-expressed from the reference's own tests
(cuburn/tests/test_profile.py:12-27), plus the --still quirk.
"""
import numpy as np

from cuburn_b200 import profile


def _get_profile(args=None):
    return profile.get_from_args(profile.add_args().parse_args(args or []))


def test_enumerate_times():
    name, prof = _get_profile()
    gprof = profile.wrap(prof, {"type": "edge"})
    frames = list(profile.enumerate_times(gprof))
    frame_times = np.linspace(0, 1 - 1 / 720., 720) + 0.5 / 720
    assert len(frames) == 720
    assert [f[0] for f in frames] == list(range(1, 721))
    assert np.allclose([f[1][0] for f in frames], frame_times, rtol=0, atol=1e-15)
    assert name == '720p'


def test_nframes_for_sharding_equal():
    name, prof = _get_profile(['-P', '720p', '--fps=1', '--duration=5', '--shard=5'])
    gprof = profile.wrap(prof, {"type": "edge"})
    frames = list(profile.enumerate_times(gprof))
    frame_times = np.linspace(0, 1 - 1 / 5., 5) + 0.5 / 5
    assert len(frames) == 1
    assert frames[0][0] == 1
    assert sorted(frames[0][1]) == sorted(frame_times)


def test_still_renders_frame_two():
    name, prof = _get_profile(['-P', '1080p', '--still'])
    gprof = profile.wrap(prof, {"type": "animation"})
    frames = profile.enumerate_times(gprof)
    assert len(frames) == 1 and frames[0][0] == 2
    assert abs(frames[0][1][0] - 1.5 / 720) < 1e-15
    assert gprof.frame_width(0.5) == 0.0
    assert (gprof.width, gprof.height) == (1920, 1080)


def test_profile_overrides_and_refscalars():
    name, prof = _get_profile(['-P', 'preview', '--spp', '100', '--width', '320'])
    gnm = {"type": "animation", "camera": {"spp": 2.0, "scale": 0.5}}
    gprof = profile.wrap(prof, gnm)
    assert gprof.width == 320 and gprof.height == 360 and gprof.skip == 1
    assert gprof.spp(0.3) == 200.0                      # profile x genome multiplier
    assert gprof.filters.logscale.scale(0.1) == 0.5     # camera.scale dragged in
    assert gprof.filter_order == ['bilateral', 'logscale', 'smearclip']
    assert gprof.filters.bilateral.spatial_std(0.5) == 6


def test_enumerate_jobs_names(tmp_path):
    args = profile.add_args().parse_args(['-o', str(tmp_path), '--suffix', '_x', '--still'])
    args.name = None
    name, prof = profile.get_from_args(args)
    gprof = profile.wrap(prof, {"type": "animation"})
    jobs = profile.enumerate_jobs(gprof, 'flame', args)
    assert len(jobs) == 1
    assert jobs[0][0] == str(tmp_path / 'flame_00002_x')
    open(jobs[0][0] + '.jpg', 'w').close()
    assert profile.enumerate_jobs(gprof, 'flame', args, resume=True) == []
