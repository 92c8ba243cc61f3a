"""
bench.py's control flow and the JSON line it prints, on the CPU: the whole N = 1 run
(headline workload, extras, roofline microbenchmark, end-to-end leg) against the recording
stand-in for the library, whose events measure 1 ms per kernel launch issued between them.
Checks the contract's keys and the arithmetic that turns times into the reported figures --
not the figures.
"""
import importlib
import json
import sys

import pytest

import fake_native

pytestmark = pytest.mark.production_schedule


def test_default_run_prints_one_line_with_the_contract_keys(built, monkeypatch, capsys):
    from cuburn_b200 import render
    lib = fake_native.install(monkeypatch)
    monkeypatch.setattr(render.Renderer, '_modrefs', {})
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--steps', '4', '--warmup', '1',
                                      '--no-cpu-baseline'])
    monkeypatch.delenv('RANK', raising=False)
    monkeypatch.delenv('WORLD_SIZE', raising=False)
    sys.modules.pop('bench', None)
    bench = importlib.import_module('bench')
    bench.main()
    fake_native.uninstall()
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith('{')]
    assert len(out) == 1
    d = json.loads(out[0])

    # the base contract
    assert d['metric'] == 'ifs_iterations_per_second' and d['unit'] == 'iterations/s'
    assert (d['n_gpus'], d['steps'], d['warmup']) == (1, 4, 3)         # warm-up is at least 3
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['dtype'] == 'f32' and d['data'] == 'synthetic'
    assert d['config']['workload'].startswith('1080p still, G6F')
    assert 'L2 flushed' in d['config']['l2'] and 'model' not in d['config']
    n = 1920 * 1080 * 2000
    assert d['config']['samples_per_step'] == n
    # on the stand-in's clock (1 ms per launch) a step takes its number of launches:
    # 41 in steady state, 44 when the frame runs the hot-bin pilot + scan
    ms = d['ms_per_step']
    assert 41.0 <= ms <= 44.0 and d['value'] == pytest.approx(n / (ms * 1e-3))
    assert ms == pytest.approx(d['gpu_launches_per_step'])
    # end to end: the same frame plus nothing that launches (copies are not kernels)
    assert 41.0 <= d['e2e']['ms_per_step'] <= 44.0
    assert d['e2e']['value'] == pytest.approx(n / (d['e2e']['ms_per_step'] * 1e-3))
    assert d['e2e']['d2h_bytes_per_step'] == 1920 * 1080 * 4
    assert 30000 < d['e2e']['h2d_bytes_per_step'] < 40000
    # 41 launches per steady-state frame; the genome's first frame (in the warm-up) carries
    # the pilot, and the hot-bin check returns every 8th frame
    assert 41.0 <= d['gpu_launches_per_step'] <= 42.0
    assert d['gpu_launches'] == round(d['gpu_launches_per_step'] * 4)
    assert d['clocks']['reasons'] == ['nvidia-smi unavailable'] or 'sm_mhz' in d['clocks']

    r = d['roofline']
    assert r['kernel'] == 'cb_iter' and r['bound'] == 'l2_atomic' and r['unit'] == 'reductions/s'
    assert r['frac'] == pytest.approx(r['achieved'] / r['peak'])
    assert 1.0 <= r['kernel_ms'] <= 4.0             # cb_iterate alone, or pilot + scan + main
    assert r['achieved'] == pytest.approx(n / (r['kernel_ms'] * 1e-3))
    assert r['peak'] == pytest.approx(148 * 4 * 256 * 2048 / 1e-3)      # one launch = 1 ms
    assert d['roofline_filters']['ms'] == pytest.approx(ms - r['kernel_ms'])
    assert r['hbm']['algorithmic_bytes_per_sample'] == 16.0 and r['hbm']['peak'] > 1000
    assert isinstance(r['traffic'], (int, float)) and 'ncu' in r['traffic_source']
    f = d['roofline_filters']
    assert f['algorithmic_bytes_per_bin'] == 804 and f['bilateral_main_pass']['bound'] == 'fma_pipe'

    x = d['extra']
    assert sorted(x) == ['anim1080', 'still4k', 'still8k']
    assert x['still4k']['scaling'] == 'strong' and x['still4k']['nbins'] == 8487424
    assert x['still4k']['samples_per_step'] == 3840 * 2160 * 4000
    assert x['still4k']['accumulate'] == 'float4 reductions'
    assert x['still8k']['accumulate'] == 'packed u64 cells'
    assert x['still8k']['e2e']['d2h_bytes_per_step'] == 7680 * 4320 * 4
    assert x['anim1080']['frames'] == 24 and x['anim1080']['frames_per_second'] > 0
    assert d['cpu_baseline'] is None                # --no-cpu-baseline

    # the timed region of the headline workload issued nothing but this library's calls,
    # and the 8K extra really went down the packed path
    assert any(it['cells'] for it in lib.iterations)
    assert 'cb_flush_packed' in lib.names()
