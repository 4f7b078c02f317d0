"""CPU checks of the measurement helpers in bench.py: the algorithmic work the rooflines are computed from
(SURVEY 8d figures), the clock-regime rule that picks the tensor-core denominator, and the parity record."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_algorithmic_flops_match_the_survey_figures():
    # conv fwd to conv5_1 at 1024^2 = 0.757 TFLOP, the 24 tensor-core 3x3 launches (conv1_2..conv5_1,
    # fwd + dgrad) = 1.5075 TFLOP; 4096^2 = 16x
    assert abs(bench.conv_flops(1024, 1024) / 1e12 - 0.757) < 1e-3
    assert abs(2 * bench.conv_flops(1024, 1024, first=1) / 1e12 - 1.5075) < 1e-3
    assert bench.conv_flops(4096, 4096) == 16 * bench.conv_flops(1024, 1024)
    # ceil-mode pooled extents on an odd canvas (75 x 101 -> 38 x 51 -> 19 x 26 -> 10 x 13 -> 5 x 7)
    assert bench.level_dims(75, 101) == [(75, 101), (38, 51), (19, 26), (10, 13), (5, 7)]


def test_algorithmic_bytes_of_the_bandwidth_bound_categories():
    ab = bench.algorithmic_bytes(1024, 1024, esz=2)
    n = 3 * 1024 * 1024
    assert ab['pixel_terms'] == 12 * n
    assert ab['optimizer'] == 192 * n                          # compact L-BFGS, m = 10: two passes of (2m + 4) vectors
    assert ab['gram'] == 2 * (64 * 1024 ** 2 + 128 * 512 ** 2 + 256 * 256 ** 2 + 512 * 128 ** 2 + 512 * 64 ** 2)
    assert ab['style_grad'] == 2 * ab['gram']
    assert abs(ab["gram"] / 1e6 - 255.9) < 0.1 and abs(ab['pool'] / 1e6 - 566.2) < 0.1


def test_regime_rule_picks_the_denominator_from_the_observed_clocks():
    assert bench.regime_of({'sm_mhz': 1965.0, 'sm_max_mhz': 1965.0, 'reasons': []}) == 'burst'
    assert bench.regime_of({'sm_mhz': 1965.0, 'sm_max_mhz': 1965.0, 'reasons': ['sw_power_cap']}) == 'sustained'
    assert bench.regime_of({'sm_mhz': 1500.0, 'sm_max_mhz': 1965.0, 'reasons': []}) == 'sustained'
    assert bench.regime_of(None) == 'sustained' and bench.regime_of({'sm_mhz': None, 'sm_max_mhz': None}) == 'sustained'
    peaks = {'bf16_tflops': 1651.3, 'bf16_tflops_sustained': 1398.1, 'hbm_gbs': 6542.1}
    cats = {'conv_tc': {'ms_per_step': 1.5, 'launch_spans_per_step': 24.0}, 'pool': {'ms_per_step': 0.13, 'launch_spans_per_step': 8.0}}
    main, all_ = bench.rooflines_of(cats, 1024, 1, peaks, 'burst', 'fp16')
    assert main['peak'] == 1651.3 and abs(main['frac'] - main['frac_of_burst_peak']) < 1e-12
    assert abs(main['achieved'] - 1.5075e12 / 1.5e-3 / 1e12) < 1.0
    main_s, _ = bench.rooflines_of(cats, 1024, 1, peaks, 'sustained', 'fp16')
    assert main_s['peak'] == 1398.1 and main_s['frac'] > main['frac']
    pool = [r for r in all_ if r.get('category') == 'pool'][0]
    assert pool['bound'] == 'hbm' and abs(pool['achieved'] - 566.2e6 / 0.13e-3 / 1e9) < 20


def test_parity_record():
    rs = np.random.RandomState(0)
    g = rs.randn(3, 8, 8)
    cpu = {'loss': 10.0, 'grad': g, 'trace': {'a_s_loss': 2.0, 'scd_grad': 4.0, 'time': 1.0, 'loss': 10.0}}
    gpu = {'loss': 10.001, 'grad': g * 1.01, 'trace': {'a_s_loss': 2.0004, 'scd_grad': 4.2, 'time': 9.0, 'loss': 10.001}}
    p = bench.parity_of(gpu, cpu)
    assert abs(p['loss_rel'] - 1e-4) < 1e-9 and abs(p['grad_rel'] - 0.01) < 1e-9
    assert p['worst_trace_key'] == 'scd_grad' and abs(p['worst_trace_rel'] - 0.05) < 1e-9
    assert p['worst_loss_trace_key'] == 'a_s_loss' and abs(p['worst_loss_trace_rel'] - 2e-4) < 1e-9
