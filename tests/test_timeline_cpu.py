"""Host logic of the timeline probe (mpmavatar_b200/timeline.py summarise) on synthetic stamps."""
import numpy as np

from mpmavatar_b200.timeline import NAMES, summarise


def _chain(n, period=100_000):
    # per substep (ns): P2G_E 0-30k, P2G_V 22k-45k, scatter 40k-47k, grid 41k-52k, G2P_V 50k-66k, G2P_E 62k-90k
    spans = {0: (0, 30_000), 2: (22_000, 45_000), 3: (40_000, 47_000), 4: (41_000, 52_000), 5: (50_000, 66_000), 7: (62_000, 90_000)}
    tl = np.full((n, 8, 2), -1, np.int64)
    for i in range(n):
        for k, (a, b) in spans.items():
            tl[i, k] = (i * period + a, i * period + b)
    return tl


def test_summarise_phases_and_period():
    r = summarise(_chain(12))
    assert abs(r["substep_us"] - 100.0) < 1e-9
    assert abs(r["p2g_union_us"] - 45.0) < 1e-9   # first start of kernels 0-2 to their last end
    assert abs(r["g2p_union_us"] - 40.0) < 1e-9   # kernels 5-7
    assert set(r["kernels"]) == {NAMES[k] for k in (0, 2, 3, 4, 5, 7)}  # traditional kernels did not run
    assert r["kernels"]["p2g_V"] == {"start_us": 22.0, "dur_us": 23.0}


def test_summarise_uses_the_median_and_skips_the_first_substeps():
    tl = _chain(12)
    tl[:4, 0, 1] += 500_000        # cold first substeps
    tl[8, 7, 1] += 30_000          # one slow substep does not move the median
    r = summarise(tl)
    assert abs(r["kernels"]["p2g_E"]["dur_us"] - 30.0) < 1e-9
    assert abs(r["g2p_union_us"] - 40.0) < 1e-9
