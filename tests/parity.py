"""Shared parity checker: CUDA result vs oracle result, bit-exact ids unless the oracle itself
says two candidates are closer than f32 can resolve (then the swap must be inside that tie group).

Tolerances (BASELINE.json north_star): ids identical (ties by id); |distance - oracle| <= 1e-5.
"""
import numpy as np

SCORE_TOL = 1e-5      # north-star tolerance for fp32 scores
TIE_EPS = 4e-7        # f64-distance gap below which fp32 cannot be expected to order two rows


def check_topk(gpu_ids, gpu_dist, o_ids, o_d32, o_d64, k_eff):
    """o_* are the oracle's top-(k_eff + margin). Returns the number of near-tie swaps accepted."""
    gpu_ids = np.asarray(gpu_ids)
    gpu_dist = np.asarray(gpu_dist, dtype=np.float32)
    assert len(gpu_ids) == k_eff, f"expected {k_eff} results, got {len(gpu_ids)}"
    assert len(set(gpu_ids.tolist())) == k_eff, "duplicate ids in result"
    # ascending (distance, id)
    for i in range(1, k_eff):
        assert (gpu_dist[i - 1], gpu_ids[i - 1]) < (gpu_dist[i], gpu_ids[i]), f"not ascending at {i}"
    if np.array_equal(gpu_ids, o_ids[:k_eff]):
        if k_eff:
            assert np.abs(gpu_dist - o_d32[:k_eff]).max() <= SCORE_TOL
        return 0
    d64 = {int(i): float(d) for i, d in zip(o_ids, o_d64)}
    swaps = 0
    for i in range(k_eff):
        g = int(gpu_ids[i])
        if g == int(o_ids[i]):
            assert abs(float(gpu_dist[i]) - float(o_d32[i])) <= SCORE_TOL
            continue
        assert g in d64, f"rank {i}: id {g} is not among the oracle's top-{len(o_ids)}"
        assert abs(d64[g] - float(o_d64[i])) < TIE_EPS, (
            f"rank {i}: gpu id {g} (d64={d64[g]!r}) vs oracle id {int(o_ids[i])} (d64={float(o_d64[i])!r}) "
            "differ by more than a near-tie")
        assert abs(float(gpu_dist[i]) - d64[g]) <= SCORE_TOL
        swaps += 1
    return swaps
