"""Shared helpers for the parity tests: margin-classified comparison of RVQ code indices.

SURVEY.md §8(c): bit-exact indices are not attainable unconditionally -- top-2 distance gaps reach one
fp32 ulp of the distance -- so a mismatch is accepted only when the ORACLE's own top-2 margin at the
first diverging stage of that frame is below `margin_tol`; later stages of the frame are excluded
(the residual differs after a flip).
"""
import numpy as np


def classify_codes(codes_test, codes_ref, margins_ref, margin_tol):
    """codes_*: [n_q, B, T] ints; margins_ref: [n_q, B, T] oracle (top1 - top2) of -dist.
    Returns dict(exact_frames, total_frames, near_tie_frames, bad_frames, worst_margin)."""
    ct = np.asarray(codes_test).astype(np.int64)
    cr = np.asarray(codes_ref).astype(np.int64)
    mg = np.asarray(margins_ref)
    assert ct.shape == cr.shape == mg.shape, (ct.shape, cr.shape, mg.shape)
    n_q, B, T = ct.shape
    diff = ct != cr
    any_diff = diff.any(axis=0)
    first = np.where(any_diff, diff.argmax(axis=0), -1)       # first diverging stage per frame
    total = B * T
    exact = int((~any_diff).sum())
    near, bad, worst = 0, 0, 0.0
    for b, t in zip(*np.nonzero(any_diff)):
        m = float(mg[first[b, t], b, t])
        worst = max(worst, m)
        if m <= margin_tol:
            near += 1
        else:
            bad += 1
    return dict(exact_frames=exact, total_frames=total, near_tie_frames=near, bad_frames=bad, worst_margin=worst,
                exact_rate=exact / total, first_stage=first)


def assert_codes_parity(codes_test, codes_ref, margins_ref, margin_tol, min_exact_rate=0.97, what=""):
    r = classify_codes(codes_test, codes_ref, margins_ref, margin_tol)
    assert r["bad_frames"] == 0, f"{what}: {r['bad_frames']} frames differ with oracle margin > {margin_tol} " \
                                  f"(worst {r['worst_margin']:.3e}); exact {r['exact_frames']}/{r['total_frames']}"
    assert r["exact_rate"] >= min_exact_rate, f"{what}: exact-match rate {r['exact_rate']:.4f} < {min_exact_rate}"
    return r
