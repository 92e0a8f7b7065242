"""Host-side mirror of the reference's glue around the hot path:
  * group_parts (demo:1304-1411) fuzzed against the REAL reference's find_most_relevant_object driven through the
    reference's own loop order, including exact IoU ties and coincident boxes (build container only);
  * STrack.multi_gmc (demo:538-554) against golden vectors produced by the reference (oracle/gen_golden.py host_glue);
  * KalmanFilter.gating_distance (demo:338-380) against the same golden file (GPU: the projection is a kernel).
"""
import os
import types

import numpy as np
import pytest

from oracle.ref_loader import load_reference, reference_available

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "host_glue.npz")


def _random_boxes(rng, n, tie_level):
    """Detector rows (classid, score, x1, y1, x2, y2) on a coarse integer lattice: tie_level > 0 forces many equal
    IoUs (identical and mirrored boxes)."""
    rows = []
    step = (1, 8, 16)[tie_level]
    for _ in range(n):
        cls = int(rng.integers(0, 4))
        x1 = int(rng.integers(0, 40)) * step; y1 = int(rng.integers(0, 30)) * step
        w = int(rng.integers(1, 8)) * 8; h = int(rng.integers(1, 10)) * 8
        rows.append((cls, float(rng.uniform(0.2, 1.0)), x1, y1, x1 + w, y1 + h))
    if tie_level and rows:
        for _ in range(n // 3):                       # exact duplicates in another class -> IoU == 1 ties
            c, s, x1, y1, x2, y2 = rows[int(rng.integers(0, len(rows)))]
            rows.append((int(rng.integers(0, 4)), float(rng.uniform(0.2, 1.0)), x1, y1, x2, y2))
    return rows


def _reference_grouping(ref, rows):
    """The reference's association block (demo:1372-1411) on the reference's own classes and helper."""
    mk = lambda cls, r, **kw: cls(trackid=0, classid=r[0], score=r[1], x1=r[2], y1=r[3], x2=r[4], y2=r[5],
                                  cx=(r[2] + r[4]) // 2, cy=(r[3] + r[5]) // 2, is_used=False, **kw)
    bodies = [mk(ref.Body, r, head=None, hand1=None, hand2=None) for r in rows if r[0] == 0]
    heads = [mk(ref.Head, r, face=None, face_landmarks=None) for r in rows if r[0] == 1]
    hands = [mk(ref.Hand, r) for r in rows if r[0] == 2]
    faces = [mk(ref.Face, r) for r in rows if r[0] == 3]
    if len(faces) > 0:
        for h in heads:
            f = ref.find_most_relevant_object(base_obj=h, target_objs=faces)
            if f is not None:
                h.face = f
    if len(heads) > 0:
        for b in bodies:
            h = ref.find_most_relevant_object(base_obj=b, target_objs=heads)
            if h is not None:
                b.head = h
    if len(hands) > 0:
        for b in bodies:
            h1 = ref.find_most_relevant_object(base_obj=b, target_objs=hands)
            if h1 is not None:
                b.hand1 = h1
            h2 = ref.find_most_relevant_object(base_obj=b, target_objs=hands)
            if h2 is not None:
                b.hand2 = h2
    return bodies


def _key(box):
    return None if box is None else (box.classid, box.score, box.x1, box.y1, box.x2, box.y2)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("tie_level", [0, 1, 2])
def test_group_parts_fuzz_against_reference(tie_level):
    from botsort_b200 import tracker as T
    ref = load_reference()
    rng = np.random.default_rng(100 + tie_level)
    for trial in range(150):
        rows = _random_boxes(rng, int(rng.integers(0, 40)), tie_level)
        mine = T.group_parts([T.Box(trackid=0, classid=r[0], score=r[1], x1=r[2], y1=r[3], x2=r[4], y2=r[5],
                                    cx=(r[2] + r[4]) // 2, cy=(r[3] + r[5]) // 2, is_used=False) for r in rows])
        theirs = _reference_grouping(ref, rows)
        assert len(mine) == len(theirs)
        for a, b in zip(mine, theirs):
            assert _key(a) == _key(b)
            assert _key(a.head) == _key(b.head), f"trial {trial}: head"
            assert _key(a.hand1) == _key(b.hand1) and _key(a.hand2) == _key(b.hand2), f"trial {trial}: hands"
            assert _key(getattr(a.head, "face", None)) == _key(getattr(b.head, "face", None)), f"trial {trial}: face"


def test_multi_gmc_matches_reference_golden():
    from botsort_b200 import tracker as T
    g = np.load(GOLDEN)
    tracks = [types.SimpleNamespace(mean=g["gd_mean"][i].copy(), covariance=g["gd_cov"][i].copy()) for i in range(len(g["gd_mean"]))]
    T.STrack.multi_gmc(tracks, g["gmc_H"])
    np.testing.assert_allclose(np.asarray([t.mean for t in tracks]), g["gmc_mean"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.asarray([t.covariance for t in tracks]), g["gmc_cov"], rtol=0, atol=1e-10)
    T.STrack.multi_gmc([], g["gmc_H"])                # empty list: no-op like the reference


@pytest.mark.gpu
def test_gating_distance_matches_reference_golden(ctx):
    from botsort_b200 import tracker as T
    g = np.load(GOLDEN)
    kf = T.KalmanFilter(ctx)
    for metric in ("maha", "gaussian"):
        for only_pos in (False, True):
            want = g[f"gd_{metric}_{int(only_pos)}"]
            for i in range(len(g["gd_mean"])):
                got = kf.gating_distance(g["gd_mean"][i], g["gd_cov"][i], g["gd_meas"][i].copy(), only_position=only_pos, metric=metric)
                np.testing.assert_allclose(got, want[i], rtol=1e-9, atol=1e-9)
    with pytest.raises(ValueError):
        kf.gating_distance(g["gd_mean"][0], g["gd_cov"][0], g["gd_meas"][0], metric="nope")
