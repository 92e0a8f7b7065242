"""TEST INFRASTRUCTURE ONLY -- decision-margin certification of synthetic frames.

"Bit-exact ids" is only well defined when no decision of the reference sits on a knife edge: a frame is
CERTIFIED when
  * no cost the reference compares lies within `margin` (1e-3) of the threshold it is compared with
    (duplicate distance 0.15, appearance gate 0.25, proximity / second-stage 0.5, unconfirmed 0.7, first
    association 0.8; scores against 0.1 / 0.4 / 0.9), and
  * the optimum of every linear assignment is unchanged under several random +-`perturb` (1e-4) perturbations of
    its cost matrix.
Two levels are reported:
  "strict"      every compared quantity (the SURVEY's rule); crowded scenes rarely pass because SOME IoU among
                thousands of overlapping pairs is always within 1e-3 of a threshold -- although the IoU arithmetic is
                bit-identical between the reference and the GPU path (float64, same operation order)
  "similarity"  only quantities that derive from the float32 feature similarity -- the ones BLAS / tensor-core
                summation order can actually move.
Nothing in the product package imports this module.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from . import oracle_np as O


def _near(a: np.ndarray, thresh: float, margin: float) -> int:
    a = np.asarray(a)
    if a.size == 0:
        return 0
    return int(np.count_nonzero(np.abs(a - thresh) < margin))


def _lap_stable(cost: np.ndarray, thresh: float, rng, perturb: float, trials: int, mask=None) -> bool:
    if cost.size == 0:
        return True
    x0, _ = O.lapjv_extended(cost, thresh)
    for _ in range(trials):
        noise = rng.uniform(-perturb, perturb, cost.shape)
        if mask is not None:
            noise = np.where(mask, noise, 0.0)
        x1, _ = O.lapjv_extended(cost + noise, thresh)
        if not np.array_equal(x0, x1):
            return False
    return True


def certify_frame(last: Dict[str, np.ndarray], margin: float = 1e-3, perturb: float = 1e-4, trials: int = 3,
                  seed: int = 0) -> Dict[str, object]:
    """`last` = OracleBoTSORT.last after a frame.  Returns {"strict": bool, "similarity": bool, "why": [...]}"""
    rng = np.random.default_rng(seed)
    why_strict: List[str] = []
    why_sim: List[str] = []
    sc = last.get("scores", np.zeros(0))
    for t in (O.TRACK_LOW_THRESH, O.TRACK_HIGH_THRESH, O.NEW_TRACK_THRESH):
        k = _near(sc.astype(np.float64), t, margin)
        if k:
            why_strict.append(f"{k} scores within {margin} of {t}")
    d1, iou1, emb1 = last.get("dists1", np.zeros((0, 0))), last.get("iou1", np.zeros((0, 0))), last.get("emb1", np.zeros((0, 0)))
    if d1.size:
        k = _near(emb1, O.APPEARANCE_THRESH, margin)
        if k:
            why_sim.append(f"{k} first-association embedding distances within {margin} of the appearance gate")
        from_emb = (d1 < iou1)                      # entries whose cost IS the embedding distance
        k = _near(d1[from_emb], O.MATCH_THRESH, margin)
        if k:
            why_sim.append(f"{k} embedding costs within {margin} of match_thresh")
        k = _near(d1[~from_emb], O.MATCH_THRESH, margin) + _near(iou1, O.PROXIMITY_THRESH, margin)
        if k:
            why_strict.append(f"{k} first-association IoU costs within {margin} of a threshold")
        if not _lap_stable(d1, O.MATCH_THRESH, rng, perturb, trials, mask=from_emb):
            why_sim.append("first assignment changes under +-%g perturbation of its embedding costs" % perturb)
        elif not _lap_stable(d1, O.MATCH_THRESH, rng, perturb, trials):
            why_strict.append("first assignment changes under +-%g perturbation" % perturb)
    d2 = last.get("dists2", np.zeros((0, 0)))
    if d2.size:
        if _near(d2, O.SECOND_THRESH, margin):
            why_strict.append("second-association IoU cost near 0.5")
        if not _lap_stable(d2, O.SECOND_THRESH, rng, perturb, trials):
            why_strict.append("second assignment unstable")
    d3, iou3, emb3 = last.get("dists3", np.zeros((0, 0))), last.get("iou3", np.zeros((0, 0))), last.get("emb3", np.zeros((0, 0)))
    if d3.size:
        k = _near(emb3, O.APPEARANCE_THRESH, margin)
        if k:
            why_sim.append(f"{k} unconfirmed embedding distances within {margin} of the appearance gate")
        from_emb3 = (d3 < iou3)
        if _near(d3[from_emb3], O.UNCONF_THRESH, margin):
            why_sim.append("unconfirmed embedding cost near 0.7")
        if _near(d3[~from_emb3], O.UNCONF_THRESH, margin) + _near(iou3, O.PROXIMITY_THRESH, margin):
            why_strict.append("unconfirmed IoU cost near a threshold")
        if not _lap_stable(d3, O.UNCONF_THRESH, rng, perturb, trials, mask=from_emb3):
            why_sim.append("unconfirmed assignment unstable under embedding perturbation")
        elif not _lap_stable(d3, O.UNCONF_THRESH, rng, perturb, trials):
            why_strict.append("unconfirmed assignment unstable")
    dd = last.get("dup_dist", np.zeros((0, 0)))
    if dd.size and _near(dd, O.DUP_IOU_DIST, margin):
        why_strict.append("duplicate IoU distance near 0.15")
    return {"similarity": len(why_sim) == 0, "strict": len(why_sim) == 0 and len(why_strict) == 0,
            "why": why_sim + why_strict}
