#!/usr/bin/env python3
"""Known-answer vectors for the per-pair Fitch functions, generated from the REFERENCE's own fastc64
code (oracle/_ref/libxmpref_fitch.so, built by `make -C oracle ref`).  TEST INFRASTRUCTURE ONLY.

    python tests/golden/make_kat.py      ->  tests/golden/fitch_kat.json

The reference ships no known-answer tests for these functions, so the vectors
are outputs of its compiled code on seeded inputs: ambiguity codes, mixed weights, stop-early limits,
empty input, lengths around the 16-byte block size.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import oracle_lib as ol  # noqa: E402


def main():
    R = ol.ref()
    rng = np.random.default_rng(20261019)
    cases = []
    for nbytes in (0, 16, 32, 48, 64, 160, 1248):
        for amb in (0.0, 0.2, 1.0):
            s1, s2, prev = (ol.random_sets(rng, nbytes, amb) for _ in range(3))
            w = ol.random_weights(rng, nbytes * 2, heavy_fraction=rng.choice([0.0, 0.3, 1.0]))
            n8 = nbytes // 8
            dest = np.zeros(nbytes, np.uint8)
            R.FitchBases_b2_w8_fastc64(ol.p(s1), ol.p(s2), ol.p(dest), n8)
            same = dest.copy()
            d2 = np.zeros(nbytes, np.uint8)
            changed = int(R.FitchBasesAndCompare_b2_w8_fastc64(ol.p(s1), ol.p(s2), ol.p(d2), ol.p(prev), n8) != 0)
            unchanged = int(R.FitchBasesAndCompare_b2_w8_fastc64(ol.p(s1), ol.p(s2), ol.p(d2), ol.p(same), n8) != 0)
            score1 = int(R.FitchScoreWeight1_b2_w8_fastc64(ol.p(s1), ol.p(s2), n8))
            score = int(R.FitchScore_b2_w8_fastc64(ol.p(s1), ol.p(s2), ol.p(w), n8))
            score_fw1 = int(R.FitchScore_fw1_b2_w8_fastc64(ol.p(s1), ol.p(s2), ol.p(w), n8))
            limit = score // 2
            early = int(R.FitchScore_fw1_stopearly_b2_w8_fastc64(ol.p(s1), ol.p(s2), ol.p(w), n8, limit))
            early1 = int(R.FitchScoreWeight1_stopearly_b2_w8_fastc64(ol.p(s1), ol.p(s2), n8, score1 // 2))
            cases.append({"nbytes": nbytes, "s1": s1.tobytes().hex(), "s2": s2.tobytes().hex(), "prev": prev.tobytes().hex(),
                          "weights": [int(x) for x in w], "bases": dest.tobytes().hex(), "changed_vs_prev": changed,
                          "changed_vs_self": unchanged, "score_weight1": score1, "score": score, "score_fw1": score_fw1,
                          "limit": limit, "score_fw1_stopearly_w8": early, "score_weight1_stopearly_w8": early1})
    json.dump(cases, open(os.path.join(HERE, "fitch_kat.json"), "w"))
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
