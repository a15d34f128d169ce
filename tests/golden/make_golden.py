#!/usr/bin/env python3
"""tests/golden/make_golden.py — (re)generates tests/golden/*.json.

Run in the BUILD container only (needs /root/reference).  Two kinds of fixture:

1. `eigen_known_answers.json` — the known-answer vectors the reference's own
   gtest holds for LogSumExp and Softmax, extracted verbatim from
   kaldi-hmm-gmm/csrc/eigen-test.cc:460-474 and :641-654 (parsed from the source
   text; nothing is recomputed).
2. `diag_gmm_closed_form.json` — seeded inputs plus expected values computed in
   float64 from the explicit Gaussian formulas that the reference's Python
   tests assert against (python/tests/test_diag_gmm.py:45-51, 327-403, 529-553;
   python/tests/test_mle_diag_gmm.py:200-252).  The reference module itself
   cannot be imported here (its extension needs Eigen, which is absent), so the
   expectations come from those closed forms, evaluated independently of
   oracle/.
"""
import json
import math
import os
import re

import numpy as np

REF = "/root/reference/kaldi-hmm-gmm"
HERE = os.path.dirname(os.path.abspath(__file__))


def _floats(s):
    return [float(x) for x in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", s)]


def eigen_known_answers():
    src = open(os.path.join(REF, "csrc/eigen-test.cc")).read()
    lse = src[src.index("TEST(Eigen, LogSumExp)"):src.index("TEST(Eigen, Addmm)")]
    blocks = re.findall(r"v <<(.*?);", lse, re.S)
    expect = re.findall(r"EXPECT_NEAR\(f, ([\d.]+), ([\de.-]+)\)", lse)
    cases = [dict(v=_floats(b), expected=float(e), tol=float(t)) for b, (e, t) in zip(blocks, expect)]
    sm = src[src.index("TEST(Eigen, TestSoftmax)"):src.index("TEST(Eigen, Op1)")]
    v = _floats(re.search(r"v <<(.*?);", sm, re.S).group(1))
    exp = _floats(re.search(r"expected <<(.*?);", sm, re.S).group(1))
    tol = float(re.search(r"EXPECT_NEAR\(expected\[i\], actual\[i\], ([\de.-]+)\)", sm).group(1))
    return dict(source="kaldi-hmm-gmm/csrc/eigen-test.cc:460-474,641-654",
                logsumexp=cases, softmax=dict(v=v, expected=exp, tol=tol))


def diag_gmm_closed_form():
    rng = np.random.default_rng(20230414)
    out = []
    for nmix, dim in [(10, 8), (8, 2), (10, 3), (3, 5), (1, 39), (17, 40)]:
        w = rng.random(nmix)
        w /= w.sum()
        mean = rng.random((nmix, dim))
        var = rng.random((nmix, dim)) * 0.9 + 0.1
        x = rng.random((4, dim))
        w32, m32, v32, x32 = (a.astype(np.float32) for a in (w, mean, var, x))
        w, mean, var, x = (a.astype(np.float64) for a in (w32, m32, v32, x32))
        # test_diag_gmm.py:45-51
        gconsts = np.log(w) - 0.5 * (dim * math.log(2 * math.pi) + np.log(var).sum(1) + (mean ** 2 / var).sum(1))
        # test_diag_gmm.py:351-373 per-component log-likes; :327-349 total
        comp = np.stack([np.log(w) + (-(xi - mean) ** 2 / (2 * var)).sum(1) - 0.5 * np.log(2 * math.pi * var).sum(1) for xi in x])
        mx = comp.max(1, keepdims=True)
        total = (np.log(np.exp(comp - mx).sum(1)) + mx[:, 0])
        post = np.exp(comp - total[:, None])  # test_diag_gmm.py:529-553
        # test_mle_diag_gmm.py:200-252: stats after accumulate_from_diag(x[0], weight 0.2)
        wt = 0.2
        occ = post[0] * wt
        out.append(dict(nmix=nmix, dim=dim, weights=w32.tolist(), means=m32.tolist(), vars=v32.tolist(), x=x32.tolist(),
                        gconsts=gconsts.tolist(), component_loglikes=comp.tolist(), loglike=total.tolist(),
                        posteriors=post.tolist(), acc_weight=wt, occ=occ.tolist(),
                        mean_acc=(occ[:, None] * x[0][None, :]).tolist(),
                        var_acc=(occ[:, None] * (x[0] ** 2)[None, :]).tolist()))
    return dict(source="closed forms asserted by python/tests/test_diag_gmm.py:45-51,327-403,529-553 and "
                       "python/tests/test_mle_diag_gmm.py:200-252, evaluated in float64", cases=out)


if __name__ == "__main__":
    json.dump(eigen_known_answers(), open(os.path.join(HERE, "eigen_known_answers.json"), "w"), indent=1)
    json.dump(diag_gmm_closed_form(), open(os.path.join(HERE, "diag_gmm_closed_form.json"), "w"))
    print("wrote golden fixtures")
