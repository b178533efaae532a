"""The alignment choices nothing in the reference pins (fgbio's Aligner is an un-vendored dependency) are
compile-time switches in BOTH the oracle (oracle/idp_oracle.c, ALIGN_POLICY) and the product (csrc/idp_internal.h, AlignPolicy).
This file keeps the build with every switch flipped parity-green, so that correcting a choice once a JVM can be consulted is a
one-line change on both sides."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from oracle import oracle_py as O

CHECK = os.path.join(H.ROOT, "tests", "policy_check.py")
# pairs on which the two settings give different alignments (found by search; e.g. CCAAC / ACACC in Glocal at 1/-1/0/-1 starts at
# target base 1 by default and at base 2 with the switches flipped), plus a few where they agree
CASES = [(b"CCAAC", b"ACACC"), (b"ACAACCA", b"CACACCAA"), (b"AACAAAC", b"CCAAAACCCCCA"), (b"CCCACACA", b"AACAACAAAACC"), (b"AAAACACA", b"AAACCA"),
         (b"CCCCC", b"AAAA"), (b"ACGTACGTAC", b"TTACGTTACGTACTT"), (b"GATTACA", b"GATTTACAGATACA")]
SCORES = [dict(match_score=1, mismatch_score=-1, gap_open=0, gap_extend=-1), dict(match_score=2, mismatch_score=-1, gap_open=-1, gap_extend=-1),
          dict(match_score=1, mismatch_score=-4, gap_open=-6, gap_extend=-1)]
_DUMP = """
import json, sys
sys.path.insert(0, %r)
from oracle import oracle_py as O
out = []
for scores in %r:
    params = O.make_params(**scores)
    for mode in (O.MODE_LOCAL, O.MODE_GLOCAL, O.MODE_GLOBAL):
        for q, t in %r:
            a = O.align(params, mode, q, t)
            out.append([a[k] for k in ("score", "query_start", "query_end", "target_start", "target_end")])
print(json.dumps(dict(bits=O.lib().orc_align_policy_bits(), out=out)))
"""


def _dump(oracle_lib=None):
    env = dict(os.environ)
    if oracle_lib:
        env["IDP_ORACLE_LIB"] = oracle_lib
    txt = subprocess.check_output([sys.executable, "-c", _DUMP % (H.ROOT, SCORES, CASES)], env=env, text=True)
    return json.loads(txt.strip().splitlines()[-1])


def test_oracle_policy_switches_are_live():
    """The flipped build is a different aligner: same scores where only ties are involved, different coordinates somewhere."""
    base, alt = _dump(), _dump(O.build_alt_policy())
    assert base["bits"] == 0 and alt["bits"] == 15
    a, b = np.array(base["out"]), np.array(alt["out"])
    assert a.shape == b.shape and (a != b).any()


@pytest.mark.gpu
@pytest.mark.parametrize("alt", [False, True])
def test_cuda_path_matches_oracle_under_both_policies(alt):
    from fg_idprimer_b200 import build as B
    env = dict(os.environ)
    if alt:
        env["IDP_LIB_PATH"] = B.build_alt_policy()
        env["IDP_ORACLE_LIB"] = O.build_alt_policy()
    out = subprocess.run([sys.executable, CHECK, "15" if alt else "0"], env=env, capture_output=True, text=True)
    assert out.returncode == 0 and "POLICY-PARITY-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
