"""PRIMARY graphs through the sm_100a kernels (CanonicalDBG semantics on device: rc-strand jump tables, node-id flips,
both-strand node map, sub-k seeding of both strands) against the oracle's CanonicalDBG restatement, which is pinned on the
reference's primary-mode goldens (tests/test_oracle_primary.py)."""
import pytest

import parity_common as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("subk", [False, True])
def test_mt_primary_gpu(subk):
    # integration_tests/test_align.py:270-330 (graph built with --mode primary)
    got = P.check_mt_primary(None, **({"min_seed_length": 10} if subk else {}))
    from test_oracle_canonical import CANONICAL, CANONICAL_SUBK
    for i, exp in (CANONICAL_SUBK if subk else CANONICAL):
        assert got[int(i)].split("\t")[:8] == exp.encode().decode("unicode_escape").split("\t")[:8]
    assert got[6].split("\t")[4] == "310" and (subk or got[5].split("\t")[4] == "22")


def test_fuzz_primary_gpu():
    for seed in range(100, 140):
        bad, info = P.fuzz_case(None, seed, primary=True)
        assert not bad, (seed, info)


def test_mode_known_answers_gpu():
    # the reference's unit tests for CANONICAL / PRIMARY graphs (tests/graph/test_aligner.cpp:1483-1631, 1773-1800)
    P.check_mode_kats(None)
