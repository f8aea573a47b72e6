"""CPU suite: the checker (the unmodified reference built into oracle/_ref) against the committed golden fixtures,
and the synthetic generator against its pinned input digests.  Keeps the oracle honest when the host compiler, the
flags or the generator change."""
import hashlib
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def _kw(name):
    kw = dict(GOLD[name]["args"])
    if "orientations" in kw:
        kw["orientations"] = tuple(kw["orientations"])
    return kw


@pytest.mark.parametrize("name", sorted(GOLD))
def test_generator_is_pinned(rb, name):
    g = make_golden.make_gof(rb, _kw(name))
    md5 = hashlib.md5(g.occupancy.tobytes() + g.geometry.tobytes() + g.attribute.tobytes() + g.patches.tobytes()).hexdigest()
    assert md5 == GOLD[name]["input_md5"], "the synthetic generator changed: regenerate tests/golden (make_golden.py)"


@pytest.mark.parametrize("name", sorted(GOLD))
def test_reference_checker_matches_golden(rb, name):
    from oracle import checker
    if not checker.have_reference():
        pytest.skip("oracle/_ref not built")
    got = make_golden.run_case(rb, checker.Reference(), checker, name, _kw(name))
    want = GOLD[name]
    assert len(got["frames"]) == len(want["frames"])
    for f, (a, b) in enumerate(zip(got["frames"], want["frames"])):
        assert a == b, f"{name} frame {f}: the reference checker no longer reproduces the golden fixture"
