"""Loader for tests/golden/qluster_*.npz (written by tests/golden/make_qluster_fixtures.py): the oracle's
qluster results at BASELINE sizes, frozen because the oracle needs minutes per fixture."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_qluster_fixtures", os.path.join(HERE, "golden", "make_qluster_fixtures.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

NAMES = [n for n in ("c2_5k", "c1_full", "c2_20k", "c2_20k_hp") if os.path.exists(os.path.join(HERE, "golden", f"qluster_{n}.npz"))]


def load(name):
    """-> (inputs tuple as gen.case_inputs returns it, fixture dict); asserts the regenerated input is the frozen one."""
    fx = dict(np.load(os.path.join(HERE, "golden", f"qluster_{name}.npz")))
    inputs = gen.case_inputs(name)
    sha = gen.input_sha(*inputs[:3])
    assert sha == fx["inputSha"].tobytes().decode(), f"fixture {name}: the seeded generator no longer reproduces the frozen input"
    return inputs, fx


def assert_result_equals_fixture(got, fx):
    """Membership, cluster order, consensus sequences, qualities, counts and chimera records, bit for bit."""
    want = fx
    assert np.array_equal(got["membership"], want["membership"])
    p = got["packed"]
    assert np.array_equal(p["offsets"], want["offsets"])
    assert np.array_equal(p["cnts"], want["cnts"])
    assert np.array_equal(p["bases"], want["seqs"])
    assert np.array_equal(p["quals"], want["quals"])
    chi = np.array([(c,) + tuple(v) for c, v in sorted(got["chimeras"].items())], np.uint32).reshape(-1, 5)
    assert np.array_equal(chi, want["chimeras"])
    assert got["stats"]["uniqueSeqs"] == int(want["nUnique"][0])
