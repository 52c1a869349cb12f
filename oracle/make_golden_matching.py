"""TEST INFRASTRUCTURE ONLY.  tests/golden/matching.npz: output of the UNMODIFIED reference
geometric_registration/common.py::build_correspondence (open3d stubbed: that file only uses it for I/O helpers) on the
seeded descriptor sets of tests/_inputs.py::matching_case, and a check of the oracle restatement against it.
Run in the build container only:  python oracle/make_golden_matching.py"""
import os
import sys
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.dont_write_bytecode = True
sys.modules["open3d"] = types.ModuleType("open3d")
sys.path.insert(0, "/root/reference/geometric_registration")
import common as ref_common  # noqa: E402  (the reference, unmodified)

import _inputs  # noqa: E402
from oracle import model_ref  # noqa: E402

out = {}
for name, (ns, nt, seed) in _inputs.MATCHING_CASES.items():
    s, t = _inputs.matching_case(ns, nt, seed)
    ref = ref_common.build_correspondence(s, t)
    mine = model_ref.build_correspondence(s, t)
    assert ref.shape == mine.shape and np.array_equal(ref, mine), name
    out[name] = ref.astype(np.int64)
    print(name, ns, nt, "->", ref.shape[0], "mutual pairs")
np.savez_compressed(os.path.join(REPO, "tests", "golden", "matching.npz"), **out)
