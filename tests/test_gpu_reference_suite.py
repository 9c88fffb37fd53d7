"""The reference's OWN unit tests, unmodified, against the GPU backend.

`make -C oracle ref` byte-compiles the reference's tests/test_beam.py, test_forward.py, test_transducer.py, test_prefix.py and their
fixture library tests/testing.py from where they lie into oracle/_ref/reftests/ (build outputs: they travel to the GPU
box with the other oracle/_ref artefacts; no reference source is copied into the repository).  Here `poreover` and
`poreover.decoding` are aliased to poreover_b200's drop-in modules in sys.modules and the suites are run with
unittest.  Expected failures, by name: the flip-flop *beam tree* cases, which are out of scope
(the reference's own flip-flop 2D test fails in the reference itself).
"""
import os
import sys
import types
import unittest

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFTESTS = os.path.join(ROOT, "oracle", "_ref", "reftests")

# flip-flop beam trees / flip-flop forward are out of scope
EXPECTED_TO_FAIL = {"test_flipflop_same", "test_fw_flipflop"}

pytestmark = pytest.mark.gpu


@pytest.fixture()
def aliased_reference_package(monkeypatch):
    if not os.path.exists(os.path.join(REFTESTS, "test_beam.pycode")):
        pytest.skip("oracle/_ref/reftests not built (make -C oracle ref needs the reference tree)")
    import poreover_b200
    import poreover_b200.decoding as dec
    pkg = types.ModuleType("poreover")
    pkg.__path__ = []  # a package: `import poreover.decoding` resolves through sys.modules
    pkg.decoding = dec
    pkg.align = __import__("poreover_b200.align", fromlist=["align"])
    monkeypatch.setitem(sys.modules, "poreover", pkg)
    monkeypatch.setitem(sys.modules, "poreover.decoding", dec)
    for name in ("decoding_cpp", "decoding_cy", "transducer", "decode", "envelope", "prefix_search"):
        monkeypatch.setitem(sys.modules, "poreover.decoding." + name, getattr(dec, name))
    monkeypatch.setitem(sys.modules, "poreover.align", pkg.align)
    if not hasattr(np, "product"):  # tests/testing.py:73 uses np.product, removed in NumPy 2
        monkeypatch.setattr(np, "product", np.prod, raising=False)
    # the byte-compiled modules, loaded under their own names ("testing" first: the suites import it)
    import importlib.machinery
    import importlib.util
    for m in ("testing", "test_beam", "test_forward", "test_transducer", "test_prefix"):
        monkeypatch.delitem(sys.modules, m, raising=False)
    for m in ("testing", "test_beam", "test_forward", "test_transducer", "test_prefix"):
        path = os.path.join(REFTESTS, m + ".pycode")
        loader = importlib.machinery.SourcelessFileLoader(m, path)
        spec = importlib.util.spec_from_loader(m, loader, origin=path)
        mod = importlib.util.module_from_spec(spec)
        mod.__file__ = os.path.join(REFTESTS, m + ".py")  # the suites locate poreover.csv next to themselves
        sys.modules[m] = mod
        loader.exec_module(mod)
    yield
    for m in ("testing", "test_beam", "test_forward", "test_transducer", "test_prefix"):
        sys.modules.pop(m, None)


@pytest.mark.parametrize("suite", ["test_beam", "test_forward", "test_transducer", "test_prefix"])
def test_reference_suite(aliased_reference_package, suite):
    tests = unittest.defaultTestLoader.loadTestsFromName(suite)
    assert tests.countTestCases() > 0
    result = unittest.TestResult()
    tests.run(result)
    bad = [(t.id(), tb) for t, tb in result.errors + result.failures]
    unexpected = [(i, tb) for i, tb in bad if i.split(".")[-1] not in EXPECTED_TO_FAIL]
    assert not unexpected, "\n\n".join("%s\n%s" % x for x in unexpected)
    ran_ok = result.testsRun - len(bad)
    assert ran_ok >= 1, "nothing passed in %s" % suite
