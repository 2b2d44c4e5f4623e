"""When the reference checkout is present (the build container; never the GPU box), re-run the full
pinning of the oracle: all 66 per-op RTL dumps of tv/, all 44 kernel-level software-model vectors and
case3_expected_result.txt, read straight from /root/reference.  The committed fixtures in
tests/golden/ are the travelling subset of exactly this check (tools/make_golden.py)."""
import importlib.util
import os

import pytest

import golden_util as G

REF = os.environ.get("ALOHA_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tv")), reason="reference checkout not present")
def test_oracle_bit_exact_on_every_reference_vector():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tools", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    checked = mod.main(check_only=True)          # asserts on the first mismatch
    assert checked == 111 == G.manifest()["vectors_checked"]
