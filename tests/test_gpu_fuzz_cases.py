"""Batches of tools/fuzz_parity.py (a randomised GPU-vs-oracle sweep) that once differed, replayed:
   549564285  600-bp reads with long poly-C stretches: anchors more than 127 bases from their overlap end — the exact
              verify kernel started its extension at the 7-bit cap instead of at the anchor (common.cuh: HLO_CAND_FAR);
   733435268  pair_rule 1: a row that fails the overhang test by clipping 102 bases holds more mismatches than the fast
              kernel's cheap reject allowed for (verify.cu: the bound is max_k + OL - min_len, not a window's worth)."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT, "tools", "fuzz_parity.py"))
fuzz = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fuzz)


@pytest.mark.parametrize("seed", [549564285, 733435268, 17, 4242, 987654321])
def test_fuzz_seed_is_identical_to_the_oracle(seed):
    try:
        same, x = fuzz.run_case(seed)
    except fuzz.api.HloError as e:
        pytest.skip("parameters outside the supported domain: %s" % e)
    assert same, (x["pkw"], len(x["rec"]), len(x["rec_o"]), x["tm"]["n_accepted"], x["st"]["accepted"])
