"""GPU: cvd_flow_consistency_masks against the oracle and the golden masks the reference's
utils/consistency.py produced.  Masks are thresholded float quantities: pixels whose value sits within rounding of a
threshold may flip between implementations, so the bar is a mismatch FRACTION (reported on failure), not equality."""
import os

import numpy as np
import pytest

from oracle import flowmask_oracle as fo
from oracle.make_golden import FLOWMASK_CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(FLOWMASK_CASES))
def test_flow_consistency_masks_match_reference(name):
    from consistent_depth_b200.utils.consistency import consistent_flow_masks
    seed, H, W, ft, ct = FLOWMASK_CASES[name]
    flows, colors = fo.synthetic_pair(seed, H, W)
    masks = consistent_flow_masks(flows, colors, ft, ct)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "flowmask.npz"))
    want = fo.consistent_flow_masks(flows, colors, ft, ct)
    for d in range(2):
        assert masks[d].shape == (H, W) and masks[d].dtype == bool
        mo, mg = float((masks[d] != want[d]).mean()), float((masks[d] != g[f"{name}_mask{d}"]).mean())
        og = float((want[d] != g[f"{name}_mask{d}"]).mean())          # oracle on this host vs the reference-generated golden
        assert mo <= 2e-3 and mg <= 2e-3, f"direction {d}: kernel vs oracle {mo:.2e}, kernel vs golden {mg:.2e}, oracle vs golden {og:.2e}"
