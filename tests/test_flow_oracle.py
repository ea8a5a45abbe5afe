"""CPU: the numpy restatement of video_to_flow (oracle/flow_oracle.py) against cv2.calcOpticalFlowFarneback
fields and the output of the reference's own lib/utils.video_to_flow (tests/golden/flow_small.pt)."""
import numpy as np
import torch

from oracle import flow_oracle as FO
from helpers import golden, flow_clip, flow_level_agreement


def test_farneback_restatement_matches_cv2_fields():
    f = golden("flow_small.pt")
    for name, case in f.items():
        B, D, S, seed = case["cfg"]
        grey = FO.gray_frames(flow_clip(B, D, S, seed).numpy())
        for i in range(2):
            got = FO.farneback(grey[0, i], grey[0, i + 1])
            want = case["cv2_flow_b0"][i].numpy()
            assert np.linalg.norm(got - want) <= 2e-6 * np.linalg.norm(want), (name, i)


def test_video_to_flow_restatement_matches_reference_output():
    f = golden("flow_small.pt")
    for name in ("s64", "s112"):
        B, D, S, seed = f[name]["cfg"]
        out, _ = FO.video_to_flow(flow_clip(B, D, S, seed).numpy())
        levels = torch.from_numpy(np.rint((out + 1) * 0.5 * 255).astype(np.uint8))
        exact, near = flow_level_agreement(levels, f[name]["levels"])
        # float rounding at the uint8 truncation of values up to +-65025 moves a byte by one level now and then
        assert exact > 0.985 and near > 0.9995, (name, exact, near)
        assert torch.equal(levels[:, :, -1], levels[:, :, -2])          # last frame repeated (lib/utils.py:125)


def test_pyramid_levels_for_the_configured_sizes():
    assert [FO.pyramid_levels(s, s) for s in (64, 112, 128)] == [1, 1, 2]
