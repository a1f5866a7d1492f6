"""vf_prepare_batch_u8 against the oracle restatement of data/nmr_dataset.py:10-52 (bit-exact: byte / index work)."""
import numpy as np
import pytest
import torch

import vf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,V,H,W,C", [(28, 24, 64, 64, 3), (3, 5, 7, 9, 3), (1, 2, 16, 16, 4)])
def test_prepare_batch_bit_exact(B, V, H, W, C):
    from view_fusion_b200 import inputs
    rng = np.random.default_rng(B * 7 + V)
    views = rng.integers(0, 256, size=(B, V, H, W, C), dtype=np.uint8)
    perm = np.stack([rng.permutation(V) for _ in range(B)])
    want = O.process_batch_u8(views, perm)
    target, cond, angle = inputs.prepare_batch(torch.from_numpy(views).cuda(), torch.from_numpy(perm))
    torch.cuda.synchronize()
    assert torch.equal(target.cpu(), torch.from_numpy(want["target"]))
    assert torch.equal(cond.cpu(), torch.from_numpy(want["cond"]))
    assert torch.equal(angle.cpu(), torch.from_numpy(want["angle"]))


def test_prepare_batch_rejects_bad_arguments():
    from view_fusion_b200 import inputs
    v = torch.zeros(2, 4, 8, 8, 3, dtype=torch.uint8)
    with pytest.raises(RuntimeError):
        inputs.prepare_batch(v, torch.zeros(2, 4, dtype=torch.long))
    with pytest.raises(ValueError):
        inputs.prepare_batch(v.cuda(), torch.zeros(2, 3, dtype=torch.long))
    with pytest.raises(ValueError):
        inputs.prepare_batch(v.cuda(), torch.full((2, 4), 4, dtype=torch.long))
    with pytest.raises(ValueError):
        inputs.prepare_batch(v.float().cuda(), torch.zeros(2, 4, dtype=torch.long))
