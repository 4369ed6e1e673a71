"""`.th` cross-compatibility in the other direction (CPU only; needs the reference tree, i.e. the authoring container):
a checkpoint written by the PRODUCT's save() is loaded by the unmodified reference and reproduces the model."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_reference = pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present")


@needs_reference
def test_product_written_checkpoint_loads_in_the_reference(tmp_path):
    import iffnerf_b200 as I
    from oracle.ref_import import import_reference
    R = import_reference()
    src = torch.load(os.path.join(GOLDEN, "c_ckpt_reference.th"), map_location="cpu", weights_only=False)
    kw = dict(src["kwargs"])
    kw.update(device="cpu")
    with contextlib.redirect_stdout(io.StringIO()):
        ours = I.TensorVMSplit(**kw)              # parameter container only: nothing renders on the CPU
        ours.load(src)
        path = str(tmp_path / "product.th")
        ours.save(path)
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        assert set(ckpt.keys()) == set(src.keys()) and ckpt["model_name"] == src["model_name"]
        assert set(ckpt["kwargs"].keys()) == set(src["kwargs"].keys())
        kr = dict(ckpt["kwargs"])
        kr.update(device="cpu")
        ref = R.TensorVMSplit(**kr)
        ref.load(ckpt)
    want = src["state_dict"]
    got = ref.state_dict()
    assert set(got.keys()) == set(want.keys())
    for k in want:
        assert torch.equal(got[k], want[k]), k
    assert np.array_equal(ckpt["alphaMask.mask"], src["alphaMask.mask"])
    assert tuple(ckpt["alphaMask.shape"]) == tuple(src["alphaMask.shape"])
    assert torch.equal(ckpt["alphaMask.aabb"], src["alphaMask.aabb"])
    assert torch.equal(ref.alphaMask.alpha_volume, ours.alphaMask.alpha_volume.cpu())
    assert ref.nSamples == ours.nSamples and torch.equal(ref.stepSize, ours.stepSize.cpu())
