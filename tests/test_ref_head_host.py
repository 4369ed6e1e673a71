"""Host side of the `Ref` head (iffnerf_b200/ref_head.py) against the unmodified reference's models/ref.py, imported
from /root/reference or the travelling copy oracle/_ref (skipped when neither exists): identical `state_dict` (keys,
order and bits — the directional-encoding table included) from one seed for several head configurations, and the
plain-tensor evaluation within fp32 rounding of the reference's."""
import pytest
import torch

from oracle import ref_import

CONFIGS = [{}, {"rgb_premultiplier": 0.7, "rgb_bias": 0.2}, {"deg_view": 3, "feature_c": 64}, {"deg_view": 5},
           {"deg_view": 1}, {"predicted_normals": False}]


@pytest.mark.skipif(ref_import.reference_root() is None, reason="no reference tree (oracle/_ref not built)")
@pytest.mark.parametrize("kw", CONFIGS, ids=lambda kw: ",".join(f"{k}={v}" for k, v in kw.items()) or "default")
def test_ref_head_state_and_function_match_reference(kw):
    ref_import.import_reference()
    from models.ref import Ref as ReferenceRef
    from iffnerf_b200.ref_head import Ref
    torch.manual_seed(5)
    theirs = ReferenceRef(27, **kw)
    torch.manual_seed(5)
    ours = Ref(27, **kw)
    sa, sb = theirs.state_dict(), ours.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert sa[k].dtype == sb[k].dtype and torch.equal(sa[k], sb[k]), k
    ours.load_state_dict(sa, strict=True)
    theirs.load_state_dict(sb, strict=True)
    g = torch.Generator().manual_seed(1)
    feat = torch.randn(2000, 27, generator=g)
    view = torch.nn.functional.normalize(torch.randn(2000, 3, generator=g), dim=-1)
    normals = None if kw.get("predicted_normals", True) else torch.nn.functional.normalize(
        torch.randn(2000, 3, generator=g), dim=-1)
    want, _ = theirs(None, view, feat, normals)
    got, _ = ours(None, view, feat, normals)
    assert (got - want).abs().max().item() <= 4e-7          # a few fp32 ulps of an O(1) colour
    if kw.get("predicted_normals", True):
        assert torch.equal(ours.compute_normals(feat), theirs.compute_normals(feat))
