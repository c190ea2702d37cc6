"""CPU: the FoldingNet-style autoencoder is state-dict compatible with the reference's and computes
the same function (needs the reference tree; skipped on machines without it)."""
import importlib.util
import os

import pytest
import torch

REF = os.path.join(os.environ.get("MOLEARN_REF", "/root/reference"), "src", "molearn", "models", "foldingnet.py")


def test_shapes_and_strides():
    from molearn_b200.models import AutoEncoder
    torch.manual_seed(0)
    net = AutoEncoder(out_points=300).eval()
    x = torch.randn(3, 300, 3)
    with torch.no_grad():
        z = net.encode(x)
        y = net.decode(z)
    assert z.shape == (3, 2)
    assert y.shape == (3, 384, 3) and y.stride() == (3 * 384, 1, 384)   # permuted view of [B,3,M]


@pytest.mark.skipif(not os.path.isfile(REF), reason="reference tree not available")
def test_matches_reference_network():
    from molearn_b200.models import AutoEncoder
    spec = importlib.util.spec_from_file_location("_ref_foldingnet", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(0)
    theirs = ref.AutoEncoder(out_points=300).eval()
    ours = AutoEncoder(out_points=300).eval()
    missing = ours.load_state_dict(theirs.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x = torch.randn(4, 300, 3)
    with torch.no_grad():
        assert torch.allclose(ours.encode(x), theirs.encode(x), rtol=1e-4, atol=1e-5)
        z = torch.randn(4, 2)
        assert torch.allclose(ours.decode(z), theirs.decode(z), rtol=1e-4, atol=1e-5)
