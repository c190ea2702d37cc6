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


def _randomise_batchnorm(net, seed=1):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(0.3 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.bias.data.copy_(0.2 * torch.randn(m.num_features, generator=g))


def test_fused_decoder_matches_plain_decode_cpu():
    """BatchNorm folding, channel padding and the channels-last glue of FusedDecoder reproduce network.decode
    (CPU, plain conv2d path; the cuDNN fused conv+bias+relu is covered by the GPU test)."""
    from molearn_b200.models import AutoEncoder, FusedDecoder
    torch.manual_seed(0)
    net = AutoEncoder(out_points=300).eval()
    _randomise_batchnorm(net)
    fused = FusedDecoder(net.decoder)
    z = torch.randn(5, 2)
    with torch.no_grad():
        want = net.decode(z)
        got = fused(z)
    assert got.shape == want.shape == (5, 384, 3)
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5 * float(want.abs().max()))
    net.train()
    with pytest.raises(RuntimeError):
        FusedDecoder(net.decoder)


@pytest.mark.skipif(not os.path.isfile(REF), reason="reference tree not available")
def test_fused_decoder_accepts_the_reference_network():
    from molearn_b200.models import fuse_decoder
    spec = importlib.util.spec_from_file_location("_ref_foldingnet2", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(0)
    theirs = ref.AutoEncoder(out_points=300).eval()
    _randomise_batchnorm(theirs)
    fused = fuse_decoder(theirs)
    assert fused is not None
    z = torch.randn(3, 2)
    with torch.no_grad():
        want = theirs.decode(z)
        assert torch.allclose(fused(z), want, rtol=1e-4, atol=1e-5 * float(want.abs().max()))
    assert fuse_decoder(torch.nn.Linear(2, 2)) is None
    assert fuse_decoder(ref.AutoEncoder(out_points=300).double().eval()) is None     # fp32 networks only
