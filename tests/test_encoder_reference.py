"""The PyTorch encoder attached by baseline.attach_encoder() (out of scope of the CUDA path, kept
for drop-in `forward(images)`) against the reference's own encoder, live (needs /root/reference)."""
import pytest
import torch

import refload

pytestmark = pytest.mark.skipif(not refload.reference_available(), reason="reference not mounted")


def test_encoder_matches_reference_cpu():
    ns = refload.load_reference_model("OSIE")
    torch.manual_seed(0)
    ref = ns.model.baseline().eval()
    from scanpaths_b200.models.baseline_attention import baseline
    ours = baseline(task="OSIE").attach_encoder()
    missing, unexpected = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not missing and not unexpected
    x = torch.randn(1, 3, 240, 320)
    with torch.no_grad():
        a = torch.relu(ref.sal_conv(ref.resnet(x)))
        b = ours.encode(x)
    assert a.shape == b.shape == (1, 512, 30, 40)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), (a - b).abs().max()
