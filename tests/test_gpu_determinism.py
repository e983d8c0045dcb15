"""GPU: the engine is bit-deterministic and reads nothing it did not write.
(1) eval pass + train forward + backward repeated on fresh workspaces give identical bits;
(2) a workspace pre-filled with zeros, NaN or 1e30 gives identical logits and gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('head', [None, 'mlp'])
@pytest.mark.parametrize('n', [7, 20, 110])
def test_repeatable_and_workspace_independent(head, n):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from b200ocl import nets
    torch.manual_seed(0)
    model = nets.Reduced_ResNet18(100) if head is None else nets.SupConResNet(head=head)
    eng = nets.engine_of(model)
    g = torch.Generator(device='cuda').manual_seed(n)
    x = torch.rand(n, 3, 32, 32, device='cuda', generator=g)
    bn0 = eng.state.bn_stats.clone()
    first = None
    for fill in (0.0, float('nan'), 1e30, 0.0):
        eng.state.bn_stats.copy_(bn0)
        feat = eng.features_eval(x).clone()
        ws = eng.new_train_workspace(n)
        ws.view(torch.float32).fill_(fill)
        out, ws = eng.forward_train(x, ws=ws)
        dout = torch.randn(out.shape, device='cuda', generator=torch.Generator(device='cuda').manual_seed(5)) / n
        eng.backward(x, dout, ws)
        cur = (feat, out.clone(), eng.state.grads.clone())
        assert not torch.isnan(cur[2]).any()
        if first is None:
            first = cur
        else:
            for a, b in zip(cur, first):
                assert torch.equal(a, b)
