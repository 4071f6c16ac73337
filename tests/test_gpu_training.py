"""The caller side of the training path: fused cross-entropy, fused Adam, WavenetTrainer."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from helpers import rel_err

pytestmark = pytest.mark.gpu


def test_fused_cross_entropy_matches_torch():
    import wavenet_training as wt
    g = torch.Generator().manual_seed(0)
    for n, c in ((1000, 256), (37, 256), (513, 100)):
        x = (torch.randn(n, c, generator=g) * 3).cuda().requires_grad_(True)
        t = torch.randint(0, c, (n,), generator=g).cuda()
        loss = wt.fused_cross_entropy(x, t)
        (loss * 2.5).backward()
        x2 = x.detach().clone().requires_grad_(True)
        want = F.cross_entropy(x2, t)
        (want * 2.5).backward()
        assert abs(float(loss) - float(want)) < 1e-6 * max(1.0, abs(float(want)))
        assert rel_err(x.grad.cpu().numpy(), x2.grad.cpu().numpy()) < 1e-6


def test_fused_adam_matches_torch_adam():
    import wavenet_training as wt
    torch.manual_seed(1)
    shapes = [(256, 256, 2), (256,), (5000,), (3, 7)]
    a = [torch.nn.Parameter(torch.randn(*s).cuda()) for s in shapes]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa = wt.FusedAdam(a, lr=3e-3, weight_decay=1e-2)
    ob = torch.optim.Adam(b, lr=3e-3, weight_decay=1e-2)
    for step in range(4):
        for p, q in zip(a, b):
            gr = torch.randn_like(p) * (step + 1)
            p.grad, q.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    for p, q in zip(a, b):
        assert rel_err(p.detach().cpu().numpy(), q.detach().cpu().numpy()) < 2e-6


@pytest.mark.parametrize("one_hot", [True, False])
def test_trainer_runs_and_learns(one_hot):
    """A few steps of WavenetTrainer on the tiny dataset: the loss of the fused path equals F.cross_entropy on the same batch,
    the index-input and one-hot items give the same step, and repeated steps on one batch reduce the loss."""
    import audio_data
    import wavenet_model as wmod
    import wavenet_training as wt
    torch.manual_seed(0)
    m = wmod.WaveNetModel(layers=4, blocks=2, dilation_channels=32, residual_channels=32, skip_channels=64, end_channels=64,
                          classes=256, output_length=16, kernel_size=2, bias=True).cuda()
    ds = audio_data.WavenetDataset(dataset_file=os.path.join(GOLDEN, "tiny_dataset.npz"),
                                   item_length=m.receptive_field + m.output_length - 1, target_length=m.output_length,
                                   test_stride=20, one_hot=one_hot)
    tr = wt.WavenetTrainer(m, ds, lr=2e-3, num_workers=0, logger=wt.Logger(log_interval=10 ** 9, validation_interval=10 ** 9))
    x, t = ds[3]
    logits = tr._logits(x.unsqueeze(0))
    want = F.cross_entropy(logits, t.view(-1).cuda())
    got = wt.fused_cross_entropy(logits, t.view(-1).cuda())
    assert abs(float(got) - float(want)) < 1e-5
    first = float(got)
    steps = tr.train(batch_size=8, epochs=1, max_steps=12)
    assert steps == 12
    with torch.no_grad():
        after = float(F.cross_entropy(tr._logits(x.unsqueeze(0)), t.view(-1).cuda()))
    assert after < first, (first, after)
    loss, acc = tr.validate()
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0 and m.training and ds.train
