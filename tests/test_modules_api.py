"""The reference-named module functions of the product package (dilate / DilatedQueue / constant_pad_1d)
against the reference's own known answers (tests/test_modules.py:8-36, tests/test_tensor_queue.py:13-50,
:103-120) and the arrays the unmodified reference produced (tests/golden/modules.npz, queue.npz)."""
import numpy as np
import pytest
import torch

import wavenet_modules as wm
from oracle import wavenet_oracle as O


def test_dilate_known_answers():
    x = torch.linspace(0, 12, steps=13).view(1, 1, 13)
    d = wm.dilate(x, 1)
    assert d.size() == (1, 1, 13) and d[0, 0, 4] == 4
    d = wm.dilate(x, 2)
    assert d.size() == (2, 1, 7) and d[1, 0, 2] == 4
    d = wm.dilate(d, 4, init_dilation=2)
    assert d.size() == (4, 1, 4) and d[3, 0, 1] == 4
    d = wm.dilate(d, 1, init_dilation=4)
    assert d.size() == (1, 1, 16) and d[0, 0, 7] == 4


def test_dilate_matches_reference_arrays(golden):
    g = golden("modules.npz")
    x = torch.from_numpy(g["x13"])
    d2 = wm.dilate(x, 2)
    d4 = wm.dilate(d2, 4, init_dilation=2)
    d1 = wm.dilate(d4, 1, init_dilation=4)
    for got, key in ((d2, "d2"), (d4, "d4"), (d1, "d1")):
        assert np.array_equal(got.numpy(), g[key])
    xm = torch.from_numpy(g["xm"])
    assert wm.dilate(xm, 2).shape == (4, 3, 3) and wm.dilate(xm, 4).shape == (8, 3, 2)
    assert np.array_equal(wm.dilate(xm, 2).numpy(), g["xm2"])
    assert np.array_equal(wm.dilate(xm, 4).numpy(), g["xm4"])


@pytest.mark.parametrize("n,c,l,d0,d1", [(1, 3, 17, 1, 4), (2, 2, 9, 1, 2), (4, 5, 8, 4, 1), (8, 1, 5, 8, 2),
                                          (3, 2, 10, 1, 8), (1, 1, 1, 1, 2)])
def test_dilate_equals_oracle_fold(n, c, l, d0, d1):
    x = torch.randn(n, c, l)
    assert torch.equal(wm.dilate(x, d1, init_dilation=d0), O.fold_time(x, d1, init_dilation=d0))
    assert torch.equal(wm.dilate(x, d1, init_dilation=d0, pad_start=False),
                       O.fold_time(x, d1, init_dilation=d0, pad_start=False))


def test_queue_known_answers(golden):
    q = wm.DilatedQueue(max_length=8, num_channels=3)
    e = torch.zeros(3)
    for _ in range(11):
        e = e + 1
        q.enqueue(e)
    assert q.data[0, 0] == 9 and q.data[0, 2] == 11 and q.data[0, 7] == 8
    q = wm.DilatedQueue(max_length=8, num_channels=1)
    e = torch.zeros(1)
    for _ in range(11):
        e = e + 1
        q.enqueue(e)
    for _ in range(9):
        d = q.dequeue(num_deq=3, dilation=2)
    assert d[0].tolist() == [5, 7, 9]
    g = golden("queue.npz")
    q = wm.DilatedQueue(max_length=12, num_channels=2)
    e = torch.zeros(2)
    for i in range(30):
        e = e + 1
        q.enqueue((e * torch.tensor([1.0, -1.0])).view(2, 1))        # the model passes (R,1) columns
        d = q.dequeue(num_deq=3, dilation=4)
        assert d[0][0] == max(i - 7, 0)
        assert np.array_equal(d.numpy(), g["combined"][i])
    assert np.array_equal(q.data.numpy(), g["final"]) and q.in_pos == g["in_pos"] and q.out_pos == g["out_pos"]
    q.reset()
    assert q.in_pos == 0 and q.out_pos == 0 and float(q.data.abs().sum()) == 0.0


def test_constant_pad(golden):
    g = golden("modules.npz")
    x = torch.arange(6.).view(2, 3)
    assert np.array_equal(wm.constant_pad_1d(x, 5, dimension=1, value=7.0).numpy(), g["pad_end"])
    assert np.array_equal(wm.constant_pad_1d(x, 5, dimension=1, pad_start=True).numpy(), g["pad_start"])
    with pytest.raises(AssertionError):
        wm.constant_pad_1d(torch.zeros(4), 3)
    # reference tests/test_tensor_queue.py:103-120: pad-at-end shape (5,3,4), gradient shape (2,3,4)
    x = torch.ones(2, 3, 4, requires_grad=True)
    y = wm.constant_pad_1d(x, 5, dimension=0, pad_start=False)
    assert y.shape == (5, 3, 4)
    y.sum().backward()
    assert x.grad.shape == (2, 3, 4) and float(x.grad.sum()) == 24.0
    x = torch.randn(2, 3, requires_grad=True)
    w = torch.randn(2, 7)
    (wm.constant_pad_1d(x, 7, dimension=1, pad_start=True) * w).sum().backward()
    assert torch.equal(x.grad, w[:, 4:])
