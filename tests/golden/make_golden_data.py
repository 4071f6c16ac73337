#!/usr/bin/env python
"""Golden items of the reference's WavenetDataset (reference audio_data.py:12-131) on a tiny synthetic dataset.npz.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_data.py
Writes tests/golden/tiny_dataset.npz (three seeded uint8 arrays) and tests/golden/dataset_items.npz (for several item /
target lengths, strides and both splits: the index sequence and the target of selected items, plus the dataset lengths).
The reference module imports librosa at top level; a stub module stands in (nothing of it is called here).
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    sys.path.insert(0, REF)
    import audio_data as ref
    rng = np.random.RandomState(3)
    tiny = os.path.join(HERE, "tiny_dataset.npz")
    np.savez(tiny, *[rng.randint(0, 256, size=n).astype(np.uint8) for n in (1500, 700, 2100)])
    out = {}
    for ci, (item_length, target_length, stride) in enumerate([(64, 16, 20), (301, 7, 1), (100, 100, 5), (33, 1, 100)]):
        for train in (True, False):
            ds = ref.WavenetDataset(dataset_file=tiny, item_length=item_length, target_length=target_length,
                                    test_stride=stride, train=train)
            n = len(ds)
            picks = sorted(set([0, 1, 2, n // 3, n // 2, n - 2, n - 1]) & set(range(n)))
            # items whose window crosses from arr_0 into arr_1 / arr_1 into arr_2
            picks += [i for i in range(n) if i not in picks][:: max(1, n // 9)]
            key = f"c{ci}_{'train' if train else 'test'}"
            out[key + "_cfg"] = np.array([item_length, target_length, stride, int(train), n])
            out[key + "_picks"] = np.array(picks, dtype=np.int64)
            if not picks:
                continue
            xs, ts = [], []
            for i in picks:
                one_hot, target = ds[i]
                assert one_hot.shape == (256, item_length) and float(one_hot.sum()) == item_length
                xs.append(one_hot.argmax(0).numpy().astype(np.uint8))
                ts.append(target.numpy())
            out[key + "_x"] = np.stack(xs)
            out[key + "_t"] = np.stack(ts)
    np.savez_compressed(os.path.join(HERE, "dataset_items.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("_x")})


if __name__ == "__main__":
    main()
