"""mu-law helpers with the reference's names (reference audio_data.py:133-158).

Only the pure functions the hot path touches are provided (``generate_fast`` ends with ``mu_law_expansion``);
the dataset / file-loading half of the reference module (librosa, .npz building) is out of scope.
Note the reference uses mu = classes (256), not classes - 1.
"""
import numpy as np


def mu_law_encoding(data, mu):
    return np.sign(data) * np.log(1 + mu * np.abs(data)) / np.log(mu + 1)


def mu_law_expansion(data, mu):
    return np.sign(data) * (np.exp(np.abs(data) * np.log(mu + 1)) - 1) / mu


def quantize_data(data, classes):
    mu_x = mu_law_encoding(data, classes)
    bins = np.linspace(-1, 1, classes)
    return np.digitize(mu_x, bins) - 1
