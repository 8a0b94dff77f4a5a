import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: g[k] for k in g.files}


def golden_weights(g):
    import oracle as O
    wc = O.init_weights(int(g["w_seed_coarse"]), float(g["w_bias_range"]))
    wf = O.init_weights(int(g["w_seed_fine"]), float(g["w_bias_range"]))
    assert np.array_equal(O.flatten_weights(wc)[::997], g["w_coarse_sample"])
    assert np.array_equal(O.flatten_weights(wf)[::997], g["w_fine_sample"])
    return wc, wf


def psnr_db(a, b):
    return float(-10.0 * np.log10(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)))


def cuda(x):
    return torch.as_tensor(np.asarray(x)).cuda()
