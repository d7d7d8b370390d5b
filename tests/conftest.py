import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def base_freq(contigs):
    allb = np.concatenate(contigs)
    cnt = np.bincount(allb, minlength=4).astype(np.float64)
    return tuple(float(x) for x in (cnt / cnt.sum()).astype(np.float32))


@pytest.fixture(scope="session")
def oracle_mod():
    """The C oracle (test infrastructure): built on demand with gcc."""
    from oracle import oracle as orc
    orc.build()
    return orc


def make_case(cfg, scale, seed):
    """(contigs, reads bases, read lengths, loaded reads block, loaded ref block, rc ref block)"""
    from damapper_b200 import synth, dazzdb
    contigs, rb, rl = synth.make_config(cfg, scale=scale, seed=seed)
    rd = dazzdb.load_block((rb, rl))
    rf = dazzdb.load_block(contigs)
    rc = dazzdb.load_block(dazzdb.revcomp_contigs(contigs))
    return contigs, rb, rl, rd, rf, rc
