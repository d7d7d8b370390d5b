import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


_CAPMAN = None
_FATAL_CB = None


def pytest_configure(config):
    global _CAPMAN
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    _CAPMAN = config.pluginmanager.getplugin("capturemanager")


def install_fatal_hook(api):
    """The library reports fatal errors through the Clean_Exit callback (map.h:39) and then exits; under
    pytest's capture that would end the run without a word.  The hook prints the message on the real
    stderr before the process goes."""
    import ctypes as C
    global _FATAL_CB

    def on_fatal(code):
        try:
            if _CAPMAN is not None:
                _CAPMAN.suspend_global_capture(in_=True)
        except Exception:
            pass
        try:
            msg = api.load().damgpu_last_error().decode(errors="replace")
        except Exception:
            msg = "?"
        sys.__stderr__.write("\n\nFATAL inside libdamgpu during a test (the process exits, code %d): %s\n" % (code, msg))
        sys.__stderr__.flush()
        os._exit(3)

    _FATAL_CB = C.CFUNCTYPE(None, C.c_int)(on_fatal)
    api.load().damgpu_set_fatal(C.cast(_FATAL_CB, C.c_void_p))


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
