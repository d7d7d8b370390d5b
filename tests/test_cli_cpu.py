"""CPU tests of the host driver's command line (damapper_b200/damapper): everything it rejects before it
touches CUDA must behave as the reference's driver does (damapper.c:583-729), plus the -G device list."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "damapper_b200", "damapper")


def _run(args, cwd, env=None):
    return subprocess.run([EXE] + list(args), cwd=cwd, env=dict(os.environ, **(env or {})), capture_output=True,
                          text=True, timeout=60)


@pytest.fixture(scope="module")
def dbs(tmp_path_factory):
    if not os.access(EXE, os.X_OK):
        pytest.skip("host driver not built (python -c 'import __graft_entry__ as g; g.build()')")
    from damapper_b200 import dazzdb, synth
    wd = str(tmp_path_factory.mktemp("cli"))
    contigs, rb, rl = synth.make_config("C1", scale=0.01, seed=5)
    dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
    w = dazzdb.StreamDBWriter(os.path.join(wd, "reads.db"))
    w.append(rb, rl)
    w.close(nblocks=3)
    return wd


def test_usage_and_flag_errors(dbs):
    p = _run([], dbs)
    assert p.returncode == 1 and "Usage: damapper" in p.stderr and "-G<int(1)>" in p.stderr
    for args, msg in [(["-N", "ref.dam", "reads.db"], "Cannot specify N flag without C also"),
                      (["-C", "-N", "-p", "ref.dam", "reads.db"], "Cannot specify both N and p flags together"),
                      (["-k33", "ref.dam", "reads.db"], "K-mer length must be 32 or less"),
                      (["-k0", "ref.dam", "reads.db"], "K-mer length must be positive"),
                      (["-e.5", "ref.dam", "reads.db"], "Average correlation must be in [.7,1.)"),
                      (["-n.5", "ref.dam", "reads.db"], "Near optimal threshold must be in [.7,1.]"),
                      (["-sx", "ref.dam", "reads.db"], "argument is not an integer"),
                      (["-q", "ref.dam", "reads.db"], "-q is an illegal option"),
                      (["-P/no/such/dir", "ref.dam", "reads.db"], "cannot open directory"),
                      (["-G0", "ref.dam", "reads.db"], "Number of GPUs must be positive")]:
        p = _run(args, dbs)
        assert p.returncode == 1, args
        assert msg in p.stderr, (args, p.stderr)


def test_reference_must_be_a_partitioned_whole_db(dbs):
    p = _run(["reads.1", "reads.2"], dbs)                  # a block as first argument
    assert p.returncode == 1 and "cannot be a block" in p.stderr
    p = _run(["nosuch.dam", "reads.1"], dbs)
    assert p.returncode == 1


def test_multi_gpu_device_list_is_checked_before_cuda(dbs):
    """-G<n> with fewer devices than workers in DAMGPU_DEVICES / CUDA_VISIBLE_DEVICES stops before any worker is
    forked (and before CUDA is touched, so this runs without a GPU); -G larger than the number of reads blocks is
    clamped to it."""
    p = _run(["-G3", "ref.dam", "reads.1", "reads.2", "reads.3"], dbs, {"DAMGPU_DEVICES": "0,1"})
    assert p.returncode == 1 and "-G3 but only 2 devices in '0,1'" in p.stderr
    p = _run(["-G8", "ref.dam", "reads.1", "reads.2"], dbs, {"DAMGPU_DEVICES": "5"})
    assert p.returncode == 1 and "-G2 but only 1 devices in '5'" in p.stderr
    p = _run(["-G2", "ref.dam", "reads.1", "reads.2"], dbs, {"CUDA_VISIBLE_DEVICES": "3", "DAMGPU_DEVICES": ""})
    assert p.returncode == 1
