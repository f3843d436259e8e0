"""The validation rules of the SPEC builder (csrc/spec.cuh: search / sweep thresholds, strict rows, length bounds, appends and
removals as operations) replayed on the CPU by tools/sim_spec_build.cpp: every insert of a window is executed twice — against
the graph as it stood at the window start (what a speculative execution sees) and in stream order (core.rs:489-599) — and
whenever the rules call the speculative execution valid, its writes must BE the sequential ones.  No GPU involved; the device
kernels are checked against the oracle's graphs in tests/test_gpu_spec_build.py."""
import os
import re
import subprocess

import numpy as np
import pytest

from redis_hnsw_b200 import data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    d = tmp_path_factory.mktemp("sim")
    exe = str(d / "sim_spec_build")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "sim_spec_build.cpp")])
    return d, exe


def _replay(sim, n, dim, m, efc, window, env_extra=()):
    d, exe = sim
    x, _ = data.lowrank(n, dim, r=8, seed=5)
    path = str(d / ("v_%d_%d.bin" % (n, dim)))
    x.tofile(path)
    env = dict(os.environ, SIM_VERIFY=str(window), SIM_WINDOWS="40")
    env.update(dict(env_extra))
    out = subprocess.run([exe, path, str(n), str(dim), str(m), str(efc), "7", str(n // 5), str(n // 2), str(n - 800)],
                         env=env, capture_output=True, text=True, check=True).stdout
    rows = re.findall(r"accepted coarse (\d+) fine (\d+) fine\+oplog (\d+) \| violations (\d+) (\d+) (\d+)", out)
    assert len(rows) == 3, out
    return np.array(rows, dtype=np.int64)


@pytest.mark.parametrize("n,dim,m,efc,window", [(6000, 32, 8, 40, 16), (5000, 64, 5, 24, 32)])
def test_accepted_speculative_executions_equal_the_sequential_ones(sim, n, dim, m, efc, window):
    r = _replay(sim, n, dim, m, efc, window)
    assert np.all(r[:, 3:] == 0), "a validation rule accepted an execution that differs from the sequential one: %s" % r
    assert np.all(r[:, 2] > r[:, 1]) and np.all(r[:, 1] > r[:, 0]), "finer rules must accept more executions: %s" % r
    assert r[:, 2].sum() > 500        # the replay saw enough accepted executions to mean something


def test_the_replay_catches_an_unsound_rule(sim):
    """Negative control: without the strict rule (a re-selected row must be unchanged) the replay has to find executions
    that were accepted but differ."""
    r = _replay(sim, 6000, 32, 8, 40, 16, env_extra={"SIM_UNSOUND": "1"}.items())
    assert r[:, 5].sum() > 0
    assert np.all(r[:, 3:5] == 0)
