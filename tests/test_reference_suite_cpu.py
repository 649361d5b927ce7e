"""The reference's OWN test files for this path (tests/test_scHPF_model.py: the estimator shell;
tests/test_inference.py: the kernels), run unmodified against this package through a shim
package called `schpf`.

Only where the reference checkout exists (the build container; the files are read in place,
never copied) and without a GPU: the estimator under test is schpf_b200's, its engine and the
function-level kernels are the CPU oracle, so what this pins is (a) the host logic of the
drop-in estimator against the reference's own expectations and (b) the oracle against the
reference's known-answer tests for the kernels.  The CUDA path meets the same
expectations in tests/test_gpu_kernels.py / test_gpu_engine.py, re-expressed so that they can
run on the GPU box, where /root/reference does not exist.

The two files are run one at a time: run together, test_inference.py::test_llh_pois[float32] fails
for the REAL reference as well (its fp32 tolerance depends on the numpy random state the other file
leaves behind), and the shimmed package reproduces exactly that outcome (59 passed, 1 failed)."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

REF_TESTS = os.path.join(os.environ.get("SCHPF_REFERENCE", "/root/reference"), "tests")

SHIM_INIT = '''
import schpf_b200
from schpf_b200 import *                                   # scHPF, HPF_Gamma, run_trials, ...
from schpf_b200 import scHPF_ as _shell, loss as _loss
from oracle_engine import OracleEngine as _OracleEngine
from oracle import hpf_numpy as _onp
_shell._engine_factory = _OracleEngine                     # no GPU here: arithmetic from the CPU oracle
_loss.compute_pois_llh = _onp.compute_pois_llh
from . import hpf_numba
__version__ = schpf_b200.__version__
'''

SHIM_KERNELS = '''
"""The reference's kernel names (schpf/hpf_numba.py), served by the CPU oracle."""
import numpy as np
from oracle.hpf_numpy import (compute_Xphi_data, compute_loading_shape_update, compute_loading_rate_update,
                              compute_capacity_rate_update, compute_pois_llh)
from oracle import hpf_numpy as _onp


def psi(x):
    return float(_onp.psi(x))


def cgammaln(x):
    return float(_onp.cgammaln(x))
'''


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference checkout not present")
@pytest.mark.parametrize("test_file", ["test_scHPF_model.py", "test_inference.py"])
def test_reference_test_file_passes_against_this_package(tmp_path, test_file):
    shim = tmp_path / "schpf"
    shim.mkdir()
    (shim / "__init__.py").write_text(textwrap.dedent(SHIM_INIT))
    (shim / "hpf_numba.py").write_text(textwrap.dedent(SHIM_KERNELS))
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1",
               PYTHONPATH=os.pathsep.join([str(tmp_path), ROOT, os.path.join(ROOT, "tests")]))
    # rootdir / confcutdir = the reference's tests directory: its conftest.py provides the fixtures
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "--rootdir", REF_TESTS,
           "--confcutdir", REF_TESTS, os.path.join(REF_TESTS, test_file)]
    r = subprocess.run(cmd, env=env, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    tail = "\n".join(r.stdout.strip().splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
