"""The reference's OWN test files for this path (tests/test_scHPF_model.py: the estimator;
tests/test_inference.py: the kernels), run unmodified on the GPU box against the CUDA path.

`baseline/_ref/tests` (the unmodified reference install, git-ignored, shipped by gpurun) provides
the two files and their conftest; a shim package called `schpf` serves them this package: the
estimator is schpf_b200's with the real `CaviEngine`, and `schpf.hpf_numba` is
`schpf_b200.hpf_cuda` -- every kernel the reference's tests call runs as sm_100a CUDA through the
C ABI.  (CPU twin with the oracle as engine: tests/test_reference_suite_cpu.py.)

The two files are run one at a time: run together, test_inference.py::test_llh_pois[float32] fails
for the REAL reference as well (its fp32 tolerance depends on the numpy random state the other
file leaves behind)."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

REF_TESTS = os.path.join(ROOT, "baseline", "_ref", "tests")

SHIM_INIT = '''
import schpf_b200
from schpf_b200 import *                                   # scHPF, HPF_Gamma, run_trials, ...
from schpf_b200 import hpf_cuda as hpf_numba               # the reference's kernel names, CUDA inside
import sys as _sys
_sys.modules[__name__ + ".hpf_numba"] = hpf_numba
__version__ = schpf_b200.__version__
'''


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="baseline/_ref/tests did not travel")
@pytest.mark.parametrize("test_file", ["test_scHPF_model.py", "test_inference.py"])
def test_reference_test_file_passes_on_the_cuda_path(tmp_path, test_file):
    shim = tmp_path / "schpf"
    shim.mkdir()
    (shim / "__init__.py").write_text(textwrap.dedent(SHIM_INIT))
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=os.pathsep.join([str(tmp_path), ROOT]))
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "--rootdir", REF_TESTS,
           "--confcutdir", REF_TESTS, os.path.join(REF_TESTS, test_file)]
    r = subprocess.run(cmd, env=env, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    tail = "\n".join(r.stdout.strip().splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
