"""Training path on the GPU (first-correct version): runs tests/train_gpu_cases.py in a subprocess.

Status: the path was written after this round's GPU minutes were spent.  Its kernel source is checked on the CPU
(tests/test_train_emu.py runs train.cuh / train_host.cuh on host threads against the oracle), but it has not met a GPU
yet -- hence a non-strict expected failure, last in the run and in its own process, so it can neither hide a fault nor
disturb the validated sampling tests.  An XPASS in the log means the cases passed on the GPU.
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first GPU run of the training path (written with no GPU minutes left; CPU emulation passes)")
def test_training_path_cases_on_the_gpu():
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "train_gpu_cases.py"), "-x", "-q", "-m", "gpu"],
                       cwd=os.path.dirname(HERE), capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0
