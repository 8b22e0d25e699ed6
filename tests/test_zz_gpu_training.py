"""Training path on the GPU (first-correct version): runs tests/train_gpu_cases.py in a subprocess.

Status: the path was written after this round's GPU minutes were spent.  Its kernel source is checked on the CPU
(tests/test_train_emu.py runs train.cuh / train_host.cuh on host threads against the oracle), but it has not met a GPU
yet -- hence a non-strict expected failure, last in the run and in its own process, so it can neither hide a fault nor
disturb the validated sampling tests.  An XPASS in the log means the cases passed on the GPU.
"""
import os
import subprocess
import sys
import warnings

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first GPU run of the training path (written with no GPU minutes left; CPU emulation passes)")
def test_training_path_cases_on_the_gpu():
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "train_gpu_cases.py"), "-x", "-q", "-m", "gpu"],
                       cwd=os.path.dirname(HERE), capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-2000:])
    tail = [ln for ln in r.stdout.strip().splitlines() if ln.strip()][-1:] or ["no output"]
    # the summary line of the inner run shows up in this run's warnings summary whatever the outcome is reported as
    warnings.warn("training path GPU cases: rc=%d, %s" % (r.returncode, tail[0].strip("= ")))
    try:
        os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
        with open(os.path.join(os.path.dirname(HERE), "gpurun_out", "training_gpu_cases.txt"), "w") as f:
            f.write(r.stdout[-20000:] + "\n--- stderr ---\n" + r.stderr[-5000:])
    except OSError:
        pass
    assert r.returncode == 0
