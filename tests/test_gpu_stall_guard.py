"""The kernels' stall guard (include/b2q.h, "Stall guard"): a pipeline wait that can never be satisfied must end in a
record + trap within seconds, and the record must be readable after the CUDA context is gone.  Runs in a child process
because the trap poisons the context of whoever launched the kernel."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("B2Q_STALL_SELFTEST") != "1",
                    reason="deliberately traps a kernel (the driver logs an Xid for it): opt-in with B2Q_STALL_SELFTEST=1; "
                           "last run: profiles/r02_stall_selftest.log")
def test_stall_guard_selftest_reports_and_traps(lib_built):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stall_selftest.py")], capture_output=True, text=True,
                       timeout=180, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "selftest OK" in r.stdout and "b2q stall guard" in r.stdout


def test_stall_report_is_empty_without_a_gpu(lib_built):
    import b200qlora as q

    assert q._lib.load().b2q_debug_stall_count() == 0 and q._lib.stall_report() == ""
