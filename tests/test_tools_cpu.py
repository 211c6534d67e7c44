"""The measurement scripts parse and answer --help without a GPU (a typo there would cost GPU minutes, not CPU seconds),
and the shell trips are syntactically valid."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script", ["bench.py", "tools/model_bench.py", "tools/table3.py", "tools/train_step.py",
                                    "tools/layer_bench.py", "tools/op_bench.py", "tools/raft_bench.py"])
def test_python_tools_answer_help(script):
    r = subprocess.run([sys.executable, os.path.join(ROOT, script), "--help"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "usage" in r.stdout.lower()


@pytest.mark.parametrize("script", ["tools/trip_r2.sh", "tools/sanitize.sh", "tools/ncu_layers.sh", "tools/sweep_tiling.sh",
                                    "stereo_toolbox_b200/csrc/build.sh", "stereo_toolbox_b200/csrc/probe/run_tmabw2.sh"])
def test_shell_scripts_parse(script):
    r = subprocess.run(["bash", "-n", os.path.join(ROOT, script)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
