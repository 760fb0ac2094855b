"""Soak test of the product path: >= 100 training steps at the BASELINE.json configs[1] shape
(64x64, T = 20, batch 256, base_ch 64 + skip ConvLSTMs) in one process, with the background weight-gradient stream,
the cooperative timestep-persistent cell kernels and the host->device prefetcher all active, asserting that no kernel
watchdog fires (b200_device_error() stays 0) and no CUDA error surfaces.

Round 1's bench aborted about once in 150 steps: the five TMA producer warps of wgrad_tc / wgrad_tc2 all waited, by
parity, on every stage of the ring, and a warp that filled none of them could fall a whole lap behind (see
profiles/r02_fault_root_cause.md).  The run is a subprocess: a device trap would take the CUDA context of the whole
pytest process with it.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _soak(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "soak.py"), *args], capture_output=True, text=True,
                       timeout=900)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert "OK " in r.stdout and "FAULT" not in r.stdout and "FLAG" not in r.stdout, tail
    return r.stdout


def test_soak_120_training_steps_config1_no_watchdog():
    out = _soak("--steps", "120", "--phase-sync", "0")
    # the background stream really ran (25 blocks per step)
    assert "bg blocks 3000" in out, out


def test_soak_forward_backward_only_60_steps():
    # the fwd+bwd-only region of bench.py (no optimizer step between backward passes)
    _soak("--steps", "60", "--phase-sync", "0", "--opt-every", "0")
