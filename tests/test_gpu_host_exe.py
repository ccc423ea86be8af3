"""The host executable `swarmdyn` end to end on the GPU: the reference's own command line (built by
the parameter layer) in, the reference's text output files out."""
import os
import subprocess

import numpy as np
import pytest

from sde_burst_coast_b200 import params as P

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, 'sde_burst_coast_b200', 'host', 'swarmdyn')


def _run(dic, cwd, extra=()):
    if not os.path.exists(EXE):
        subprocess.run(['make', '-C', os.path.dirname(EXE)], check=True)
    cmd = P.dic2swarmdyn_command(dic, exe=EXE).split() + list(extra)
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith('Go')
    return r


def _rows(path):
    return np.array([[float(v) for v in line.split()] for line in open(path) if line.strip()])


def test_c1_tank_run_writes_reference_files(tmp_path):
    d = P.get_base_params(1, 2, mode='burst_coast')
    P.selection_line_parameter(d, 2)
    d.update(out_h5=0, output_mode=1, N=8)
    _run(d, tmp_path, ['-Z', '11'])
    sw = _rows(tmp_path / 'swarm_xx.dat')
    part = _rows(tmp_path / 'part_xx.dat')
    assert sw.shape[1] == 16 and part.shape[1] == 7
    n_out = sw.shape[0]
    assert n_out == len([s for s in range(3000) if s % 30 == 0 and s >= 1000])
    assert part.shape[0] == n_out * 8
    assert np.all(sw[:, 0] == 8)
    assert np.all((sw[:, 1] >= 0) & (sw[:, 1] <= 1 + 1e-6))            # polarisation
    assert np.all(np.hypot(part[:, 0], part[:, 1]) <= 24.5 + 2e-3)  # inside the tank (values are printed with 6 significant digits)
    assert np.all(np.isnan(sw[:, 15]))                                    # ND quirk (swarmdyn.cpp:270)
    assert not (tmp_path / 'pred_xx.dat').exists()
    # same seed -> same files; different seed -> different trajectory
    other = tmp_path / 'b'
    other.mkdir()
    _run(d, other, ['-Z', '11'])
    assert open(other / 'part_xx.dat').read() == open(tmp_path / 'part_xx.dat').read()
    third = tmp_path / 'c'
    third.mkdir()
    _run(d, third, ['-Z', '12'])
    assert open(third / 'part_xx.dat').read() != open(tmp_path / 'part_xx.dat').read()


def test_natpred_run_kills_and_writes_predator(tmp_path):
    d = P.get_base_params(0.5, 3, mode='natPred')
    P.selection_line_parameter(d, 2)
    d.update(out_h5=0, output_mode=1, N=30, pred_speed0=20, kill_rate=50, N_confu=40)
    _run(d, tmp_path, ['-Z', '5'])
    pred = _rows(tmp_path / 'pred_xx.dat')
    net = _rows(tmp_path / 'swarm_fishNet_xx.dat')
    assert pred.shape[1] == 6 and net.shape[1] == 4
    assert np.allclose(np.hypot(pred[:, 2], pred[:, 3]), 20, rtol=1e-5)   # constant predator speed
    assert net[-1, 3] >= net[0, 3] and net[-1, 3] > 0                      # Ndead grows


def test_ensemble_flag_runs_replicates(tmp_path):
    d = P.get_base_params(0.2, 0.5, mode='burst_coast')
    P.selection_line_parameter(d, 2)
    d.update(out_h5=0, output_mode=0, N=16, IC=0)
    _run(d, tmp_path, ['-Z', '3', '-y', '4'])
    a = _rows(tmp_path / 'swarm_xx.dat')
    b = _rows(tmp_path / 'swarm_xx_3.dat')
    assert a.shape == b.shape and a.shape[1] == 16 and not np.allclose(a[:, 3], b[:, 3])


def test_voronoi_flag(tmp_path):
    d = P.get_base_params(0.2, 1, mode='burst_coast')
    P.selection_line_parameter(d, 2)
    d.update(out_h5=0, output_mode=0, N=8)
    _run(d, tmp_path, ['-Z', '3', '-V', '1'])
    a = _rows(tmp_path / 'swarm_xx.dat')
    other = tmp_path / 'metric'
    other.mkdir()
    _run(d, other, ['-Z', '3'])
    b = _rows(other / 'swarm_xx.dat')
    assert a.shape == b.shape and not np.allclose(a[-1, 3:5], b[-1, 3:5])   # different neighbour rule, different run
