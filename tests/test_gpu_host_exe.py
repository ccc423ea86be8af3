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


def _reference_main(argv, cwd):
    """Run the reference's own main() (oracle/_ref, swarmdyn.cpp:33-97 with the shipped Voronoi drivers
    and its clock-seeded mt19937 stream) in a child process."""
    import sys
    code = ("import ctypes\nlib = ctypes.CDLL(%r)\nargv = %r\n"
            "arr = (ctypes.c_char_p * len(argv))(*[a.encode() for a in argv])\n"
            "lib.orc_ref_swarmdyn_main.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p)]\n"
            "lib.orc_ref_swarmdyn_main(len(argv), arr)\n") % (os.path.join(ROOT, 'oracle', '_ref', 'liborc_ref.so'), argv)
    subprocess.run([sys.executable, '-c', code], cwd=cwd, check=True, stdout=subprocess.DEVNULL, timeout=600)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'liborc_ref.so')), reason='oracle/_ref not built')
def test_c1_statistics_match_the_shipped_reference_binary(tmp_path):
    """Config C1 end to end against the reference's OWN main(): same argv, Voronoi neighbour rule on both
    sides, text output; time- and ensemble-averaged observables of /swarm agree within 4 standard errors."""
    d = P.get_base_params(5, 15, mode='burst_coast')
    P.selection_line_parameter(d, 2)
    d.update(out_h5=0, output_mode=0, N=8)
    cols = {'pol_order': 1, 'avg_speed': 7, 'NND': 13, 'IID': 14}
    ref_runs = []
    for k in range(10):
        rd = tmp_path / ('ref%d' % k)
        rd.mkdir()
        _reference_main(P.dic2swarmdyn_command(d).split(), rd)
        ref_runs.append(_rows(rd / 'swarm_xx.dat'))
    R = 96
    _run(d, tmp_path, ['-Z', '21', '-y', str(R), '-V', '1'])
    ours = [_rows(tmp_path / ('swarm_xx.dat' if r == 0 else 'swarm_xx_%d.dat' % r)) for r in range(R)]
    assert ours[0].shape == ref_runs[0].shape
    for name, c in cols.items():
        a = np.array([run[:, c].mean() for run in ref_runs])
        b = np.array([run[:, c].mean() for run in ours])
        se = np.sqrt(a.var(ddof=1) / a.size + b.var(ddof=1) / b.size)
        assert abs(a.mean() - b.mean()) <= 4 * se + 0.01 * abs(a.mean()), (name, a.mean(), b.mean(), se)


def test_h5_output_matches_text_output(tmp_path):
    """-J 1 (out_xx.h5 through the built-in writer) holds exactly what -J 0 writes as text, in the
    reference's dataset layout: /part (outsteps, N, 7), /pred, /swarm (outsteps, 16), /swarm_fishNet."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import h5mini_read
    d = P.get_base_params(0.5, 2, mode='natPred')
    P.selection_line_parameter(d, 2)
    d.update(output_mode=1, N=30, pred_speed0=20, kill_rate=50, N_confu=40)
    a, b = tmp_path / 'h5', tmp_path / 'txt'
    a.mkdir()
    b.mkdir()
    _run(dict(d, out_h5=1), a, ['-Z', '9'])
    _run(dict(d, out_h5=0), b, ['-Z', '9'])
    f = h5mini_read.File(str(a / 'out_xx.h5'))
    assert sorted(f.links()) == ['part', 'pred', 'swarm', 'swarm_fishNet']
    part, pred, swarm, net = (f.read('/' + k) for k in ('part', 'pred', 'swarm', 'swarm_fishNet'))
    sw_txt = _rows(b / 'swarm_xx.dat')
    assert swarm.shape == sw_txt.shape and part.shape == (swarm.shape[0], 30, 7)
    assert np.allclose(swarm[:, :15], sw_txt[:, :15], rtol=2e-5, atol=1e-6, equal_nan=True)   # text has 6 digits
    assert pred.shape[1:] == (1, 6) and net.shape == (pred.shape[0], 4)
    pred_txt = _rows(b / 'pred_xx.dat')
    assert np.allclose(pred[:, 0, :], pred_txt, rtol=2e-5, atol=1e-6)
    # fileID other than "xx": datasets under /<fileID>/
    c = tmp_path / 'grp'
    c.mkdir()
    _run(dict(d, out_h5=1, fileID='run3'), c, ['-Z', '9'])
    g = h5mini_read.File(str(c / 'out_run3.h5'))
    assert list(g.links()) == ['run3'] and np.array_equal(g.read('/run3/part'), part)


def test_python_batch_helper_and_executable_write_the_same_rows(tmp_path):
    """run_ensemble (sde_burst_coast_b200/ensemble.py) and `swarmdyn -y R -Z seed` drive the same library calls -
    device initial conditions, step schedule, device observables - so their /swarm rows agree to the six significant
    digits of the text files, replicate by replicate; a shard of the ensemble reproduces its replicates bit for bit."""
    from sde_burst_coast_b200.ensemble import run_ensemble
    d = P.get_base_params(0.6, 1.2, mode='natPred')
    P.selection_line_parameter(d, 2)
    d.update(out_h5=0, output_mode=1, N=30, pred_speed0=20, kill_rate=50, N_confu=40, IC=3)
    R, seed = 5, 2024
    _run(d, tmp_path, ['-Z', str(seed), '-y', str(R)])
    res = run_ensemble(d, R, seed=seed, want_particles=True)
    assert res['swarm'].shape[1:] == (R, 16) and res['swarm_fishNet'].shape[2] == 4
    for r in range(R):
        suffix = 'xx' if r == 0 else 'xx_%d' % r
        sw = _rows(tmp_path / ('swarm_%s.dat' % suffix))
        assert sw.shape == res['swarm'][:, r].shape
        mine = res['swarm'][:, r]
        ok = np.isclose(sw, mine, rtol=2e-5, atol=1e-9) | (np.isnan(sw) & np.isnan(mine))
        assert ok.all(), (r, np.argwhere(~ok)[:5])
        net = _rows(tmp_path / ('swarm_fishNet_%s.dat' % suffix))
        assert np.allclose(net, res['swarm_fishNet'][:, r], rtol=2e-5, atol=1e-9)
    shard = run_ensemble(d, 2, seed=seed, replicate_offset=3)
    assert np.array_equal(shard['swarm'], res['swarm'][:, 3:5], equal_nan=True)
