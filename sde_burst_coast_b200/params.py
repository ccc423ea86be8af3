"""Parameter layer of the burst-coast hot path.

Mirrors the reference's two parameter layers so its driver scripts keep working:
  * the Python presets of SwarmDynByPy.py (`get_base_params` :45-123, `selectionLineParameter`
    :24-42, `dic2swarmdyn_command` :126-171), and
  * the derived quantities of `InitSystemParameters` (src/settings.cpp:79-145) and
    `ParseParameters` (src/input_output.cpp:116-158),
and packs the hot fields into the C-ABI struct `sbc_params` (include/sbc.h).
"""
import ctypes
import math

POSSIBLE_MODES = ['burst_coast', 'sinFisher', 'mulFisher', 'natPred', 'natPredNoConfu', 'exploration']

NEIGHBOUR_METRIC, NEIGHBOUR_VORONOI = 0, 1
PATH_AUTO, PATH_ALLPAIRS, PATH_CELLLIST = 0, 1, 2
PAIR_ACTIVE_ROWS, PAIR_ALL_ROWS = 0, 1


class SbcParams(ctypes.Structure):
    """ctypes image of `sbc_params` (include/sbc.h); field order must match the header."""
    _fields_ = [
        ('sizeL', ctypes.c_double), ('dt', ctypes.c_double), ('beta', ctypes.c_double),
        ('alphaTurn', ctypes.c_double), ('soc_strength', ctypes.c_double),
        ('env_strength', ctypes.c_double), ('prob_social', ctypes.c_double),
        ('burst_rate', ctypes.c_double), ('rep_range', ctypes.c_double),
        ('alg_range', ctypes.c_double), ('att_range', ctypes.c_double),
        ('flee_range', ctypes.c_double), ('pred_time', ctypes.c_double),
        ('pred_speed0', ctypes.c_double), ('kill_range', ctypes.c_double),
        ('kill_rate', ctypes.c_double), ('noisep', ctypes.c_double),
        ('sim_time', ctypes.c_double), ('trans_time', ctypes.c_double),
        ('burst_steps', ctypes.c_uint32), ('BC', ctypes.c_int32), ('N_confu', ctypes.c_int32),
        ('pred_kill', ctypes.c_int32), ('pred_move', ctypes.c_int32),
        ('neighbour_mode', ctypes.c_int32), ('path', ctypes.c_int32), ('reserved', ctypes.c_int32),
    ]


def selection_line_parameter(dic, SL):
    """Parameters in which the three selection lines differ (SwarmDynByPy.py:24-42)."""
    assert SL in (0, 1, 2), 'Unvalid parameter for selection line: {}'.format(SL)
    line = {'burst_duration': [0.089, 0.091, 0.097],
            'burst_rate': [5.7, 6.3, 4.7],
            'prob_social': [0.5, 0.71, 0.74],
            'rep_range': [3.92, 4.44, 4.71],
            'att_range': [15.84, 18.45, 19.5],
            'soc_strength': [215.1, 234.4, 193],
            'env_strength': [215.1, 234.4, 193]}
    for k, v in line.items():
        dic[k] = v[SL]
    dic['alg_range'] = line['rep_range'][SL]  # no orientation zone
    return dic


def get_base_params(pred_time, record_time, mode=None, trans_time=None):
    """Scenario presets with the reference's values (SwarmDynByPy.py:45-123)."""
    mode = 'pred_prey' if mode is None else mode
    trans_time = 0 if trans_time is None else trans_time
    assert mode in POSSIBLE_MODES, 'mode {} not known'.format(mode)
    p = dict(path='./', fileID='xx', out_h5=1, trans_time=pred_time + trans_time,
             time=pred_time + record_time + trans_time, dt=0.001, output=0.03, output_mode=0,
             IC=3, BC=6, N=8, Npred=0, Dphi=0.02, beta=2.51, rep_range=0.5, alg_range=16.2,
             att_range=30, soc_strength=110.82, burst_rate=3.3, flee_range=7, N_confu=4,
             pred_kill=1, pred_move=0, pred_time=pred_time, pred_speed0=15, kill_range=5,
             kill_rate=1, burst_duration=0.121, prob_social=0.8, alphaTurn=1, size=24.5)
    p['env_strength'] = p['soc_strength']
    if mode in ('natPred', 'natPredNoConfu', 'sinFisher', 'mulFisher'):
        p.update(N=30, BC=0, size=100, Npred=1)
        if mode == 'sinFisher':
            p['pred_move'] = 0
        elif mode == 'natPred':
            p.update(pred_move=1, pred_kill=3)
        elif mode == 'natPredNoConfu':
            p.update(pred_move=1, pred_kill=2)
        elif mode == 'mulFisher':
            p['Npred'] = int(p['size'] / 2)
    return p


_FLAG_OF = [('size', 'L', '%g'), ('N', 'N', '%d'), ('Npred', 'n', '%d'), ('Dphi', 'D', '%g'),
            ('beta', 'b', '%g'), ('flee_range', 'f', '%g'), ('pred_time', 'e', '%g'),
            ('pred_speed0', 'S', '%g'), ('rep_range', 'h', '%g'), ('alg_range', 'a', '%g'),
            ('att_range', 'r', '%g'), ('soc_strength', 'H', '%g'), ('burst_rate', 'A', '%g'),
            ('env_strength', 'R', '%g'), ('output', 'o', '%g'), ('out_h5', 'J', '%d'),
            ('dt', 'd', '%g'), ('time', 't', '%g'), ('BC', 'B', '%d'), ('IC', 'I', '%d'),
            ('path', 'l', '%s'), ('output_mode', 'm', '%d'), ('N_confu', 'c', '%d'),
            ('trans_time', 'T', '%g'), ('pred_kill', 'x', '%d'), ('pred_move', 'X', '%d'),
            ('kill_rate', 'O', '%g'), ('kill_range', 'G', '%g'), ('fileID', 'E', '%s'),
            ('prob_social', 'Q', '%g'), ('burst_duration', 'u', '%g'), ('alphaTurn', 'Y', '%g')]


def dic2swarmdyn_command(dic, exe='./swarmdyn'):
    """argv string with the reference's flag letters and order (SwarmDynByPy.py:126-171)."""
    for key, val in dic.items():
        if 'time' in key:
            assert val >= 0, '{} has negative value {}'.format(key, val)
    cmd = exe
    for key, flag, fmt in _FLAG_OF:
        cmd += (' -%s ' + fmt) % (flag, dic[key])
    return cmd


def tank_params(N=8, SL=2, record_time=40):
    """BASELINE config C1/C2: burst-coast circular tank, selection line `SL`."""
    p = get_base_params(0, record_time, mode='burst_coast')
    selection_line_parameter(p, SL)
    p.update(N=N, Npred=0, pred_time=record_time + 2)
    return p


def make_sbc_params(dic, neighbour_mode=NEIGHBOUR_METRIC, path=PATH_AUTO):
    """Dictionary (reference key names) -> `sbc_params`, applying the derivations of
    InitSystemParameters (settings.cpp:79-145)."""
    dt = float(dic['dt'])
    sim_time = float(dic['time'])
    trans_time = float(dic.get('trans_time', 0))
    pred_time = float(dic.get('pred_time', 0))
    npred = int(dic.get('Npred', 0))
    if trans_time >= sim_time:
        trans_time = sim_time - max(float(dic.get('output', dt)), dt)
    if trans_time < 0:
        trans_time = 0
    if pred_time >= sim_time or pred_time < 0:
        npred = 0
    if npred == 0:
        pred_time = sim_time + 2
    sizeL = float(dic['size'])
    BC = int(dic['BC'])
    if BC < 0:
        sizeL = float(int(dic['N']) ** 2)
    sp = SbcParams()
    sp.sizeL = sizeL
    sp.dt = dt
    sp.beta = float(dic['beta'])
    sp.alphaTurn = float(dic['alphaTurn'])
    sp.soc_strength = float(dic['soc_strength'])
    sp.env_strength = float(dic['env_strength'])
    sp.prob_social = float(dic['prob_social'])
    sp.burst_rate = float(dic['burst_rate'])
    sp.rep_range = float(dic['rep_range'])
    sp.alg_range = float(dic['alg_range'])
    sp.att_range = float(dic['att_range'])
    sp.flee_range = float(dic['flee_range'])
    sp.pred_time = pred_time
    sp.pred_speed0 = float(dic['pred_speed0'])
    sp.kill_range = float(dic['kill_range'])
    sp.kill_rate = float(dic['kill_rate'])
    sp.noisep = math.sqrt(2 * dt * float(dic['Dphi']))
    sp.sim_time = sim_time
    sp.trans_time = trans_time
    sp.burst_steps = int(float(dic['burst_duration']) / dt)
    sp.BC = BC
    sp.N_confu = int(dic['N_confu'])
    sp.pred_kill = int(dic['pred_kill'])
    sp.pred_move = int(dic['pred_move'])
    sp.neighbour_mode = neighbour_mode
    sp.path = path
    return sp, npred
