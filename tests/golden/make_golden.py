#!/usr/bin/env python
"""Generate the committed golden fixtures FROM THE REFERENCE ITSELF (run in the build container,
where /root/reference is mounted; the GPU box only sees the fixtures):

  swarmdyn_commands.json  - SwarmDynByPy.get_base_params / selectionLineParameter /
                            dic2swarmdyn_command of the reference (imported from /root/reference)
  parameters/<case>.dat   - parameters.dat written by the reference's own main() (oracle/_ref build,
                            swarmdyn.cpp:33-97 compiled in place) for each of those command lines
  oracle_vectors.npz      - states, three-zone pair results and stepped states computed by the
                            reference's own translation units (liborc_ref.so) on seeded inputs

usage: python tests/golden/make_golden.py
"""
import ctypes
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, '/root/reference')


def commands():
    import SwarmDynByPy as ref
    out = []
    for mode in ref.possible_modes if hasattr(ref, 'possible_modes') else ['burst_coast', 'sinFisher', 'mulFisher', 'natPred', 'natPredNoConfu']:
        for SL in (0, 1, 2):
            try:
                dic = ref.get_base_params(20, 20, mode=mode)
            except Exception as e:  # a mode the reference itself rejects
                out.append(dict(mode=mode, SL=SL, error=str(e)))
                continue
            ref.selectionLineParameter(dic, SL)  # in place, returns None
            cmd = ref.dic2swarmdyn_command(dic)
            out.append(dict(mode=mode, SL=SL, dic={k: (v if not isinstance(v, np.generic) else v.item()) for k, v in dic.items()}, command=cmd))
    return out


def reference_parameters_dat(cmd, outdir):
    """Run the reference main() (renamed by -Dmain=) in a child process: text output, no files but parameters.dat."""
    argv = cmd.split()
    # short, silent run: the parameter echo does not depend on the run length beyond what we keep
    code = r'''
import ctypes, sys
lib = ctypes.CDLL(%r)
argv = %r
arr = (ctypes.c_char_p * len(argv))(*[a.encode() for a in argv])
lib.orc_ref_swarmdyn_main.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p)]
lib.orc_ref_swarmdyn_main(len(argv), arr)
''' % (os.path.join(ROOT, 'oracle', '_ref', 'liborc_ref.so'), argv)
    subprocess.run([sys.executable, '-c', code], cwd=outdir, check=True, stdout=subprocess.DEVNULL, timeout=600)
    return open(os.path.join(outdir, 'parameters.dat')).read()


def short_command(cmd):
    """Same flags, but 0.2 s of simulated time, no output files (-m 2), text mode (-J 0)."""
    t = cmd.split()
    def setf(flag, val):
        t[t.index(flag) + 1] = str(val)
    setf('-t', 0.2); setf('-T', 0.1); setf('-e', 0.1); setf('-m', 2); setf('-J', 0)
    return ' '.join(t)


def oracle_vectors():
    from helpers import make_params, random_state
    from oracle import pyorc
    cases = [('tank8', 'burst_coast', dict(BC=6), 8), ('tank30_alg', 'burst_coast', dict(BC=6, alg_range=10.0), 30),
             ('tank5_64', 'burst_coast', dict(BC=5), 64), ('open40', 'burst_coast', dict(BC=-1), 40),
             ('periodic200', 'natPred', dict(BC=0, size=100.0, Npred=0, alg_range=10.0), 200)]
    out = {}
    for name, mode, over, N in cases:
        sp, _, _ = make_params(mode, N, **over)
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        st = random_state(N, sp, rng, spread=50.0 if sp.BC == -1 else None)
        w = pyorc.World(sp, N, seed=77, replicate=2, kind='reference')
        w.set_state(**st)
        for k, v in st.items():
            out['%s/state/%s' % (name, k)] = v
        for flags in (0, 1):
            res = w.pair_interactions(flags)
            for k, v in res.items():
                out['%s/pair%d/%s' % (name, flags, k)] = v
        us, ud = rng.random(N), np.asarray(rng.random(N), dtype=np.float32).astype(np.float64)
        kn = rng.integers(1, 600, N).astype(np.uint32)
        out['%s/inj/us' % name], out['%s/inj/ud' % name], out['%s/inj/kn' % name] = us, ud, kn
        w.inject(us, ud, kn)
        w.step(0, 1)
        for k, v in w.get_state().items():
            out['%s/step1/%s' % (name, k)] = v
        w.step(1, 199)   # Philox protocol decisions, seed 77, replicate 2
        for k, v in w.get_state().items():
            out['%s/step200/%s' % (name, k)] = v
        w.close()
    return out


def main():
    from oracle import pyorc
    pyorc.build()
    assert pyorc.have_reference(), 'needs /root/reference (oracle/_ref build)'
    cmds = commands()
    os.makedirs(os.path.join(HERE, 'parameters'), exist_ok=True)
    for c in cmds:
        if 'command' not in c:
            continue
        c['short_command'] = short_command(c['command'])
        with tempfile.TemporaryDirectory() as d:
            txt = reference_parameters_dat(c['short_command'], d)
        name = '%s_SL%d' % (c['mode'], c['SL'])
        open(os.path.join(HERE, 'parameters', name + '.dat'), 'w').write(txt)
    json.dump(cmds, open(os.path.join(HERE, 'swarmdyn_commands.json'), 'w'), indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, 'oracle_vectors.npz'), **oracle_vectors())
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
