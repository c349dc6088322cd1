"""Runs the reference's OWN test files, unmodified, against an implementation of the env API.

    python tests/reference_suite.py --target dropin      (needs a GPU: wurm_b200 behind the names the tests import)
    python tests/reference_suite.py --target reference   (CPU: the reference itself under the oracle/reference_loader shims)

The test files are read where the reference lies (`$WURM_REFERENCE_PATH`, `/root/reference`, or the git-ignored copy
`baseline/_ref/` that `__graft_entry__.build()` ships to the GPU box) -- they are never copied into this repo's history.
They import `wurm.envs`, `wurm.utils`, `wurm.vis`, `config` and `matplotlib.pyplot`; for the drop-in run those names
are bound to `wurm_b200.envs`, `wurm_b200.utils`, `wurm_b200.config` and empty stand-ins for the two plotting
modules (`visualise = False` in the files: never called).  Nothing else is touched: classes, keyword arguments,
attribute writes (`env.envs = ...`, `env.heads[...] = 1`, `env.boost_cost_prob = 0`), return structures and exception
types are exercised exactly as the reference's authors wrote them.

Prints one JSON object {"file::test": "ok" | "FAIL: ...", ...}.  Run in a process of its own (module aliases, and for
--target reference the torch shims, are process-wide); tests/test_reference_suite*.py do that.
"""
import argparse
import importlib.util
import io
import json
import os
import sys
import types
import unittest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FILES = ['test_single_snake_env.py', 'test_multi_snake_env.py', 'test_simple_gridworld.py']


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__path__ = []
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def bind_dropin():
    import wurm_b200.config
    import wurm_b200.envs
    import wurm_b200.utils
    wurm = _stub('wurm')
    sys.modules['wurm.envs'] = wurm.envs = wurm_b200.envs
    sys.modules['wurm.utils'] = wurm.utils = wurm_b200.utils
    sys.modules['wurm.vis'] = wurm.vis = _stub('wurm.vis', plot_envs=lambda *a, **k: None)
    sys.modules['config'] = wurm_b200.config
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except ImportError:
            mpl = _stub('matplotlib')
            mpl.pyplot = _stub('matplotlib.pyplot')


def bind_reference():
    from oracle import reference_loader as rl
    rl.load(record=False, device='cpu')          # installs the shims, stubs gym / matplotlib, puts the tree on sys.path
    import wurm.vis  # noqa: F401  (imports matplotlib.pyplot: stubbed)


def run(target):
    from oracle import reference_loader as rl
    ref = rl.find_reference()
    if ref is None:
        return {'error': 'reference tree not found ($WURM_REFERENCE_PATH, /root/reference, baseline/_ref)'}
    (bind_dropin if target == 'dropin' else bind_reference)()
    results = {}
    for fname in FILES:
        path = os.path.join(ref, 'tests', fname)
        spec = importlib.util.spec_from_file_location('reference_' + fname[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        try:
            spec.loader.exec_module(mod)
        except Exception as exc:
            results[fname + '::import'] = f'FAIL: {type(exc).__name__}: {exc}'
            continue
        suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
        for case in suite:
            for test in case:
                name = f'{fname}::{test._testMethodName}'
                res = unittest.TestResult()
                out = io.StringIO()
                saved = sys.stdout
                sys.stdout = out                      # the reference's tests print a lot
                try:
                    test.run(res)
                finally:
                    sys.stdout = saved
                bad = res.errors + res.failures
                results[name] = 'ok' if not bad else 'FAIL: ' + bad[0][1].strip().splitlines()[-1][:300]
                if bad and os.environ.get('WURM_SUITE_VERBOSE'):
                    sys.stderr.write(f'--- {name}\n{bad[0][1]}\n')
    return results


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--target', choices=['dropin', 'reference'], required=True)
    args = ap.parse_args()
    print(json.dumps(run(args.target)))
