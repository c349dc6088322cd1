"""CPU: the reference's own test files run against the REFERENCE ITSELF under the five interpreter shims of
oracle/reference_loader.py -- the evidence that the shims restore the semantics the reference was written for (and the
pass set the drop-in is held to in tests/test_reference_suite_gpu.py).  Skipped where no reference tree exists."""
import pytest

from test_reference_suite_gpu import REFERENCE_TESTS, run_suite

# needs gym's SimpleImageViewer even for mode='rgb_array' (multi_snake.py:229-231); gym is not installed
NEEDS_GYM = {'test_multi_snake_env.py::test_boost_rendering'}


@pytest.fixture(scope='module')
def results():
    from oracle import reference_loader as rl
    if rl.find_reference() is None:
        pytest.skip('no reference tree on this machine')
    return run_suite('reference')


@pytest.mark.parametrize('name', [n for n in REFERENCE_TESTS if n not in NEEDS_GYM])
def test_reference_passes_its_own_test_under_the_shims(results, name):
    assert results.get(name) == 'ok', results.get(name)
