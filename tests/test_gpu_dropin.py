"""T6 as a pytest: tests/t6_dropin.py in a fresh interpreter (the splice replaces modules in sys.modules, which must
not leak into the other tests).  Skipped where the reference copy did not travel (baseline/_ref)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_unmodified_reference_callers_run_on_the_device_path():
    sys.path.insert(0, HERE)
    import t6_dropin
    if not t6_dropin.available():
        pytest.skip('no reference copy under baseline/_ref (made by __graft_entry__.build() in the build container)')
    r = subprocess.run([sys.executable, os.path.join(HERE, 't6_dropin.py')], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:]+r.stderr[-3000:]
    assert 'failures: none' in r.stdout
