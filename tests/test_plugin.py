"""The drop-in boundary exercised for real: the unmodified reference build (baseline/_ref) runs with this
repository's CUDA pair-force path plugged in as a HOOMD-blue component (plugin/hoomd_b200_plugin, built by
baseline/build_plugin.sh like example_plugins/pair_plugin) and reproduces its own hoomd.md.pair.LJ.
Needs the reference install and the plugin binary (both travel to the GPU box as built files); skipped
where they were not built."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    return (os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "hoomd"))
            and glob.glob(os.path.join(ROOT, "plugin", "hoomd_b200_plugin", "_b200_plugin*.so")))


def test_plugin_sources_bind_declared_symbols():
    """Every symbol the component binds with dlsym is declared in include/hoomd_b200.h."""
    import re
    src = open(os.path.join(ROOT, "plugin", "hoomd_b200_plugin", "module.cc")).read()
    hdr = open(os.path.join(ROOT, "include", "hoomd_b200.h")).read()
    names = re.findall(r'HB2_SYM\(\w+, "(hb2_\w+)"\)', src)
    assert len(names) >= 15
    for n in names:
        assert re.search(r"\b" + n + r"\(", hdr), n


@pytest.mark.gpu
def test_reference_runs_with_the_b200_force():
    if not _built():
        pytest.skip("reference install or plugin binary not built (baseline/build_ref.sh, baseline/build_plugin.sh)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "plugin_check.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "plugin ok" in r.stdout
